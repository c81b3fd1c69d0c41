"""One ASDQE forward (64 pairs, 3x512x512, bf16) for an ncu launch list."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rethink_acoustic_image_enhancement_b200 as pk
from oracle import synth
a = pk.DenoiseRatePredictor(); a.load_state_dict(synth.asdqe_state_dict(), strict=False); a = a.cuda().eval().set_precision("bf16")
lq = torch.rand(64, 3, 512, 512, device="cuda"); gt = torch.rand(64, 3, 512, 512, device="cuda")
with torch.no_grad():
    for _ in range(2): a(lq, gt)
torch.cuda.synchronize()
