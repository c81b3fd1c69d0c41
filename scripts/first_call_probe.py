"""First forward of a process vs later ones (same input): must be bit-identical."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
import rethink_acoustic_image_enhancement_b200 as pk
S, B = int(sys.argv[1]), int(sys.argv[2])
kw = dict(inp_channels=1, out_channels=1, LayerNorm_type="BiasFree", static="train")
sd = synth.teacher_state_dict(seed=0, temp_scale=4.0, **kw)
m = pk.KDLAE_teacher(**kw); m.load_state_dict(sd); m = m.cuda().eval().set_precision(sys.argv[3] if len(sys.argv) > 3 else "bf16")
img = synth.seeded_tensor("probe.img", (B, 1, S, S), 0, "sonar").cuda()
x = {"img": img, "denoise_rate": torch.full((B, 1, 1, 1), 0.6, device="cuda")}
m.micro_batch = B
with torch.no_grad():
    outs = [m(x) for _ in range(4)]
    torch.cuda.synchronize()
for i in range(1, 4):
    print(f"S={S} B={B} mode={os.environ.get('KDLAE_FUSE_PWDW', '6')} run{i} vs run0: hq {float((outs[i]['hq'] - outs[0]['hq']).abs().max()):.3e} "
          f"sr {float((outs[i]['sr'] - outs[0]['sr']).abs().max()):.3e}", flush=True)
