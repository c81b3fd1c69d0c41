"""Per-launch CUDA-event times of one KDLAE-S / ASDQE / KDLAE-T forward (kdlae_profile_launches): class, ms, GB/s, TFLOP/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rethink_acoustic_image_enhancement_b200 as pk
from rethink_acoustic_image_enhancement_b200 import _lib
from oracle import synth
which = sys.argv[1] if len(sys.argv) > 1 else "student"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
with torch.no_grad():
    if which == "student":
        m = pk.KDLAE_student(residual=True); m.load_state_dict(synth.student_state_dict()); m = m.cuda().eval().set_precision("bf16")
        x = torch.rand(B, 5, 512, 512, device="cuda"); f = lambda: m(x)
    elif which == "asdqe":
        m = pk.DenoiseRatePredictor(); m.load_state_dict(synth.asdqe_state_dict(), strict=False); m = m.cuda().eval().set_precision("bf16")
        lq, gt = torch.rand(B, 3, 512, 512, device="cuda"), torch.rand(B, 3, 512, 512, device="cuda"); f = lambda: m(lq, gt)
    else:
        kw = dict(inp_channels=1, out_channels=1, LayerNorm_type="BiasFree", static="train")
        m = pk.KDLAE_teacher(**kw); m.load_state_dict(synth.teacher_state_dict(seed=0, **kw)); m = m.cuda().eval().set_precision("bf16")
        x = {"img": torch.rand(B, 1, 512, 512, device="cuda"), "denoise_rate": torch.full((B, 1, 1, 1), 0.6, device="cuda")}; f = lambda: m(x)
    m.micro_batch = B
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize()
    total = e0.elapsed_time(e1)
    _lib.profile_begin(); f(); recs = _lib.profile_launches(); _lib.profile_end()
print(f"{which} batch {B}: {total:.3f} ms per forward ({B / total * 1e3:.1f} units/s); {len(recs)} launches, sum {sum(r[1] for r in recs):.3f} ms")
for i, (c, ms, fl, by) in enumerate(recs):
    print(f"{i:3d} {c:34s} {ms * 1e3:9.1f} us  {by / ms / 1e6 if ms > 0 else 0:8.0f} GB/s  {fl / ms / 1e9 if ms > 0 else 0:8.1f} TFLOP/s")
