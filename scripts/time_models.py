"""Device-resident throughput of KDLAE-S and ASDQE (BASELINE configs 3 and 4 shapes, reduced batch) on one B200."""
import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rethink_acoustic_image_enhancement_b200 as pk
from rethink_acoustic_image_enhancement_b200 import _lib
from oracle import synth

def timeit(f, n=3):
    for _ in range(2): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

out = {}
for prec in ("bf16",):
    s = pk.KDLAE_student(residual=True); s.load_state_dict(synth.student_state_dict()); s = s.cuda().eval().set_precision(prec)
    x = torch.rand(32, 5, 512, 512, device="cuda")
    with torch.no_grad():
        ms = timeit(lambda: s(x))
        _lib.profile_begin(); s(x); prof = _lib.profile_end()
    out[f"student_{prec}"] = dict(stacks_per_s=32 / ms * 1e3, ms=ms, tflops=32 * 1.4881e11 / ms / 1e9,
                                  classes={k: round(v["ms"], 2) for k, v in prof.items()})
    a = pk.DenoiseRatePredictor(); a.load_state_dict(synth.asdqe_state_dict(), strict=False); a = a.cuda().eval().set_precision(prec)
    lq = torch.rand(64, 3, 512, 512, device="cuda"); gt = torch.rand(64, 3, 512, 512, device="cuda")
    with torch.no_grad():
        ms = timeit(lambda: a(lq, gt))
        _lib.profile_begin(); a(lq, gt); prof = _lib.profile_end()
    out[f"asdqe_{prec}"] = dict(pairs_per_s=64 / ms * 1e3, ms=ms, tflops=64 * 2.1368e11 / ms / 1e9,
                                classes={k: round(v["ms"], 2) for k, v in prof.items()})
print(json.dumps(out, indent=1))
