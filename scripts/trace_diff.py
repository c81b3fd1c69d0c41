"""Stage-by-stage checksum diff of two KDLAE-T forwards that should be bit-identical (kdlae_debug_trace_*)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
import rethink_acoustic_image_enhancement_b200 as pk
from rethink_acoustic_image_enhancement_b200 import _lib
lib = _lib.load()

def traced(f):
    _lib.check(lib.kdlae_debug_trace_begin(), "trace_begin")
    out = f()
    n = 8192
    sums = (C.c_ulonglong * (2 * n))()
    tags = C.create_string_buffer(1 << 18)
    k = lib.kdlae_debug_trace_end(sums, n, tags, len(tags))
    return out, [(t, sums[2 * i], sums[2 * i + 1]) for i, t in enumerate(tags.value.decode().split("\n")[:k])]

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
kw = dict(inp_channels=1, out_channels=1, LayerNorm_type="BiasFree", static="train")
sd = synth.teacher_state_dict(seed=0, temp_scale=4.0, **kw)
m = pk.KDLAE_teacher(**kw); m.load_state_dict(sd); m = m.cuda().eval().set_precision("bf16")
img = synth.seeded_tensor("probe.img", (1, 1, S, S), 0, "sonar").cuda()
x = {"img": img, "denoise_rate": torch.full((1, 1, 1, 1), 0.6, device="cuda")}
m.micro_batch = 1
with torch.no_grad():
    m(x)
    (ws,) = m._engine._ws.values()
    runs = []
    for fill in (0, 0xFF, 0, 0xFF):
        ws.fill_(fill)
        out, tr = traced(lambda: m(x))
        runs.append((fill, out, tr))
base = runs[0][2]
for fill, out, tr in runs[1:]:
    bad = [(i, a[0]) for i, (a, b) in enumerate(zip(base, tr)) if a != b]
    print(f"fill {fill:#x}: {len(tr)} trace points, first differing: {bad[:6]}, total differing {len(bad)}; "
          f"hq diff {float((out['hq'] - runs[0][1]['hq']).abs().max()):.3e}")
    if bad:
        i = bad[0][0]
        print("  context:", [t[0] for t in base[max(0, i - 12):i + 1]], "index", i)
