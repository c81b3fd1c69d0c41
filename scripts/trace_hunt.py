"""Hunt for sporadic nondeterminism: many traced forwards of the same input; report the first stage whose checksum deviates."""
import ctypes as C, os, sys, collections
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

S, B, R = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
kw = dict(inp_channels=1, out_channels=1, LayerNorm_type="BiasFree", static="train")
sd = synth.teacher_state_dict(seed=0, temp_scale=4.0, **kw)
m = pk.KDLAE_teacher(**kw); m.load_state_dict(sd); m = m.cuda().eval().set_precision("bf16")
img = synth.seeded_tensor("probe.img", (B, 1, S, S), 0, "sonar").cuda()
x = {"img": img, "denoise_rate": torch.full((B, 1, 1, 1), 0.6, device="cuda")}
m.micro_batch = B
junk = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
side = torch.cuda.Stream()
ma = torch.randn(4096, 4096, device="cuda", dtype=torch.bfloat16)
traces = []
with torch.no_grad():
    for i in range(R):
        if i % 3 == 1:
            junk.fill_(i & 0xFF)
        if i % 4 >= 2:
            with torch.cuda.stream(side):
                for _ in range(20):
                    torch.mm(ma, ma)
        if i % 5 == 4:
            torch.cuda.synchronize()
        _, tr = traced(lambda: m(x))
        traces.append(tr)
keys = [tuple((a, b) for _, a, b in t) for t in traces]
cnt = collections.Counter(keys)
major = cnt.most_common(1)[0][0]
print(f"S={S} B={B}: {len(cnt)} distinct traces over {R} runs; majority count {cnt[major]}")
for r, k in enumerate(keys):
    if k != major:
        i = next(j for j, (a, b) in enumerate(zip(k, major)) if a != b)
        tags = [t[0] for t in traces[r]]
        # which block: count "blk.in" before i
        nb = sum(1 for t in tags[:i + 1] if t == "blk.in")
        print(f"  run {r}: first deviating point {i} = {tags[i]} (block #{nb}); previous points {tags[max(0, i - 3):i]}; deviating points total "
              f"{sum(1 for a, b in zip(k, major) if a != b)}")
