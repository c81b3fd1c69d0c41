"""Which part of the teacher workspace is read before it is written?  Poison one slice at a time with NaN bytes."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
import rethink_acoustic_image_enhancement_b200 as pk

kw = dict(inp_channels=1, out_channels=1, LayerNorm_type="BiasFree", static="train")
sd = synth.teacher_state_dict(seed=0, temp_scale=4.0, **kw)
m = pk.KDLAE_teacher(**kw); m.load_state_dict(sd); m = m.cuda().eval().set_precision("bf16")
S = 512
img = synth.seeded_tensor("probe.img", (1, 1, S, S), 0, "sonar").cuda()
x = {"img": img, "denoise_rate": torch.full((1, 1, 1, 1), 0.6, device="cuda")}
m.micro_batch = 1
with torch.no_grad():
    m(x)
    (ws,) = m._engine._ws.values()
    ws.zero_()
    base = m(x)
    n = ws.numel()
    # layout (teacher.cu ws_layout, mb = 1, bf16): element counts * 2 bytes, 256-aligned
    P1, d = S * S, 48
    names = ["x1", "d1", "x2", "x3", "x4", "d3", "d2", "s0", "bufA", "bufB", "o1", "rstd", "mu", "gram", "mb"]
    sizes = [P1 * d * 2, P1 * 2 * d * 2, P1 // 4 * 2 * d * 2, P1 // 16 * 4 * d * 2, P1 // 64 * 8 * d * 2, P1 // 16 * 4 * d * 2,
             P1 // 4 * 2 * d * 2, P1 * 4 * d * 2]
    print("workspace bytes", n, "first buffers end at", sum(sizes))
    K = 64
    hits = []
    for i in range(K):
        ws.zero_()
        lo, hi = n * i // K, n * (i + 1) // K
        ws[lo:hi].fill_(0xFF)
        out = m(x)
        dh, ds = float((out["hq"] - base["hq"]).abs().max()), float((out["sr"] - base["sr"]).abs().max())
        if dh > 0 or ds > 0 or not torch.isfinite(out["sr"]).all():
            hits.append((i, lo, hi, dh, ds))
    print(json.dumps(hits))
