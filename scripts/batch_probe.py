"""Batch-position / workspace-content invariance probe: identical images must give bit-identical outputs at every batch position,
whatever the (uninitialised) workspace held before the call."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
import rethink_acoustic_image_enhancement_b200 as pk

kw = dict(inp_channels=1, out_channels=1, LayerNorm_type="BiasFree", static="train")
sd = synth.teacher_state_dict(seed=0, temp_scale=4.0, **kw)
m = pk.KDLAE_teacher(**kw); m.load_state_dict(sd); m = m.cuda().eval().set_precision("bf16")

def run(img, B, fill):
    S = img.shape[-1]
    m.micro_batch = B
    x = {"img": img.expand(B, 1, S, S).contiguous(), "denoise_rate": torch.full((B, 1, 1, 1), 0.6, device="cuda")}
    with torch.no_grad():
        m(x)                                    # sizes the workspace
        for ws in m._engine._ws.values():
            ws.fill_(fill)
        return m(x)

for S in (128, 512):
    img = synth.seeded_tensor("probe.img", (1, 1, S, S), 0, "sonar").cuda()
    base = run(img, 1, 0)
    for B, fill in ((1, 0xFF), (1, 0x3C), (4, 0), (4, 0xFF), (5, 0x3C)):
        out = run(img, B, fill)
        res = {}
        for k in ("hq", "sr"):
            res[k] = [float((out[k][b] - base[k][0]).abs().max()) for b in range(B)]
            res[k + "_finite"] = bool(torch.isfinite(out[k]).all())
        print(S, B, hex(fill), json.dumps(res), flush=True)
