"""Run-to-run determinism probe: the same forward repeated must be bit-identical.  KDLAE_FUSE_PWDW is read per forward."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
import rethink_acoustic_image_enhancement_b200 as pk

kw = dict(inp_channels=1, out_channels=1, LayerNorm_type="BiasFree", static="train")
sd = synth.teacher_state_dict(seed=0, temp_scale=4.0, **kw)
m = pk.KDLAE_teacher(**kw); m.load_state_dict(sd); m = m.cuda().eval().set_precision("bf16")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
R = int(sys.argv[3]) if len(sys.argv) > 3 else 12
img = synth.seeded_tensor("probe.img", (B, 1, S, S), 0, "sonar").cuda()
x = {"img": img, "denoise_rate": torch.full((B, 1, 1, 1), 0.6, device="cuda")}
m.micro_batch = B
for mode in ("7", "0", "3", "6", "9"):
    os.environ["KDLAE_FUSE_PWDW"] = mode
    with torch.no_grad():
        base = m(x)
        bad = {"hq": 0, "sr": 0}
        mx = {"hq": 0.0, "sr": 0.0}
        for _ in range(R):
            out = m(x)
            for k in bad:
                d = float((out[k] - base[k]).abs().max())
                bad[k] += d > 0
                mx[k] = max(mx[k], d)
    print(f"S={S} B={B} mode={mode} mismatching runs of {R}: {bad} max {mx}", flush=True)
