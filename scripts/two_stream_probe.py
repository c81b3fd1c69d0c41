"""Two KDLAE-T forwards side by side on two CUDA streams (each on KDLAE_SM_LIMIT SMs) vs one after the other on all SMs.
Run as:  KDLAE_SM_LIMIT=74 python scripts/two_stream_probe.py   and   python scripts/two_stream_probe.py --serial"""
import os, sys, json, argparse, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rethink_acoustic_image_enhancement_b200 as pk
from oracle import synth

ap = argparse.ArgumentParser()
ap.add_argument("--serial", action="store_true")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--offset-ms", type=float, default=0.0)
args = ap.parse_args()
KW = dict(inp_channels=1, out_channels=1, LayerNorm_type="BiasFree", static="train", params="cat")
sd = synth.teacher_state_dict(seed=0, **{k: v for k, v in KW.items() if k != "params"})
models = []
for _ in range(2):
    m = pk.KDLAE_teacher(**KW); m.load_state_dict(sd); m = m.cuda().eval().set_precision("bf16"); m.micro_batch = 8
    models.append(m)
xs = [torch.rand(args.batch, 1, 512, 512, device="cuda") for _ in range(2)]
rate = torch.rand(args.batch, 1, 1, 1, device="cuda")
streams = [torch.cuda.Stream(), torch.cuda.Stream()]

def step():
    if args.serial:
        for m, x in zip(models, xs):
            m({"img": x, "denoise_rate": rate})
    else:
        cur = torch.cuda.current_stream()
        for m, x, s in zip(models, xs, streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                if s is streams[1] and args.offset_ms > 0:
                    torch.cuda._sleep(int(args.offset_ms * 1.9e6))      # bounded delay: de-phase the two forwards
                m({"img": x, "denoise_rate": rate})
        for s in streams:
            cur.wait_stream(s)

with torch.no_grad():
    for _ in range(2): step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(3): step()
    b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print(json.dumps({"mode": "serial" if args.serial else "two streams", "sm_limit": os.environ.get("KDLAE_SM_LIMIT"), "offset_ms": args.offset_ms,
                  "images_per_s": 2 * args.batch / ms * 1e3, "ms": ms}))
