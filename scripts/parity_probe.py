"""Parity of the CUDA KDLAE-T path against the CPU oracle at the BENCHMARK shape (1x512x512, static='train').

    python scripts/parity_probe.py [--size 512] [--temp 4.0] [--kind sonar] [--fp32]

Prints PSNR / max-abs of hq and sr for the bf16 path (and optionally the fp32 path).  Test infrastructure: imports oracle/.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import oracle  # noqa: E402
from oracle import synth  # noqa: E402
import rethink_acoustic_image_enhancement_b200 as pk  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--temp", type=float, default=4.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--kind", default="sonar")
    ap.add_argument("--ln", default="BiasFree")
    ap.add_argument("--channels", type=int, default=1)
    ap.add_argument("--fp32", action="store_true")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    kw = dict(inp_channels=a.channels, out_channels=a.channels, LayerNorm_type=a.ln, static="train")
    sd = synth.teacher_state_dict(seed=a.seed, temp_scale=a.temp, **kw)
    S = a.size
    img = synth.seeded_tensor("probe.img", (1, a.channels, S, S), a.seed, a.kind)
    rate = torch.full((1, 1, S, S), 0.6)
    t0 = time.perf_counter()
    with torch.no_grad():
        hq_ref, sr_ref = oracle.teacher_forward(sd, img, rate)
    t_cpu = time.perf_counter() - t0
    m = pk.KDLAE_teacher(**kw)
    m.load_state_dict(sd)
    m = m.to("cuda:0").eval()
    res = {"size": S, "temp": a.temp, "kind": a.kind, "ln": a.ln, "channels": a.channels, "oracle_s": t_cpu,
           "cores": os.cpu_count()}
    for prec in (["bf16", "fp32"] if a.fp32 else ["bf16"]):
        with torch.no_grad():
            out = m.set_precision(prec)({"img": img.cuda(), "denoise_rate": rate.cuda()})
        torch.cuda.synchronize()
        hq, sr = out["hq"].cpu(), out["sr"].cpu()
        res[prec] = {"psnr_hq": synth.psnr(hq, hq_ref), "psnr_sr": synth.psnr(sr, sr_ref),
                     "maxabs_hq": (hq - hq_ref).abs().max().item(), "maxabs_sr": (sr - sr_ref).abs().max().item(),
                     "finite": bool(torch.isfinite(hq).all() and torch.isfinite(sr).all())}
    print(json.dumps(res))
    if a.out:
        with open(a.out, "a") as fh:
            fh.write(json.dumps(res) + "\n")


if __name__ == "__main__":
    main()
