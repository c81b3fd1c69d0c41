"""Whole KDLAE-T training step (BASELINE configs[4]) through the module in train() mode.

    python scripts/train_step_bench.py [--batch 2] [--size 128] [--steps 5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/train_step_bench.py

One step = what ImageCleanModel.optimize_parameters does (image_restoration_model.py:198-224) on the full-size teacher
(dim 48, blocks 4/6/6/8, 1 channel, BiasFree, static='train'; 26.9 M parameters): zero_grad, forward on a batch of lq crops with a
denoise-rate map, L1LossSr against hq and 2x sr targets, backward, clip_grad_norm_(0.01) + AdamW - and, with N > 1 ranks, the DDP
gradient all-reduce (25 MB buckets on a communication stream, launched from gradient hooks while the backward still runs).  Every
arithmetic kernel is from libkdlae_b200.so (fp32 CUDA-core kernels by default, --tf32 puts the 1x1-conv GEMMs on tcgen05: DESIGN.md).  KDLAET.yml trains
on 128 x 128 crops with 1-6 crops per GPU.  Timed on the device with CUDA events, max over ranks; rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import rethink_acoustic_image_enhancement_b200 as pk  # noqa: E402
from oracle import synth  # noqa: E402
from rethink_acoustic_image_enhancement_b200 import _lib  # noqa: E402
from rethink_acoustic_image_enhancement_b200.metrics import L1LossSr  # noqa: E402
from rethink_acoustic_image_enhancement_b200.training import BucketedAllReducer, FlatAdamW, get_matmul_precision, set_matmul_precision  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--tf32", action="store_true", help="1x1-conv GEMMs (forward, dgrad, wgrad) on tcgen05 in TF32 (training.set_matmul_precision)")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        saved = os.dup(1); os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier(); torch.cuda.synchronize()
        sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    if a.tf32:
        set_matmul_precision("tf32")
    kw = dict(inp_channels=1, out_channels=1, LayerNorm_type="BiasFree", static="train")
    m = pk.KDLAE_teacher(**kw)
    m.load_state_dict(synth.teacher_state_dict(seed=0, temp_scale=1.0, **kw), strict=True)
    m = m.to(dev).train()
    params = list(m.parameters())
    opt = FlatAdamW(params, lr=1e-5, weight_decay=0.5e-4)                   # KDLAET.yml optim_g
    red = BucketedAllReducer(opt.grad, bucket_bytes=25 << 20)
    red.attach(opt.params, opt.offsets)
    crit = L1LossSr(loss_weight=1.0)
    B, S = a.batch, a.size
    img = synth.seeded_tensor(f"train.img.{rank}", (B, 1, S, S), 1, "sonar").to(dev)
    rate = torch.rand(B, 1, 1, 1, generator=torch.Generator().manual_seed(rank)).expand(B, 1, S, S).contiguous().to(dev)
    gt = {"hq": synth.seeded_tensor(f"train.hq.{rank}", (B, 1, S, S), 2, "sonar").to(dev),
          "sr": synth.seeded_tensor(f"train.sr.{rank}", (B, 1, 2 * S, 2 * S), 3, "sonar").to(dev)}
    losses = []

    def step():
        opt.zero_grad()
        out = m({"img": img, "denoise_rate": rate})
        loss = crit(out, gt)
        loss.backward()                    # gradient hooks launch the buckets on the communication stream
        red.wait()
        opt.step()
        losses.append(loss.detach())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    barrier()
    n0 = _lib.load().kdlae_launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(a.steps):
        step()
    t1.record()
    barrier()
    ms = torch.tensor([t0.elapsed_time(t1) / a.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = (_lib.load().kdlae_launch_count() - n0) // a.steps
    if rank == 0:
        ls = [float(l) for l in losses]
        print(json.dumps({"metric": "KDLAE-T training step images/sec (CUDA forward + backward + clip + AdamW" +
                          (" + bucketed NCCL all-reduce)" if world > 1 else ")"),
                          "value": round(B * world / (float(ms) / 1e3), 3), "unit": "images/s", "n_gpus": world, "ms_per_step": round(float(ms), 2),
                          "per_gpu_batch": B, "crop": S, "dtype": "f32", "matmul": get_matmul_precision(), "params": sum(p.numel() for p in params),
                          "grad_bytes": opt.grad.numel() * 4, "buckets": len(red.bounds), "gpu_launches_per_step": int(launches),
                          "loss_first": round(ls[0], 6), "loss_last": round(ls[-1], 6),
                          "peak_mem_gb": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 2)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
