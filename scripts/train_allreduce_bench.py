"""DDP gradient all-reduce of the KDLAE-T training step (BASELINE configs[4]) over NCCL / NVLink.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/train_allreduce_bench.py

The flat fp32 gradient buffer has the size of the whole KDLAE-T parameter set (26 874 300 floats = 107.5 MB).  Measured, per step,
on the device (CUDA events, max over ranks):
  * backward_ms   - forward + backward of a stack of CUDA-trained TransformerBlocks (training.transformer_block_train, C = 96, 256x256
                    crops) whose parameters own the TAIL of the flat buffer (the rest stands in for the layers whose backward is not
                    built yet and is all-reduced as soon as the step starts);
  * allreduce_ms  - the bucketed all-reduce alone (25 MB buckets, ReduceOp.AVG), bus bandwidth = 2 (N-1)/N * bytes / time;
  * overlap_ms    - both together: buckets are launched on a communication stream from the gradient hooks during backward.
Rank 0 prints one JSON line.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import synth  # noqa: E402
from rethink_acoustic_image_enhancement_b200.training import BucketedAllReducer, FlatAdamW, transformer_block_train  # noqa: E402
from rethink_acoustic_image_enhancement_b200.metrics import L1LossSr  # noqa: E402

N_PARAMS = 26874300          # KDLAE_teacher(inp=out=1, static='train') (tests/test_host.py)


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        saved = os.dup(1); os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier(); torch.cuda.synchronize()
        sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    C, heads, nblk, B, S = 96, 1, 2, 2, 256
    sd = {}
    for i in range(nblk):
        synth._block(sd, f"blk{i}", C, 2.66, False, False, seed=i, temp_scale=2.0, heads=heads)
    params = {k: torch.nn.Parameter(v.to(dev)) for k, v in sd.items()}
    n_blk = sum(p.numel() for p in params.values())
    filler = torch.nn.Parameter(torch.zeros(N_PARAMS - n_blk, device=dev))         # the layers without a CUDA backward yet
    opt = FlatAdamW([filler] + list(params.values()))                                # block parameters own the tail of the buffer
    red = BucketedAllReducer(opt.grad, bucket_bytes=25 << 20)
    blk_params = list(params.values())
    red.attach(blk_params, opt.offsets[1:])
    filler_buckets = [b for b, (lo, hi) in enumerate(red.bounds) if red._need[b] == 0]
    crit = L1LossSr()
    g = torch.Generator(device="cpu").manual_seed(rank)
    x = torch.randn(B, C, S, S, generator=g).to(dev)
    gt = torch.rand(B, C, S, S, generator=g).to(dev)

    def fwd_bwd():
        h = x
        for i in range(nblk):
            h = transformer_block_train(h, params, f"blk{i}")
        loss = crit({"hq": h, "sr": None}, {"hq": gt})
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(f, n=5):
        for _ in range(2):
            f()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            f()
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b) / n], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # 1. backward alone (hooks detached: no communication)
    for hnd in red._hooks:
        hnd.remove()
    red._hooks.clear()

    def step_compute():
        opt.zero_grad()
        fwd_bwd()
    backward_ms = timed(step_compute)

    # 2. all-reduce alone
    def step_comm():
        red.all_reduce()
    allreduce_ms = timed(step_comm)

    # 3. overlapped: filler buckets go out at the start of the step, block buckets from the gradient hooks
    red.attach(blk_params, opt.offsets[1:])

    def step_overlap():
        opt.zero_grad()
        for b in filler_buckets:
            red._launch(b)
        fwd_bwd()
        red.wait()
        opt.step()
    overlap_ms = timed(step_overlap)
    loss = float(fwd_bwd().item()); red.wait()
    if rank == 0:
        nbytes = opt.grad.numel() * 4
        bus = (2.0 * (world - 1) / world) * nbytes / (allreduce_ms / 1e3) / 1e9 if world > 1 else None
        print(json.dumps({"bench": "KDLAE-T training-step gradient all-reduce (flat fp32, 25 MB buckets, NCCL AVG)", "n_gpus": world,
                          "gradient_bytes": nbytes, "buckets": len(red.bounds), "allreduce_ms": allreduce_ms, "bus_bandwidth_GBs": bus,
                          "backward_ms": backward_ms, "overlap_ms": overlap_ms, "serial_ms": backward_ms + allreduce_ms,
                          "compute": f"{nblk} CUDA-trained TransformerBlocks, C={C}, batch {B} x {S}x{S}, L1LossSr, fused clip+AdamW in the overlapped step",
                          "loss": loss}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
