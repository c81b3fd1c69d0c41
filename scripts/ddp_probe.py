"""The drop-in teacher under torch's own DistributedDataParallel (how the reference wraps it: basicsr/models/base_model.py:76-82).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P scripts/ddp_probe.py

Each rank runs a training step (module in train() mode -> CUDA forward with saves, L1LossSr, CUDA backward) on its own batch, once
bare and once wrapped in DDP.  Checked: the DDP gradients are identical on every rank and equal the average of the ranks' bare
gradients (DDP's NCCL all-reduce runs from the autograd hooks of the parameters that the CUDA backward fills).
Rank 0 prints one JSON line.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.nn.parallel import DistributedDataParallel as DDP  # noqa: E402

import rethink_acoustic_image_enhancement_b200 as pk  # noqa: E402
from oracle import synth  # noqa: E402
from rethink_acoustic_image_enhancement_b200.metrics import L1LossSr  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    saved = os.dup(1); os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier(); torch.cuda.synchronize()
    sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    kw = dict(inp_channels=1, out_channels=1, dim=16, num_blocks=[1, 1, 1, 2], num_refinement_blocks=1, heads=[1, 2, 4, 8],
              LayerNorm_type="BiasFree", static="train")
    m = pk.KDLAE_teacher(**kw)
    m.load_state_dict(synth.teacher_state_dict(seed=3, temp_scale=2.0, **kw), strict=True)
    m = m.to(dev).train()
    crit = L1LossSr()
    B, S = 2, 32
    x = {"img": synth.seeded_tensor(f"ddp.img.{rank}", (B, 1, S, S), 1, "sonar").to(dev), "denoise_rate": torch.full((B, 1, S, S), 0.5, device=dev)}
    gt = {"hq": synth.seeded_tensor(f"ddp.hq.{rank}", (B, 1, S, S), 2, "sonar").to(dev),
          "sr": synth.seeded_tensor(f"ddp.sr.{rank}", (B, 1, 2 * S, 2 * S), 3, "sonar").to(dev)}
    # bare step: local gradients, averaged by hand
    crit(m(x), gt).backward()
    bare = torch.cat([p.grad.flatten() for p in m.parameters()]).clone()
    dist.all_reduce(bare, op=dist.ReduceOp.AVG)
    for p in m.parameters():
        p.grad = None
    # the same step under DDP
    ddp = DDP(m, device_ids=[local])
    crit(ddp(x), gt).backward()
    torch.cuda.synchronize()
    g = torch.cat([p.grad.flatten() for p in m.parameters()])
    err = float((g - bare).abs().max() / bare.abs().max())
    chk = torch.tensor([float(g.double().sum()), float(g.double().abs().sum())], device=dev, dtype=torch.float64)
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    same = all(torch.equal(allc[0], c) for c in allc)
    if rank == 0:
        print(json.dumps({"ddp_world": world, "params": int(g.numel()), "ddp_vs_hand_averaged_rel_err": err, "identical_on_all_ranks": bool(same),
                          "ok": bool(same and err < 1e-6)}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
