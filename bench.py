#!/usr/bin/env python
"""KDLAE-T images/sec @ 1x512x512 on N B200s (BASELINE.json configs[1]) - one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one KDLAE-T bf16 forward (hq 512x512 + sr 1024x1024) over a batch of 64 synthetic single-channel
512x512 images PER GPU (weak scaling: independent images, no data-path collective).
  value    : whole-job images/s with the inputs already resident in HBM (CUDA events, max over ranks)
  e2e      : same metric through the drop-in nn.Module with pinned-HOST inputs/outputs (H2D + forward + D2H timed)
  roofline : dominant kernel class, algorithmic bytes-or-flops / CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline : the CPU oracle (torch CPU restatement of the reference forward) timed on the host cores
--impl reference times that CPU implementation as the reference arm (the reference is pure PyTorch-CPU-capable
code; its own modules cannot travel to the GPU box, the oracle is pinned to them by tests/golden).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "KDLAE-T images/sec @1x512x512"
UNIT = "images/s"
FLOPS_PER_IMG_512 = 1.9119e12   # algorithmic FLOPs of one 1x1x512x512 forward, static='train'
MODEL_KW = dict(inp_channels=1, out_channels=1, LayerNorm_type="BiasFree", static="train", params="cat")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1400.0, source="fallback")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def cpu_oracle_rate(size: int, threads: int, repeats: int = 1):
    """images/s-equivalent of the CPU oracle: one fp32 forward at size x size, scaled by (size/512)^2 (FLOPs ~ H*W)."""
    import oracle
    from oracle import synth
    torch.set_num_threads(threads)
    sd = synth.teacher_state_dict(seed=0, **{k: v for k, v in MODEL_KW.items() if k != "params"})
    img = synth.seeded_tensor("bench.cpu.img", (1, 1, size, size), 0)
    rate = torch.full((1, 1, size, size), 0.6)
    best = float("inf")
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            oracle.teacher_forward(sd, img, rate)
            best = min(best, time.perf_counter() - t0)
    scale = (size * size) / (512.0 * 512.0)
    return scale / best, best


def time_uint8_pipeline(pk, model, dev, B, S):
    """uint8 HWC host images -> CUDA pre-pass -> KDLAE-T -> CUDA post-pass -> uint8 hq / sr on the host,
    in chunks of one micro-batch on a copy-in / compute / copy-out stream triple (1 byte per element over PCIe each way)."""
    chunk = min(B, 16)
    n = (B + chunk - 1) // chunk
    img_h = torch.randint(0, 256, (B, S, S, 1), dtype=torch.uint8).pin_memory()
    hq_h = torch.empty(B, S, S, 1, dtype=torch.uint8).pin_memory()
    sr_h = torch.empty(B, 2 * S, 2 * S, 1, dtype=torch.uint8).pin_memory()
    rates = torch.rand(B, device=dev)
    img_d = [torch.empty(chunk, S, S, 1, dtype=torch.uint8, device=dev) for _ in range(n)]
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def step():
        cur = torch.cuda.current_stream()
        s_in.wait_stream(cur)
        evs = []
        with torch.cuda.stream(s_in):
            for c in range(n):
                img_d[c].copy_(img_h[c * chunk:(c + 1) * chunk], non_blocking=True)
                ev = torch.cuda.Event(); ev.record(s_in); evs.append(ev)
        for c in range(n):
            cur.wait_event(evs[c])
            hq, sr = pk.teacher_infer_uint8(model, img_d[c], rates[c * chunk:(c + 1) * chunk])
            done = torch.cuda.Event(); done.record(cur)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                hq_h[c * chunk:(c + 1) * chunk].copy_(hq, non_blocking=True)
                sr_h[c * chunk:(c + 1) * chunk].copy_(sr, non_blocking=True)
                hq.record_stream(s_out); sr.record_stream(s_out)
        cur.wait_stream(s_out)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        step()
        torch.cuda.current_stream().synchronize()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    return {"value": B / ms * 1e3, "batch": B, "ms_per_step": ms, "h2d_bytes_per_step": int(img_h.numel() + 4 * B),
            "d2h_bytes_per_step": int(hq_h.numel() + sr_h.numel())}


def time_other_models(pk, synth, dev, pk_):
    """Device-resident throughput of the other two forwards of the path (BASELINE configs 3 and 4 shapes, bf16)."""
    def timeit(f, n=3):
        for _ in range(2):
            f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            f()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    out = {}
    with torch.no_grad():
        s = pk.KDLAE_student(residual=True)
        s.load_state_dict(synth.student_state_dict())
        s = s.to(dev).eval().set_precision("bf16")
        x = torch.rand(32, 5, 512, 512, device=dev)
        ms = timeit(lambda: s(x))
        out["KDLAE-S 5x512x512 stacks/s"] = {"value": 32 / ms * 1e3, "batch": 32, "tflops": 32 * 1.4881e11 / ms / 1e9,
                                             "tensor_frac": 32 * 1.4881e11 / ms / 1e9 / pk_["bf16_tflops"]}
        del s, x
        a = pk.DenoiseRatePredictor()
        a.load_state_dict(synth.asdqe_state_dict(), strict=False)
        a = a.to(dev).eval().set_precision("bf16")
        lq, gt = torch.rand(64, 3, 512, 512, device=dev), torch.rand(64, 3, 512, 512, device=dev)
        ms = timeit(lambda: a(lq, gt))
        out["ASDQE 3x512x512 pairs/s"] = {"value": 64 / ms * 1e3, "batch": 64, "tflops": 64 * 2.1368e11 / ms / 1e9,
                                          "tensor_frac": 64 * 2.1368e11 / ms / 1e9 / pk_["bf16_tflops"]}
    return out


def run_reference(args, rank):
    """Reference arm: the reference's CPU implementation of the path (oracle port) on all host cores."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # size the per-step sample so the whole run stays within a few minutes
    _, t64 = cpu_oracle_rate(64, cores)           # also the warm-up of the thread pool
    _, t128 = cpu_oracle_rate(128, cores)
    budget = 150.0 / max(1, args.steps + args.warmup)
    size = 128
    for cand in (512, 256):
        if t128 * (cand / 128.0) ** 2 <= budget:
            size = cand
            break
    for _ in range(args.warmup):
        cpu_oracle_rate(size, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_rate(size, cores)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = ((size * size) / (512.0 * 512.0)) / dt
    sample = f"1 image of 1x{size}x{size} per step (FLOPs scale with H*W; images/s quoted in 512x512 equivalents)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "KDLAE-T forward, 1x512x512 images, CPU oracle (reference algorithm, torch CPU ops)",
                       "per_gpu_batch": args.batch, "size": args.size},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--micro-batch", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-only", action="store_true",
                    help="for ncu launch lists: warm-up + timed device steps only (no e2e leg, no CPU baseline, no event profile); "
                         "the printed line is not a bench value")
    ap.add_argument("--extras", action="store_true", help="also time KDLAE-S (config 3) and ASDQE (config 4) on rank 0")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from oracle import synth
    import rethink_acoustic_image_enhancement_b200 as pk
    from rethink_acoustic_image_enhancement_b200 import _lib

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its banner ("NCCL version ...") straight to the process's stdout when the communicator is created; stdout
        # carries exactly one JSON line (the contract), so fd 1 points at stderr while NCCL initialises
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()                       # creates the communicator (lazy) while stdout is still redirected
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    B, S = args.batch, args.size
    model = pk.KDLAE_teacher(**MODEL_KW)
    model.load_state_dict(synth.teacher_state_dict(seed=0, **{k: v for k, v in MODEL_KW.items() if k != "params"}))
    model = model.to(dev).eval().set_precision(args.precision)
    if args.micro_batch:
        model.micro_batch = args.micro_batch

    g = torch.Generator().manual_seed(1234 + rank)
    img_h = torch.rand(B, 1, S, S, generator=g).pin_memory()
    rate_h = torch.rand(B, 1, 1, 1, generator=g).expand(B, 1, S, S).contiguous().pin_memory()
    img_d, rate_d = img_h.to(dev), rate_h.to(dev)
    hq_h = torch.empty(B, 1, S, S).pin_memory()
    sr_h = torch.empty(B, 1, 2 * S, 2 * S).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from rethink_acoustic_image_enhancement_b200.sharding import max_over_ranks as _mor

    def max_over_ranks(ms: float) -> float:
        return _mor(ms, dev)

    def step_device():
        with torch.no_grad():
            return model({"img": img_d, "denoise_rate": rate_d})

    # e2e: host (pinned) -> device -> module.forward -> host, in chunks of one micro-batch so that the copy engines move chunk k+1
    # in and chunk k-1 out while chunk k computes (the usual way to feed an inference module from host memory)
    chunk = min(B, 16)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    nchunks = (B + chunk - 1) // chunk
    im_d = [torch.empty_like(img_d[c * chunk:(c + 1) * chunk]) for c in range(nchunks)]     # device staging, reused every step
    rt_d = [torch.empty_like(rate_d[c * chunk:(c + 1) * chunk]) for c in range(nchunks)]

    def step_e2e():
        cur = torch.cuda.current_stream()
        with torch.no_grad():
            s_in.wait_stream(cur)                      # last step's kernels have consumed the staging buffers
            evs = []
            with torch.cuda.stream(s_in):
                for c in range(nchunks):
                    im_d[c].copy_(img_h[c * chunk:(c + 1) * chunk], non_blocking=True)
                    rt_d[c].copy_(rate_h[c * chunk:(c + 1) * chunk], non_blocking=True)
                    ev = torch.cuda.Event(); ev.record(s_in); evs.append(ev)
            for c in range(nchunks):
                cur.wait_event(evs[c])
                out = model({"img": im_d[c], "denoise_rate": rt_d[c]})
                done = torch.cuda.Event(); done.record(cur)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(done)
                    hq_h[c * chunk:(c + 1) * chunk].copy_(out["hq"], non_blocking=True)
                    sr_h[c * chunk:(c + 1) * chunk].copy_(out["sr"], non_blocking=True)
                    out["hq"].record_stream(s_out); out["sr"].record_stream(s_out)
            cur.wait_stream(s_out)

    # ---- device-resident throughput ("value") ----
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    if args.profile_only:
        if rank == 0:
            print(json.dumps({"profile_only": True, "ms_per_step": ms_total / args.steps, "gpu_launches": int(launches)}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the public module with host buffers ("e2e") ----
    for _ in range(2):          # un-timed: lets the caching allocator reach its steady state for the per-chunk outputs
        step_e2e()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        step_e2e()
        torch.cuda.current_stream().synchronize()   # the step's result is read on the host every step
    f1.record()
    barrier()
    ms_e2e = max_over_ranks(f0.elapsed_time(f1))

    # ---- per-kernel-class CUDA-event profile of one more identical step (roofline) ----
    prof = None
    if rank == 0:
        _lib.profile_begin()
        step_device()
        prof = _lib.profile_end()
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk_ = peaks()
    imgs = B * world * args.steps
    value = imgs / (ms_total / 1e3)
    e2e_value = imgs / (ms_e2e / 1e3)
    tot_ms = sum(v["ms"] for v in prof.values())
    classes = {}
    for name, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        sec = v["ms"] / 1e3
        classes[name] = {"ms_per_step": round(v["ms"], 3), "share": round(v["ms"] / tot_ms, 4), "launches": v["launches"],
                         "algorithmic_GBs": round(v["bytes"] / sec / 1e9, 1), "algorithmic_TFLOPs": round(v["flops"] / sec / 1e12, 2),
                         "hbm_frac": round(v["bytes"] / sec / 1e9 / pk_["hbm_gbs"], 4),
                         "tensor_frac": round(v["flops"] / sec / 1e12 / pk_["bf16_tflops"], 4) if "tcgen05" in name else None}
    top = next(iter(classes))
    tv = prof[top]
    hbm_bound = classes[top]["hbm_frac"] >= (classes[top]["tensor_frac"] or 0.0)
    per_launch_s = tv["ms"] / 1e3 / tv["launches"]
    if hbm_bound:
        roof = {"kernel": top, "bound": "hbm", "achieved": tv["bytes"] / tv["launches"] / per_launch_s / 1e9, "peak": pk_["hbm_gbs"],
                "unit": "GB/s"}
    else:
        roof = {"kernel": top, "bound": "tensor", "achieved": tv["flops"] / tv["launches"] / per_launch_s / 1e12,
                "peak": pk_["bf16_tflops"], "unit": "TFLOP/s"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["algorithmic_bytes_per_launch"] = tv["bytes"] / tv["launches"]
    roof["traffic"] = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")   # dram__bytes_read+write per launch, from an ncu capture
    if os.path.exists(tpath) and S == 512:
        with open(tpath) as fh:
            tj = json.load(fh)
        same_mb = min(B, 16) == tj.get("_micro_batch") and (args.micro_batch or tj.get("_micro_batch")) == tj.get("_micro_batch")
        if top in tj and same_mb:                                  # captured at the same micro-batch (launch = same tensors)
            roof["traffic"] = tj[top]["dram_bytes_per_launch"]
            roof["traffic_source"] = tj["_provenance"]
    roof["peak_source"] = pk_["source"] + (" sustained (kernel timed inside a long step)" if pk_["source"] == "measured" else "")
    if top.startswith("fused_conv1x1_dwconv3x3"):
        roof["note"] = ("fused tcgen05 1x1 -> FFMA2 depthwise (+GELU gate) kernels (pwdw_t.cu for qkv, pwdw_f2.cu for the GDFN): the 3C / 2h "
                        "wide intermediate never reaches HBM, so the HBM fraction is low by design (ncu DRAM traffic = algorithmic bytes); "
                        "ncu: smsp__issue_active 57-60 %, FMA pipe 28-30 % - bound by CUDA-core issue / dependency latency of the "
                        "depthwise + GELU code, not by memory (profiles/r01_summary.md)")
    roof["avg_launch_ms"] = per_launch_s * 1e3
    roof["share_of_step"] = classes[top]["share"]

    scale = (S * S) / (512.0 * 512.0)
    whole_tflops = value * FLOPS_PER_IMG_512 * scale / 1e12 / world
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": f"KDLAE-T bf16 batched inference (BASELINE configs[1]): batch {B} of 1x{S}x{S} per GPU, "
                               "hq + sr outputs, random-init key-seeded weights",
                   "per_gpu_batch": B, "global_batch": B * world, "size": S, "micro_batch": model.micro_batch or "auto",
                   "parallelism": f"batch-sharded x{world}, no collective",
                   "l2": "working set per step (>8 GB of activations) far exceeds the 126 MB L2; no explicit flush"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(img_h.numel() * 4 + rate_h.numel() * 4),
                "d2h_bytes_per_step": int(hq_h.numel() * 4 + sr_h.numel() * 4), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "kernel_classes": classes,
        "model_tflops_per_gpu": whole_tflops,
        "model_tensor_frac": whole_tflops / pk_["bf16_tflops"],
    }
    if args.extras:
        line["other_models"] = time_other_models(pk, synth, dev, pk_)
        line["other_models"]["KDLAE-T uint8 pipeline e2e images/s"] = time_uint8_pipeline(pk, model, dev, B, S)
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cpu_oracle_rate(64, cores)
        v, t = cpu_oracle_rate(256 if cores >= 8 else 128, cores)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"one fp32 oracle forward of 1x1x{256 if cores >= 8 else 128}^2 ({t:.1f} s), scaled by H*W to 512x512 equivalents"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
