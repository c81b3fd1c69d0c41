#!/usr/bin/env python
"""KDLAE-T images/sec @ 1x512x512 on N B200s (BASELINE.json configs[1]) - one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one KDLAE-T bf16 forward (hq 512x512 + sr 1024x1024) over a batch of 64 synthetic single-channel
512x512 images PER GPU (weak scaling: independent images, no data-path collective).
  value    : whole-job images/s with the inputs already resident in HBM (CUDA events, max over ranks)
  e2e      : same metric through the drop-in nn.Module with pinned-HOST inputs/outputs (H2D + forward + D2H timed)
  roofline : dominant kernel class, algorithmic bytes-or-flops / CUDA-event duration vs MEASURED_PEAKS.json
  parity   : PSNR (device reduction, kdlae_psnr) of image 0 of the timed batch against one live CPU-oracle forward of the same
             1x512x512 image - the metric is "images/s; PSNR vs ref"
  strong_scaling : the same global batch of 64 split over the N ranks (BASELINE configs[1] read literally: 64 / N per GPU)
  other_configs  : BASELINE configs[2] (KDLAE-S-FLS, 128 stacks of 5x512x512) and configs[3] (1024 images through KDLAE-S-US
             stacks and ASDQE scoring), sharded over the N ranks, each with its own roofline
  cpu_baseline : the reference algorithm on the host cores: the unmodified reference module when baseline/_ref/ holds it
             (kind "reference"), else the CPU oracle port pinned to it by tests/golden (kind "port")
--impl reference times that CPU implementation as the reference arm.
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


_REF_MODEL = None


def reference_teacher():
    """The UNMODIFIED reference KDLAE_teacher when an install / copy of the reference lives under baseline/_ref/ (git-ignored;
    BASELINE.md section 3), else None.  /root/reference itself is never read: it does not exist on the GPU box."""
    global _REF_MODEL
    if _REF_MODEL is not None:
        return _REF_MODEL or None
    _REF_MODEL = False
    base = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(base):
        for dirpath, _dirs, files in os.walk(base):
            if "KDLAE_model.py" in files:
                try:
                    sys.path.insert(0, dirpath)
                    from KDLAE_model import KDLAE_teacher  # type: ignore
                    from oracle import synth
                    m = KDLAE_teacher(**MODEL_KW)
                    m.load_state_dict(synth.teacher_state_dict(seed=0, **{k: v for k, v in MODEL_KW.items() if k != "params"}),
                                      strict=True)
                    _REF_MODEL = m.eval()
                except Exception as e:      # a broken copy must not take the bench down: fall back to the pinned port
                    print(f"[bench] baseline/_ref present but unusable ({e!r}); using the oracle port", file=sys.stderr)
                break
    return _REF_MODEL or None


def cpu_kind() -> str:
    return "reference" if reference_teacher() is not None else "port"


def cpu_oracle_rate(size: int, threads: int, repeats: int = 1, img=None, rate_value: float = 0.6):
    """images/s-equivalent of the reference algorithm on the CPU: one fp32 forward at size x size, scaled by (size/512)^2
    (FLOPs ~ H*W).  Returns (rate, seconds, (hq, sr) of the last forward)."""
    import oracle
    from oracle import synth
    torch.set_num_threads(threads)
    sd = synth.teacher_state_dict(seed=0, **{k: v for k, v in MODEL_KW.items() if k != "params"})
    if img is None:
        img = synth.seeded_tensor("bench.cpu.img", (1, 1, size, size), 0)
    rate = torch.full((1, 1, size, size), rate_value)
    ref = reference_teacher()
    best, out = float("inf"), None
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            if ref is not None:
                o = ref({"img": img, "denoise_rate": rate})
                out = (o["hq"], o["sr"])
            else:
                out = oracle.teacher_forward(sd, img, rate)
            best = min(best, time.perf_counter() - t0)
    scale = (size * size) / (512.0 * 512.0)
    return scale / best, best, out


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


def roofline_of(prof, pk_):
    """Roofline object of the kernel class with the largest share of a profiled pass (CUDA events per launch, kdlae_profile_*)."""
    tot = sum(v["ms"] for v in prof.values())
    name, v = max(prof.items(), key=lambda kv: kv[1]["ms"])
    per = v["ms"] / 1e3 / v["launches"]
    gbs, tfl = v["bytes"] / v["launches"] / per / 1e9, v["flops"] / v["launches"] / per / 1e12
    hbm_frac, tensor_frac = gbs / pk_["hbm_gbs"], (tfl / pk_["bf16_tflops"] if "tcgen05" in name else 0.0)
    if hbm_frac >= tensor_frac:
        r = {"kernel": name, "bound": "hbm", "achieved": gbs, "peak": pk_["hbm_gbs"], "unit": "GB/s", "frac": hbm_frac}
    else:
        r = {"kernel": name, "bound": "tensor", "achieved": tfl, "peak": pk_["bf16_tflops"], "unit": "TFLOP/s", "frac": tensor_frac}
    r.update(hbm_frac=hbm_frac, tensor_frac=tensor_frac, avg_launch_ms=per * 1e3, share_of_step=v["ms"] / tot, traffic=None,
             algorithmic_bytes_per_launch=v["bytes"] / v["launches"], algorithmic_flops_per_launch=v["flops"] / v["launches"])
    return r


def time_other_configs(pk, synth, dev, pk_, world, rank, barrier, max_over_ranks):
    """BASELINE configs[2] and configs[3] at their own batch sizes, sharded over the ranks (no collective), device-resident,
    bf16: (3) KDLAE-S-FLS, 128 stacks of 5x512x512; (4) 1024 single-channel 512x512 images -> KDLAE-S-US on 5-frame stacks ->
    ASDQE scores of the (origin, denoised) pairs with the channel replicated to 3 (ASDQE_test.py:60-61 `.convert('RGB')`)."""
    from rethink_acoustic_image_enhancement_b200 import _lib
    from rethink_acoustic_image_enhancement_b200.sharding import shard_slice

    def timed(f, n=3):
        for _ in range(2):
            f()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            f()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b) / n)

    out = {}
    with torch.no_grad():
        stu = pk.KDLAE_student(residual=True)
        stu.load_state_dict(synth.student_state_dict())
        stu = stu.to(dev).eval().set_precision("bf16")
        # ---- configs[2]
        sl = shard_slice(128, rank, world)
        n3 = sl.stop - sl.start
        x = torch.rand(n3, 5, 512, 512, device=dev)
        ms = timed(lambda: stu(x))
        entry = {"metric": "KDLAE-S-FLS 5x512x512 stacks/sec", "value": 128 / ms * 1e3, "unit": "stacks/s", "global_batch": 128,
                 "per_gpu_batch": n3, "ms_per_step": ms, "dtype": "bf16",
                 "model_tflops_per_gpu": n3 * 1.4881e11 / ms / 1e9, "model_tensor_frac": n3 * 1.4881e11 / ms / 1e9 / pk_["bf16_tflops"]}
        if rank == 0:
            _lib.profile_begin(); stu(x); entry["roofline"] = roofline_of(_lib.profile_end(), pk_)
        barrier()
        out["configs[2]"] = entry
        del x
        # ---- configs[3]
        asd = pk.DenoiseRatePredictor()
        asd.load_state_dict(synth.asdqe_state_dict(), strict=False)
        asd = asd.to(dev).eval().set_precision("bf16")
        sl = shard_slice(1024, rank, world)
        n4 = sl.stop - sl.start
        nst = (n4 + 4) // 5
        frames = torch.rand(nst, 5, 512, 512, device=dev)

        def pipeline():
            den = stu(frames)
            lq = frames.view(nst * 5, 1, 512, 512)[:n4].expand(n4, 3, 512, 512)
            gt = den.view(nst * 5, 1, 512, 512)[:n4].expand(n4, 3, 512, 512)
            return asd(lq, gt)

        ms = timed(pipeline, n=2)
        flops = nst * 1.4881e11 + n4 * 2.1368e11
        entry = {"metric": "ASDQE-scored images/sec (KDLAE-S-US denoise + ASDQE score)", "value": 1024 / ms * 1e3, "unit": "images/s",
                 "global_batch": 1024, "per_gpu_batch": n4, "stacks_per_gpu": nst, "ms_per_step": ms, "dtype": "bf16",
                 "model_tflops_per_gpu": flops / ms / 1e9, "model_tensor_frac": flops / ms / 1e9 / pk_["bf16_tflops"]}
        asd_ms = timed(lambda: asd(frames.view(nst * 5, 1, 512, 512)[:n4].expand(n4, 3, 512, 512),
                                   frames.view(nst * 5, 1, 512, 512)[:n4].expand(n4, 3, 512, 512)), n=2)
        entry["asdqe_only_pairs_per_s"] = 1024 / asd_ms * 1e3
        entry["asdqe_only_tensor_frac"] = n4 * 2.1368e11 / asd_ms / 1e9 / pk_["bf16_tflops"]
        if rank == 0:
            _lib.profile_begin(); pipeline(); entry["roofline"] = roofline_of(_lib.profile_end(), pk_)
        barrier()
        out["configs[3]"] = entry
        del frames, stu, asd
    torch.cuda.empty_cache()
    try:
        out["configs[4]"] = time_training_step(pk, synth, dev, world, rank, timed)
    except Exception as e:                      # never lose the headline line to the secondary workload
        out["configs[4]"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    return out


def time_training_step(pk, synth, dev, world, rank, timed, batch=2, crop=256):
    """BASELINE configs[4]: one KDLAE-T basicsr training step (image_restoration_model.py:198-224) on 256x256 crops through the
    module in train() mode - CUDA forward with saves, L1LossSr, CUDA backward, clip_grad_norm_(0.01) + AdamW - with the DDP
    gradient all-reduce (25 MB buckets launched from gradient hooks on a communication stream: NCCL over NVLink) when N > 1.
    fp32 CUDA-core kernels: a correctness-first slice, every gradient checked against the oracle."""
    from rethink_acoustic_image_enhancement_b200.metrics import L1LossSr
    from rethink_acoustic_image_enhancement_b200.training import BucketedAllReducer, FlatAdamW
    kw = dict(inp_channels=1, out_channels=1, LayerNorm_type="BiasFree", static="train")
    m = pk.KDLAE_teacher(**kw)
    m.load_state_dict(synth.teacher_state_dict(seed=0, temp_scale=1.0, **kw), strict=True)
    m = m.to(dev).train()
    opt = FlatAdamW(list(m.parameters()), lr=1e-5, weight_decay=0.5e-4)           # KDLAET.yml optim_g
    red = BucketedAllReducer(opt.grad, bucket_bytes=25 << 20)
    red.attach(opt.params, opt.offsets)
    crit = L1LossSr(loss_weight=1.0)
    img = synth.seeded_tensor(f"bench.train.img.{rank}", (batch, 1, crop, crop), 1, "sonar").to(dev)
    rate = torch.full((batch, 1, crop, crop), 0.6, device=dev)
    gt = {"hq": synth.seeded_tensor(f"bench.train.hq.{rank}", (batch, 1, crop, crop), 2, "sonar").to(dev),
          "sr": synth.seeded_tensor(f"bench.train.sr.{rank}", (batch, 1, 2 * crop, 2 * crop), 3, "sonar").to(dev)}
    losses = []

    def step():
        opt.zero_grad()
        loss = crit(m({"img": img, "denoise_rate": rate}), gt)
        loss.backward()
        red.wait()
        opt.step()
        losses.append(loss.detach())

    from rethink_acoustic_image_enhancement_b200 import training
    ms = timed(step, n=3)
    ls = [float(l) for l in losses]
    training.set_matmul_precision("tf32")        # 1x1-conv GEMMs (forward, dgrad, wgrad) on tcgen05 in TF32: torch's default conv numerics
    try:
        ms_tf32 = timed(step, n=3)
    finally:
        training.set_matmul_precision("fp32")
    del m, opt, red
    torch.cuda.empty_cache()
    return {"metric": "KDLAE-T training step images/sec (forward + L1-Shadow loss + backward + clip + AdamW" +
                      (" + bucketed NCCL all-reduce)" if world > 1 else ")"),
            "value": batch * world / ms * 1e3, "unit": "images/s", "global_batch": batch * world, "per_gpu_batch": batch, "crop": crop,
            "ms_per_step": ms, "dtype": "f32", "grad_bytes": 26874300 * 4, "allreduce": "25 MB buckets, overlapped with backward" if world > 1 else None,
            "loss_first": ls[0], "loss_last": ls[-1],
            "tf32_matmul": {"value": batch * world / ms_tf32 * 1e3, "unit": "images/s", "ms_per_step": ms_tf32,
                            "what": "same step with the dense-conv GEMMs (1x1 forward, dgrad, wgrad; 3x3 forward, dgrad) and the Gram reductions on tcgen05 kind::tf32 (gemm_tf32.cu)"}}


def run_reference(args, rank):
    """Reference arm: the reference's CPU implementation of the path (oracle port) on all host cores."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # size the per-step sample so the whole run stays within a few minutes
    cpu_oracle_rate(64, cores)                    # also the warm-up of the thread pool
    _, t128, _ = cpu_oracle_rate(128, cores)
    budget = 240.0 / max(1, args.steps + args.warmup)
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
            "config": {"workload": "KDLAE-T forward, 1x512x512 images, reference algorithm on the host CPU (torch CPU ops): "
                                   + ("unmodified reference module from baseline/_ref" if cpu_kind() == "reference" else
                                      "CPU oracle port pinned to the reference by tests/golden"),
                       "per_gpu_batch": args.batch, "size": args.size, "sample_size": size, "same_size_as_metric": size == 512},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": cpu_kind(), "sample": sample},
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
    ap.add_argument("--extras", action="store_true", help="also time the uint8 pre/post pipeline (SURVEY 8f N2) on rank 0")
    ap.add_argument("--no-other-configs", action="store_true", help="skip BASELINE configs[2] / configs[3] (KDLAE-S, ASDQE)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (global batch 64 split over the ranks)")
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
    # denoise_rate: one value per image (KDLAE_T.ipynb cell 5, paired_image_dataset.py:961); the drop-in module takes it as
    # [B,1,1,1] and broadcasts inside its kernel, so the H x W map is never built or copied
    rate_h = torch.rand(B, 1, 1, 1, generator=g).pin_memory()
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

    # ---- strong scaling: BASELINE configs[1] read literally - ONE batch of 64 split over the ranks (64 / N images per GPU) ----
    strong = None
    if not args.no_strong and not args.profile_only:
        from rethink_acoustic_image_enhancement_b200.sharding import shard_slice
        sl = shard_slice(B, rank, world)
        si, sr_ = img_d[sl.start:sl.stop], rate_d[sl.start:sl.stop]

        def step_strong():
            with torch.no_grad():
                return model({"img": si, "denoise_rate": sr_})
        for _ in range(3):
            step_strong()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            step_strong()
        g1.record()
        barrier()
        ms_strong = max_over_ranks(g0.elapsed_time(g1))
        strong = {"value": B * args.steps / (ms_strong / 1e3), "unit": UNIT, "scaling": "strong", "global_batch": B,
                  "per_gpu_batch": sl.stop - sl.start, "ms_per_step": ms_strong / args.steps}

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

    pk_ = peaks()
    other = None
    if not args.no_other_configs and S == 512:
        im_d.clear(); rt_d.clear()
        model._engine._ws.clear()           # hand the teacher's workspace back before the student / ASDQE batches
        torch.cuda.empty_cache()
        other = time_other_configs(pk, synth, dev, pk_, world, rank, barrier, max_over_ranks)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

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
    roof["hbm_frac"], roof["tensor_frac"] = classes[top]["hbm_frac"], classes[top]["tensor_frac"]
    roof["algorithmic_bytes_per_launch"] = tv["bytes"] / tv["launches"]
    roof["algorithmic_flops_per_launch"] = tv["flops"] / tv["launches"]
    roof["traffic"] = None
    # what ncu says limits the class (issue slots / FMA pipe of the CUDA-core depthwise + GELU code), from the committed
    # `--set full` capture of the same command (profiles/r02_issue.json, written by scripts/ncu_issue.py)
    ipath = os.path.join(ROOT, "profiles", "r02_issue.json")
    if os.path.exists(ipath):
        with open(ipath) as fh:
            ij = json.load(fh)
        if top in ij:
            roof["issue"] = ij[top]
            if max(roof["hbm_frac"], roof["tensor_frac"] or 0.0) < 0.5 and ij[top].get("issue_frac", 0.0) >= 0.5:
                roof["bound"] = "issue"       # neither roofline binds: instruction issue does (see roof["issue"])
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")   # dram__bytes_read+write per launch, from an ncu capture
    if not os.path.exists(tpath):
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath) and S == 512:
        with open(tpath) as fh:
            tj = json.load(fh)
        same_mb = min(B, 16) == tj.get("_micro_batch") and (args.micro_batch or tj.get("_micro_batch")) == tj.get("_micro_batch")
        if top in tj and same_mb:                                  # captured at the same micro-batch (launch = same tensors)
            roof["traffic"] = tj[top]["dram_bytes_per_launch"]
            roof["traffic_source"] = tj["_provenance"]
    roof["peak_source"] = pk_["source"] + (" sustained (kernel timed inside a long step)" if pk_["source"] == "measured" else "")
    if top.startswith("fused_conv1x1_dwconv3x3"):
        roof["note"] = ("fused tcgen05 1x1 -> FFMA2 depthwise (+GELU gate) kernels (pwdw_t.cu, both the qkv and the GDFN pair of every "
                        "block with C <= 128): the 3C / 2h wide intermediate never reaches HBM, so the HBM fraction is low by design (ncu DRAM traffic = "
                        "algorithmic bytes) and the tensor pipe only carries the small 1x1 (K = C); the class is bound by CUDA-core "
                        "instruction issue / FMA-pipe occupancy of the depthwise + GELU code - see roofline.issue (ncu --set full, "
                        "profiles/r02_issue.json, profiles/r02_summary.md)")
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
        "strong_scaling": strong,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(img_h.numel() * 4 + rate_h.numel() * 4),
                "d2h_bytes_per_step": int(hq_h.numel() * 4 + sr_h.numel() * 4), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "kernel_classes": classes,
        "model_tflops_per_gpu": whole_tflops,
        "model_tensor_frac": whole_tflops / pk_["bf16_tflops"],
    }
    if other is not None:
        line["other_configs"] = other
    if args.extras:
        line["uint8_pipeline_e2e"] = time_uint8_pipeline(pk, model, dev, B, S)
    if not args.no_cpu_baseline:
        # one live forward of the reference algorithm on image 0 of the timed batch: the CPU baseline AND the parity reference
        from rethink_acoustic_image_enhancement_b200 import metrics as pm
        cores = os.cpu_count() or 1
        cpu_oracle_rate(64, cores)
        csize = S if cores >= 8 else min(S, 128)
        img0 = img_h[:1, :, :csize, :csize].contiguous()
        v, t, (hq_ref, sr_ref) = cpu_oracle_rate(csize, cores, img=img0, rate_value=float(rate_h[0]))
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": cpu_kind(),
                                "sample": f"one fp32 forward of the reference algorithm on image 0 of the batch at 1x{csize}x{csize} "
                                          f"({t:.1f} s)" + ("" if csize == 512 else ", scaled by H*W to 512x512 equivalents")}
        with torch.no_grad():
            got = model({"img": img0.to(dev), "denoise_rate": rate_d[:1]})
            p_hq = float(pm.psnr_batch(got["hq"], hq_ref.to(dev))[0])
            p_sr = float(pm.psnr_batch(got["sr"], sr_ref.to(dev))[0])
            finite = bool(torch.isfinite(got["hq"]).all() and torch.isfinite(got["sr"]).all())
        line["parity"] = {"psnr_hq": p_hq, "psnr_sr": p_sr, "shape": f"1x{csize}x{csize}", "gate_db": 50.0, "finite": finite,
                          "reference": cpu_kind(), "how": "device PSNR (kdlae_psnr) of the bf16 CUDA forward of image 0 of the "
                                                          "timed batch vs one live CPU forward of the reference algorithm"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
