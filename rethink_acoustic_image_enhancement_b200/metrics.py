"""Validation metric and training loss on the device, behind the C ABI.

Mirrors, at the call boundary, the reference's
  * ``calculate_psnr(img1, img2, crop_border, input_order='HWC', test_y_channel=False)``
    (Train/basicsr/metrics/psnr_ssim.py:9-70) - here for CHW / NCHW CUDA tensors, as one reduction kernel per batch instead of
    ``.cpu().numpy()`` per image;
  * ``L1LossSr`` (Train/basicsr/models/losses/losses.py:135-194) - forward value and d loss / d pred in one pass;
  * the metric part of ``ImageCleanModel.nondist_validation`` (Train/basicsr/models/image_restoration_model.py:264-348) for the
    dict-input teacher.
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional

import torch

from . import _lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _as_nchw(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (there is no CPU path), got {t.device}")
    if t.dim() == 2:
        t = t[None, None]
    elif t.dim() == 3:
        t = t[None]
    if t.dim() != 4:
        raise RuntimeError(f"{name}: expected [H,W], [C,H,W] or [B,C,H,W], got {tuple(t.shape)}")
    return t.detach().to(torch.float32).contiguous()


def psnr_batch(img1: torch.Tensor, img2: torch.Tensor, crop_border: int = 0, as_uint8: bool = False) -> torch.Tensor:
    """Per-image PSNR [B] (float64, on the device) of two [B,C,H,W] CUDA tensors in [0,1].

    ``as_uint8`` first converts both images like ``tensor2img`` (clamp, *255, round) - the ``use_image`` branch of the
    reference's validation loop; the peak is then 255 unless the first image's maximum is <= 1 (psnr_ssim.py:68-69)."""
    a, b = _as_nchw(img1, "psnr"), _as_nchw(img2, "psnr")
    if a.shape != b.shape:
        raise AssertionError(f"Image shapes are differnet: {tuple(img1.shape)}, {tuple(img2.shape)}.")   # psnr_ssim.py:31-32
    B, C, H, W = a.shape
    lib = _lib.load()
    with torch.cuda.device(a.device):
        out = torch.empty((B, 2), dtype=torch.float64, device=a.device)
        scratch = torch.empty(lib.kdlae_psnr_scratch_bytes(B), dtype=torch.uint8, device=a.device)
        _lib.check(lib.kdlae_psnr(a.data_ptr(), b.data_ptr(), B, C, H, W, int(crop_border), int(bool(as_uint8)), out.data_ptr(),
                                  scratch.data_ptr(), _stream()), "kdlae_psnr")
    mse, mx = out[:, 0], out[:, 1]
    peak = torch.where(mx <= 1.0, torch.ones_like(mx), torch.full_like(mx, 255.0))
    return torch.where(mse == 0, torch.full_like(mse, float("inf")), 20.0 * torch.log10(peak / torch.sqrt(mse)))


def calculate_psnr(img1: torch.Tensor, img2: torch.Tensor, crop_border: int, input_order: str = "CHW",
                   test_y_channel: bool = False) -> float:
    """Drop-in for the tensor branch of ``calculate_psnr`` (psnr_ssim.py:37-50: a 4-D tensor is squeezed / its first image is
    used, tensors are CHW)."""
    if input_order not in ("HWC", "CHW"):
        raise ValueError(f'Wrong input_order {input_order}. Supported input_orders are "HWC" and "CHW"')
    if test_y_channel:
        raise NotImplementedError("calculate_psnr(test_y_channel=True) is not built: the acoustic images are single channel")
    a, b = _as_nchw(img1, "calculate_psnr"), _as_nchw(img2, "calculate_psnr")
    return float(psnr_batch(a[:1], b[:1], crop_border)[0].item())


class _L1SrFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hq, hq_gt, sr, sr_gt, loss_weight):
        lib = _lib.load()
        dev = hq.device
        with torch.cuda.device(dev):
            h, hg = hq.detach().float().contiguous(), hq_gt.detach().float().contiguous()
            s = sg = None
            if sr is not None:
                s, sg = sr.detach().float().contiguous(), sr_gt.detach().float().contiguous()
            loss = torch.empty((), dtype=torch.float32, device=dev)
            g_hq = torch.empty_like(h)
            g_sr = torch.empty_like(s) if s is not None else None
            scratch = torch.empty(lib.kdlae_l1_sr_scratch_bytes(), dtype=torch.uint8, device=dev)
            _lib.check(lib.kdlae_l1_sr_loss(h.data_ptr(), hg.data_ptr(), h.numel(), None if s is None else s.data_ptr(),
                                            None if sg is None else sg.data_ptr(), 0 if s is None else s.numel(),
                                            float(loss_weight), loss.data_ptr(), g_hq.data_ptr(),
                                            None if g_sr is None else g_sr.data_ptr(), None, scratch.data_ptr(), _stream()),
                       "kdlae_l1_sr_loss")
        ctx.save_for_backward(g_hq, g_sr if g_sr is not None else torch.empty(0, device=dev))
        ctx.has_sr = g_sr is not None
        ctx.dtypes = (hq.dtype, None if sr is None else sr.dtype)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        g_hq, g_sr = ctx.saved_tensors
        return ((g_hq * grad_out).to(ctx.dtypes[0]), None, (g_sr * grad_out).to(ctx.dtypes[1]) if ctx.has_sr else None, None, None)


class L1LossSr(torch.nn.Module):
    """Drop-in for losses.py:135-194 (``reduction='mean'``, no element weights): the loss value and the gradient with respect
    to ``pred['hq']`` / ``pred['sr']`` come out of one fused pass per output (kdlae_l1_sr_loss)."""

    def __init__(self, loss_weight: float = 1.0, reduction: str = "mean"):
        super().__init__()
        if reduction not in ("none", "mean", "sum"):
            raise ValueError(f"Unsupported reduction mode: {reduction}. Supported ones are: ['none', 'mean', 'sum']")
        if reduction != "mean":
            raise NotImplementedError("L1LossSr: only reduction='mean' (the shipped KDLAET.yml setting) is built")
        self.loss_weight, self.reduction = loss_weight, reduction

    def forward(self, pred: Dict[str, Optional[torch.Tensor]], target: Dict[str, torch.Tensor], weight=None, **kwargs):
        if weight is not None:
            raise NotImplementedError("L1LossSr: element weights are not built (the training loop never passes them)")
        hq, sr = pred["hq"], pred.get("sr")
        if not hq.is_cuda:
            raise RuntimeError("L1LossSr: expected CUDA tensors (there is no CPU path)")
        return _L1SrFn.apply(hq, target["hq"], sr, target["sr"] if sr is not None else None, self.loss_weight)


def pad_test(model, lq: Dict[str, torch.Tensor], window_size: int = 8) -> Dict[str, Optional[torch.Tensor]]:
    """``ImageCleanModel.pad_test`` (image_restoration_model.py:226-237) for the dict-input teacher: reflect-pad ``lq['img']`` (and a
    full-size rate map) at the bottom / right to a multiple of ``window_size``, run the forward, crop ``hq`` back and ``sr`` back
    at twice the size.  (The reference calls ``self.lq.size()`` on the dict there and cannot run this hook for KDLAE-T; the
    padding itself is torch data movement on the device.)"""
    img, rate = lq["img"], lq["denoise_rate"]
    _, _, h, w = img.shape
    ph, pw = (-h) % window_size, (-w) % window_size
    if ph or pw:
        img = torch.nn.functional.pad(img, (0, pw, 0, ph), "reflect")
        if rate.dim() == 4 and rate.shape[-2:] == (h, w):
            rate = torch.nn.functional.pad(rate, (0, pw, 0, ph), "reflect")
    out = model({"img": img, "denoise_rate": rate})
    hq = out["hq"][:, :, :h, :w]
    sr = out["sr"][:, :, :2 * h, :2 * w] if out.get("sr") is not None else None
    return {"hq": hq, "sr": sr}


def validate(model, batches: Iterable[dict], crop_border: int = 0, use_image: bool = True, key: str = "hq",
             window_size: int = 0) -> Dict[str, float]:
    """Metric part of ``nondist_validation`` (image_restoration_model.py:264-348) for the dict-input teacher: for every batch
    ``{'lq': {'img','denoise_rate'}, 'gt': {'hq','sr'}}`` run the forward and accumulate PSNR **on the device** - no per-image
    ``.cpu()``, no ``torch.cuda.empty_cache()`` per iteration (:299); one host read at the end."""
    total: Optional[torch.Tensor] = None
    cnt = 0
    was_training = model.training
    model.eval()
    try:
        with torch.no_grad():
            for data in batches:
                out = pad_test(model, data["lq"], window_size) if window_size else model(data["lq"])   # val.window_size (:283-287)
                p = psnr_batch(out[key], data["gt"][key].to(out[key].device), crop_border, as_uint8=use_image)
                total = p.sum() if total is None else total + p.sum()
                cnt += p.numel()
    finally:
        model.train(was_training)
    return {"psnr": float(total.item()) / max(cnt, 1) if total is not None else math.nan, "count": cnt}


__all__: List[str] = ["psnr_batch", "calculate_psnr", "L1LossSr", "pad_test", "validate"]
