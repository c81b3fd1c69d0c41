"""CPU restatement of the reference's validation metric and training loss.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinned by tests/golden/metrics.npz, which oracle/make_golden_metrics.py
freezes from the unmodified reference functions (calculate_psnr, tensor2img, L1LossSr imported from /root/reference/Train).
"""
from __future__ import annotations

import numpy as np
import torch


def tensor2img_u8(t: torch.Tensor) -> np.ndarray:
    """Train/basicsr/utils/img_util.py:67-94 for a [C,H,W] tensor, min_max=(0,1), out_type=uint8, no colour swap:
    clamp -> *255 -> numpy round (half to even) -> uint8, HWC (HW for one channel)."""
    x = t.detach().float().cpu().clamp(0, 1).numpy().transpose(1, 2, 0)
    if x.shape[2] == 1:
        x = x[:, :, 0]
    return (x * 255.0).round().astype(np.uint8)


def calculate_psnr(img1, img2, crop_border: int) -> float:
    """Train/basicsr/metrics/psnr_ssim.py:9-70 (test_y_channel=False).  Tensors are [C,H,W] / [1,C,H,W] (:37-50);
    ndarrays are HWC or HW."""
    def prep(x):
        if torch.is_tensor(x):
            if x.dim() == 4:
                x = x[0]
            x = x.detach().cpu().numpy().transpose(1, 2, 0)
        if x.ndim == 2:
            x = x[..., None]
        return x.astype(np.float64)
    a, b = prep(img1), prep(img2)
    assert a.shape == b.shape
    if crop_border != 0:
        a = a[crop_border:-crop_border, crop_border:-crop_border, ...]
        b = b[crop_border:-crop_border, crop_border:-crop_border, ...]
    mse = np.mean((a - b) ** 2)
    if mse == 0:
        return float("inf")
    max_value = 1.0 if a.max() <= 1 else 255.0
    return float(20.0 * np.log10(max_value / np.sqrt(mse)))


def l1_loss_sr(pred: dict, target: dict, loss_weight: float = 1.0) -> torch.Tensor:
    """Train/basicsr/models/losses/losses.py:153-194, reduction='mean', weight=None (autograd-capable torch CPU ops)."""
    def shadow(p, t):
        pb = torch.where(p > 0.1, torch.ones_like(p), torch.zeros_like(p))
        tb = torch.where(t > 0.1, torch.ones_like(t), torch.zeros_like(t))
        return loss_weight * (pb - tb).abs().mean()
    hl = loss_weight * (pred["hq"] - target["hq"]).abs().mean()
    hs = shadow(pred["hq"], target["hq"])
    if pred.get("sr") is not None:
        sl = loss_weight * (pred["sr"] - target["sr"]).abs().mean()
        ss = shadow(pred["sr"], target["sr"])
    else:
        sl, ss = 0, 0
    return 0.5 * hl + 0.25 * sl + 0.25 * (hs + ss)
