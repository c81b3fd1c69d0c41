"""uint8 image pipeline around KDLAE_teacher.forward.

Mirrors the inference recipe of the reference (KDLAE/KDLAE_T.ipynb cell 5): uint8 HWC image -> float/255 -> NCHW ->
reflect-pad bottom/right to a multiple of 8 -> constant denoise-rate map -> model -> clamp(0,1) -> crop ->
skimage.img_as_ubyte (rint(x*255)) -> zero the pixels whose source is exactly 0 in every channel (the sr mask is the
2x2-repeated one).  The two passes run as CUDA kernels behind the C ABI (kdlae_preprocess_u8 / kdlae_postprocess_u8), so
one byte per element crosses PCIe in each direction instead of four and no H x W rate map is ever built on the host.
"""
from typing import Optional, Tuple, Union

import torch

from . import _lib


def padded_size(h: int, w: int, multiple: int = 8) -> Tuple[int, int]:
    """Size the notebook pads to (cell 5: `((h+m)//m)*m` when h % m != 0, else h)."""
    H = ((h + multiple) // multiple) * multiple if h % multiple else h
    W = ((w + multiple) // multiple) * multiple if w % multiple else w
    return H, W


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def preprocess_u8(images: torch.Tensor, denoise_rate: Union[float, torch.Tensor], multiple: int = 8, rate_map: bool = False):
    """images [B,h,w,c] uint8 (CUDA, HWC) -> (img [B,c,H,W] fp32, rate): rate is the [B,1,1,1] per-image value the drop-in
    teacher broadcasts inside its kernel, or with rate_map=True the materialised [B,1,H,W] map the reference module needs."""
    if images.dtype != torch.uint8 or images.dim() != 4 or not images.is_cuda:
        raise RuntimeError("preprocess_u8: expected a CUDA uint8 tensor [B,h,w,c] (there is no CPU fallback)")
    images = images.contiguous()
    B, h, w, c = images.shape
    H, W = padded_size(h, w, multiple)
    dev = images.device
    rates = torch.as_tensor(denoise_rate, dtype=torch.float32, device=dev).reshape(-1)
    if rates.numel() == 1:
        rates = rates.expand(B)
    if rates.numel() != B:
        raise RuntimeError(f"preprocess_u8: denoise_rate must be a scalar or have {B} entries, got {rates.numel()}")
    rates = rates.contiguous()
    with torch.cuda.device(dev):
        img = torch.empty((B, c, H, W), dtype=torch.float32, device=dev)
        rmap = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev) if rate_map else None
        _lib.check(_lib.load().kdlae_preprocess_u8(images.data_ptr(), B, h, w, c, rates.data_ptr(), img.data_ptr(),
                                                   None if rmap is None else rmap.data_ptr(), H, W, _stream()), "kdlae_preprocess_u8")
    return img, (rmap if rate_map else rates.view(B, 1, 1, 1))


def postprocess_u8(pred: torch.Tensor, images: torch.Tensor, scale: int = 1) -> torch.Tensor:
    """pred [B,c,Hp,Wp] fp32 (CUDA) + the uint8 source [B,h,w,c] -> uint8 [B,h*scale,w*scale,c]."""
    if pred.dtype != torch.float32 or pred.dim() != 4 or not pred.is_cuda:
        raise RuntimeError("postprocess_u8: expected a CUDA fp32 tensor [B,c,Hp,Wp]")
    pred, images = pred.contiguous(), images.contiguous()
    B, h, w, c = images.shape
    if pred.shape[0] != B or pred.shape[1] != c:
        raise RuntimeError(f"postprocess_u8: prediction {tuple(pred.shape)} does not match source {tuple(images.shape)}")
    with torch.cuda.device(pred.device):
        out = torch.empty((B, h * scale, w * scale, c), dtype=torch.uint8, device=pred.device)
        _lib.check(_lib.load().kdlae_postprocess_u8(pred.data_ptr(), images.data_ptr(), B, h, w, c, pred.shape[2], pred.shape[3], scale,
                                                    out.data_ptr(), _stream()), "kdlae_postprocess_u8")
    return out


def teacher_infer_uint8(model, images: torch.Tensor, denoise_rate: Union[float, torch.Tensor] = 1.0,
                        multiple: int = 8) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """uint8 in, uint8 out: (hq [B,h,w,c], sr [B,2h,2w,c] or None) exactly as KDLAE_T.ipynb cell 5 produces them."""
    img, rate = preprocess_u8(images, denoise_rate, multiple)
    with torch.no_grad():
        pred = model({"img": img, "denoise_rate": rate})
    hq = postprocess_u8(pred["hq"], images, 1)
    sr = postprocess_u8(pred["sr"], images, 2) if pred.get("sr") is not None else None
    return hq, sr
