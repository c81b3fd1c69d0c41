"""TEST INFRASTRUCTURE ONLY - CPU restatement of the uint8 pre / post-processing of the reference's inference recipe
(KDLAE/KDLAE_T.ipynb cell 5; the cell is notebook code, not an importable function, so this restatement uses the same
ATen / numpy operations in the same order: astype(float32)/255, permute, F.pad(..., 'reflect'), torch.ones * rate,
torch.clamp, crop, skimage.img_as_ubyte == rint(x*255) for floats in [0,1], zero mask, np.repeat for the sr mask).
Pinned: oracle/make_golden_prepost.py EXECUTES the cell's own source (read from the notebook at run time) on synthetic PNG
files and froze its inputs / outputs in tests/golden/prepost.npz; this restatement reproduces them bit for bit
(tests/test_prepost.py).  One caveat: scikit-image (requirements.txt:9, unpinned) is not installed in the build image, so the
cell ran with a restatement of `img_as_ubyte`'s published float -> uint8 conversion (skimage/util/dtype.py `_convert`)."""
import numpy as np
import torch
import torch.nn.functional as F


def padded_size(h, w, m=8):
    H, W = ((h + m) // m) * m, ((w + m) // m) * m          # cell 5: H,W = ((h+m)//m)*m
    return (H if h % m else h), (W if w % m else w)          # padh = H-h if h % m != 0 else 0


def preprocess_u8(images_u8: np.ndarray, denoise_rate, m: int = 8):
    """images [B,h,w,c] uint8 -> (img [B,c,H,W] fp32, rate map [B,1,H,W] fp32)."""
    B, h, w, c = images_u8.shape
    x = torch.from_numpy(images_u8.astype(np.float32) / 255.0).permute(0, 3, 1, 2)     # load_image_as_tensor
    H, W = padded_size(h, w, m)
    x = F.pad(x, (0, W - w, 0, H - h), "reflect") if (H > h or W > w) else x
    rates = torch.as_tensor(denoise_rate, dtype=torch.float32).reshape(-1)
    rates = rates.expand(B) if rates.numel() == 1 else rates
    alpha = torch.ones((B, 1, H, W)) * rates.view(B, 1, 1, 1)
    return x.contiguous(), alpha


def img_as_ubyte(x: np.ndarray) -> np.ndarray:
    """skimage.util.img_as_ubyte for float input in [0,1]: multiply by 255, rint, clip, cast."""
    y = np.multiply(x, 255.0, dtype=np.float32)
    np.rint(y, out=y)
    np.clip(y, 0, 255, out=y)
    return y.astype(np.uint8)


def postprocess_u8(pred: torch.Tensor, images_u8: np.ndarray, scale: int = 1) -> np.ndarray:
    """pred [B,c,Hp,Wp] fp32 -> uint8 [B,h*scale,w*scale,c] (clamp, crop, img_as_ubyte, blind-zone mask)."""
    B, h, w, c = images_u8.shape
    r = torch.clamp(pred, 0, 1)[:, :, : h * scale, : w * scale]
    out = img_as_ubyte(r.permute(0, 2, 3, 1).cpu().numpy())
    mask = np.all(images_u8 == 0, axis=-1)                      # every channel of the source pixel is 0
    if scale == 2:
        mask = np.repeat(np.repeat(mask, 2, axis=1), 2, axis=2)
    out[mask] = 0
    return out
