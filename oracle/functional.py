"""Functional CPU restatement of the reference forwards (torch CPU ops, any float dtype).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function takes the
reference ``state_dict`` (same keys/shapes as the reference nn.Modules) and
plain tensors; there are no modules here.  Citations are to /root/reference.

The arithmetic follows the reference exactly: biased channel variance with
eps 1e-5 inside the sqrt, BiasFree LN does not centre x, L2 normalisation with
eps 1e-12 over the pixel axis, erf GELU, eval-mode BatchNorm with eps 1e-5,
bilinear align_corners=True up-sampling, zero temporal/spatial padding.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------
# KDLAE-T  (KDLAE/KDLAE_model.py:32-336)
# --------------------------------------------------------------------------

def _channel_layernorm(x: Tensor, sd: SD, prefix: str) -> Tensor:
    """KDLAE_model.py:38-83 — LayerNorm over channels for every pixel.

    BiasFree (:50-52): x / sqrt(var + 1e-5) * w (x is NOT centred).
    WithBias (:67-70): (x - mu) / sqrt(var + 1e-5) * w + b.
    """
    w = sd[prefix + ".body.weight"].view(1, -1, 1, 1)
    var = x.var(dim=1, keepdim=True, unbiased=False)
    rstd = 1.0 / torch.sqrt(var + 1e-5)
    bkey = prefix + ".body.bias"
    if bkey in sd:
        mu = x.mean(dim=1, keepdim=True)
        return (x - mu) * rstd * w + sd[bkey].view(1, -1, 1, 1)
    return x * rstd * w


def _mdta(x: Tensor, sd: SD, p: str, heads: int) -> Tensor:
    """KDLAE_model.py:124-145 — transposed (channel) attention."""
    b, c, h, w = x.shape
    qkv = F.conv2d(x, sd[p + ".qkv.weight"], sd.get(p + ".qkv.bias"))
    qkv = F.conv2d(qkv, sd[p + ".qkv_dwconv.weight"], sd.get(p + ".qkv_dwconv.bias"),
                   padding=1, groups=3 * c)
    q, k, v = qkv.view(b, 3, heads, c // heads, h * w).unbind(dim=1)
    qn = q / q.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    kn = k / k.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    logits = torch.matmul(qn, kn.transpose(-1, -2)) * sd[p + ".temperature"].view(1, heads, 1, 1)
    attn = torch.softmax(logits, dim=-1)
    out = torch.matmul(attn, v).reshape(b, c, h, w)
    return F.conv2d(out, sd[p + ".project_out.weight"], sd.get(p + ".project_out.bias"))


def _gdfn(x: Tensor, sd: SD, p: str) -> Tensor:
    """KDLAE_model.py:101-106 — gated depthwise feed-forward."""
    t = F.conv2d(x, sd[p + ".project_in.weight"], sd.get(p + ".project_in.bias"))
    t = F.conv2d(t, sd[p + ".dwconv.weight"], sd.get(p + ".dwconv.bias"), padding=1, groups=t.shape[1])
    hidden = t.shape[1] // 2
    g = F.gelu(t[:, :hidden]) * t[:, hidden:]
    return F.conv2d(g, sd[p + ".project_out.weight"], sd.get(p + ".project_out.bias"))


def _blocks(x: Tensor, sd: SD, prefix: str, n: int, heads: int) -> Tensor:
    """KDLAE_model.py:159-163 repeated n times."""
    for i in range(n):
        p = f"{prefix}.{i}"
        x = x + _mdta(_channel_layernorm(x, sd, p + ".norm1"), sd, p + ".attn", heads)
        x = x + _gdfn(_channel_layernorm(x, sd, p + ".norm2"), sd, p + ".ffn")
    return x


def _count_blocks(sd: SD, prefix: str) -> int:
    n = 0
    while f"{prefix}.{n}.attn.temperature" in sd:
        n += 1
    return n


def teacher_forward(sd: SD, img: Tensor, denoise_rate: Tensor,
                    heads: Sequence[int] = (1, 2, 4, 8),
                    static: Optional[str] = None, params: str = "cat"):
    """KDLAE_teacher.forward, KDLAE_model.py:270-336 (dual_pixel_task=False).

    Returns (hq, sr); sr is None when the SR head is absent (static != 'train').
    """
    if static is None:
        static = "train" if "cen.weight" in sd else "no"
    conv3 = lambda t, key, **kw: F.conv2d(t, sd[key + ".weight"], sd.get(key + ".bias"), padding=kw.pop("padding", 1), **kw)
    nb = lambda pre: _count_blocks(sd, pre)

    e1 = _blocks(conv3(img, "patch_embed.proj"), sd, "encoder_level1", nb("encoder_level1"), heads[0])
    e2 = _blocks(F.pixel_unshuffle(conv3(e1, "down1_2.body.0"), 2), sd, "encoder_level2", nb("encoder_level2"), heads[1])
    e3 = _blocks(F.pixel_unshuffle(conv3(e2, "down2_3.body.0"), 2), sd, "encoder_level3", nb("encoder_level3"), heads[2])
    lat = _blocks(F.pixel_unshuffle(conv3(e3, "down3_4.body.0"), 2), sd, "latent", nb("latent"), heads[3])

    d3 = torch.cat([F.pixel_shuffle(conv3(lat, "up4_3.body.0"), 2), e3], dim=1)
    d3 = F.conv2d(d3, sd["reduce_chan_level3.weight"], sd.get("reduce_chan_level3.bias"))
    d3 = _blocks(d3, sd, "decoder_level3", nb("decoder_level3"), heads[2])

    d2 = torch.cat([F.pixel_shuffle(conv3(d3, "up3_2.body.0"), 2), e2], dim=1)
    d2 = F.conv2d(d2, sd["reduce_chan_level2.weight"], sd.get("reduce_chan_level2.bias"))
    d2 = _blocks(d2, sd, "decoder_level2", nb("decoder_level2"), heads[1])

    d1 = torch.cat([F.pixel_shuffle(conv3(d2, "up2_1.body.0"), 2), e1], dim=1)
    d1 = _blocks(d1, sd, "decoder_level1", nb("decoder_level1"), heads[0])
    d1 = _blocks(d1, sd, "refinement", nb("refinement"), heads[0])

    out = conv3(d1, "output")
    if params == "cat":  # :315-319
        out = torch.cat([out, denoise_rate], dim=1)
        out = conv3(out, "output_param", padding=2, dilation=2)
        out = _blocks(out, sd, "refinement_out", nb("refinement_out"), heads[0])
        out = conv3(out, "output2")
    hq = out + img

    sr = None
    if static == "train":  # :324-329
        s = F.pixel_shuffle(conv3(conv3(hq, "cen"), "upen.body.0"), 2)
        s = _blocks(s, sd, "enhance", nb("enhance"), heads[0])
        sr = conv3(s, "outputen")
    return hq, sr


# --------------------------------------------------------------------------
# KDLAE-S  (KDLAE/KDLAE_model.py:340-430)
# --------------------------------------------------------------------------

def _conv3d_pair(x: Tensor, sd: SD, p: str) -> Tensor:
    """_create_conv_block, KDLAE_model.py:386-393: (Conv3d 3x3x3 pad 1 + ReLU) x 2."""
    x = F.relu(F.conv3d(x, sd[p + ".0.weight"], sd[p + ".0.bias"], padding=1))
    return F.relu(F.conv3d(x, sd[p + ".2.weight"], sd[p + ".2.bias"], padding=1))


def student_forward(sd: SD, x: Tensor, residual: bool = True) -> Tensor:
    """KDLAE_student.forward, KDLAE_model.py:395-430.  x: [B, F, H, W] -> [B, F, H, W]."""
    levels = 0
    while f"encoders.{levels}.0.weight" in sd:
        levels += 1
    x5 = x.unsqueeze(1)
    cur, skips = x5, []
    for i in range(levels):
        s = _conv3d_pair(cur, sd, f"encoders.{i}")
        skips.append(s)
        cur = F.max_pool3d(s, kernel_size=(1, 2, 2))
    cur = _conv3d_pair(cur, sd, "st_fusion")
    for i in range(levels):
        cur = F.conv_transpose3d(cur, sd[f"upconv_layers.{i}.weight"], sd[f"upconv_layers.{i}.bias"], stride=(1, 2, 2))
        cur = cur + skips[levels - 1 - i]
        cur = _conv3d_pair(cur, sd, f"decoders.{i}")
    out = F.conv3d(cur, sd["out_conv.weight"], sd["out_conv.bias"])
    if residual:
        out = out + x5
    return out.squeeze(1)


# --------------------------------------------------------------------------
# ASDQE  (ASDQE/ASDQE_model.py:20-171)
# --------------------------------------------------------------------------

def _conv_bn_relu(x: Tensor, sd: SD, conv: str, bn: str) -> Tensor:
    """ASDQE_model.py:24-31 one conv+BN(eval)+ReLU stage."""
    y = F.conv2d(x, sd[conv + ".weight"], sd[conv + ".bias"], padding=1)
    y = F.batch_norm(y, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"], sd[bn + ".bias"],
                     training=False, eps=1e-5)
    return F.relu(y)


def _double_conv(x: Tensor, sd: SD, p: str) -> Tensor:
    p = p + ".double_conv"
    return _conv_bn_relu(_conv_bn_relu(x, sd, p + ".0", p + ".1"), sd, p + ".3", p + ".4")


def _up(x1: Tensor, x2: Tensor, sd: SD, p: str) -> Tensor:
    """ASDQE_model.py:60-66: bilinear x2 (align_corners), pad to x2's size, cat([x2, x1])."""
    x1 = F.interpolate(x1, scale_factor=2, mode="bilinear", align_corners=True)
    dy, dx = x2.shape[2] - x1.shape[2], x2.shape[3] - x1.shape[3]
    x1 = F.pad(x1, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
    return _double_conv(torch.cat([x2, x1], dim=1), sd, p + ".conv")


def _pad16(x: Tensor, multiple: int) -> Tensor:
    """pad_to_multiple, ASDQE_model.py:113-121 (zero pad bottom/right)."""
    h, w = x.shape[-2:]
    ph, pw = (-h) % multiple, (-w) % multiple
    return F.pad(x, (0, pw, 0, ph)) if (ph or pw) else x


def asdqe_trunk(sd: SD, lq: Tensor, gt: Tensor) -> Tensor:
    """Stems + U-Net up to ``enhanced_feat`` (ASDQE_model.py:158-167)."""
    multiple = sd["lq_extractor.double_conv.0.weight"].shape[0]  # unet_multiple = dim (:129)
    lq, gt = _pad16(lq, multiple), _pad16(gt, multiple)
    feat = torch.cat([_double_conv(lq, sd, "lq_extractor"),
                      _double_conv(gt, sd, "gt_extractor"),
                      _double_conv(lq - gt, sd, "diff_extractor")], dim=1)
    x1 = _double_conv(feat, sd, "unet.inc")
    x2 = _double_conv(F.max_pool2d(x1, 2), sd, "unet.down1.maxpool_conv.1")
    x3 = _double_conv(F.max_pool2d(x2, 2), sd, "unet.down2.maxpool_conv.1")
    x4 = _double_conv(F.max_pool2d(x3, 2), sd, "unet.down3.maxpool_conv.1")
    u = _up(x4, x3, sd, "unet.up1")
    u = _up(u, x2, sd, "unet.up2")
    u = _up(u, x1, sd, "unet.up3")
    return F.conv2d(u, sd["unet.outc.conv.weight"], sd["unet.outc.conv.bias"])


def asdqe_forward(sd: SD, lq: Tensor, gt: Tensor) -> Tensor:
    """DenoiseRatePredictor.forward in eval mode (ASDQE_model.py:143-171) -> [B, 1]."""
    f = asdqe_trunk(sd, lq, gt).mean(dim=(2, 3))
    f = F.relu(F.linear(f, sd["regressor.2.weight"], sd["regressor.2.bias"]))
    f = F.relu(F.linear(f, sd["regressor.5.weight"], sd["regressor.5.bias"]))
    return torch.tanh(F.linear(f, sd["regressor.8.weight"], sd["regressor.8.bias"]))


def flops_teacher(h: int, w: int, ic: int = 1, static_train: bool = True) -> float:
    """Algorithmic FLOPs (2*MAC) of one KDLAE-T forward, SURVEY.md section 8(d): 1.9119e12 at 512x512."""
    base = 1.9119e12 if static_train else 1.5241e12
    return base * (h * w) / (512.0 * 512.0)


__all__ = ["teacher_forward", "student_forward", "asdqe_forward", "asdqe_trunk", "flops_teacher", "math"]
