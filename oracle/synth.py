"""Key-seeded synthetic weights and inputs for the oracle and the parity tests.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Every tensor is generated from ``crc32(key) ^ seed`` with its own CPU
``torch.Generator`` so the values depend only on (key, shape, seed) - not on
module construction order, device or torch's global RNG.  Distributions mimic
the reference's default initialisers (Conv/Linear: U(+-1/sqrt(fan_in))) but
make the otherwise-degenerate parts non-trivial: LN
weights != 1, attention temperatures scaled, BatchNorm running statistics and
affine parameters randomised.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import Dict, Sequence

import torch

Tensor = torch.Tensor


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def seeded_tensor(key: str, shape: Sequence[int], seed: int = 0, kind: str = "uniform01") -> Tensor:
    """Deterministic fp32 tensor: 'uniform01' in [0,1), 'normal', or 'sonar' (45% exact zeros)."""
    g = _gen(key, seed)
    if kind == "uniform01":
        return torch.rand(tuple(shape), generator=g)
    if kind == "normal":
        return torch.randn(tuple(shape), generator=g)
    if kind == "sonar":
        x = torch.rand(tuple(shape), generator=g)
        m = torch.rand(tuple(shape), generator=g) < 0.45
        return x.masked_fill(m, 0.0)
    raise ValueError(kind)


def _fill(sd: "OrderedDict[str, Tensor]", key: str, shape, seed: int, mode: str, **kw) -> None:
    g = _gen(key, seed)
    shape = tuple(shape)
    if mode == "fan_in":  # conv / linear weight or bias, bound 1/sqrt(fan_in)
        bound = kw.get("gain", 1.0) / math.sqrt(kw["fan_in"])
        t = (torch.rand(shape, generator=g) * 2 - 1) * bound
    elif mode == "around1":
        t = 1.0 + kw.get("spread", 0.1) * torch.randn(shape, generator=g)
    elif mode == "small":
        t = kw.get("spread", 0.1) * torch.randn(shape, generator=g)
    elif mode == "pos":
        t = 0.5 + torch.rand(shape, generator=g)
    elif mode == "temp":
        t = kw.get("scale", 1.0) * (0.5 + torch.rand(shape, generator=g))
    else:
        raise ValueError(mode)
    sd[key] = t.float()


def _conv(sd, key, cout, cin_per_group, k, seed, bias=False, nd=2, gain=1.0):
    ks = (k,) * nd if isinstance(k, int) else tuple(k)
    fan_in = cin_per_group * math.prod(ks)
    _fill(sd, key + ".weight", (cout, cin_per_group) + ks, seed, "fan_in", fan_in=fan_in, gain=gain)
    if bias:
        _fill(sd, key + ".bias", (cout,), seed, "fan_in", fan_in=fan_in)


def _block(sd, p, dim, ffn_factor, bias, ln_bias, seed, temp_scale, heads):
    hidden = int(dim * ffn_factor)
    def norm(n):
        _fill(sd, f"{p}.{n}.body.weight", (dim,), seed, "around1")
        if ln_bias:
            _fill(sd, f"{p}.{n}.body.bias", (dim,), seed, "small")

    norm("norm1")
    _fill(sd, f"{p}.attn.temperature", (heads, 1, 1), seed, "temp", scale=temp_scale)
    _conv(sd, f"{p}.attn.qkv", dim * 3, dim, 1, seed, bias)
    _conv(sd, f"{p}.attn.qkv_dwconv", dim * 3, 1, 3, seed, bias)
    _conv(sd, f"{p}.attn.project_out", dim, dim, 1, seed, bias)
    norm("norm2")
    _conv(sd, f"{p}.ffn.project_in", hidden * 2, dim, 1, seed, bias)
    _conv(sd, f"{p}.ffn.dwconv", hidden * 2, 1, 3, seed, bias)
    _conv(sd, f"{p}.ffn.project_out", dim, hidden, 1, seed, bias)


def teacher_state_dict(inp_channels=1, out_channels=1, dim=48, num_blocks=(4, 6, 6, 8),
                       num_refinement_blocks=4, heads=(1, 2, 4, 8), ffn_expansion_factor=2.66,
                       bias=False, LayerNorm_type="BiasFree", static="train", seed=0,
                       temp_scale=1.0) -> "OrderedDict[str, Tensor]":
    """All tensors of KDLAE_teacher.state_dict() (KDLAE_model.py:205-268), key-seeded."""
    sd: "OrderedDict[str, Tensor]" = OrderedDict()
    lnb = LayerNorm_type != "BiasFree"
    blk = lambda pre, n, d, hd: [_block(sd, f"{pre}.{i}", d, ffn_expansion_factor, bias, lnb, seed, temp_scale, hd) for i in range(n)]
    _conv(sd, "patch_embed.proj", dim, inp_channels, 3, seed, bias)
    blk("encoder_level1", num_blocks[0], dim, heads[0])
    _conv(sd, "down1_2.body.0", dim // 2, dim, 3, seed)
    blk("encoder_level2", num_blocks[1], dim * 2, heads[1])
    _conv(sd, "down2_3.body.0", dim, dim * 2, 3, seed)
    blk("encoder_level3", num_blocks[2], dim * 4, heads[2])
    _conv(sd, "down3_4.body.0", dim * 2, dim * 4, 3, seed)
    blk("latent", num_blocks[3], dim * 8, heads[3])
    _conv(sd, "up4_3.body.0", dim * 16, dim * 8, 3, seed)
    _conv(sd, "reduce_chan_level3", dim * 4, dim * 8, 1, seed, bias)
    blk("decoder_level3", num_blocks[2], dim * 4, heads[2])
    _conv(sd, "up3_2.body.0", dim * 8, dim * 4, 3, seed)
    _conv(sd, "reduce_chan_level2", dim * 2, dim * 4, 1, seed, bias)
    blk("decoder_level2", num_blocks[1], dim * 2, heads[1])
    _conv(sd, "up2_1.body.0", dim * 4, dim * 2, 3, seed)
    blk("decoder_level1", num_blocks[0], dim * 2, heads[0])
    blk("refinement", num_refinement_blocks, dim * 2, heads[0])
    _conv(sd, "output", out_channels, dim * 2, 3, seed, bias)
    _conv(sd, "output_param", dim * 2, out_channels + 1, 3, seed, bias)
    blk("refinement_out", num_refinement_blocks, dim * 2, heads[0])
    _conv(sd, "output2", out_channels, dim * 2, 3, seed, bias)
    if static == "train":
        _conv(sd, "cen", dim * 2, out_channels, 3, seed, bias)
        _conv(sd, "upen.body.0", dim * 4, dim * 2, 3, seed)
        blk("enhance", num_refinement_blocks, dim, heads[0])
        _conv(sd, "outputen", out_channels, dim, 3, seed, bias)
    return sd


def student_state_dict(inp_channels=1, out_channels=1, hidden_channels=(16, 32, 64), seed=0) -> "OrderedDict[str, Tensor]":
    """All tensors of KDLAE_student.state_dict() (KDLAE_model.py:341-393), key-seeded."""
    sd: "OrderedDict[str, Tensor]" = OrderedDict()
    levels = len(hidden_channels) - 1

    def pair(p, cin, cout):
        _conv(sd, p + ".0", cout, cin, 3, seed, True, nd=3)
        _conv(sd, p + ".2", cout, cout, 3, seed, True, nd=3)

    cin = inp_channels
    for i in range(levels):
        pair(f"encoders.{i}", cin, hidden_channels[i])
        cin = hidden_channels[i]
    pair("st_fusion", cin, hidden_channels[-1])
    order = list(enumerate(range(levels - 1, -1, -1)))
    for j, i in order:  # all ConvTranspose3d first, then all decoder blocks (ModuleList registration order)
        cup = hidden_channels[-1] if i == levels - 1 else hidden_channels[i + 1]
        key = f"upconv_layers.{j}"
        _fill(sd, key + ".weight", (cup, hidden_channels[i], 1, 2, 2), seed, "fan_in", fan_in=hidden_channels[i] * 4)
        _fill(sd, key + ".bias", (hidden_channels[i],), seed, "fan_in", fan_in=hidden_channels[i] * 4)
    for j, i in order:
        pair(f"decoders.{j}", hidden_channels[i], hidden_channels[i])
    _conv(sd, "out_conv", out_channels, hidden_channels[0], 1, seed, True, nd=3)
    return sd


def asdqe_state_dict(in_channels=3, dim=16, seed=0, mlp_gain=4.0) -> "OrderedDict[str, Tensor]":
    """All tensors of DenoiseRatePredictor.state_dict() (ASDQE_model.py:127-156), key-seeded.

    BatchNorm statistics/affine are randomised and the regressor weights are scaled by
    ``mlp_gain`` so that scores differ visibly between inputs.
    """
    sd: "OrderedDict[str, Tensor]" = OrderedDict()

    def dconv(p, cin, cout):
        p = p + ".double_conv"
        for ci, (ck, bk) in zip((cin, cout), (("0", "1"), ("3", "4"))):
            _conv(sd, f"{p}.{ck}", cout, ci, 3, seed, True, gain=2.0)
            _fill(sd, f"{p}.{bk}.weight", (cout,), seed, "pos")
            _fill(sd, f"{p}.{bk}.bias", (cout,), seed, "small")
            _fill(sd, f"{p}.{bk}.running_mean", (cout,), seed, "small")
            _fill(sd, f"{p}.{bk}.running_var", (cout,), seed, "pos")
            sd[f"{p}.{bk}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    for stem in ("lq_extractor", "gt_extractor", "diff_extractor"):
        dconv(stem, in_channels, dim)
    c = dim * 3
    dconv("unet.inc", c, 64)
    dconv("unet.down1.maxpool_conv.1", 64, 128)
    dconv("unet.down2.maxpool_conv.1", 128, 256)
    dconv("unet.down3.maxpool_conv.1", 256, 256)
    dconv("unet.up1.conv", 512, 128)
    dconv("unet.up2.conv", 256, 64)
    dconv("unet.up3.conv", 128, 64)
    _conv(sd, "unet.outc.conv", c, 64, 1, seed, True)
    for name, (fo, fi) in (("regressor.2", (256, c)), ("regressor.5", (64, 256)), ("regressor.8", (1, 64))):
        _fill(sd, name + ".weight", (fo, fi), seed, "fan_in", fan_in=fi, gain=mlp_gain)
        _fill(sd, name + ".bias", (fo,), seed, "fan_in", fan_in=fi)
    return sd


def psnr(a: Tensor, b: Tensor) -> float:
    """20*log10(1/sqrt(mse)) on [0,1] data - Train/basicsr/metrics/psnr_ssim.py:66-70 with max=1."""
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return float("inf") if mse == 0 else 20.0 * math.log10(1.0 / math.sqrt(mse))


def to_dtype(sd: Dict[str, Tensor], dtype) -> Dict[str, Tensor]:
    return OrderedDict((k, v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items())
