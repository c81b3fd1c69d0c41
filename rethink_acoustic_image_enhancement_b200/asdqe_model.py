"""Drop-in DenoiseRatePredictor (ASDQE/ASDQE_model.py:123-171) running on libkdlae_b200.so.

Same constructor, state_dict (148 entries incl. BatchNorm running stats) and forward(lq, gt) -> [B,1].
The forward implements eval-mode semantics (BatchNorm running statistics, Dropout = identity), which is
how the reference scores images (ASDQE_test.py:83).  Sub-modules only hold parameters.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from ._engine import _ptr
from .kdlae_model import _ConvParams, _FusedModule


class _BNParams(nn.Module):
    def __init__(self, n: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n))
        self.bias = nn.Parameter(torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class _LinearParams(nn.Module):
    def __init__(self, fin: int, fout: int):
        super().__init__()
        lin = nn.Linear(fin, fout)  # initialiser only; never called
        self.weight = nn.Parameter(lin.weight.detach().clone())
        self.bias = nn.Parameter(lin.bias.detach().clone())


class DoubleConv(nn.Module):  # ASDQE_model.py:20-31: indices 0,1 / 3,4 carry parameters (2, 5 are ReLU)
    def __init__(self, in_channels: int, out_channels: int) -> None:
        super().__init__()
        seq = nn.Sequential()
        seq.add_module("0", _ConvParams(in_channels, out_channels, (3, 3), bias=True))
        seq.add_module("1", _BNParams(out_channels))
        seq.add_module("3", _ConvParams(out_channels, out_channels, (3, 3), bias=True))
        seq.add_module("4", _BNParams(out_channels))
        self.double_conv = seq


class Down(nn.Module):  # :36-43  maxpool_conv = Sequential(MaxPool2d, DoubleConv) -> key "maxpool_conv.1"
    def __init__(self, cin: int, cout: int) -> None:
        super().__init__()
        seq = nn.Sequential()
        seq.add_module("1", DoubleConv(cin, cout))
        self.maxpool_conv = seq


class Up(nn.Module):  # :48-58 (bilinear=True)
    def __init__(self, cin: int, cout: int) -> None:
        super().__init__()
        self.conv = DoubleConv(cin, cout)


class OutConv(nn.Module):  # :68-72
    def __init__(self, cin: int, cout: int) -> None:
        super().__init__()
        self.conv = _ConvParams(cin, cout, (1, 1), bias=True)


class UNet(nn.Module):  # :77-95 (bilinear=True -> factor 2)
    def __init__(self, inp_channels: int, out_channels: int) -> None:
        super().__init__()
        self.n_channels, self.out_channels, self.bilinear = inp_channels, out_channels, True
        self.inc = DoubleConv(inp_channels, 64)
        self.down1 = Down(64, 128)
        self.down2 = Down(128, 256)
        self.down3 = Down(256, 256)
        self.up1 = Up(512, 128)
        self.up2 = Up(256, 64)
        self.up3 = Up(128, 64)
        self.outc = OutConv(64, out_channels)


class DenoiseRatePredictor(_FusedModule):
    """Drop-in for DenoiseRatePredictor (ASDQE/ASDQE_model.py:123-171)."""

    def __init__(self, in_channels: int = 3, dim: int = 16) -> None:
        super().__init__("asdqe")
        if not (1 <= in_channels <= 4) or dim % 8 or 3 * dim > 64:
            raise NotImplementedError("DenoiseRatePredictor: built for in_channels <= 4 and dim in {8, 16} (shipped: 3, 16)")
        self.unet_multiple = dim
        self.lq_extractor = DoubleConv(in_channels, dim)
        self.gt_extractor = DoubleConv(in_channels, dim)
        self.diff_extractor = DoubleConv(in_channels, dim)
        self.unet = UNet(dim * 3, dim * 3)
        reg = nn.Sequential()  # :143-154: indices 2, 5, 8 are the Linear layers
        reg.add_module("2", _LinearParams(dim * 3, 256))
        reg.add_module("5", _LinearParams(256, 64))
        reg.add_module("8", _LinearParams(64, 1))
        self.regressor = reg
        self.regressor[-1].bias.data.fill_(0.0)  # :156 (regressor[-2] is the last Linear in the reference)
        cfg = _lib.AsdqeCfg()
        cfg.in_channels, cfg.dim = in_channels, dim
        self._cfg = cfg
        self._in_channels = in_channels

    def _run(self, lq: torch.Tensor, gt: torch.Tensor, want_feat: bool):
        eng = self._engine
        eng.require_cuda(lq, "DenoiseRatePredictor")
        self._no_autograd("DenoiseRatePredictor", lq, gt)
        if lq.shape != gt.shape or lq.dim() != 4 or lq.shape[1] != self._in_channels:
            raise RuntimeError(f"DenoiseRatePredictor: expected lq/gt [B,{self._in_channels},H,W] of equal shape, got "
                               f"{tuple(lq.shape)} and {tuple(gt.shape)}")
        B, _, H, W = lq.shape
        lib, cfg, dev = _lib.load(), self._cfg, lq.device
        prec = _lib.PRECISIONS[self.precision]
        with torch.cuda.device(dev):
            lqc = lq.detach().to(torch.float32).contiguous()
            gtc = gt.detach().to(device=dev, dtype=torch.float32).contiguous()
            tensors = [v if v.is_floating_point() else None for v in eng.tensors(self)]   # num_batches_tracked -> NULL
            nbytes = lib.kdlae_asdqe_packed_bytes(cfg, prec)

            def pack(arr, n, blob):
                _lib.check(lib.kdlae_asdqe_pack(cfg, arr, n, blob.data_ptr(), blob.numel(), prec, eng.stream()), "kdlae_asdqe_pack")

            packed = eng.packed(tensors, dev, prec, nbytes, pack)
            one = lib.kdlae_asdqe_workspace_bytes(cfg, 1, H, W, prec)
            mb = min(B, self.micro_batch or eng.pick_micro_batch(B, one, dev, cap=32))
            ws = eng.workspace(dev, lib.kdlae_asdqe_workspace_bytes(cfg, mb, H, W, prec))
            score = torch.empty((B, 1), dtype=torch.float32, device=dev)
            feat: Optional[torch.Tensor] = None
            if want_feat:
                m = self.unet_multiple
                feat = torch.empty((B, 3 * cfg.dim, (H + m - 1) // m * m, (W + m - 1) // m * m), dtype=torch.float32, device=dev)
            _lib.check(lib.kdlae_asdqe_forward(cfg, packed.data_ptr(), lqc.data_ptr(), gtc.data_ptr(), score.data_ptr(), _ptr(feat),
                                               B, H, W, mb, ws.data_ptr(), ws.numel(), prec, eng.stream()), "kdlae_asdqe_forward")
        return score, feat

    def forward(self, lq: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
        return self._run(lq, gt, False)[0]

    def forward_with_features(self, lq: torch.Tensor, gt: torch.Tensor):
        """(score [B,1], enhanced_feat [B,3*dim,Hp,Wp]) - the U-Net output of ASDQE_model.py:167, for parity tests."""
        return self._run(lq, gt, True)


__all__ = ["DenoiseRatePredictor"]
