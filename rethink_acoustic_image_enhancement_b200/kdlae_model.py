"""Drop-in KDLAE_teacher / KDLAE_student with the reference's constructors, state_dict and forward.

Mirrors KDLAE/KDLAE_model.py (reference) at the nn.Module boundary only:
  * constructor kwargs            KDLAE_model.py:205-218 (teacher), :341-342 (student)
  * state_dict() keys/shapes/order (483 tensors for the shipped teacher, 26 for the student)
  * forward({'img','denoise_rate'}) -> {'hq','sr'}  (:270-336)   /   forward(x[B,F,H,W]) (:395-430)

The sub-modules below only *hold parameters* under the reference's names; none of them computes.
The forward is one call into libkdlae_b200.so (hand-written sm_100a CUDA).  No cuDNN, no CPU path.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib
from ._engine import Engine, default_precision, _ptr


# ------------------------------------------------------------------------------------------
# parameter holders (same names / shapes / default initialisers as the reference's layers)
# ------------------------------------------------------------------------------------------
class _ConvParams(nn.Module):
    """Holds `weight` (and `bias`) shaped like nn.ConvNd; initialised like nn.ConvNd.reset_parameters."""

    def __init__(self, cin: int, cout: int, ksize: Sequence[int], groups: int = 1, bias: bool = False, transposed: bool = False):
        super().__init__()
        ksize = tuple(ksize)
        shape = (cin, cout // groups, *ksize) if transposed else (cout, cin // groups, *ksize)
        self.weight = nn.Parameter(torch.empty(shape))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if bias:
            fan_in = shape[1] * math.prod(ksize)
            bound = 1.0 / math.sqrt(fan_in) if fan_in > 0 else 0.0
            self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))
        else:
            self.register_parameter("bias", None)

    def forward(self, *a, **k):  # pragma: no cover - never used
        raise RuntimeError("parameter holder: compute happens in the fused CUDA forward of the parent module")


class _LNBody(nn.Module):
    def __init__(self, dim: int, with_bias: bool):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        if with_bias:
            self.bias = nn.Parameter(torch.zeros(dim))


class LayerNorm(nn.Module):  # KDLAE_model.py:73-83
    def __init__(self, dim: int, LayerNorm_type: str):
        super().__init__()
        self.body = _LNBody(dim, LayerNorm_type != "BiasFree")


class FeedForward(nn.Module):  # KDLAE_model.py:89-99
    def __init__(self, dim: int, ffn_expansion_factor: float, bias: bool):
        super().__init__()
        hidden = int(dim * ffn_expansion_factor)
        self.project_in = _ConvParams(dim, hidden * 2, (1, 1), bias=bias)
        self.dwconv = _ConvParams(hidden * 2, hidden * 2, (3, 3), groups=hidden * 2, bias=bias)
        self.project_out = _ConvParams(hidden, dim, (1, 1), bias=bias)


class Attention(nn.Module):  # KDLAE_model.py:112-120
    def __init__(self, dim: int, num_heads: int, bias: bool):
        super().__init__()
        self.num_heads = num_heads
        self.temperature = nn.Parameter(torch.ones(num_heads, 1, 1))
        self.qkv = _ConvParams(dim, dim * 3, (1, 1), bias=bias)
        self.qkv_dwconv = _ConvParams(dim * 3, dim * 3, (3, 3), groups=dim * 3, bias=bias)
        self.project_out = _ConvParams(dim, dim, (1, 1), bias=bias)


class TransformerBlock(nn.Module):  # KDLAE_model.py:150-157
    def __init__(self, dim, num_heads, ffn_expansion_factor, bias, LayerNorm_type):
        super().__init__()
        self.norm1 = LayerNorm(dim, LayerNorm_type)
        self.attn = Attention(dim, num_heads, bias)
        self.norm2 = LayerNorm(dim, LayerNorm_type)
        self.ffn = FeedForward(dim, ffn_expansion_factor, bias)


class OverlapPatchEmbed(nn.Module):  # KDLAE_model.py:169-173
    def __init__(self, in_c=3, embed_dim=48, bias=False):
        super().__init__()
        self.proj = _ConvParams(in_c, embed_dim, (3, 3), bias=bias)


class Downsample(nn.Module):  # KDLAE_model.py:182-187  (conv n -> n/2, PixelUnshuffle(2))
    def __init__(self, n_feat):
        super().__init__()
        self.body = nn.ModuleList([_ConvParams(n_feat, n_feat // 2, (3, 3))])


class Upsample(nn.Module):  # KDLAE_model.py:192-197  (conv n -> 2n, PixelShuffle(2))
    def __init__(self, n_feat):
        super().__init__()
        self.body = nn.ModuleList([_ConvParams(n_feat, n_feat * 2, (3, 3))])


def _stage(n, dim, heads, ffn, bias, ln):
    return nn.Sequential(*[TransformerBlock(dim, heads, ffn, bias, ln) for _ in range(n)])


class _FusedModule(nn.Module):
    """Common runtime: precision switch, packed-weight cache invalidation on load_state_dict."""

    def __init__(self, kind: str):
        super().__init__()
        self._engine = Engine(kind)
        self.precision = default_precision()
        self.micro_batch: Optional[int] = None   # None = pick from free HBM

    def set_precision(self, precision: str) -> "_FusedModule":
        """'bf16' (tcgen05 tensor-core path, default) or 'fp32' (reference-grade FFMA path, <= 1e-4)."""
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {list(_lib.PRECISIONS)}")
        self.precision = precision
        return self

    def _load_from_state_dict(self, *args, **kwargs):
        self._engine.invalidate()
        return super()._load_from_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._engine.invalidate()
        return super()._apply(fn, *args, **kwargs)

    def train(self, mode: bool = True):
        """The packed kernel weights are keyed on the parameters' (data_ptr, version counter).  An in-place update through
        ``param.data`` (the reference's own ``model_ema``, base_model.py) does not move either, so every ``train()`` / ``eval()``
        call - which the reference's validation hooks make before each forward - also drops the packed copy."""
        self._engine.invalidate()
        return super().train(mode)

    def refresh_weights(self) -> "_FusedModule":
        """Force a re-pack of the kernel weights at the next forward (after modifying parameters through ``.data``)."""
        self._engine.invalidate()
        return self

    def _wants_autograd(self, *inputs: Optional[torch.Tensor]) -> bool:
        if not torch.is_grad_enabled():
            return False
        wants_param_grad = self.training and any(p.requires_grad for p in self.parameters())
        return wants_param_grad or any(t is not None and t.requires_grad for t in inputs)

    def _no_autograd(self, name: str, *inputs: Optional[torch.Tensor]) -> None:
        """The fused forward records no autograd graph.  Refuse loudly wherever the reference would have propagated a
        gradient: training mode with trainable parameters, or (in any mode, e.g. a frozen teacher used as a loss term on a
        student's output) an input that requires grad - silently returning a constant would zero that gradient."""
        if self._wants_autograd(*inputs):
            raise NotImplementedError(
                f"{name}: the fused CUDA forward has no backward yet (training step = SURVEY.md section 8f N1); "
                "call it under torch.no_grad() (and model.eval()) for inference and validation, and detach inputs that "
                "require grad")


# ------------------------------------------------------------------------------------------
# KDLAE-T
# ------------------------------------------------------------------------------------------
class KDLAE_teacher(_FusedModule):
    """Drop-in for KDLAE_teacher (KDLAE/KDLAE_model.py:204-336); same ctor, state_dict and forward."""

    def __init__(self,
                 inp_channels=3,
                 out_channels=3,
                 dim=48,
                 num_blocks=[4, 6, 6, 8],
                 num_refinement_blocks=4,
                 heads=[1, 2, 4, 8],
                 ffn_expansion_factor=2.66,
                 bias=False,
                 LayerNorm_type='WithBias',
                 dual_pixel_task=False,
                 static="train",
                 params='cat'):
        super().__init__("teacher")
        if bias:
            raise NotImplementedError("KDLAE_teacher(bias=True) is not built: every shipped config uses bias=False")
        if dual_pixel_task:
            raise NotImplementedError("dual_pixel_task=True is broken in the reference itself (out_hq undefined, "
                                      "KDLAE_model.py:305-321) and is not supported")
        self.params = params
        ln, ffn = LayerNorm_type, ffn_expansion_factor
        self.patch_embed = OverlapPatchEmbed(inp_channels, dim)
        self.encoder_level1 = _stage(num_blocks[0], dim, heads[0], ffn, bias, ln)
        self.down1_2 = Downsample(dim)
        self.encoder_level2 = _stage(num_blocks[1], dim * 2, heads[1], ffn, bias, ln)
        self.down2_3 = Downsample(dim * 2)
        self.encoder_level3 = _stage(num_blocks[2], dim * 4, heads[2], ffn, bias, ln)
        self.down3_4 = Downsample(dim * 4)
        self.latent = _stage(num_blocks[3], dim * 8, heads[3], ffn, bias, ln)
        self.up4_3 = Upsample(dim * 8)
        self.reduce_chan_level3 = _ConvParams(dim * 8, dim * 4, (1, 1), bias=bias)
        self.decoder_level3 = _stage(num_blocks[2], dim * 4, heads[2], ffn, bias, ln)
        self.up3_2 = Upsample(dim * 4)
        self.reduce_chan_level2 = _ConvParams(dim * 4, dim * 2, (1, 1), bias=bias)
        self.decoder_level2 = _stage(num_blocks[1], dim * 2, heads[1], ffn, bias, ln)
        self.up2_1 = Upsample(dim * 2)
        self.decoder_level1 = _stage(num_blocks[0], dim * 2, heads[0], ffn, bias, ln)
        self.refinement = _stage(num_refinement_blocks, dim * 2, heads[0], ffn, bias, ln)
        self.dual_pixel_task = dual_pixel_task
        self.output = _ConvParams(dim * 2, out_channels, (3, 3), bias=bias)
        self.output_param = _ConvParams(out_channels + 1, dim * 2, (3, 3), bias=bias)
        self.refinement_out = _stage(num_refinement_blocks, dim * 2, heads[0], ffn, bias, ln)
        self.output2 = _ConvParams(dim * 2, out_channels, (3, 3), bias=bias)
        self.static = static
        if self.static == "train":
            hc = dim * 2
            self.cen = _ConvParams(out_channels, hc, (3, 3), bias=bias)
            self.upen = Upsample(hc)
            self.enhance = _stage(num_refinement_blocks, hc // 2, heads[0], ffn, bias, ln)
            self.outputen = _ConvParams(hc // 2, out_channels, (3, 3), bias=bias)

        cfg = _lib.TeacherCfg()
        cfg.inp_channels, cfg.out_channels, cfg.dim = inp_channels, out_channels, dim
        cfg.num_blocks = (_lib.C.c_int * 4)(*num_blocks)
        cfg.num_refinement_blocks = num_refinement_blocks
        cfg.heads = (_lib.C.c_int * 4)(*heads)
        cfg.hidden = (_lib.C.c_int * 4)(*[int(dim * 2 ** l * ffn_expansion_factor) for l in range(4)])
        cfg.ln_with_bias = int(LayerNorm_type != "BiasFree")
        cfg.sr_head = int(static == "train")
        cfg.params_cat = int(params == "cat")
        self._cfg = cfg
        self._io = (inp_channels, out_channels)

    def _forward_train(self, inp_img: torch.Tensor, denoise_rate: torch.Tensor) -> Dict[str, Optional[torch.Tensor]]:
        """Training mode (or an input that requires grad): the fp32 CUDA forward-with-saves + CUDA backward of training.py
        (image_restoration_model.py:198-224 calls the net exactly like this).  BiasFree LayerNorm only - the
        shipped KDLAET.yml setting."""
        from . import training
        if self._cfg.ln_with_bias:
            raise NotImplementedError("KDLAE_teacher: the CUDA backward is built for LayerNorm_type='BiasFree' (the shipped training "
                                      "configuration); WithBias models run inference only")
        with torch.cuda.device(inp_img.device):
            hq, sr = training.teacher_train_forward(dict(self.named_parameters()), inp_img, denoise_rate,
                                                    static=self.static, mode=self.params)
        return {"hq": hq, "sr": sr}

    def forward(self, input: Dict[str, torch.Tensor]) -> Dict[str, Optional[torch.Tensor]]:
        inp_img = input["img"]
        denoise_rate = input["denoise_rate"]
        eng = self._engine
        eng.require_cuda(inp_img, "KDLAE_teacher")
        if inp_img.dim() != 4 or inp_img.shape[1] != self._io[0]:
            raise RuntimeError(f"KDLAE_teacher: expected img [B,{self._io[0]},H,W], got {tuple(inp_img.shape)}")
        B, _, H, W = inp_img.shape
        if H % 8 or W % 8:
            raise RuntimeError(f"KDLAE_teacher: H and W must be multiples of 8 (pixel_unshuffle), got {H}x{W}")
        if self._wants_autograd(inp_img, denoise_rate):
            return self._forward_train(inp_img, denoise_rate)
        lib, cfg, dev = _lib.load(), self._cfg, inp_img.device
        prec = _lib.PRECISIONS[self.precision]
        with torch.cuda.device(dev):
            img = inp_img.detach().to(torch.float32).contiguous()
            rate, per_image = None, 0
            if cfg.params_cat:
                dr = denoise_rate.detach()
                if dr.dim() != 4 or dr.shape[1] != 1:
                    raise RuntimeError(f"KDLAE_teacher: expected denoise_rate [B,1,H,W] (or [B,1,1,1]), got {tuple(dr.shape)}")
                # one value per image ([B,1,1,1], or a map expanded from it): hand the kernel B floats, never the H x W map
                if dr.shape[2:] == (1, 1) or (dr.stride(2) == 0 and dr.stride(3) == 0):
                    rate = dr[:, 0, 0, 0].to(device=dev, dtype=torch.float32).expand(B).contiguous()
                    per_image = 1
                else:
                    rate = dr.to(device=dev, dtype=torch.float32).expand(B, 1, H, W).contiguous()
            tensors = eng.tensors(self)
            nbytes = lib.kdlae_teacher_packed_bytes(cfg, prec)

            def pack(arr, n, blob):
                _lib.check(lib.kdlae_teacher_pack(cfg, arr, n, blob.data_ptr(), blob.numel(), prec, eng.stream()),
                           "kdlae_teacher_pack")

            packed = eng.packed(tensors, dev, prec, nbytes, pack)
            one = lib.kdlae_teacher_workspace_bytes(cfg, 1, H, W, prec)
            mb = self.micro_batch or eng.pick_micro_batch(B, one, dev, cap=16)
            mb = min(mb, B)
            ws_bytes = lib.kdlae_teacher_workspace_bytes(cfg, mb, H, W, prec)
            ws = eng.workspace(dev, ws_bytes)
            hq = torch.empty((B, self._io[1], H, W), dtype=torch.float32, device=dev)
            sr = torch.empty((B, self._io[1], 2 * H, 2 * W), dtype=torch.float32, device=dev) if cfg.sr_head else None
            _lib.check(lib.kdlae_teacher_forward(cfg, packed.data_ptr(), img.data_ptr(), _ptr(rate), per_image, hq.data_ptr(),
                                                 _ptr(sr), B, H, W, mb, ws.data_ptr(), ws.numel(), prec, eng.stream()),
                       "kdlae_teacher_forward")
            if inp_img.dtype != torch.float32 and inp_img.is_floating_point():   # the reference returns the input's dtype
                hq = hq.to(inp_img.dtype)
                sr = sr.to(inp_img.dtype) if sr is not None else None
        return {"hq": hq, "sr": sr}


# Alias instantiated by the reference's KDLAE-T yaml (restormer_arch.py:566, KDLAET.yml:66)
RestormerSuperResolutionParam2 = KDLAE_teacher


# ------------------------------------------------------------------------------------------
# KDLAE-S
# ------------------------------------------------------------------------------------------
class KDLAE_student(_FusedModule):
    """Drop-in for KDLAE_student (KDLAE/KDLAE_model.py:340-430)."""

    def __init__(self, inp_channels=1, out_channels=1, residual=False, hidden_channels=[16, 32, 64], kernel_size=3):
        super().__init__("student")
        if inp_channels != 1 or out_channels != 1:
            raise NotImplementedError("KDLAE_student: forward() unsqueezes a singleton feature axis (KDLAE_model.py:397); "
                                      "only inp_channels = out_channels = 1 is meaningful")
        if kernel_size != 3 or len(hidden_channels) != 3:
            raise NotImplementedError("KDLAE_student: only kernel_size=3 and three hidden_channels entries are built "
                                      "(the shipped KDLAE-S-US / KDLAE-S-FLS configuration)")
        self.residual = residual
        self.num_levels = len(hidden_channels) - 1

        def block(cin, cout):  # _create_conv_block (:386-393): indices 0 and 2 carry parameters (1, 3 are ReLU)
            seq = nn.Sequential()
            seq.add_module("0", _ConvParams(cin, cout, (3, 3, 3), bias=True))
            seq.add_module("2", _ConvParams(cout, cout, (3, 3, 3), bias=True))
            return seq

        self.encoders = nn.ModuleList()
        cin = inp_channels
        for i in range(self.num_levels):
            self.encoders.append(block(cin, hidden_channels[i]))
            cin = hidden_channels[i]
        self.st_fusion = block(cin, hidden_channels[-1])
        self.upconv_layers = nn.ModuleList()
        self.decoders = nn.ModuleList()
        for i in range(self.num_levels - 1, -1, -1):
            cup = hidden_channels[-1] if i == self.num_levels - 1 else hidden_channels[i + 1]
            self.upconv_layers.append(_ConvParams(cup, hidden_channels[i], (1, 2, 2), bias=True, transposed=True))
            self.decoders.append(block(hidden_channels[i], hidden_channels[i]))
        self.out_conv = _ConvParams(hidden_channels[0], out_channels, (1, 1, 1), bias=True)

        cfg = _lib.StudentCfg()
        cfg.hidden = (_lib.C.c_int * 3)(*hidden_channels)
        cfg.residual = int(bool(residual))
        self._cfg = cfg

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        eng = self._engine
        eng.require_cuda(x, "KDLAE_student")
        self._no_autograd("KDLAE_student", x)
        if x.dim() != 4:
            raise RuntimeError(f"KDLAE_student: expected x [B,F,H,W], got {tuple(x.shape)}")
        B, F, H, W = x.shape
        if H % 4 or W % 4:
            raise RuntimeError(f"KDLAE_student: H and W must be multiples of 4 (skip-add shapes), got {H}x{W}")
        lib, cfg, dev = _lib.load(), self._cfg, x.device
        prec = _lib.PRECISIONS[self.precision]
        with torch.cuda.device(dev):
            xin = x.detach().to(torch.float32).contiguous()
            tensors = eng.tensors(self)
            nbytes = lib.kdlae_student_packed_bytes(cfg, prec)

            def pack(arr, n, blob):
                _lib.check(lib.kdlae_student_pack(cfg, arr, n, blob.data_ptr(), blob.numel(), prec, eng.stream()),
                           "kdlae_student_pack")

            packed = eng.packed(tensors, dev, prec, nbytes, pack)
            one = lib.kdlae_student_workspace_bytes(cfg, 1, F, H, W, prec)
            mb = min(B, self.micro_batch or eng.pick_micro_batch(B, one, dev, cap=32))
            ws = eng.workspace(dev, lib.kdlae_student_workspace_bytes(cfg, mb, F, H, W, prec))
            y = torch.empty_like(xin)
            _lib.check(lib.kdlae_student_forward(cfg, packed.data_ptr(), xin.data_ptr(), y.data_ptr(), B, F, H, W, mb,
                                                 ws.data_ptr(), ws.numel(), prec, eng.stream()), "kdlae_student_forward")
            if x.dtype != torch.float32 and x.is_floating_point():
                y = y.to(x.dtype)
        return y


__all__: List[str] = ["KDLAE_teacher", "KDLAE_student", "RestormerSuperResolutionParam2"]
