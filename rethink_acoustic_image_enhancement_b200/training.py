"""First slice of the KDLAE-T training step on the device, behind the C ABI.

Mirrors, at the call boundary, what ``ImageCleanModel.optimize_parameters`` does around the network
(Train/basicsr/models/image_restoration_model.py:198-224) and what DDP does for it (base_model.py:76-82):

  * ``transformer_block_train`` = ``mdta_block_train`` + ``gdfn_block_train`` - a whole BiasFree TransformerBlock
                              (KDLAE_model.py:159-163) whose forward AND backward are CUDA kernels;
  * ``gdfn_block_train``    - the GDFN half of a TransformerBlock (``x + ffn(norm2(x))``, KDLAE_model.py:50-52,101-106,163) as an
                              autograd function whose forward AND backward are CUDA kernels (kdlae_gdfn_forward_train / _backward);
  * ``L1LossSr``            - in ``metrics.py`` (fused loss value + gradient);
  * ``FlatAdamW``           - parameters, gradients and moments in flat fp32 buffers; ``clip_grad_norm_(params, 0.01)`` + AdamW as
                              one reduction + one fused update (kdlae_grad_norm_sq / kdlae_adamw_step);
  * ``BucketedAllReducer``  - the DDP gradient all-reduce: the flat gradient buffer is cut into buckets that are all-reduced
                              (NCCL over NVLink on the GPU box, gloo in the CPU tests) on a side stream as soon as every
                              gradient of a bucket has been produced, so communication overlaps the rest of the backward.
PyTorch supplies tensors, streams, autograd bookkeeping and torch.distributed; the arithmetic is in libkdlae_b200.so.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib


def set_matmul_precision(mode: str) -> None:
    """'fp32' (default): every GEMM of the training step on the CUDA cores in fp32.  'tf32': the 1x1-conv GEMMs (forward,
    dgrad and wgrad) on tcgen05 in TF32 with fp32 accumulation - what ``torch.backends.cudnn.allow_tf32`` (PyTorch's default) gives the
    reference's ``nn.Conv2d`` layers on this GPU.  Process-wide (``kdlae_set_train_matmul_tf32``)."""
    if mode not in ("fp32", "tf32"):
        raise ValueError("set_matmul_precision: mode must be 'fp32' or 'tf32'")
    _lib.check(_lib.load().kdlae_set_train_matmul_tf32(int(mode == "tf32")), "kdlae_set_train_matmul_tf32")


def get_matmul_precision() -> str:
    return "tf32" if _lib.load().kdlae_train_matmul_tf32() else "fp32"


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ceil8(v: int) -> int:
    return (v + 7) // 8 * 8


# ------------------------------------------------------------------------------------------------------------------
# GDFN half of a TransformerBlock with a CUDA backward
# ------------------------------------------------------------------------------------------------------------------
class _GdfnFn(torch.autograd.Function):
    """x is NHWC ([B,H,W,C] fp32); parameters in the reference's shapes."""

    @staticmethod
    def forward(ctx, x, gamma, w_in, w_dw, w_out):
        if not x.is_cuda:
            raise RuntimeError("gdfn_block_train: expected CUDA tensors (there is no CPU path)")
        lib = _lib.load()
        B, H, W, C = x.shape
        h = w_out.shape[1]
        hp = _ceil8(h)
        dev = x.device
        with torch.cuda.device(dev):
            # layout plumbing (no arithmetic): reference parameter shapes -> packed layouts of the forward path
            xh = x.detach().float().contiguous()
            g = gamma.detach().float().contiguous()
            w2 = w_in.detach().float().view(2 * h, C)
            win = torch.zeros(2 * hp, C, device=dev)
            win[:h] = w2[:h]
            win[hp:hp + h] = w2[h:]
            d2 = w_dw.detach().float().view(2 * h, 9)
            wdw = torch.zeros(9, 2 * hp, device=dev)
            wdw[:, :h] = d2[:h].t()
            wdw[:, hp:hp + h] = d2[h:].t()
            wout = torch.zeros(C, hp, device=dev)
            wout[:, :h] = w_out.detach().float().view(C, h)
            out = torch.empty_like(xh)
            ws = torch.empty(lib.kdlae_gdfn_train_ws_floats(B, H, W, C, hp), dtype=torch.float32, device=dev)
            _lib.check(lib.kdlae_gdfn_forward_train(xh.data_ptr(), g.data_ptr(), win.data_ptr(), wdw.data_ptr(), wout.data_ptr(),
                                                    out.data_ptr(), B, H, W, C, hp, ws.data_ptr(), _stream()), "kdlae_gdfn_forward_train")
        ctx.save_for_backward(xh, g, win, wdw, wout, ws)
        ctx.dims = (B, C, H, W, h, hp)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        xh, g, win, wdw, wout, ws = ctx.saved_tensors
        B, C, H, W, h, hp = ctx.dims
        lib = _lib.load()
        dev = xh.device
        with torch.cuda.device(dev):
            dout = grad_out.detach().float().contiguous()
            dx = torch.empty_like(xh)
            dg = torch.empty(C, device=dev)
            dwin, dwdw, dwout = torch.empty_like(win), torch.empty_like(wdw), torch.empty_like(wout)
            _lib.check(lib.kdlae_gdfn_backward(xh.data_ptr(), g.data_ptr(), win.data_ptr(), wdw.data_ptr(), wout.data_ptr(), dout.data_ptr(),
                                               dx.data_ptr(), dg.data_ptr(), dwin.data_ptr(), dwdw.data_ptr(), dwout.data_ptr(),
                                               B, H, W, C, hp, ws.data_ptr(), _stream()), "kdlae_gdfn_backward")
        d_w_in = torch.cat([dwin[:h], dwin[hp:hp + h]]).view(2 * h, C, 1, 1)
        d_w_dw = torch.cat([dwdw[:, :h].t(), dwdw[:, hp:hp + h].t()]).reshape(2 * h, 1, 3, 3)
        d_w_out = dwout[:, :h].reshape(C, h, 1, 1)
        return dx, dg, d_w_in, d_w_dw, d_w_out


def gdfn_block_train(x: torch.Tensor, norm_weight: torch.Tensor, project_in_weight: torch.Tensor, dwconv_weight: torch.Tensor,
                     project_out_weight: torch.Tensor) -> torch.Tensor:
    """``x + FeedForward(BiasFree_LayerNorm(x))`` (KDLAE_model.py:163) with forward and backward in CUDA (fp32 path).
    x [B,C,H,W]; parameters in the reference's shapes: norm2.body.weight [C], ffn.project_in.weight [2h,C,1,1],
    ffn.dwconv.weight [2h,1,3,3], ffn.project_out.weight [C,h,1,1]."""
    return _nchw(_GdfnFn.apply(_nhwc(x), norm_weight, project_in_weight, dwconv_weight, project_out_weight))


class _MdtaFn(torch.autograd.Function):
    """x is NHWC ([B,H,W,C] fp32); parameters in the reference's shapes."""

    @staticmethod
    def forward(ctx, x, gamma, temperature, w_qkv, w_dw, w_proj):
        if not x.is_cuda:
            raise RuntimeError("mdta_block_train: expected CUDA tensors (there is no CPU path)")
        lib = _lib.load()
        B, H, W, C = x.shape
        heads = temperature.numel()
        dev = x.device
        with torch.cuda.device(dev):
            xh = x.detach().float().contiguous()
            g = gamma.detach().float().contiguous()
            tp = temperature.detach().float().reshape(heads).contiguous()
            wq = w_qkv.detach().float().view(3 * C, C).contiguous()
            wd = w_dw.detach().float().view(3 * C, 9).t().contiguous()          # [9][3C]
            wp = w_proj.detach().float().view(C, C).contiguous()
            out = torch.empty_like(xh)
            ws = torch.empty(lib.kdlae_mdta_train_ws_floats(B, H, W, C, heads), dtype=torch.float32, device=dev)
            _lib.check(lib.kdlae_mdta_forward_train(xh.data_ptr(), g.data_ptr(), wq.data_ptr(), wd.data_ptr(), wp.data_ptr(), tp.data_ptr(),
                                                    out.data_ptr(), B, H, W, C, heads, ws.data_ptr(), _stream()), "kdlae_mdta_forward_train")
        ctx.save_for_backward(xh, g, tp, wq, wd, wp, ws)
        ctx.dims = (B, C, H, W, heads, tuple(temperature.shape))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        xh, g, tp, wq, wd, wp, ws = ctx.saved_tensors
        B, C, H, W, heads, tshape = ctx.dims
        lib = _lib.load()
        dev = xh.device
        with torch.cuda.device(dev):
            dout = grad_out.detach().float().contiguous()
            dx = torch.empty_like(xh)
            dg, dtp = torch.empty(C, device=dev), torch.empty(heads, device=dev)
            dwq, dwd, dwp = torch.empty_like(wq), torch.empty_like(wd), torch.empty_like(wp)
            _lib.check(lib.kdlae_mdta_backward(xh.data_ptr(), g.data_ptr(), wq.data_ptr(), wd.data_ptr(), wp.data_ptr(), tp.data_ptr(),
                                               dout.data_ptr(), dx.data_ptr(), dg.data_ptr(), dwq.data_ptr(), dwd.data_ptr(), dwp.data_ptr(),
                                               dtp.data_ptr(), B, H, W, C, heads, ws.data_ptr(), _stream()), "kdlae_mdta_backward")
        return (dx, dg, dtp.view(tshape), dwq.view(3 * C, C, 1, 1),
                dwd.t().reshape(3 * C, 1, 3, 3), dwp.view(C, C, 1, 1))


def mdta_block_train(x: torch.Tensor, norm_weight: torch.Tensor, temperature: torch.Tensor, qkv_weight: torch.Tensor,
                     qkv_dwconv_weight: torch.Tensor, project_out_weight: torch.Tensor) -> torch.Tensor:
    """``x + Attention(BiasFree_LayerNorm(x))`` (KDLAE_model.py:162) with forward and backward in CUDA (fp32 path).
    Parameters in the reference's shapes: norm1.body.weight [C], attn.temperature [heads,1,1], attn.qkv.weight [3C,C,1,1],
    attn.qkv_dwconv.weight [3C,1,3,3], attn.project_out.weight [C,C,1,1]."""
    return _nchw(_MdtaFn.apply(_nhwc(x), norm_weight, temperature, qkv_weight, qkv_dwconv_weight, project_out_weight))


def _block_nhwc(x: torch.Tensor, p: dict, prefix: str) -> torch.Tensor:
    x = _MdtaFn.apply(x, p[prefix + ".norm1.body.weight"], p[prefix + ".attn.temperature"], p[prefix + ".attn.qkv.weight"],
                      p[prefix + ".attn.qkv_dwconv.weight"], p[prefix + ".attn.project_out.weight"])
    return _GdfnFn.apply(x, p[prefix + ".norm2.body.weight"], p[prefix + ".ffn.project_in.weight"], p[prefix + ".ffn.dwconv.weight"],
                         p[prefix + ".ffn.project_out.weight"])


def transformer_block_train(x: torch.Tensor, p: dict, prefix: str) -> torch.Tensor:
    """One BiasFree TransformerBlock (KDLAE_model.py:159-163) in training mode from a dict of the reference's parameters
    (x and the result are NCHW)."""
    return _nchw(_block_nhwc(_nhwc(x), p, prefix))


# ------------------------------------------------------------------------------------------------------------------
# dense convolutions outside the blocks, and the whole KDLAE-T training forward
# ------------------------------------------------------------------------------------------------------------------
class _ConvFn(torch.autograd.Function):
    """Conv2d(Cin, Cout, k, stride 1, padding dilation * (k // 2), bias=False) on NHWC fp32 with a CUDA backward."""

    @staticmethod
    def forward(ctx, x, weight, dilation):
        if not x.is_cuda:
            raise RuntimeError("conv_train: expected CUDA tensors (there is no CPU path)")
        lib = _lib.load()
        B, H, W, Cin = x.shape
        Cout, _, k, _ = weight.shape
        dev = x.device
        with torch.cuda.device(dev):
            xh = x.detach().float().contiguous()
            w = weight.detach().float().permute(0, 2, 3, 1).reshape(Cout, k * k, Cin).contiguous()      # [Cout][tap][Cin]
            out = torch.empty((B, H, W, Cout), dtype=torch.float32, device=dev)
            _lib.check(lib.kdlae_conv_train_forward(xh.data_ptr(), w.data_ptr(), out.data_ptr(), B, H, W, Cin, Cout, k, int(dilation),
                                                    _stream()), "kdlae_conv_train_forward")
        ctx.save_for_backward(xh, w)
        ctx.dims = (B, H, W, Cin, Cout, k, int(dilation))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        xh, w = ctx.saved_tensors
        B, H, W, Cin, Cout, k, dil = ctx.dims
        lib = _lib.load()
        dev = xh.device
        with torch.cuda.device(dev):
            dout = grad_out.detach().float().contiguous()
            dx = torch.empty_like(xh) if ctx.needs_input_grad[0] else None
            dw = torch.empty_like(w)
            ws = torch.empty(lib.kdlae_conv_train_ws_floats(B, H, W, Cin, Cout, k), dtype=torch.float32, device=dev)
            _lib.check(lib.kdlae_conv_train_backward(xh.data_ptr(), w.data_ptr(), dout.data_ptr(), None if dx is None else dx.data_ptr(),
                                                     dw.data_ptr(), B, H, W, Cin, Cout, k, dil, ws.data_ptr(), _stream()),
                       "kdlae_conv_train_backward")
        # contiguous in the parameter's own layout (DDP's bucket views and the flat gradient buffer expect that)
        return dx, dw.view(Cout, k, k, Cin).permute(0, 3, 1, 2).contiguous(), None


def conv_train(x: torch.Tensor, weight: torch.Tensor, dilation: int = 1) -> torch.Tensor:
    """``F.conv2d(x, weight, None, padding=dilation * (k // 2), dilation=dilation)`` for NCHW x with forward and backward in CUDA."""
    return _nchw(_ConvFn.apply(_nhwc(x), weight, dilation))


def _nhwc(x: torch.Tensor) -> torch.Tensor:
    return x.permute(0, 2, 3, 1)


def _nchw(x: torch.Tensor) -> torch.Tensor:
    return x.permute(0, 3, 1, 2).contiguous()


def _unshuffle2(x: torch.Tensor) -> torch.Tensor:
    """nn.PixelUnshuffle(2) on NHWC (pure data movement): out[.., c*4 + 2*dy + dx] = in[2y+dy, 2x+dx, c]."""
    B, H, W, C = x.shape
    return x.reshape(B, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 5, 2, 4).reshape(B, H // 2, W // 2, C * 4)


def _shuffle2(x: torch.Tensor) -> torch.Tensor:
    """nn.PixelShuffle(2) on NHWC."""
    B, H, W, C4 = x.shape
    C = C4 // 4
    return x.reshape(B, H, W, C, 2, 2).permute(0, 1, 4, 2, 5, 3).reshape(B, 2 * H, 2 * W, C)


def teacher_train_forward(params: dict, img: torch.Tensor, denoise_rate: torch.Tensor, static: str = "train", mode: str = "cat"):
    """KDLAE_teacher.forward (KDLAE_model.py:270-336, BiasFree LayerNorm, bias=False) in training mode: every convolution and
    every TransformerBlock runs a CUDA forward that saves what its CUDA backward needs; PixelShuffle / PixelUnshuffle, the
    channel concats and the two adds are torch data movement that autograd differentiates.  ``params``: name -> parameter in the
    reference's state_dict naming (``dict(model.named_parameters())``).  Returns (hq, sr) like the reference's dict entries."""
    def nb(prefix):
        n = 0
        while f"{prefix}.{n}.norm1.body.weight" in params:
            n += 1
        return n

    def blocks(x, prefix):
        for i in range(nb(prefix)):
            x = _block_nhwc(x, params, f"{prefix}.{i}")
        return x

    conv = lambda t, key, dil=1: _ConvFn.apply(t, params[key + ".weight"], dil)
    x0 = _nhwc(img.float()).contiguous()
    e1 = blocks(conv(x0, "patch_embed.proj"), "encoder_level1")
    e2 = blocks(_unshuffle2(conv(e1, "down1_2.body.0")), "encoder_level2")
    e3 = blocks(_unshuffle2(conv(e2, "down2_3.body.0")), "encoder_level3")
    lat = blocks(_unshuffle2(conv(e3, "down3_4.body.0")), "latent")
    d3 = conv(torch.cat([_shuffle2(conv(lat, "up4_3.body.0")), e3], dim=-1), "reduce_chan_level3")
    d3 = blocks(d3, "decoder_level3")
    d2 = conv(torch.cat([_shuffle2(conv(d3, "up3_2.body.0")), e2], dim=-1), "reduce_chan_level2")
    d2 = blocks(d2, "decoder_level2")
    d1 = blocks(torch.cat([_shuffle2(conv(d2, "up2_1.body.0")), e1], dim=-1), "decoder_level1")
    d1 = blocks(d1, "refinement")
    out = conv(d1, "output")
    if mode == "cat":
        rate = denoise_rate.float()
        if rate.dim() == 4 and rate.shape[-2:] != img.shape[-2:]:
            rate = rate.expand(img.shape[0], 1, img.shape[2], img.shape[3])
        out = conv(torch.cat([out, _nhwc(rate)], dim=-1), "output_param", 2)
        out = conv(blocks(out, "refinement_out"), "output2")
    hq = out + x0
    sr = None
    if static == "train":
        s = _shuffle2(conv(conv(hq, "cen"), "upen.body.0"))
        sr = _nchw(conv(blocks(s, "enhance"), "outputen"))
    return _nchw(hq), sr


# ------------------------------------------------------------------------------------------------------------------
# flat parameters + fused clip-norm / AdamW
# ------------------------------------------------------------------------------------------------------------------
class FlatAdamW:
    """AdamW over ONE flat fp32 buffer (parameters are re-pointed to views of it, gradients accumulate into views of a flat
    gradient buffer), with ``torch.nn.utils.clip_grad_norm_(params, max_norm)`` fused into the update.
    Defaults follow Train/Denoising/Options/paper202508/KDLAET.yml (AdamW, lr 3e-4, betas (0.9, 0.999), weight_decay 1e-4) and
    image_restoration_model.py:218-219 (clip at 0.01)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 3e-4, betas: Sequence[float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-4, max_norm: float = 0.01):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatAdamW: no trainable parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.offsets: List[int] = []
        off = 0
        for p in self.params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.detach().reshape(-1))
            p.data = self.flat[off:off + k].view_as(p)
            p.grad = self.grad[off:off + k].view_as(p)
            self.offsets.append(off)
            off += k
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = lr, tuple(betas), eps, weight_decay, max_norm
        self.steps = 0
        if dev.type == "cuda":
            self._norm_sq = torch.zeros(1, dtype=torch.float64, device=dev)
            self._scratch = torch.empty(1024, dtype=torch.float64, device=dev)

    def zero_grad(self) -> None:
        self.grad.zero_()

    def grad_norm(self) -> torch.Tensor:
        """Total gradient norm (device scalar) as clip_grad_norm_ computes it."""
        self._require_cuda()
        _lib.check(_lib.load().kdlae_grad_norm_sq(self.grad.data_ptr(), self.grad.numel(), self._norm_sq.data_ptr(),
                                                  self._scratch.data_ptr(), _stream()), "kdlae_grad_norm_sq")
        return self._norm_sq.sqrt()

    def step(self) -> None:
        self._require_cuda()
        lib = _lib.load()
        self.steps += 1
        with torch.cuda.device(self.flat.device):
            clip = self.max_norm is not None and self.max_norm > 0
            if clip:
                _lib.check(lib.kdlae_grad_norm_sq(self.grad.data_ptr(), self.grad.numel(), self._norm_sq.data_ptr(),
                                                  self._scratch.data_ptr(), _stream()), "kdlae_grad_norm_sq")
            _lib.check(lib.kdlae_adamw_step(self.flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                                            self.exp_avg_sq.data_ptr(), self.flat.numel(), self.lr, self.betas[0], self.betas[1],
                                            self.eps, self.weight_decay, self.steps, float(self.max_norm or 0.0),
                                            self._norm_sq.data_ptr() if clip else None, _stream()), "kdlae_adamw_step")
            # the kernel wrote the parameters through raw pointers: bump their autograd version counters (one multi-tensor no-op),
            # so that everything keyed on them - the packed-weight cache of the fused inference forward - sees the update
            with torch.no_grad():
                torch._foreach_add_(self.params, 0.0)

    def _require_cuda(self) -> None:
        if self.flat.device.type != "cuda":
            raise RuntimeError("FlatAdamW: the fused update runs in CUDA only (there is no CPU path)")


# ------------------------------------------------------------------------------------------------------------------
# DDP gradient all-reduce over the flat gradient buffer
# ------------------------------------------------------------------------------------------------------------------
class BucketedAllReducer:
    """Average a flat gradient buffer over the ranks in buckets (base_model.py:76-82 wraps the net in DDP; this is the same
    collective, issued by hand over the flat buffer).  Buckets are filled from the END of the parameter list - the order in
    which backward produces gradients - and each bucket's all-reduce is launched on a communication stream as soon as its last
    gradient has been accumulated (``attach``), or all at once (``all_reduce``)."""

    def __init__(self, flat_grad: torch.Tensor, bucket_bytes: int = 25 << 20, group: Optional[dist.ProcessGroup] = None):
        self.flat, self.group = flat_grad, group
        n, per = flat_grad.numel(), max(1, bucket_bytes // flat_grad.element_size())
        self.bounds: List[tuple] = []
        hi = n
        while hi > 0:                       # bucket 0 = the tail of the buffer (first gradients to be ready)
            lo = max(0, hi - per)
            self.bounds.append((lo, hi))
            hi = lo
        self.comm_stream = torch.cuda.Stream(flat_grad.device) if flat_grad.is_cuda else None
        self._pending: List = []
        self._ready = [0] * len(self.bounds)
        self._need = [0] * len(self.bounds)
        self._hooks: List = []

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def _launch(self, b: int) -> None:
        lo, hi = self.bounds[b]
        chunk = self.flat[lo:hi]
        if self.world == 1:
            return
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream(self.flat.device))
            with torch.cuda.stream(self.comm_stream):
                self._pending.append(dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
        else:                               # gloo (CPU tests): no AVG reduction
            w = dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._pending.append((w, chunk))

    def all_reduce(self) -> None:
        for b in range(len(self.bounds)):
            self._launch(b)
        self.wait()

    def attach(self, params: Sequence[torch.nn.Parameter], offsets: Sequence[int]) -> None:
        """Launch each bucket from autograd as soon as all gradients that end inside it have been accumulated."""
        self._need = [0] * len(self.bounds)
        owners = []
        for p, off in zip(params, offsets):
            end = off + p.numel()
            bs = [i for i, (lo, hi) in enumerate(self.bounds) if lo < end and off < hi]     # every bucket the gradient overlaps
            owners.append(bs)
            for b in bs:
                self._need[b] += 1
        for p, bs in zip(params, owners):
            def hook(_p, bs=bs):
                for b in bs:
                    self._ready[b] += 1
                    if self._ready[b] == self._need[b]:
                        self._launch(b)
            self._hooks.append(p.register_post_accumulate_grad_hook(hook))

    def wait(self) -> None:
        for w in self._pending:
            if isinstance(w, tuple):
                w[0].wait()
                w[1].div_(self.world)
            else:
                w.wait()
        self._pending.clear()
        self._ready = [0] * len(self.bounds)
        if self.comm_stream is not None:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.comm_stream)


__all__ = ["gdfn_block_train", "mdta_block_train", "transformer_block_train", "conv_train", "teacher_train_forward", "FlatAdamW",
           "BucketedAllReducer", "set_matmul_precision", "get_matmul_precision"]
