"""GPU: single fused stages through the C ABI against torch-CPU float64 restatements of the same op."""
import os
import pytest
import torch
import torch.nn.functional as F

from rethink_acoustic_image_enhancement_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _conv_ref(a_nhwc, w_ntc, ksize, row_scale, bias, relu, res):
    """float64 CPU reference. a: [n,H,W,C], w: [N, k*k, C]."""
    a = a_nhwc.double().cpu().permute(0, 3, 1, 2)
    N, taps, C = w_ntc.shape
    w = w_ntc.double().cpu().view(N, ksize, ksize, C).permute(0, 3, 1, 2)
    y = F.conv2d(a, w, padding=ksize // 2).permute(0, 2, 3, 1)
    if row_scale is not None:
        y = y * row_scale.double().cpu().view(*y.shape[:3], 1)
    if bias is not None:
        y = y + bias.double().cpu()
    if res is not None:
        y = y + res.double().cpu().view_as(y)
    if relu:
        y = y.clamp_min(0)
    return y


def _run_conv(lib, a, w, ksize, row_scale, bias, relu, res, prec, force_simt=0):
    n, H, W, C = a.shape
    N = w.shape[0]
    out = torch.full((n, H, W, N), float("nan"), dtype=a.dtype, device=DEV)
    st = lib.kdlae_conv_gemm(a.data_ptr(), C, w.data_ptr(), N, n, H, W, ksize,
                             None if row_scale is None else row_scale.data_ptr(),
                             None if bias is None else bias.data_ptr(), relu,
                             None if res is None else res.data_ptr(), out.data_ptr(), prec, force_simt, _stream())
    _lib.check(st, "kdlae_conv_gemm")
    torch.cuda.synchronize()
    return out


CASES = [  # (n, H, W, C, N, ksize)
    (2, 24, 20, 48, 144, 1),    # qkv of level 1; rows not a multiple of 128
    (1, 16, 16, 96, 48, 1),
    (3, 8, 8, 128, 512, 1),     # N split in 2 chunks
    (1, 8, 24, 384, 1152, 1),   # latent qkv: 5 chunks, K = 6 stages
    (2, 24, 20, 48, 24, 3),     # down1_2 shape, N not a multiple of 16
    (1, 16, 40, 96, 192, 3),    # up2_1 shape
    (1, 8, 8, 64, 64, 3),       # image smaller than one spatial tile
]


@pytest.mark.parametrize("case", CASES)
def test_conv_gemm_fp32_simt(lib, case):
    n, H, W, C, N, k = case
    g = torch.Generator().manual_seed(hash(case) & 0xFFFF)
    a = torch.randn(n, H, W, C, generator=g).to(DEV)
    w = (torch.randn(N, k * k, C, generator=g) / (C * k * k) ** 0.5).to(DEV)
    rs = (0.5 + torch.rand(n, H, W, generator=g)).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(n, H, W, N, generator=g).to(DEV)
    out = _run_conv(lib, a, w, k, rs, bias, 1, res, 0)
    ref = _conv_ref(a, w, k, rs, bias, 1, res)
    assert (out.double().cpu() - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("force_simt", [0, 1])
def test_conv_gemm_bf16(lib, case, force_simt):
    """force_simt=0: tcgen05/TMEM/TMA kernel; 1: CUDA-core kernel with bf16 storage."""
    n, H, W, C, N, k = case
    g = torch.Generator().manual_seed(hash(case) & 0xFFFF)
    a = torch.randn(n, H, W, C, generator=g).to(DEV).bfloat16()
    w = (torch.randn(N, k * k, C, generator=g) / (C * k * k) ** 0.5).to(DEV).bfloat16()
    rs = (0.5 + torch.rand(n, H, W, generator=g)).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(n, H, W, N, generator=g).to(DEV).bfloat16()
    for relu, use_rs, use_res in ((0, False, False), (1, True, True)):
        out = _run_conv(lib, a, w, k, rs if use_rs else None, bias if use_rs else None, relu, res if use_res else None, 1, force_simt)
        ref = _conv_ref(a, w, k, rs if use_rs else None, bias if use_rs else None, relu, res if use_res else None)
        err = (out.double().cpu() - ref).abs().max().item()
        scale = ref.abs().max().item()
        assert torch.isfinite(out).all()
        assert err < 1.2e-2 * scale, f"max err {err} vs scale {scale}"  # one bf16 rounding of the output (2^-8 relative)


@pytest.mark.parametrize("prec,dtype", [(0, torch.float32), (1, torch.bfloat16)])
def test_ln_stats(lib, prec, dtype):
    rows, C = 1000, 96
    x = (torch.randn(rows, C) * 2 + 0.7).to(DEV).to(dtype)
    rstd = torch.empty(rows, device=DEV)
    mu = torch.empty(rows, device=DEV)
    _lib.check(lib.kdlae_ln_stats(x.data_ptr(), C, rows, rstd.data_ptr(), mu.data_ptr(), prec, _stream()), "ln_stats")
    torch.cuda.synchronize()
    xd = x.double().cpu()
    ref_mu = xd.mean(-1)
    ref_rstd = 1.0 / torch.sqrt(xd.var(-1, unbiased=False) + 1e-5)
    assert (mu.double().cpu() - ref_mu).abs().max().item() < 1e-5
    assert ((rstd.double().cpu() - ref_rstd) / ref_rstd).abs().max().item() < 1e-5


@pytest.mark.parametrize("prec,dtype,tol", [(0, torch.float32, 1e-5), (1, torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("gate", [0, 1])
def test_dwconv3x3(lib, prec, dtype, tol, gate):
    n, H, W, C = 2, 10, 13, 64
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, H, W, C, generator=g).to(DEV).to(dtype)
    w = torch.randn(C, 1, 3, 3, generator=g) / 3
    w9c = w.view(C, 9).t().contiguous().to(DEV)
    Co = C // 2 if gate else C
    out = torch.empty(n, H, W, Co, dtype=dtype, device=DEV)
    _lib.check(lib.kdlae_dwconv3x3(x.data_ptr(), out.data_ptr(), w9c.data_ptr(), n, H, W, C, gate, prec, _stream()), "dwconv")
    torch.cuda.synchronize()
    y = F.conv2d(x.double().cpu().permute(0, 3, 1, 2), w.double(), padding=1, groups=C)
    if gate:
        y = F.gelu(y[:, :Co]) * y[:, Co:]
    ref = y.permute(0, 2, 3, 1)
    assert (out.double().cpu() - ref).abs().max().item() < tol * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("shape", [(1, 40, 72, 144, 0), (2, 16, 33, 288, 0), (1, 24, 64, 256, 1), (1, 8, 8, 1024, 1), (3, 9, 5, 48, 0)])
def test_dwconv3x3_bf16_tma_tiles(lib, shape):
    """TMA-staged tile kernel: several tiles, ragged edges, partial 64-channel blocks, gate halves."""
    n, H, W, C, gate = shape
    g = torch.Generator().manual_seed(11)
    x = torch.randn(n, H, W, C, generator=g).to(DEV).bfloat16()
    w = torch.randn(C, 1, 3, 3, generator=g) / 3
    w9c = w.view(C, 9).t().contiguous().to(DEV)
    Co = C // 2 if gate else C
    out = torch.full((n, H, W, Co), float("nan"), dtype=torch.bfloat16, device=DEV)
    _lib.check(lib.kdlae_dwconv3x3(x.data_ptr(), out.data_ptr(), w9c.data_ptr(), n, H, W, C, gate, 1, _stream()), "dwconv")
    torch.cuda.synchronize()
    y = F.conv2d(x.double().cpu().permute(0, 3, 1, 2), w.double(), padding=1, groups=C)
    if gate:
        y = F.gelu(y[:, :Co]) * y[:, Co:]
    ref = y.permute(0, 2, 3, 1)
    assert torch.isfinite(out).all()
    assert (out.double().cpu() - ref).abs().max().item() < 1.2e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("prec,dtype", [(0, torch.float32), (1, torch.bfloat16)])
@pytest.mark.parametrize("shape", [(2, 64 * 64, 48, 1), (1, 48 * 40, 96, 2), (1, 4096 + 64, 96, 1), (2, 30 * 30, 384, 8)])
def test_mdta_gram(lib, prec, dtype, shape):
    """q k^T and squared norms over all pixels; bf16 = tcgen05 MN-major kernel, fp32 = CUDA-core kernel."""
    nimg, HW, C, heads = shape
    ch = C // heads
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(nimg, HW, 3 * C, generator=g).to(DEV).to(dtype)
    psz = ch * ch + 2 * ch
    gram = torch.empty(nimg, heads, psz, device=DEV)
    scratch = torch.empty(lib.kdlae_mdta_gram_scratch_floats(nimg, HW, C, heads), device=DEV)
    _lib.check(lib.kdlae_mdta_gram(qkv.data_ptr(), 3 * C, nimg, HW, C, heads, gram.data_ptr(), scratch.data_ptr(), prec, _stream()),
               "mdta_gram")
    torch.cuda.synchronize()
    x = qkv.double().cpu()
    q = x[..., :C].view(nimg, HW, heads, ch).permute(0, 2, 3, 1)          # [n, h, ch, HW]
    k = x[..., C:2 * C].view(nimg, HW, heads, ch).permute(0, 2, 3, 1)
    G = q @ k.transpose(-1, -2)
    got = gram.double().cpu()
    tol = 1e-4 * HW ** 0.5 if prec == 0 else 1e-3 * HW ** 0.5
    assert (got[..., :ch * ch].view(nimg, heads, ch, ch) - G).abs().max().item() < tol
    assert (got[..., ch * ch:ch * ch + ch] - (q * q).sum(-1)).abs().max().item() < tol * 4
    assert (got[..., ch * ch + ch:] - (k * k).sum(-1)).abs().max().item() < tol * 4


@pytest.mark.parametrize("shape", [(1, 40, 72, 48, 144, 0), (2, 16, 33, 96, 288, 0), (1, 24, 64, 48, 256, 1), (1, 36, 36, 96, 512, 1),
                                   (3, 9, 5, 48, 144, 0), (1, 6, 30, 128, 384, 0), (2, 13, 61, 16, 48, 1)])
def test_fused_conv1x1_dwconv3x3_cuda_core(lib, shape):
    """k_pwdw_f2 (tcgen05 1x1 -> packed-FFMA2 depthwise 3x3 -> GELU gate): checked against float64, and bit-for-bit against
    the unfused pair kdlae_conv_gemm + kdlae_dwconv3x3 (same rounding points, same FMA order)."""
    n, H, W, C, Nt, gate = shape
    g = torch.Generator().manual_seed(19)
    x = torch.randn(n, H, W, C, generator=g).to(DEV).bfloat16()
    rstd = (0.5 + torch.rand(n, H, W, generator=g)).to(DEV)
    w1 = (torch.randn(Nt, C, generator=g) / C ** 0.5).to(DEV).bfloat16()
    wd = torch.randn(Nt, 1, 3, 3, generator=g) / 3
    w9c = wd.view(Nt, 9).t().contiguous().to(DEV)
    Co = Nt // 2 if gate else Nt
    out = torch.full((n, H, W, Co), float("nan"), dtype=torch.bfloat16, device=DEV)
    _lib.check(lib.kdlae_pwdw_f2(x.data_ptr(), rstd.data_ptr(), w1.data_ptr(), Nt, w9c.data_ptr(), out.data_ptr(), n, H, W, C, gate,
                                 _stream()), "pwdw_f2")
    torch.cuda.synchronize()
    t = (x.double().cpu() @ w1.double().cpu().t()) * rstd.double().cpu().unsqueeze(-1)
    t = t.float().bfloat16().double()                      # the kernel keeps t as a bf16 smem tile
    y = F.conv2d(t.permute(0, 3, 1, 2), wd.double(), padding=1, groups=Nt)
    if gate:
        y = F.gelu(y[:, :Co]) * y[:, Co:]
    ref = y.permute(0, 2, 3, 1)
    assert torch.isfinite(out).all()
    err = (out.double().cpu() - ref).abs().max().item()
    assert err < 1.5e-2 * max(1.0, ref.abs().max().item())
    # unfused pair
    tt = torch.empty(n, H, W, Nt, dtype=torch.bfloat16, device=DEV)
    out2 = torch.empty_like(out)
    _lib.check(lib.kdlae_conv_gemm(x.data_ptr(), C, w1.data_ptr(), Nt, n, H, W, 1, rstd.data_ptr(), None, 0, None, tt.data_ptr(), 1, 0,
                                   _stream()), "conv_gemm")
    _lib.check(lib.kdlae_dwconv3x3(tt.data_ptr(), out2.data_ptr(), w9c.data_ptr(), n, H, W, Nt, gate, 1, _stream()), "dwconv")
    torch.cuda.synchronize()
    assert torch.equal(out, out2), f"fused vs unfused differ by {(out.float() - out2.float()).abs().max().item():.3e}"


@pytest.mark.parametrize("shape", [(1, 40, 72, 48, 144, 0), (2, 16, 33, 96, 288, 0), (1, 24, 64, 48, 256, 1), (1, 36, 36, 96, 512, 1),
                                   (3, 9, 5, 48, 144, 0), (1, 6, 30, 128, 384, 0), (2, 13, 61, 16, 48, 1), (1, 64, 64, 96, 512, 1)])
def test_fused_conv1x1_dwconv3x3_transposed(lib, shape):
    """k_pwdw_t (W1 . X^T on tcgen05, depthwise inputs read from TMEM in fp32): against float64 with t kept unrounded."""
    n, H, W, C, Nt, gate = shape
    g = torch.Generator().manual_seed(23)
    x = torch.randn(n, H, W, C, generator=g).to(DEV).bfloat16()
    rstd = (0.5 + torch.rand(n, H, W, generator=g)).to(DEV)
    w1 = (torch.randn(Nt, C, generator=g) / C ** 0.5).to(DEV).bfloat16()
    wd = torch.randn(Nt, 1, 3, 3, generator=g) / 3
    w9c = wd.view(Nt, 9).t().contiguous().to(DEV)
    Co = Nt // 2 if gate else Nt
    out = torch.full((n, H, W, Co), float("nan"), dtype=torch.bfloat16, device=DEV)
    _lib.check(lib.kdlae_pwdw_t(x.data_ptr(), rstd.data_ptr(), w1.data_ptr(), Nt, w9c.data_ptr(), out.data_ptr(), n, H, W, C, gate,
                                _stream()), "pwdw_t")
    torch.cuda.synchronize()
    t = (x.double().cpu() @ w1.double().cpu().t()) * rstd.double().cpu().unsqueeze(-1)
    y = F.conv2d(t.permute(0, 3, 1, 2), wd.double(), padding=1, groups=Nt)
    if gate:
        y = F.gelu(y[:, :Co]) * y[:, Co:]
    ref = y.permute(0, 2, 3, 1)
    assert torch.isfinite(out).all()
    err = (out.double().cpu() - ref).abs().max().item()
    print(f"pwdw_t {shape}: max err {err:.3e} (ref max {ref.abs().max().item():.2f})")
    assert err < 6e-3 * max(1.0, ref.abs().max().item())      # one bf16 rounding of the result
