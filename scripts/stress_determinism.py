"""Run-to-run determinism stress of the single fused stages (C ABI): same inputs, perturbed timing (L2 flushes, a competing
stream keeping SMs busy, cold/hot clocks) -> outputs must stay bit-identical.  Finds timing-dependent races."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rethink_acoustic_image_enhancement_b200 import _lib
lib = _lib.load()
DEV = "cuda"
st = lambda: torch.cuda.current_stream().cuda_stream
side = torch.cuda.Stream()
junk = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
ma = torch.randn(4096, 4096, device=DEV, dtype=torch.bfloat16)


def perturb(i):
    if i % 3 == 1:
        junk.fill_(i & 0xFF)                      # flush L2
    if i % 4 >= 2:
        with torch.cuda.stream(side):             # competing work on the SMs
            for _ in range(1 + i % 3):
                torch.mm(ma, ma)
    if i % 7 == 6:
        torch.cuda.synchronize()


def stress(name, fn, out, reps=40):
    fn(); torch.cuda.synchronize()
    base = out.clone()
    bad, mx = 0, 0.0
    for i in range(reps):
        out.fill_(float("nan")) if out.is_floating_point() else None
        perturb(i)
        fn()
        torch.cuda.synchronize()
        same = torch.equal(out.view(torch.int16) if out.dtype == torch.bfloat16 else out.view(torch.int32),
                           base.view(torch.int16) if base.dtype == torch.bfloat16 else base.view(torch.int32))
        if not same:
            bad += 1
            mx = max(mx, float((out.float() - base.float()).abs().nan_to_num(1e9).max()))
    print(f"{name}: {bad}/{reps} mismatching runs, max |d| {mx:.3e}", flush=True)


g = torch.Generator().manual_seed(1)
for (n, H, W, C, Nt, gate) in [(2, 256, 256, 48, 144, 0), (2, 256, 256, 96, 288, 0), (2, 256, 256, 48, 256, 1), (2, 256, 256, 96, 512, 1),
                               (1, 512, 512, 96, 512, 1), (1, 512, 512, 48, 144, 0)]:
    x = torch.randn(n, H, W, C, generator=g).to(DEV).bfloat16()
    rstd = (0.5 + torch.rand(n, H, W, generator=g)).to(DEV)
    w1 = (torch.randn(Nt, C, generator=g) / C ** 0.5).to(DEV).bfloat16()
    w9c = (torch.randn(9, Nt, generator=g) / 3).to(DEV)
    Co = Nt // 2 if gate else Nt
    out = torch.empty(n, H, W, Co, dtype=torch.bfloat16, device=DEV)
    tag = f"{n}x{H}x{W} {C}->{Nt} gate={gate}"
    stress("pwdw_t  " + tag, lambda: _lib.check(lib.kdlae_pwdw_t(x.data_ptr(), rstd.data_ptr(), w1.data_ptr(), Nt, w9c.data_ptr(), out.data_ptr(), n, H, W, C, gate, st()), "t"), out)
    stress("pwdw_f2 " + tag, lambda: _lib.check(lib.kdlae_pwdw_f2(x.data_ptr(), rstd.data_ptr(), w1.data_ptr(), Nt, w9c.data_ptr(), out.data_ptr(), n, H, W, C, gate, st()), "f2"), out)
    tt = torch.randn(n, H, W, Nt, generator=g).to(DEV).bfloat16()
    stress("dwconv  " + tag, lambda: _lib.check(lib.kdlae_dwconv3x3(tt.data_ptr(), out.data_ptr(), w9c.data_ptr(), n, H, W, Nt, gate, 1, st()), "dw"), out)

for (n, H, W, C, N, k) in [(2, 256, 256, 48, 48, 1), (2, 256, 256, 128, 48, 1), (2, 256, 256, 256, 96, 1), (2, 128, 128, 192, 576, 1),
                           (2, 64, 64, 384, 768, 3), (2, 256, 256, 48, 24, 3), (1, 512, 512, 96, 192, 3)]:
    a = torch.randn(n, H, W, C, generator=g).to(DEV).bfloat16()
    w = (torch.randn(N, k * k, C, generator=g) / (C * k * k) ** 0.5).to(DEV).bfloat16()
    rs = (0.5 + torch.rand(n, H, W, generator=g)).to(DEV)
    res = torch.randn(n, H, W, N, generator=g).to(DEV).bfloat16()
    out = torch.empty(n, H, W, N, dtype=torch.bfloat16, device=DEV)
    stress(f"conv_gemm {n}x{H}x{W} {C}->{N} k{k} res", lambda: _lib.check(lib.kdlae_conv_gemm(a.data_ptr(), C, w.data_ptr(), N, n, H, W, k, None, None, 0, res.data_ptr(), out.data_ptr(), 1, 0, st()), "g"), out)
    if k == 1:
        stress(f"conv_gemm {n}x{H}x{W} {C}->{N} k{k} rowscale", lambda: _lib.check(lib.kdlae_conv_gemm(a.data_ptr(), C, w.data_ptr(), N, n, H, W, k, rs.data_ptr(), None, 0, None, out.data_ptr(), 1, 0, st()), "g"), out)

for (nimg, HW, C, heads) in [(2, 65536, 48, 1), (2, 65536, 96, 1), (2, 16384, 96, 2), (2, 4096, 192, 4), (1, 262144, 96, 1)]:
    ch = C // heads
    qkv = torch.randn(nimg, HW, 3 * C, generator=g).to(DEV).bfloat16()
    gram = torch.empty(nimg, heads, ch * ch + 2 * ch, device=DEV)
    scratch = torch.empty(lib.kdlae_mdta_gram_scratch_floats(nimg, HW, C, heads), device=DEV)
    stress(f"mdta_gram {nimg}x{HW} C={C} heads={heads}", lambda: _lib.check(lib.kdlae_mdta_gram(qkv.data_ptr(), 3 * C, nimg, HW, C, heads, gram.data_ptr(), scratch.data_ptr(), 1, st()), "gram"), gram)
