"""Fused 1x1 -> depthwise 3x3 (pwdw_f2) against the unfused pair (conv_gemm + dwconv3x3) on KDLAE-T shapes, batch 8."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rethink_acoustic_image_enhancement_b200 import _lib
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
SHAPES = [(8, 512, 48, 144, 0), (8, 512, 48, 256, 1), (8, 512, 96, 288, 0), (8, 512, 96, 512, 1), (8, 256, 96, 288, 0), (8, 256, 96, 512, 1)]
only = os.environ.get("PW_ONLY")
if os.environ.get("PW_SHAPES"):
    SHAPES = [SHAPES[int(i)] for i in os.environ["PW_SHAPES"].split(",")]
res = {}
for (n, S, C, Nt, gate) in SHAPES:
    x = torch.randn(n, S, S, C, device="cuda").bfloat16()
    rstd = torch.rand(n, S, S, device="cuda") + 0.5
    w1 = (torch.randn(Nt, C, device="cuda") / C ** 0.5).bfloat16()
    w9c = (torch.randn(9, Nt, device="cuda") / 3).contiguous()
    Co = Nt // 2 if gate else Nt
    out = torch.empty(n, S, S, Co, dtype=torch.bfloat16, device="cuda")
    out2 = torch.empty_like(out)
    out3 = torch.empty_like(out)
    tt = torch.empty(n, S, S, Nt, dtype=torch.bfloat16, device="cuda")
    def fused():
        _lib.check(lib.kdlae_pwdw_f2(x.data_ptr(), rstd.data_ptr(), w1.data_ptr(), Nt, w9c.data_ptr(), out.data_ptr(), n, S, S, C, gate, st), "f")
    def transposed():
        _lib.check(lib.kdlae_pwdw_t(x.data_ptr(), rstd.data_ptr(), w1.data_ptr(), Nt, w9c.data_ptr(), out3.data_ptr(), n, S, S, C, gate, st), "t")
    def unfused():
        _lib.check(lib.kdlae_conv_gemm(x.data_ptr(), C, w1.data_ptr(), Nt, n, S, S, 1, rstd.data_ptr(), None, 0, None, tt.data_ptr(), 1, 0, st), "g")
        _lib.check(lib.kdlae_dwconv3x3(tt.data_ptr(), out2.data_ptr(), w9c.data_ptr(), n, S, S, Nt, gate, 1, st), "d")
    row = {}
    for name, f in (("fused", fused), ("transposed", transposed), ("unfused", unfused)):
        if only and name != only: continue
        for _ in range(3): f()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda.synchronize(); a.record()
        for _ in range(5): f()
        b.record(); torch.cuda.synchronize()
        row[name + "_us"] = round(a.elapsed_time(b) / 5 * 1e3, 1)
    if not only:
        row["identical"] = bool(torch.equal(out, out2))
        row["t_vs_f2_maxdiff"] = (out3.float() - out.float()).abs().max().item()
    res[str((n, S, C, Nt, gate))] = row
    print((n, S, C, Nt, gate), row, flush=True)
