"""1x1 tcgen05 GEMM stages of KDLAE-T level 1/2 at a given KDLAE_SM_LIMIT: is a stage HBM-bound (time flat when SMs are
removed) or SM-bound (time grows as 1/SMs)?"""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rethink_acoustic_image_enhancement_b200 import _lib
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
res = {}
for name, S, C, N, has_res in [("L1 qkv 48->144", 512, 48, 144, 0), ("L1 project_in 48->256", 512, 48, 256, 0),
                               ("L1 project_out 128->48 +res", 512, 128, 48, 1), ("L2 qkv 96->288", 256, 96, 288, 0),
                               ("L2 project_out 256->96 +res", 256, 256, 96, 1), ("L1 attn apply 96->96 +res", 512, 96, 96, 1),
                               ("L1 project_out 256->96 +res", 512, 256, 96, 1)]:
    n = 8
    x = torch.randn(n, S, S, C, device="cuda").bfloat16()
    w = (torch.randn(N, C, device="cuda") / C ** 0.5).bfloat16()
    rs = torch.rand(n * S * S, device="cuda") + 0.5
    r = torch.randn(n, S, S, N, device="cuda").bfloat16() if has_res else None
    out = torch.empty(n, S, S, N, dtype=torch.bfloat16, device="cuda")
    f = lambda: _lib.check(lib.kdlae_conv_gemm(x.data_ptr(), C, w.data_ptr(), N, n, S, S, 1, None if has_res else rs.data_ptr(), None, 0,
                                               r.data_ptr() if has_res else None, out.data_ptr(), 1, 0, st), "gemm")
    for _ in range(3): f()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize(); a.record()
    for _ in range(10): f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    nbytes = n * S * S * (C + N * (2 if has_res else 1)) * 2
    res[name] = dict(us=round(ms * 1e3, 1), GBs=round(nbytes / ms / 1e6))
print(json.dumps({"sm_limit": os.environ.get("KDLAE_SM_LIMIT", "all"), "tc_stages": os.environ.get("KDLAE_TC_STAGES", "auto"), **res}))
