"""conv3x3_tc on ASDQE / KDLAE shapes: us, clocks per 120-pixel tile, TFLOP/s (see profiles/r01_summary.md for the bring-up
bisection that used temporary KDLAE_C3_DEBUG switches: skip stores / epilogue math / MMAs / roles)."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rethink_acoustic_image_enhancement_b200 import _lib
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
res = {}
SH = [(32, 512, 64, 64), (32, 512, 16, 16), (32, 256, 128, 128), (16, 512, 96, 192)]
if os.environ.get("C3_SHAPES"):
    SH = [SH[int(i)] for i in os.environ["C3_SHAPES"].split(",")]
for (n, S, C, N) in SH:
    x = torch.randn(n, S, S, C, device="cuda").bfloat16()
    w = (torch.randn(N, 9, C, device="cuda") / (9 * C) ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    out = torch.empty(n, S, S, N, dtype=torch.bfloat16, device="cuda")
    f = lambda: _lib.check(lib.kdlae_conv_gemm(x.data_ptr(), C, w.data_ptr(), N, n, S, S, 3, None, bias.data_ptr(), 1, None, out.data_ptr(), 1, 0, st), "c3")
    for _ in range(3): f()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize(); a.record()
    for _ in range(5): f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    tiles = n * ((S + 29) // 30) * ((S + 3) // 4)
    res[f"{n}x{S}^2 {C}->{N}"] = dict(us=round(ms * 1e3, 1), clk_per_tile=round(ms * 1e-3 * 1.9e9 / (tiles / 148)), tflops=round(2 * n * S * S * 9 * C * N / ms / 1e9, 1))
print(json.dumps({"debug": os.environ.get("KDLAE_C3_DEBUG", "0"), **res}))
