"""Stand-alone depthwise 3x3 kernel (dwconv_f2.cu, packed FFMA2) on KDLAE-T shapes (batch 8): algorithmic GB/s."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rethink_acoustic_image_enhancement_b200 import _lib

lib = _lib.load()
SHAPES = [(8, 512, 512, 144, 0), (8, 512, 512, 256, 1), (8, 256, 256, 288, 0), (8, 256, 256, 512, 1),
          (8, 128, 128, 576, 0), (8, 128, 128, 1024, 1), (8, 64, 64, 1152, 0), (8, 64, 64, 2048, 1)]
st = torch.cuda.current_stream().cuda_stream
res = {}
for (n, H, W, C, gate) in SHAPES:
    x = torch.randn(n, H, W, C, device="cuda").bfloat16()
    w9c = (torch.randn(9, C, device="cuda") / 3).contiguous()
    Co = C // 2 if gate else C
    out = torch.empty(n, H, W, Co, dtype=torch.bfloat16, device="cuda")
    nbytes = n * H * W * (C + Co) * 2
    f = lambda: _lib.check(lib.kdlae_dwconv3x3(x.data_ptr(), out.data_ptr(), w9c.data_ptr(), n, H, W, C, gate, 1, st), "dw")
    for _ in range(3): f()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize(); a.record()
    for _ in range(10): f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    res[str((n, H, W, C, gate))] = dict(us=round(ms * 1e3, 1), GBs=round(nbytes / ms / 1e6, 0))
print(json.dumps(res, indent=1))
