"""Bring-up probe of the TF32 wgrad (gemm_tf32.cu k_wgrad_tf32) through conv_train on 1x1 convs: dw vs float64."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.nn.functional as F
from rethink_acoustic_image_enhancement_b200 import training
training.set_matmul_precision("tf32")
for (B, Cin, H, W, Cout) in [(1, 32, 8, 8, 32), (2, 48, 40, 36, 144), (1, 96, 33, 17, 96), (1, 384, 16, 16, 1024), (3, 512, 9, 7, 96)]:
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, Cin, H, W, generator=g); w = torch.randn(Cout, Cin, 1, 1, generator=g) / Cin ** 0.5
    dout = torch.randn(B, Cout, H, W, generator=g)
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    F.conv2d(xr, wr).backward(dout.double())
    xc, wc = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    training.conv_train(xc, wc).backward(dout.cuda()); torch.cuda.synchronize()
    a, b = wc.grad.double().cpu().view(Cout, Cin), wr.grad.view(Cout, Cin)
    print((B, Cin, H, W, Cout), "rel", float((a - b).abs().max() / b.abs().max()), "got", [round(v, 3) for v in a[0, :4].tolist()],
          "ref", [round(v, 3) for v in b[0, :4].tolist()], flush=True)
