"""TF32 tensor-core GEMMs of the training step (gemm_tf32.cu) on KDLAE-T shapes through conv_train: forward, dgrad + wgrad, in
fp32 (CUDA cores) and TF32 (tcgen05) mode; ms and TFLOP/s per pass.  usage: tf32_gemm_probe.py [fp32|tf32|both]"""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rethink_acoustic_image_enhancement_b200 import training
modes = ["fp32", "tf32"] if len(sys.argv) < 2 or sys.argv[1] == "both" else [sys.argv[1]]
SHAPES = [(2, 96, 256, 256, 512, 1), (2, 256, 256, 256, 96, 1), (2, 96, 256, 256, 288, 1), (2, 96, 256, 256, 192, 3), (2, 384, 32, 32, 1152, 1)]
for mode in modes:
    training.set_matmul_precision(mode)
    for (B, Cin, H, W, Cout, k) in SHAPES:
        x = torch.randn(B, Cin, H, W, device="cuda", requires_grad=True)
        w = (torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5).requires_grad_(True)
        dout = torch.randn(B, Cout, H, W, device="cuda")
        def fwd():
            return training.conv_train(x, w)
        def fwd_bwd():
            x.grad = None; w.grad = None
            training.conv_train(x, w).backward(dout)
        res = {}
        for name, f in (("fwd", fwd), ("fwd_bwd", fwd_bwd)):
            for _ in range(2): f()
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            torch.cuda.synchronize(); a.record()
            for _ in range(5): f()
            b.record(); torch.cuda.synchronize()
            res[name] = a.elapsed_time(b) / 5
        flops = 2.0 * B * H * W * Cin * Cout * k * k
        print(json.dumps({"mode": mode, "shape": [B, Cin, H, W, Cout, k], "fwd_ms": round(res["fwd"], 3), "bwd_ms": round(res["fwd_bwd"] - res["fwd"], 3),
                          "fwd_tflops": round(flops / res["fwd"] / 1e9, 1), "bwd_tflops": round(2 * flops / (res["fwd_bwd"] - res["fwd"]) / 1e9, 1)}), flush=True)
