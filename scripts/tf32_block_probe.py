import sys, torch
sys.path.insert(0, '/root/repo')
from oracle import functional as ofn, synth
from rethink_acoustic_image_enhancement_b200 import training
DEV='cuda'
for mode in ('fp32','tf32'):
    training.set_matmul_precision(mode)
    for shape, heads, ts in [((2,48,64,64),1,4.0), ((1,96,24,40),2,4.0), ((1,96,24,40),2,1.0), ((1,128,8,12),8,4.0)]:
        B,C,H,W = shape
        sd = {}
        synth._block(sd, "blk", C, 2.66, False, False, seed=5, temp_scale=ts, heads=heads)
        x = synth.seeded_tensor("train.xb", shape, 5, "normal"); dout = synth.seeded_tensor("train.doutb", shape, 6, "normal")
        ref_p = {k: v.double().requires_grad_(True) for k, v in sd.items()}
        xr = x.double().requires_grad_(True)
        p1 = {("stage.0" + k[3:]): v for k, v in ref_p.items()}
        out_ref = ofn._blocks(xr, p1, "stage", 1, heads); out_ref.backward(dout.double())
        cu_p = {k: v.to(DEV).requires_grad_(True) for k, v in sd.items()}
        xc = x.to(DEV).requires_grad_(True)
        out = training.transformer_block_train(xc, cu_p, "blk"); out.backward(dout.to(DEV)); torch.cuda.synchronize()
        rel = lambda a, b: float((a.double().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-30))
        errs = {"out": rel(out.detach(), out_ref.detach()), "dx": rel(xc.grad, xr.grad)}
        for k in sd: errs[k[4:]] = rel(cu_p[k].grad, ref_p[k].grad)
        print(mode, shape, heads, ts, {k: f"{v:.1e}" for k, v in errs.items()})
