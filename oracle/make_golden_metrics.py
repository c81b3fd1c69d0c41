#!/usr/bin/env python
"""Freeze golden vectors for the validation metric / training loss from the UNMODIFIED reference functions.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden_metrics.py      # writes tests/golden/metrics.npz

Imports calculate_psnr (Train/basicsr/metrics/psnr_ssim.py), tensor2img (utils/img_util.py) and L1LossSr
(models/losses/losses.py) from /root/reference/Train with the two absent third-party modules (lmdb, skimage.metrics - unused
by these functions) stubbed, runs them on seeded inputs, checks oracle/metrics.py against them and stores inputs + outputs.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("KDLAE_REFERENCE_DIR", "/root/reference")
sys.path.insert(0, os.path.join(REF, "Train"))
for name in ("lmdb", "skimage", "skimage.metrics"):
    sys.modules.setdefault(name, types.ModuleType(name))

from basicsr.metrics.psnr_ssim import calculate_psnr  # type: ignore  # noqa: E402
from basicsr.models.losses.losses import L1LossSr  # type: ignore  # noqa: E402
from basicsr.utils.img_util import tensor2img  # type: ignore  # noqa: E402

from oracle import metrics as om, synth  # noqa: E402


def main():
    out = {}
    # ---- PSNR: tensor branch (float, peak 1) and use_image branch (tensor2img uint8, peak 255), with and without crop ----
    for ci, (c, h, w) in enumerate([(1, 40, 56), (3, 33, 47)]):
        gt = synth.seeded_tensor(f"metric.gt.{ci}", (2, c, h, w), 0, "sonar")
        pred = (gt + 0.05 * synth.seeded_tensor(f"metric.noise.{ci}", (2, c, h, w), 1, "normal")).clamp(-0.1, 1.2)
        pred[:, :, :2, :3] = gt[:, :, :2, :3] + torch.tensor([0.5, 1.5, 2.5]) / 255.0      # rounding ties of tensor2img
        out[f"psnr{ci}_pred"], out[f"psnr{ci}_gt"] = pred.numpy(), gt.numpy()
        vals = []
        for b in range(2):
            for crop in (0, 4):
                v_t = calculate_psnr(pred[b:b + 1], gt[b:b + 1], crop)
                # tensor2img clamps IN PLACE when handed a CPU fp32 tensor (.float().detach().cpu() are no-ops there): give it copies
                p8, g8 = tensor2img([pred[b:b + 1].clone()], rgb2bgr=False), tensor2img([gt[b:b + 1].clone()], rgb2bgr=False)
                v_i = calculate_psnr(p8, g8, crop)
                assert abs(v_t - om.calculate_psnr(pred[b:b + 1], gt[b:b + 1], crop)) < 1e-9
                assert np.array_equal(p8, om.tensor2img_u8(pred[b]))
                assert abs(v_i - om.calculate_psnr(om.tensor2img_u8(pred[b]), om.tensor2img_u8(gt[b]), crop)) < 1e-9
                vals.append([b, crop, v_t, v_i])
        out[f"psnr{ci}_vals"] = np.array(vals, dtype=np.float64)
    # ---- L1LossSr: value + autograd gradient, with and without the sr term ----
    crit = L1LossSr(loss_weight=1.0, reduction="mean")
    crit2 = L1LossSr(loss_weight=0.7, reduction="mean")
    hq_gt = synth.seeded_tensor("loss.hq_gt", (2, 1, 24, 32), 0, "sonar")
    sr_gt = synth.seeded_tensor("loss.sr_gt", (2, 1, 48, 64), 0, "sonar")
    hq = (hq_gt + 0.1 * synth.seeded_tensor("loss.hq_n", (2, 1, 24, 32), 2, "normal"))
    sr = (sr_gt + 0.1 * synth.seeded_tensor("loss.sr_n", (2, 1, 48, 64), 3, "normal"))
    hq[0, 0, 0, :5] = hq_gt[0, 0, 0, :5]                                                  # exact zeros of the difference: sign(0) = 0
    out.update(loss_hq=hq.numpy(), loss_hq_gt=hq_gt.numpy(), loss_sr=sr.numpy(), loss_sr_gt=sr_gt.numpy())
    for tag, c, with_sr in (("a", crit, True), ("b", crit2, True), ("c", crit, False)):
        p_hq, p_sr = hq.clone().requires_grad_(True), sr.clone().requires_grad_(True)
        pred = {"hq": p_hq, "sr": p_sr if with_sr else None}
        loss = c(pred, {"hq": hq_gt, "sr": sr_gt})
        loss.backward()
        o_hq, o_sr = hq.clone().requires_grad_(True), sr.clone().requires_grad_(True)
        ol = om.l1_loss_sr({"hq": o_hq, "sr": o_sr if with_sr else None}, {"hq": hq_gt, "sr": sr_gt}, c.loss_weight)
        ol.backward()
        assert abs(float(loss.detach()) - float(ol.detach())) < 1e-7 and torch.equal(p_hq.grad, o_hq.grad)
        out[f"loss_{tag}"] = np.array([float(loss.detach()), c.loss_weight, float(with_sr)], dtype=np.float64)
        out[f"loss_{tag}_ghq"] = p_hq.grad.numpy()
        if with_sr:
            assert torch.equal(p_sr.grad, o_sr.grad)
            out[f"loss_{tag}_gsr"] = p_sr.grad.numpy()
    path = os.path.join(ROOT, "tests", "golden", "metrics.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
