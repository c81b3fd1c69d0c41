"""Validation metric (N3) and training loss (N1): the oracle against the golden vectors frozen from the unmodified reference
functions (CPU), and the CUDA kernels behind the C ABI against both (GPU)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import metrics as om
from conftest import GOLDEN

DEV = "cuda:0"


@pytest.fixture(scope="module")
def gold():
    z = np.load(os.path.join(GOLDEN, "metrics.npz"))
    return {k: z[k] for k in z.files}


def test_oracle_psnr_matches_reference_golden(gold):
    for ci in (0, 1):
        pred, gt = torch.from_numpy(gold[f"psnr{ci}_pred"]), torch.from_numpy(gold[f"psnr{ci}_gt"])
        for b, crop, v_t, v_i in gold[f"psnr{ci}_vals"]:
            b, crop = int(b), int(crop)
            assert abs(om.calculate_psnr(pred[b:b + 1], gt[b:b + 1], crop) - v_t) < 1e-9
            assert abs(om.calculate_psnr(om.tensor2img_u8(pred[b]), om.tensor2img_u8(gt[b]), crop) - v_i) < 1e-9


def test_oracle_l1_loss_sr_matches_reference_golden(gold):
    hq, hq_gt = torch.from_numpy(gold["loss_hq"]), torch.from_numpy(gold["loss_hq_gt"])
    sr, sr_gt = torch.from_numpy(gold["loss_sr"]), torch.from_numpy(gold["loss_sr_gt"])
    for tag in "abc":
        val, lw, with_sr = gold[f"loss_{tag}"]
        p_hq, p_sr = hq.clone().requires_grad_(True), sr.clone().requires_grad_(True)
        loss = om.l1_loss_sr({"hq": p_hq, "sr": p_sr if with_sr else None}, {"hq": hq_gt, "sr": sr_gt}, float(lw))
        loss.backward()
        assert abs(float(loss.detach()) - val) < 1e-7
        assert np.array_equal(p_hq.grad.numpy(), gold[f"loss_{tag}_ghq"])
        if with_sr:
            assert np.array_equal(p_sr.grad.numpy(), gold[f"loss_{tag}_gsr"])


@pytest.mark.gpu
def test_device_psnr_matches_reference_golden(gold):
    from rethink_acoustic_image_enhancement_b200 import metrics as pm
    for ci in (0, 1):
        pred, gt = torch.from_numpy(gold[f"psnr{ci}_pred"]).to(DEV), torch.from_numpy(gold[f"psnr{ci}_gt"]).to(DEV)
        vals = gold[f"psnr{ci}_vals"]
        for crop in (0, 4):
            p_t = pm.psnr_batch(pred, gt, crop).cpu().numpy()
            p_i = pm.psnr_batch(pred, gt, crop, as_uint8=True).cpu().numpy()
            for b, c, v_t, v_i in vals:
                if int(c) != crop:
                    continue
                assert abs(p_t[int(b)] - v_t) < 1e-4, (ci, b, crop, p_t[int(b)], v_t)    # fp32 differences, double sums
                assert abs(p_i[int(b)] - v_i) < 1e-9, (ci, b, crop, p_i[int(b)], v_i)    # integer data: exact
        assert abs(pm.calculate_psnr(pred, gt, 4) - vals[1][2]) < 1e-4                   # 4-D tensor -> first image (:40-43)
    same = pm.psnr_batch(gt, gt)
    assert torch.isinf(same).all()
    with pytest.raises(RuntimeError, match="CUDA"):
        pm.psnr_batch(gt.cpu(), gt.cpu())


@pytest.mark.gpu
def test_device_l1_loss_sr_value_and_gradient(gold):
    from rethink_acoustic_image_enhancement_b200 import metrics as pm
    hq, hq_gt = torch.from_numpy(gold["loss_hq"]).to(DEV), torch.from_numpy(gold["loss_hq_gt"]).to(DEV)
    sr, sr_gt = torch.from_numpy(gold["loss_sr"]).to(DEV), torch.from_numpy(gold["loss_sr_gt"]).to(DEV)
    for tag in "abc":
        val, lw, with_sr = gold[f"loss_{tag}"]
        crit = pm.L1LossSr(loss_weight=float(lw))
        p_hq, p_sr = hq.clone().requires_grad_(True), sr.clone().requires_grad_(True)
        loss = crit({"hq": p_hq, "sr": p_sr if with_sr else None}, {"hq": hq_gt, "sr": sr_gt})
        (2.0 * loss).backward()                         # upstream gradient != 1 goes through backward()
        assert abs(loss.item() - val) < 1e-6
        assert torch.allclose(p_hq.grad.cpu(), 2.0 * torch.from_numpy(gold[f"loss_{tag}_ghq"]), rtol=1e-6, atol=0)
        assert (p_hq.grad[0, 0, 0, :5] == 0).all()      # sign(0) = 0, as torch
        if with_sr:
            assert torch.allclose(p_sr.grad.cpu(), 2.0 * torch.from_numpy(gold[f"loss_{tag}_gsr"]), rtol=1e-6, atol=0)
        else:
            assert p_sr.grad is None


@pytest.mark.gpu
def test_validation_loop_psnr_on_device():
    """validate(): the metric half of nondist_validation for the dict-input teacher, against the oracle forward + oracle PSNR."""
    import oracle
    from oracle import synth
    import rethink_acoustic_image_enhancement_b200 as pk
    from rethink_acoustic_image_enhancement_b200 import metrics as pm
    kw = dict(inp_channels=1, out_channels=1, LayerNorm_type="BiasFree", static="train")
    sd = synth.teacher_state_dict(seed=2, temp_scale=3.0, **kw)
    m = pk.KDLAE_teacher(**kw)
    m.load_state_dict(sd)
    m = m.to(DEV).eval().set_precision("fp32")
    batches, ref = [], []
    for i in range(2):
        img = synth.seeded_tensor(f"val.img.{i}", (2, 1, 32, 40), 2, "sonar")
        gt = (img * 0.9).clamp(0, 1)
        rate = torch.full((2, 1, 1, 1), 0.6)
        batches.append({"lq": {"img": img.to(DEV), "denoise_rate": rate.to(DEV)}, "gt": {"hq": gt.to(DEV)}})
        with torch.no_grad():
            hq_ref, _ = oracle.teacher_forward(sd, img, rate.expand(2, 1, 32, 40))
        for b in range(2):
            ref.append(om.calculate_psnr(om.tensor2img_u8(hq_ref[b]), om.tensor2img_u8(gt[b]), 2))
    res = pm.validate(m, batches, crop_border=2, use_image=True)
    assert res["count"] == 4
    # uint8 quantisation can flip a pixel by one level where the fp32 paths differ by 1e-6: tolerance on the mean PSNR
    assert abs(res["psnr"] - float(np.mean(ref))) < 0.02, (res, ref)
    assert not m.training


@pytest.mark.gpu
def test_pad_test_hook_matches_oracle_on_a_ragged_image():
    """pad_test (image_restoration_model.py:226-237): a 36 x 44 image is reflect-padded to 40 x 48, run, and cropped back -
    against the same recipe around the oracle forward (fp32 path), through validate(window_size=8) as well."""
    import torch.nn.functional as F
    import oracle
    from oracle import synth
    import rethink_acoustic_image_enhancement_b200 as pk
    from rethink_acoustic_image_enhancement_b200 import metrics as pm
    kw = dict(inp_channels=1, out_channels=1, LayerNorm_type="BiasFree", static="train")
    sd = synth.teacher_state_dict(seed=6, temp_scale=2.0, **kw)
    m = pk.KDLAE_teacher(**kw)
    m.load_state_dict(sd)
    m = m.to(DEV).eval().set_precision("fp32")
    img = synth.seeded_tensor("padtest.img", (1, 1, 36, 44), 6, "sonar")
    rate = torch.full((1, 1, 36, 44), 0.4)
    with torch.no_grad():
        got = pm.pad_test(m, {"img": img.to(DEV), "denoise_rate": rate.to(DEV)}, 8)
        pi, pr = F.pad(img, (0, 4, 0, 4), "reflect"), F.pad(rate, (0, 4, 0, 4), "reflect")
        hq_ref, sr_ref = oracle.teacher_forward(sd, pi, pr)
    assert got["hq"].shape == (1, 1, 36, 44) and got["sr"].shape == (1, 1, 72, 88)
    assert float((got["hq"].cpu() - hq_ref[:, :, :36, :44]).abs().max()) < 1e-4
    assert float((got["sr"].cpu() - sr_ref[:, :, :72, :88]).abs().max()) < 1e-4
    gt = (img * 0.8).clamp(0, 1)
    res = pm.validate(m, [{"lq": {"img": img.to(DEV), "denoise_rate": rate.to(DEV)}, "gt": {"hq": gt.to(DEV)}}], window_size=8)
    ref = om.calculate_psnr(om.tensor2img_u8(hq_ref[0, :, :36, :44]), om.tensor2img_u8(gt[0]), 0)
    assert res["count"] == 1 and abs(res["psnr"] - ref) < 0.02
