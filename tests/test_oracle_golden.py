"""CPU: the oracle restatement reproduces every golden fixture frozen from the unmodified reference."""
import pytest
import torch

import oracle
from oracle import synth
from conftest import load_golden

TOL = 2e-5  # fp32 noise between two orderings of the same arithmetic (reference fp32-vs-fp64 is 2.4e-7)


@pytest.mark.parametrize("name", ["teacher_c1_biasfree_64", "teacher_c3_withbias_32x48", "teacher_c1_nosr_40x24",
                                  "teacher_c3_biasfree_72x88"])
def test_teacher_oracle_matches_reference_fixture(name, manifest):
    case, g = manifest[name], load_golden(name)
    kw = case["kwargs"]
    sd = synth.teacher_state_dict(seed=case["seed"], temp_scale=case["temp_scale"], **kw)
    assert len(sd) == case["n_keys"]
    b, h, w = case["shape"]
    rate = g["rate"].view(b, 1, 1, 1).expand(b, 1, h, w)
    with torch.no_grad():
        hq, sr = oracle.teacher_forward(sd, g["img"], rate, static=kw["static"])
    assert (hq - g["hq"]).abs().max().item() < TOL
    if "sr" in g:
        assert (sr - g["sr"]).abs().max().item() < TOL
    else:
        assert sr is None
    # the network must actually change the image, otherwise parity against it proves nothing
    assert (g["hq"] - g["img"]).abs().mean().item() > 1e-3


@pytest.mark.parametrize("name", ["student_f5_32x40", "student_f7_16x16", "student_f1_nores_8x12"])
def test_student_oracle_matches_reference_fixture(name, manifest):
    case, g = manifest[name], load_golden(name)
    sd = synth.student_state_dict(seed=case["seed"])
    with torch.no_grad():
        y = oracle.student_forward(sd, g["x"], residual=case["residual"])
    assert (y - g["y"]).abs().max().item() < 1e-5


@pytest.mark.parametrize("name", ["asdqe_48x40", "asdqe_32x32"])
def test_asdqe_oracle_matches_reference_fixture(name, manifest):
    case, g = manifest[name], load_golden(name)
    sd = synth.asdqe_state_dict(seed=case["seed"])
    with torch.no_grad():
        s = oracle.asdqe_forward(sd, g["lq"], g["gt"])
        f = oracle.asdqe_trunk(sd, g["lq"], g["gt"])
    assert (s - g["score"]).abs().max().item() < 1e-5
    assert (f - g["feat"]).abs().max().item() < 1e-4


def test_asdqe_fixture_scores_are_not_degenerate(manifest):
    s = manifest["asdqe_48x40"]["scores"] + manifest["asdqe_32x32"]["scores"]
    assert max(s) - min(s) > 0.05


def test_seeded_tensors_are_order_independent():
    a = synth.seeded_tensor("k", (4, 5), 3)
    torch.manual_seed(123)
    torch.rand(10)
    b = synth.seeded_tensor("k", (4, 5), 3)
    assert torch.equal(a, b)
    assert not torch.equal(a, synth.seeded_tensor("k2", (4, 5), 3))


def test_psnr_definition():
    a = torch.zeros(10, 10)
    b = torch.full((10, 10), 0.1)
    assert abs(synth.psnr(a, b) - 20.0) < 1e-4  # 20*log10(1/0.1)
