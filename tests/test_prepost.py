"""uint8 pre / post-processing: oracle restatement on CPU against the outputs of the reference's own notebook
cell (tests/golden/prepost.npz, frozen by oracle/make_golden_prepost.py), CUDA passes bit-exact against both on the GPU."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import prepost as op

DEV = "cuda"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prepost.npz")


def _golden_cases():
    z = np.load(GOLDEN)
    for ci in range(int(z["n_cases"])):
        yield ci, {k[len(f"c{ci}_"):]: z[k] for k in z.files if k.startswith(f"c{ci}_")}


def test_oracle_prepost_reproduces_the_reference_cell():
    """oracle/prepost.py against what KDLAE_T.ipynb cell 5 itself produced (padded input, rate map shape, uint8 hq / sr)."""
    n = 0
    for ci, c in _golden_cases():
        img, rate = c["img"], float(c["rate"])
        x, alpha = op.preprocess_u8(img[None], rate)
        assert np.array_equal(x.numpy(), c["x"]), f"case {ci}: padded input"
        assert tuple(alpha.shape) == tuple(c["alpha_shape"]) and torch.all(alpha == np.float32(rate))
        assert np.array_equal(op.postprocess_u8(torch.from_numpy(c["pred_hq"]), img[None], 1)[0], c["hq_u8"]), f"case {ci}: hq"
        assert np.array_equal(op.postprocess_u8(torch.from_numpy(c["pred_sr"]), img[None], 2)[0], c["sr_u8"]), f"case {ci}: sr"
        n += 1
    assert n == 4


@pytest.mark.gpu
def test_prepost_kernels_reproduce_the_reference_cell():
    """The CUDA passes against the reference cell's frozen outputs, bit for bit."""
    import rethink_acoustic_image_enhancement_b200 as pk
    for ci, c in _golden_cases():
        img, rate = c["img"], float(c["rate"])
        img_d = torch.from_numpy(img[None].copy()).to(DEV)
        x, a = pk.preprocess_u8(img_d, torch.tensor([rate]), rate_map=True)
        assert np.array_equal(x.cpu().numpy(), c["x"]), f"case {ci}: padded input"
        assert tuple(a.shape) == tuple(c["alpha_shape"]) and torch.all(a == np.float32(rate))
        hq = pk.postprocess_u8(torch.from_numpy(c["pred_hq"]).to(DEV), img_d, 1)
        sr = pk.postprocess_u8(torch.from_numpy(c["pred_sr"]).to(DEV), img_d, 2)
        assert np.array_equal(hq.cpu().numpy()[0], c["hq_u8"]), f"case {ci}: hq"
        assert np.array_equal(sr.cpu().numpy()[0], c["sr_u8"]), f"case {ci}: sr"


def _sonar_u8(B, h, w, c, seed):
    g = np.random.default_rng(seed)
    img = g.integers(0, 256, size=(B, h, w, c), dtype=np.uint8)
    img[g.random((B, h, w)) < 0.45] = 0          # blind zone: 45 % exact zeros, as in Sample/MDD
    return img


@pytest.mark.parametrize("shape", [(2, 37, 50, 3), (1, 64, 64, 1), (1, 9, 15, 1)])
def test_oracle_prepost_follows_the_notebook_ops(shape):
    """The restatement against the literal sequence of cell 5 written out with plain torch / numpy calls."""
    B, h, w, c = shape
    img = _sonar_u8(B, h, w, c, 0)
    x, alpha = op.preprocess_u8(img, 0.6)
    H, W = ((h + 8) // 8) * 8, ((w + 8) // 8) * 8
    padh, padw = (H - h if h % 8 else 0), (W - w if w % 8 else 0)
    ref = F.pad(torch.from_numpy(img.astype(np.float32) / 255.0).permute(0, 3, 1, 2), (0, padw, 0, padh), "reflect")
    assert x.shape == ref.shape and torch.equal(x, ref)
    assert alpha.shape == (B, 1, h + padh, w + padw) and torch.all(alpha == 0.6)
    pred = torch.rand(B, c, h + padh, w + padw, generator=torch.Generator().manual_seed(1)) * 1.4 - 0.2
    out = op.postprocess_u8(pred, img, 1)
    r = torch.clamp(pred, 0, 1)[:, :, :h, :w].permute(0, 2, 3, 1).numpy()
    exp = np.clip(np.rint(r * np.float32(255.0)), 0, 255).astype(np.uint8)
    exp[np.all(img == 0, axis=-1)] = 0
    assert np.array_equal(out, exp)
    assert out[np.all(img == 0, axis=-1)].max(initial=0) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 37, 50, 3), (1, 64, 64, 1), (3, 9, 15, 1), (1, 120, 33, 4)])
def test_prepost_kernels_bit_exact(shape):
    import rethink_acoustic_image_enhancement_b200 as pk
    B, h, w, c = shape
    img = _sonar_u8(B, h, w, c, 3)
    rates = np.linspace(0.1, 0.9, B).astype(np.float32)
    x_ref, a_ref = op.preprocess_u8(img, rates)
    img_d = torch.from_numpy(img).to(DEV)
    x, a = pk.preprocess_u8(img_d, torch.from_numpy(rates), rate_map=True)
    assert torch.equal(x.cpu(), x_ref) and torch.equal(a.cpu(), a_ref)
    x2, r2 = pk.preprocess_u8(img_d, torch.from_numpy(rates))        # default: per-image rates, no H x W map
    assert torch.equal(x2, x) and r2.shape == (B, 1, 1, 1) and torch.equal(r2.cpu().view(-1), torch.from_numpy(rates))
    for scale in (1, 2):
        g = torch.Generator().manual_seed(5 + scale)
        pred = torch.rand(B, c, x.shape[2] * scale, x.shape[3] * scale, generator=g) * 1.4 - 0.2
        pred[0, 0, 0, :4] = torch.tensor([0.5 / 255, 1.5 / 255, 2.5 / 255, 254.5 / 255])      # rint ties: half to even
        out = pk.postprocess_u8(pred.to(DEV), img_d, scale)
        assert np.array_equal(out.cpu().numpy(), op.postprocess_u8(pred, img, scale))


@pytest.mark.gpu
def test_teacher_infer_uint8_pipeline():
    """uint8 -> uint8 through the CUDA passes + fused forward vs the oracle pipeline around the oracle forward (fp32 path:
    identical up to rounding ties, so at most 1 LSB on a handful of pixels)."""
    import rethink_acoustic_image_enhancement_b200 as pk
    from oracle import functional as ofn, synth
    kw = dict(inp_channels=1, out_channels=1, LayerNorm_type="BiasFree", static="train")
    sd = synth.teacher_state_dict(seed=4, **kw)
    m = pk.KDLAE_teacher(**kw)
    m.load_state_dict(sd)
    m = m.to(DEV).eval().set_precision("fp32")
    img = _sonar_u8(2, 45, 52, 1, 9)
    hq, sr = pk.teacher_infer_uint8(m, torch.from_numpy(img).to(DEV), 0.6)
    x, alpha = op.preprocess_u8(img, 0.6)
    with torch.no_grad():
        ref_hq, ref_sr = ofn.teacher_forward(sd, x, alpha, heads=[1, 2, 4, 8], static="train", params="cat")
    hq_ref, sr_ref = op.postprocess_u8(ref_hq, img, 1), op.postprocess_u8(ref_sr, img, 2)
    assert hq.shape == (2, 45, 52, 1) and sr.shape == (2, 90, 104, 1)
    for got, exp in ((hq, hq_ref), (sr, sr_ref)):
        d = np.abs(got.cpu().numpy().astype(np.int16) - exp.astype(np.int16))
        assert d.max() <= 1 and (d > 0).mean() < 1e-3
    assert hq.cpu().numpy()[np.all(img == 0, axis=-1)].max(initial=0) == 0
