#!/usr/bin/env python
"""Pin the oracle against the unmodified reference and freeze golden fixtures.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py            # writes tests/golden/*.npz + MANIFEST.json

For every case it (1) builds the reference nn.Module from /root/reference,
(2) loads the key-seeded synthetic state_dict with strict=True (which proves the
key/shape layout of oracle.synth equals the reference's), (3) runs the reference
forward and the oracle restatement on the same inputs, records their max-abs
difference, and (4) stores inputs + reference outputs as small fp32 fixtures.
The GPU box has no /root/reference; tests there use these fixtures and the oracle.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("KDLAE_REFERENCE_DIR", "/root/reference")

import oracle  # noqa: E402
from oracle import synth  # noqa: E402

TEACHER_CASES = {
    # name: (ctor kwargs, B, H, W, input kind, temp_scale, seed)
    "teacher_c1_biasfree_64": (dict(inp_channels=1, out_channels=1, LayerNorm_type="BiasFree", static="train"), 1, 64, 64, "uniform01", 1.0, 0),
    "teacher_c3_withbias_32x48": (dict(inp_channels=3, out_channels=3, LayerNorm_type="WithBias", static="train"), 2, 32, 48, "sonar", 8.0, 1),
    "teacher_c1_nosr_40x24": (dict(inp_channels=1, out_channels=1, LayerNorm_type="BiasFree", static="no"), 1, 40, 24, "sonar", 6.0, 2),
    # the pretrained-checkpoint layout (KDLAET.yml / KDLAE_T.ipynb: 3 channels, BiasFree): the fused 1x1 -> depthwise kernels
    # with 3-plane input / output convs, at a size with ragged 30x6 / 14x6 tiles
    "teacher_c3_biasfree_72x88": (dict(inp_channels=3, out_channels=3, LayerNorm_type="BiasFree", static="train"), 1, 72, 88, "sonar", 2.0, 3),
}
STUDENT_CASES = {
    "student_f5_32x40": (2, 5, 32, 40, True, 0),
    "student_f7_16x16": (1, 7, 16, 16, True, 1),
    "student_f1_nores_8x12": (1, 1, 8, 12, False, 2),
}
ASDQE_CASES = {
    "asdqe_48x40": (2, 48, 40, 0),   # pads to 48x48
    "asdqe_32x32": (1, 32, 32, 1),
}


def main() -> None:
    sys.path.insert(0, os.path.join(REF, "KDLAE"))
    sys.path.insert(0, os.path.join(REF, "ASDQE"))
    from KDLAE_model import KDLAE_teacher, KDLAE_student  # type: ignore
    from ASDQE_model import DenoiseRatePredictor  # type: ignore

    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.manual_seed(0)
    manifest = {"reference": "yangtaihong59/Rethink_Acoustic_Image_Enhancement @ /root/reference",
                "torch": torch.__version__, "cases": {}}

    with torch.no_grad():
        for name, (kw, b, h, w, kind, ts, seed) in TEACHER_CASES.items():
            ref = KDLAE_teacher(**kw).eval()
            sd = synth.teacher_state_dict(seed=seed, temp_scale=ts, **kw)
            ref.load_state_dict(sd, strict=True)
            assert list(ref.state_dict().keys()) == list(sd.keys()), "key ORDER differs from reference"
            img = synth.seeded_tensor(name + ".img", (b, kw["inp_channels"], h, w), seed, kind)
            rate = synth.seeded_tensor(name + ".rate", (b, 1, 1, 1), seed).expand(b, 1, h, w).contiguous()
            r = ref({"img": img, "denoise_rate": rate})
            hq, sr = oracle.teacher_forward(sd, img, rate, static=kw["static"])
            d_hq = (r["hq"] - hq).abs().max().item()
            d_sr = (r["sr"] - sr).abs().max().item() if sr is not None else 0.0
            assert (r["sr"] is None) == (sr is None)
            arrs = dict(img=img.numpy(), rate=rate[:, :, 0, 0].numpy(), hq=r["hq"].numpy())
            if sr is not None:
                arrs["sr"] = r["sr"].numpy()
            np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrs)
            manifest["cases"][name] = dict(kind="teacher", kwargs=kw, shape=[b, h, w], input=kind, temp_scale=ts,
                                           seed=seed, oracle_vs_reference_maxabs=dict(hq=d_hq, sr=d_sr), n_keys=len(sd))
            print(name, "oracle-vs-reference max|d|", d_hq, d_sr, "keys", len(sd))
            assert d_hq < 2e-5 and d_sr < 2e-5

        for name, (b, f, h, w, res, seed) in STUDENT_CASES.items():
            ref = KDLAE_student(inp_channels=1, out_channels=1, residual=res, hidden_channels=[16, 32, 64]).eval()
            sd = synth.student_state_dict(seed=seed)
            ref.load_state_dict(sd, strict=True)
            assert list(ref.state_dict().keys()) == list(sd.keys())
            x = synth.seeded_tensor(name + ".x", (b, f, h, w), seed)
            r = ref(x)
            o = oracle.student_forward(sd, x, residual=res)
            d = (r - o).abs().max().item()
            np.savez_compressed(os.path.join(out_dir, name + ".npz"), x=x.numpy(), y=r.numpy())
            manifest["cases"][name] = dict(kind="student", shape=[b, f, h, w], residual=res, seed=seed,
                                           oracle_vs_reference_maxabs=d, n_keys=len(sd))
            print(name, "oracle-vs-reference max|d|", d)
            assert d < 1e-5

        for name, (b, h, w, seed) in ASDQE_CASES.items():
            ref = DenoiseRatePredictor().eval()
            sd = synth.asdqe_state_dict(seed=seed)
            ref.load_state_dict(sd, strict=True)
            assert list(ref.state_dict().keys()) == list(sd.keys())
            lq = synth.seeded_tensor(name + ".lq", (b, 3, h, w), seed)
            gt = (lq + 0.2 * synth.seeded_tensor(name + ".gt", (b, 3, h, w), seed, "normal")).clamp(0, 1)
            r = ref(lq, gt)
            pad = lambda t: torch.nn.functional.pad(t, (0, (-w) % 16, 0, (-h) % 16))
            feats_ref = ref.unet(torch.cat([ref.lq_extractor(pad(lq)), ref.gt_extractor(pad(gt)),
                                            ref.diff_extractor(pad(lq) - pad(gt))], 1))
            o = oracle.asdqe_forward(sd, lq, gt)
            feats = oracle.asdqe_trunk(sd, lq, gt)
            d, df = (r - o).abs().max().item(), (feats_ref - feats).abs().max().item()
            np.savez_compressed(os.path.join(out_dir, name + ".npz"), lq=lq.numpy(), gt=gt.numpy(),
                                score=r.numpy(), feat=feats_ref.numpy())
            manifest["cases"][name] = dict(kind="asdqe", shape=[b, h, w], seed=seed, scores=r.flatten().tolist(),
                                           oracle_vs_reference_maxabs=dict(score=d, feat=df), n_keys=len(sd))
            print(name, "oracle-vs-reference max|d|", d, df, "scores", r.flatten().tolist())
            assert d < 1e-5 and df < 1e-4

    with open(os.path.join(out_dir, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    print("wrote", out_dir)


if __name__ == "__main__":
    main()
