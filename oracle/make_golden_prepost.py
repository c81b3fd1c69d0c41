#!/usr/bin/env python
"""Freeze golden vectors for the uint8 pre / post-processing by EXECUTING the reference's own notebook cell.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference and cv2):

    python oracle/make_golden_prepost.py      # writes tests/golden/prepost.npz

The reference's inference recipe is cell 5 of KDLAE/KDLAE_T.ipynb - notebook code, not an importable function.  This script
reads the cell's source from the notebook at run time (nothing of it is stored in this repository), cuts the numeric part
(`load_image_as_tensor` and the `with torch.no_grad():` block up to the plotting code) and executes it unmodified on synthetic
PNG files written with cv2, with
  * `device` = cpu,
  * `model_restoration` = a deterministic stand-in (the cell only needs `pred['hq']`, `pred['sr']` of the right shapes; the
    stand-in returns values outside [0, 1] and exact rounding ties so that the clamp / crop / uint8 conversion are exercised),
  * `img_as_ubyte`: scikit-image (requirements.txt:9, unpinned) is NOT installed in this image, so the cell gets a restatement
    of its published float -> uint8 conversion (skimage/util/dtype.py `_convert`: multiply by 255 in float32, np.rint, np.clip
    to [0, 255], astype(uint8); inputs outside [-1, 1] raise).  Everything else the cell runs is the reference's code.
It then checks oracle/prepost.py against what the cell produced and stores inputs + outputs.
"""
import json
import os
import sys
import tempfile
import textwrap

import cv2
import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("KDLAE_REFERENCE_DIR", "/root/reference")

from oracle import prepost as op  # noqa: E402


def img_as_ubyte(image):
    image = np.asarray(image)
    if image.dtype == np.uint8:
        return image
    if np.min(image) < -1.0 or np.max(image) > 1.0:
        raise ValueError("Images of type float must be between -1 and 1.")
    out = np.multiply(image, 255, dtype=np.float32)
    np.rint(out, out=out)
    np.clip(out, 0, 255, out=out)
    return out.astype(np.uint8)


def reference_cell_source():
    nb = json.load(open(os.path.join(REF, "KDLAE", "KDLAE_T.ipynb")))
    cells = ["".join(c["source"]) for c in nb["cells"] if c["cell_type"] == "code"]
    cell = next(c for c in cells if "load_image_as_tensor" in c and "img_as_ubyte" in c)
    lines = cell.split("\n")
    a = next(i for i, l in enumerate(lines) if l.startswith("def load_image_as_tensor"))
    b = next(i for i, l in enumerate(lines) if l.strip().startswith("def prepare_image"))
    return "\n".join(lines[a:b])


class StandInModel:
    """Deterministic stand-in for the network: the cell's pre / post-processing does not depend on what produced `pred`."""

    def __init__(self, seed):
        self.seed = seed
        self.seen = None

    def __call__(self, inp):
        img, rate = inp["img"], inp["denoise_rate"]
        self.seen = (img.clone(), rate.clone())
        g = torch.Generator().manual_seed(self.seed)
        hq = img * 1.3 - 0.15 + 0.2 * (torch.rand(img.shape, generator=g) - 0.5) * rate
        hq[0, 0, 0, :4] = torch.tensor([0.5, 1.5, 2.5, 254.5]) / 255.0            # rint ties: half to even
        sr = F.interpolate(hq, scale_factor=2, mode="nearest") * 0.95 + 0.1 * (torch.rand(
            (img.shape[0], img.shape[1], img.shape[2] * 2, img.shape[3] * 2), generator=g) - 0.5)
        self.pred = {"hq": hq, "sr": sr}
        return self.pred


def main():
    src = reference_cell_source()
    code = compile(src, "KDLAE_T.ipynb:cell5", "exec")
    rng = np.random.default_rng(11)
    out = {}
    cases = [(37, 50, 3, 0.6), (64, 64, 1, 1.0), (9, 15, 1, 0.25), (40, 33, 3, 0.0)]
    with tempfile.TemporaryDirectory() as td:
        for ci, (h, w, c, rate) in enumerate(cases):
            img = rng.integers(0, 256, size=(h, w, c), dtype=np.uint8)
            img[rng.random((h, w)) < 0.4] = 0                          # blind zone: exact zeros in every channel
            path = os.path.join(td, f"case{ci}.png")
            # the cell converts BGR -> RGB after cv2.imread: write the file so that it reads back as `img`
            cv2.imwrite(path, img[:, :, ::-1] if c == 3 else img[:, :, 0])
            model = StandInModel(100 + ci)
            ns = dict(torch=torch, F=F, np=np, cv2=cv2, img_as_ubyte=img_as_ubyte, device=torch.device("cpu"), denoise_rate=rate,
                      img_multiple_of=8, lq_path=path, model_restoration=model, print=lambda *a, **k: None)
            exec(code, ns)
            x_in, alpha = model.seen
            hq_u8, sr_u8 = ns["restored_np"], ns["restored_sr_np"]
            # the restatement against the cell
            x_o, a_o = op.preprocess_u8(img[None], rate)
            assert torch.equal(x_o, x_in) and torch.equal(a_o, alpha), f"case {ci}: preprocess differs"
            hq_o = op.postprocess_u8(model.pred["hq"], img[None], 1)[0]
            sr_o = op.postprocess_u8(model.pred["sr"], img[None], 2)[0]
            assert np.array_equal(hq_o, hq_u8) and np.array_equal(sr_o, sr_u8), f"case {ci}: postprocess differs"
            out[f"c{ci}_img"] = img
            out[f"c{ci}_rate"] = np.float32(rate)
            out[f"c{ci}_x"] = x_in.numpy()
            out[f"c{ci}_alpha_shape"] = np.array(alpha.shape)
            out[f"c{ci}_pred_hq"] = model.pred["hq"].numpy().astype(np.float32)
            out[f"c{ci}_pred_sr"] = model.pred["sr"].numpy().astype(np.float32)
            out[f"c{ci}_hq_u8"] = hq_u8
            out[f"c{ci}_sr_u8"] = sr_u8
            print(f"case {ci}: {h}x{w}x{c} rate {rate}: padded {tuple(x_in.shape)}, hq {hq_u8.shape}, sr {sr_u8.shape} - restatement identical")
    out["n_cases"] = np.int64(len(cases))
    dst = os.path.join(ROOT, "tests", "golden", "prepost.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
