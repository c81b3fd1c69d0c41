"""CPU oracle for the KDLAE-T / KDLAE-S / ASDQE forward path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or the timed
CPU baseline.  The product path (``rethink_acoustic_image_enhancement_b200``)
never imports this package and fails loudly when its CUDA library is missing.

Parity pin: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against the reference *itself*:
``oracle/make_golden.py`` imports the unmodified reference modules from
``/root/reference`` in the build container, checks that this restatement
reproduces their outputs (max-abs 0 to ~1e-6, see tests/golden/MANIFEST.json)
and freezes inputs/outputs as fixtures under ``tests/golden/``.
"""
from .functional import (  # noqa: F401
    teacher_forward,
    student_forward,
    asdqe_forward,
    asdqe_trunk,
)
from .synth import (  # noqa: F401
    teacher_state_dict,
    student_state_dict,
    asdqe_state_dict,
    seeded_tensor,
    psnr,
)
