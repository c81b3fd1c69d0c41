"""B200-native (sm_100a) forward path of KDLAE-T / KDLAE-S / ASDQE behind the reference's nn.Module API.

    from rethink_acoustic_image_enhancement_b200 import KDLAE_teacher, KDLAE_student, DenoiseRatePredictor

These classes keep the reference's constructors, state_dict layout and forward signatures
(KDLAE/KDLAE_model.py, ASDQE/ASDQE_model.py) and run the forward in hand-written CUDA through the
C ABI in include/kdlae_b200.h.  The CUDA library is mandatory: nothing here falls back to PyTorch ops.
"""
from .kdlae_model import KDLAE_teacher, KDLAE_student, RestormerSuperResolutionParam2  # noqa: F401
from .asdqe_model import DenoiseRatePredictor  # noqa: F401
from . import _lib  # noqa: F401
from .pipeline import teacher_infer_uint8, preprocess_u8, postprocess_u8  # noqa: F401

__all__ = ["KDLAE_teacher", "KDLAE_student", "RestormerSuperResolutionParam2", "DenoiseRatePredictor",
           "teacher_infer_uint8", "preprocess_u8", "postprocess_u8"]
