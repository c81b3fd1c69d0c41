"""ctypes binding of libkdlae_b200.so (the C ABI declared in include/kdlae_b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libkdlae_b200.so")

PREC_FP32 = 0
PREC_BF16 = 1
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16}


class TeacherCfg(C.Structure):
    _fields_ = [
        ("inp_channels", C.c_int), ("out_channels", C.c_int), ("dim", C.c_int),
        ("num_blocks", C.c_int * 4), ("num_refinement_blocks", C.c_int),
        ("heads", C.c_int * 4), ("hidden", C.c_int * 4),
        ("ln_with_bias", C.c_int), ("sr_head", C.c_int), ("params_cat", C.c_int),
    ]


class StudentCfg(C.Structure):
    _fields_ = [("hidden", C.c_int * 3), ("residual", C.c_int)]


class AsdqeCfg(C.Structure):
    _fields_ = [("in_channels", C.c_int), ("dim", C.c_int)]


_PP = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); must list every symbol declared in include/kdlae_b200.h
SIGNATURES = {
    "kdlae_abi_version": (C.c_int, []),
    "kdlae_last_error": (C.c_char_p, []),
    "kdlae_device_check": (C.c_int, [C.c_int]),
    "kdlae_launch_count": (C.c_ulonglong, []),
    "kdlae_profile_num_classes": (C.c_int, []),
    "kdlae_profile_class_name": (C.c_char_p, [C.c_int]),
    "kdlae_profile_begin": (C.c_int, []),
    "kdlae_profile_end": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.POINTER(C.c_longlong)]),
    "kdlae_profile_launches": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                         C.POINTER(C.c_double)]),
    "kdlae_teacher_num_tensors": (C.c_int, [C.POINTER(TeacherCfg)]),
    "kdlae_teacher_packed_bytes": (C.c_size_t, [C.POINTER(TeacherCfg), C.c_int]),
    "kdlae_teacher_pack": (C.c_int, [C.POINTER(TeacherCfg), _PP, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "kdlae_teacher_workspace_bytes": (C.c_size_t, [C.POINTER(TeacherCfg), C.c_int, C.c_int, C.c_int, C.c_int]),
    "kdlae_teacher_forward": (C.c_int, [C.POINTER(TeacherCfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "kdlae_student_packed_bytes": (C.c_size_t, [C.POINTER(StudentCfg), C.c_int]),
    "kdlae_student_pack": (C.c_int, [C.POINTER(StudentCfg), _PP, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "kdlae_student_workspace_bytes": (C.c_size_t, [C.POINTER(StudentCfg), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "kdlae_student_forward": (C.c_int, [C.POINTER(StudentCfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "kdlae_asdqe_packed_bytes": (C.c_size_t, [C.POINTER(AsdqeCfg), C.c_int]),
    "kdlae_asdqe_pack": (C.c_int, [C.POINTER(AsdqeCfg), _PP, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "kdlae_asdqe_workspace_bytes": (C.c_size_t, [C.POINTER(AsdqeCfg), C.c_int, C.c_int, C.c_int, C.c_int]),
    "kdlae_asdqe_forward": (C.c_int, [C.POINTER(AsdqeCfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "kdlae_conv_gemm": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "kdlae_mdta_gram_scratch_floats": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "kdlae_mdta_gram": (C.c_int, [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_void_p]),
    "kdlae_ln_stats": (C.c_int, [C.c_void_p, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "kdlae_dwconv3x3": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p]),
    "kdlae_pwdw_f2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.c_void_p]),
    "kdlae_preprocess_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_int, C.c_void_p]),
    "kdlae_postprocess_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p]),
    "kdlae_pwdw_t": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                               C.c_int, C.c_int, C.c_void_p]),
    "kdlae_gdfn_train_ws_floats": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "kdlae_gdfn_forward_train": (C.c_int, [C.c_void_p] * 6 + [C.c_int] * 5 + [C.c_void_p, C.c_void_p]),
    "kdlae_gdfn_backward": (C.c_int, [C.c_void_p] * 11 + [C.c_int] * 5 + [C.c_void_p, C.c_void_p]),
    "kdlae_mdta_train_ws_floats": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "kdlae_mdta_forward_train": (C.c_int, [C.c_void_p] * 7 + [C.c_int] * 5 + [C.c_void_p, C.c_void_p]),
    "kdlae_mdta_backward": (C.c_int, [C.c_void_p] * 13 + [C.c_int] * 5 + [C.c_void_p, C.c_void_p]),
    "kdlae_set_train_matmul_tf32": (C.c_int, [C.c_int]),
    "kdlae_train_matmul_tf32": (C.c_int, []),
    "kdlae_conv_train_ws_floats": (C.c_size_t, [C.c_int] * 6),
    "kdlae_conv_train_forward": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 7 + [C.c_void_p]),
    "kdlae_conv_train_backward": (C.c_int, [C.c_void_p] * 5 + [C.c_int] * 7 + [C.c_void_p, C.c_void_p]),
    "kdlae_grad_norm_sq": (C.c_int, [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p]),
    "kdlae_adamw_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_float, C.c_float, C.c_float,
                                   C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "kdlae_debug_trace_begin": (C.c_int, []),
    "kdlae_debug_trace_end": (C.c_int, [C.POINTER(C.c_ulonglong), C.c_int, C.c_char_p, C.c_int]),
    "kdlae_psnr_scratch_bytes": (C.c_size_t, [C.c_int]),
    "kdlae_psnr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                             C.c_void_p]),
    "kdlae_l1_sr_scratch_bytes": (C.c_size_t, []),
    "kdlae_l1_sr_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_long, C.c_float, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (once) and bind every declared symbol. Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m rethink_acoustic_image_enhancement_b200.build` "
            "(there is no CPU or PyTorch fallback for the KDLAE/ASDQE forward path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.kdlae_abi_version() != 3:
        raise RuntimeError("libkdlae_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().kdlae_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (status {status}): {msg}")


def launch_count() -> int:
    return int(load().kdlae_launch_count())


def profile_begin() -> None:
    check(load().kdlae_profile_begin(), "kdlae_profile_begin")


def profile_launches(max_launches: int = 1 << 16) -> list:
    """[(class name, ms, flops, bytes)] per launch since profile_begin(), in launch order; call before profile_end()."""
    lib = load()
    cls = (C.c_int * max_launches)()
    ms, fl, by = (C.c_double * max_launches)(), (C.c_double * max_launches)(), (C.c_double * max_launches)()
    n = lib.kdlae_profile_launches(max_launches, cls, ms, fl, by)
    return [(lib.kdlae_profile_class_name(cls[i]).decode(), ms[i], fl[i], by[i]) for i in range(n)]


def profile_end() -> dict:
    """{class name: dict(ms, flops, bytes, launches)} for every kernel class that launched since profile_begin()."""
    lib = load()
    n = lib.kdlae_profile_num_classes()
    ms, fl, by = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)()
    ln = (C.c_longlong * n)()
    check(lib.kdlae_profile_end(n, ms, fl, by, ln), "kdlae_profile_end")
    out = {}
    for i in range(n):
        if ln[i]:
            out[lib.kdlae_profile_class_name(i).decode()] = dict(ms=ms[i], flops=fl[i], bytes=by[i], launches=int(ln[i]))
    return out
