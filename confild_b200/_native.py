"""ctypes binding of libconfild_cnf.so (include/confild_cnf.h).

The product path has NO fallback: if the library is missing or a call fails, a
RuntimeError is raised with the library's own error text.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional

from . import build as _build

PREC_FP32 = 0
PREC_BF16X3 = 1
PREC_FP16 = 2
PREC_F16F8 = 3
PRECISIONS = {"fp32": PREC_FP32, "bf16x3": PREC_BF16X3, "fp16": PREC_FP16, "f16f8": PREC_F16F8}

#: every symbol include/confild_cnf.h declares (checked by tests/test_cabi.py)
EXPORTS = [
    "cnf_abi_version", "cnf_last_error", "cnf_tc_supported", "cnf_param_count", "cnf_packed_bytes",
    "cnf_pack_weights", "cnf_film_shift", "cnf_stash_bytes", "cnf_forward", "cnf_forward_gather", "cnf_backward",
    "cnf_film_shift_backward", "cnf_film_shift_backward_scaled", "cnf_forward_loss", "cnf_query_launch",
    "cnf_set_debug_knob", "cnf_group_norm_scratch_bytes", "cnf_group_norm_nhwc_bf16",
]
ABI_VERSION = 3
LOSS_PARTIALS = 4096  # CNF_LOSS_PARTIALS


class CnfDims(ctypes.Structure):
    _fields_ = [("cin", ctypes.c_int32), ("L", ctypes.c_int32), ("H", ctypes.c_int32),
                ("nl", ctypes.c_int32), ("cout", ctypes.c_int32)]


class CnfSensorLoss(ctypes.Structure):
    """``cnf_sensor_loss`` of include/confild_cnf.h."""
    _fields_ = [("d_y_meas", ctypes.c_void_p), ("d_mask", ctypes.c_void_p), ("mask_kind", ctypes.c_int32),
                ("y_scale", ctypes.c_float * 4), ("y_offset", ctypes.c_float * 4),
                ("d_gy", ctypes.c_void_p), ("d_partials", ctypes.c_void_p), ("d_norm", ctypes.c_void_p),
                ("d_extra_sq", ctypes.c_void_p)]


_lib: Optional[ctypes.CDLL] = None


def lib_path() -> str:
    return os.environ.get("CONFILD_CNF_LIB", _build.LIB_PATH)


def load() -> ctypes.CDLL:
    """Load (once) the in-tree shared library; raise loudly when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if "CONFILD_CNF_LIB" not in os.environ and _build.is_stale():
        # the library is a build artefact (git-ignored): (re)compile it when it is missing or older than its sources
        # and nvcc is around.  build() serialises concurrent ranks on a file lock and replaces the file atomically.
        try:
            _build.build()
        except Exception as e:  # noqa: BLE001
            if os.path.exists(path):
                raise RuntimeError(
                    f"{path} is older than its sources and rebuilding it failed ({e}); rebuild with "
                    "`python -m confild_b200.build --force`") from e
            raise RuntimeError(
                f"{path} not found and building it failed ({e}); build it with `python -m confild_b200.build` "
                "(there is no CPU or PyTorch fallback for the CNF decode path)") from e
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: build it with `python -m confild_b200.build` "
            "(there is no CPU or PyTorch fallback for the CNF decode path)")
    lib = ctypes.CDLL(path)
    vp, i64, sz, i32, f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_size_t, ctypes.c_int, ctypes.c_float
    dp = ctypes.POINTER(CnfDims)
    lib.cnf_abi_version.restype = i32
    lib.cnf_abi_version.argtypes = []
    lib.cnf_last_error.restype = ctypes.c_char_p
    lib.cnf_last_error.argtypes = []
    lib.cnf_tc_supported.restype = i32
    lib.cnf_tc_supported.argtypes = [dp]
    lib.cnf_param_count.restype = i32
    lib.cnf_param_count.argtypes = [dp, ctypes.POINTER(sz)]
    lib.cnf_packed_bytes.restype = i32
    lib.cnf_packed_bytes.argtypes = [dp, ctypes.POINTER(sz)]
    lib.cnf_pack_weights.restype = i32
    lib.cnf_pack_weights.argtypes = [dp, vp, f32, vp, sz, vp]
    lib.cnf_film_shift.restype = i32
    lib.cnf_film_shift.argtypes = [dp, vp, vp, i64, vp, vp]
    lib.cnf_stash_bytes.restype = i32
    lib.cnf_stash_bytes.argtypes = [dp, i32, i64, i64, ctypes.POINTER(sz)]
    lib.cnf_forward.restype = i32
    lib.cnf_forward.argtypes = [dp, vp, i32, vp, i64, vp, vp, i64, i64, vp, sz, vp]
    lib.cnf_forward_gather.restype = i32
    lib.cnf_forward_gather.argtypes = [dp, vp, i32, vp, i64, vp, ctypes.POINTER(vp), i32, i64, i64, vp]
    lib.cnf_backward.restype = i32
    lib.cnf_backward.argtypes = [dp, vp, i32, vp, vp, sz, vp, i64, i64, vp]
    lib.cnf_film_shift_backward.restype = i32
    lib.cnf_film_shift_backward.argtypes = [dp, vp, vp, i64, vp, vp]
    lib.cnf_film_shift_backward_scaled.restype = i32
    lib.cnf_film_shift_backward_scaled.argtypes = [dp, vp, vp, i64, vp, vp, vp]
    lib.cnf_set_debug_knob.restype = i32
    lib.cnf_set_debug_knob.argtypes = [ctypes.c_char_p, i32]
    lib.cnf_forward_loss.restype = i32
    lib.cnf_forward_loss.argtypes = [dp, vp, i32, vp, i64, vp, vp, i64, i64, vp, sz, ctypes.POINTER(CnfSensorLoss), vp]
    lib.cnf_query_launch.restype = i32
    lib.cnf_query_launch.argtypes = [dp, i32, i64, i64, ctypes.POINTER(i64), i32]
    lib.cnf_group_norm_scratch_bytes.restype = sz
    lib.cnf_group_norm_scratch_bytes.argtypes = [i64]
    lib.cnf_group_norm_nhwc_bf16.restype = i32
    lib.cnf_group_norm_nhwc_bf16.argtypes = [vp, vp, i64, vp, vp, vp, vp, i64, i64, i32, i32, f32, i32, vp]
    if lib.cnf_abi_version() != ABI_VERSION:
        raise RuntimeError(f"{path}: ABI version {lib.cnf_abi_version()} != {ABI_VERSION}; rebuild with "
                           "`python -m confild_b200.build --force`")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().cnf_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


#: defaults of the debug knobs (cnf_set_debug_knob); the environment variables of the same names set the initial values
KNOB_DEFAULTS = {"CNF_TC2": 1, "CNF_TC_STAGES": 0, "CNF_TC_PACKED": -1, "CNF_TC_CLUSTER": 1, "CNF_GN_CLUSTER": 256}


def set_knob(name: str, value: int) -> None:
    check(load().cnf_set_debug_knob(name.encode(), int(value)), "cnf_set_debug_knob")


def dims(cin: int, L: int, H: int, nl: int, cout: int) -> CnfDims:
    return CnfDims(int(cin), int(L), int(H), int(nl), int(cout))


def param_count(d: CnfDims) -> int:
    n = ctypes.c_size_t(0)
    check(load().cnf_param_count(ctypes.byref(d), ctypes.byref(n)), "cnf_param_count")
    return int(n.value)


def packed_bytes(d: CnfDims) -> int:
    n = ctypes.c_size_t(0)
    check(load().cnf_packed_bytes(ctypes.byref(d), ctypes.byref(n)), "cnf_packed_bytes")
    return int(n.value)


def stash_bytes(d: CnfDims, precision: int, T: int, P: int) -> int:
    n = ctypes.c_size_t(0)
    check(load().cnf_stash_bytes(ctypes.byref(d), precision, T, P, ctypes.byref(n)), "cnf_stash_bytes")
    return int(n.value)


def tc_supported(d: CnfDims) -> bool:
    return bool(load().cnf_tc_supported(ctypes.byref(d)))


def query_launch(d: CnfDims, precision: int, T: int, P: int) -> List[int]:
    vals = (ctypes.c_int64 * 7)()
    check(load().cnf_query_launch(ctypes.byref(d), precision, T, P, vals, 7), "cnf_query_launch")
    return [int(v) for v in vals]
