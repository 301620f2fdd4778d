"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol the header declares.
No compute entry point is called here (no GPU)."""
import ctypes
import os
import re

import pytest

from confild_b200 import _native, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _native.load()


def test_header_symbols_are_exported(lib):
    header = open(os.path.join(ROOT, "include", "confild_cnf.h")).read()
    declared = set(re.findall(r"^\s*(?:const\s+char\s*\*|int|size_t)\s+(cnf_\w+)\s*\(", header, flags=re.M))
    assert declared == set(_native.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_sizes_and_argument_checks(lib):
    d = _native.dims(2, 128, 128, 10, 3)
    assert _native.param_count(d) == 346115  # reference module parameter count at case1
    assert _native.param_count(_native.dims(3, 384, 384, 15, 3)) == 4579587
    assert _native.packed_bytes(d) > 4 * 346115
    assert _native.tc_supported(d)
    assert not _native.tc_supported(_native.dims(2, 32, 64, 2, 3))
    assert _native.stash_bytes(d, _native.PREC_BF16X3, 4, 100) == 4 * 128 * 11 * 128 * 2  # rows padded to whole tiles
    assert _native.stash_bytes(d, _native.PREC_FP32, 4, 100) == 4 * 100 * 11 * 128 * 4
    n = ctypes.c_size_t(0)
    bad = _native.dims(0, 128, 128, 10, 3)
    assert lib.cnf_param_count(ctypes.byref(bad), ctypes.byref(n)) == 1
    assert b"non-positive" in lib.cnf_last_error()
    # NULL device pointers are rejected before any CUDA call
    assert lib.cnf_forward(d, None, 1, None, 0, None, None, 1, 1, None, 0, None) == 1
    assert lib.cnf_backward(d, None, 1, None, None, 0, None, 1, 1, None) == 1
    assert lib.cnf_film_shift(d, None, None, 1, None, None) == 1
    # the sampler-side GroupNorm entry point: scratch size, NULL pointers and unsupported channel counts
    assert lib.cnf_group_norm_scratch_bytes(2) == 2 * 128 * 64 * 2 * 4 and lib.cnf_group_norm_scratch_bytes(0) == 0
    assert lib.cnf_group_norm_nhwc_bf16(None, None, 0, None, None, None, None, 2, 64, 128, 32, 1e-5, 1, None) == 1
    fake = ctypes.c_void_p(0x1000)  # non-NULL, 16-byte aligned, never dereferenced: the shape check fails first
    assert lib.cnf_group_norm_nhwc_bf16(fake, None, 0, fake, fake, fake, fake, 2, 64, 20, 4, 1e-5, 1, None) == 2
    assert b"multiple of 8" in lib.cnf_last_error()
    assert lib.cnf_group_norm_nhwc_bf16(fake, None, 0, fake, fake, fake, fake, 0, 64, 128, 32, 1e-5, 1, None) == 0  # N = 0


def test_sass_has_tcgen05_and_bulk_tma():
    build.build()
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", build.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass      # tcgen05.mma kind::f16
    assert "LDTM" in sass         # tcgen05.ld
    assert "UBLKCP" in sass       # cp.async.bulk (1-D TMA)
    assert "MUFU.SIN" in sass
    assert "UTCQMMA" in sass      # tcgen05.mma kind::f8f6f4 (the f16f8 precision's correction products)
    assert "FHFMA" in sass        # mixed-precision fma.rn.f32.f16 / .bf16 (operand residuals)
