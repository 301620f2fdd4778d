"""CPU: host-side logic added in round 2 -- mask canonicalisation and sensor-row gathering of the fused DPS loss, the
decoder's chunk plan bounds, the build staleness check and the run-time debug knobs (no compute calls)."""
import os

import pytest
import torch

import confild_b200 as cb
from confild_b200 import _native, build, dps
from confild_b200.inference_function import _chunk_plan


def test_mask_kinds():
    T, P, cout = 4, 10, 3
    m, k = dps._mask_kind(torch.ones(P), T, P, cout)
    assert k == 1 and m.shape == (P,)
    m, k = dps._mask_kind(torch.ones(P, 1), T, P, cout)
    assert k == 1 and m.shape == (P,)
    m, k = dps._mask_kind(torch.ones(1, P, 1), T, P, cout)
    assert k == 1
    m, k = dps._mask_kind(torch.ones(P, cout), T, P, cout)
    assert k == 2 and m.shape == (P, cout)
    m, k = dps._mask_kind(torch.ones(T, P, cout), T, P, cout)
    assert k == 3 and m.shape == (T, P, cout)
    m, k = dps._mask_kind(torch.ones(T, 1, 1), T, P, cout)  # per-frame weights broadcast to the full field
    assert k == 3 and m.shape == (T, P, cout) and m.is_contiguous()


def test_sensor_rows_gathers_kept_rows():
    P = 12
    coords = torch.arange(P * 2, dtype=torch.float32).reshape(P, 2)
    mask = torch.zeros(P)
    mask[[1, 5, 7]] = 1.0
    field = torch.arange(3 * P * 2, dtype=torch.float32).reshape(3, P, 2)
    cs, idx, fs = cb.sensor_rows(coords, mask, field)
    assert idx.tolist() == [1, 5, 7]
    assert torch.equal(cs, coords[[1, 5, 7]]) and torch.equal(fs, field[:, [1, 5, 7]])


@pytest.mark.parametrize("T,P,cout", [(1, 10, 3), (1024, 65536, 3), (4096, 131072, 3), (384, 10, 3), (7, 1000, 4)])
def test_chunk_plan_covers_all_frames_within_budget(T, P, cout):
    plan = _chunk_plan(T, P, cout, 16)
    assert sum(plan) == T and all(n >= 1 for n in plan)
    budget = max(16, (128 << 20) // (P * cout * 4))
    assert max(plan) <= budget


def test_build_is_fresh_and_lock_file_is_ignored():
    build.build()
    assert not build.is_stale()
    assert os.path.exists(build.LIB_PATH)
    gi = open(os.path.join(os.path.dirname(build.PKG_DIR), ".gitignore")).read()
    assert "confild_b200/_obj/" in gi and "*.so" in gi


def test_debug_knob_entry_point():
    lib = _native.load()
    for name, value in _native.KNOB_DEFAULTS.items():
        _native.set_knob(name, value)
    assert lib.cnf_set_debug_knob(b"CNF_NO_SUCH_KNOB", 1) == 1
    assert b"unknown debug knob" in lib.cnf_last_error()
    assert lib.cnf_set_debug_knob(None, 1) == 1


def test_shared_cuda_runtime_no_embedded_runtime_names():
    """The library links the shared CUDA runtime: no runtime entry-point names are embedded in it."""
    data = open(build.LIB_PATH, "rb").read()
    assert b"MemcpyBatch" not in data


def test_auto_precision_policy_resolution():
    """precision="auto" (the default) resolves to f16f8 on the tensor-core shapes and to fp32 elsewhere."""
    m = cb.SIRENAutodecoder_film(2, 128, 3, 10, 128)
    assert m.precision == "auto" and m.resolved_precision == "f16f8"
    assert cb.SIRENAutodecoder_film(3, 384, 3, 15, 384).resolved_precision == "f16f8"
    assert cb.SIRENAutodecoder_film(2, 32, 3, 2, 64).resolved_precision == "fp32"      # H without tensor-core kernels
    assert cb.SIRENAutodecoder_film(2, 16, 3, 70, 128).resolved_precision == "fp32"    # deeper than the scale table
    assert cb.SIRENAutodecoder_film(2, 128, 3, 10, 128, precision="bf16x3").resolved_precision == "bf16x3"
    with pytest.raises(ValueError):
        cb.SIRENAutodecoder_film(2, 128, 3, 10, 128, precision="int4")._precision_code()


def test_dps_loop_cache_is_keyed_on_identity_and_version():
    """dps._cached: a hit needs the SAME tensor objects at the SAME in-place version; a new object (even at a recycled
    address) or an in-place update rebuilds the entry."""
    from confild_b200 import dps

    dps._CACHE.clear()
    calls = []

    def make(tag):
        def f():
            calls.append(tag)
            return len(calls)
        return f

    a, b = torch.zeros(4), torch.ones(4)
    v1 = dps._cached("k", (a, b), (1,), make("first"))
    assert dps._cached("k", (a, b), (1,), make("hit")) == v1 and calls == ["first"]
    assert dps._cached("k", (a, b), (2,), make("other-extra")) != v1          # different shape key
    a.add_(1.0)                                                                # in-place: version bump
    v2 = dps._cached("k", (a, b), (1,), make("after-inplace"))
    assert v2 != v1 and calls[-1] == "after-inplace"
    assert dps._cached("k", (a, b), (1,), make("hit2")) == v2 and calls[-1] == "after-inplace"
    c = a.clone()                                                              # equal content, different object
    assert dps._cached("k", (c, b), (1,), make("clone")) != v2 and calls[-1] == "clone"
    key_id = id(c)
    del c
    d = torch.zeros(4)                                                         # may reuse the freed object's id
    got = dps._cached("k", (d, b), (1,), make("new-object"))
    assert calls[-1] == "new-object" or id(d) != key_id, got
    for i in range(dps._CACHE_MAX + 5):                                        # bounded
        dps._cached("fill", (torch.zeros(1),), (i,), make("x"))
    assert len(dps._CACHE) <= dps._CACHE_MAX
