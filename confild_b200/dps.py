"""Fused DPS measurement distance and its latent gradient (second half).

One guided-sampling step of the reference evaluates (guided_diffusion/condition_methods.py:28-33)

    difference = measurement - operator.forward(x_0_hat)      # operator.forward = [mask *] denormalize(decode(...))
    norm       = torch.linalg.norm(difference)
    norm_grad  = torch.autograd.grad(norm, x_prev)[0]

i.e. decode, four element-wise passes over ``(T, P, cout)`` and their autograd mirror images around the two CNF
kernels.  ``measurement_norm`` computes the same ``norm`` with the residual, its squared sum and the backward seed
produced inside the decode kernel's head (``cnf_forward_loss``), runs the chain backward right away and scales by
``1/norm`` on the device (``cnf_film_shift_backward_scaled``): K1 -> forward+loss -> finalize -> backward -> K4, no
PyTorch element-wise kernel and no host synchronisation.  The returned scalar is differentiable with respect to the
latents (and through them to ``x_prev`` / the U-Net), so ``torch.autograd.grad(norm, x_prev)`` works unchanged.

``sensor_rows`` is the helper for the dense-grid operators (Case2-style mask over the full grid): it gathers the
coordinates / measurements of the rows the mask keeps, so that decode + backward run on the sensor rows only.
"""
from __future__ import annotations

import ctypes
import weakref
from typing import Optional, Tuple

import torch

from . import _native
from .folding import _affine_of_normalizer
from .nf_networks import SIRENAutodecoder_film, canonicalize


def _mask_kind(mask: torch.Tensor, T: int, P: int, cout: int) -> Tuple[torch.Tensor, int]:
    """Canonical device layout of the mask: kind 1 (P), 2 (P,cout) or 3 (T,P,cout)."""
    m = mask.to(torch.float32)
    if m.numel() == P:  # per point: (P,), (P,1), (1,P,1) ...
        return m.reshape(P).contiguous(), 1
    if m.numel() == P * cout and m.shape[-1] == cout:
        return m.reshape(P, cout).contiguous(), 2
    return m.expand(torch.broadcast_shapes(tuple(m.shape), (T, P, cout))).reshape(T, P, cout).contiguous(), 3


#: skip the backward stash for masked-out rows when at most this fraction of the points carries a non-zero weight
ZERO_ROW_SKIP_MAX_FRACTION = 0.5

# A guided-sampling loop calls measurement_norm a thousand times with the SAME measurement, mask and coordinate tensors:
# everything derived from them alone (canonical layouts, the indices of the kept rows -- whose `nonzero` costs a host
# synchronisation --, the gathered rows, the dropped rows' measurement energy) is cached.  Entries are keyed on the
# identity of the caller's tensor objects and validated by weak reference and in-place version counter, so a recycled
# address or an in-place update can never produce a stale hit.
_CACHE: dict = {}
_CACHE_MAX = 32


def _cached(name: str, tensors, extra: tuple, make):
    key = (name,) + tuple(id(t) for t in tensors) + extra
    ent = _CACHE.get(key)
    if ent is not None:
        refs, versions, val = ent
        if all(r() is t and getattr(t, "_version", 0) == v for r, v, t in zip(refs, versions, tensors)):
            return val
        del _CACHE[key]
    val = make()
    try:
        refs = [weakref.ref(t) for t in tensors]
    except TypeError:  # a key object that cannot be weakly referenced: do not cache
        return val
    for k in [k for k, (rs, _, _) in _CACHE.items() if any(r() is None for r in rs)]:
        del _CACHE[k]  # entries whose source tensors are gone would only pin device memory
    if len(_CACHE) >= _CACHE_MAX:
        _CACHE.clear()
    _CACHE[key] = (refs, [getattr(t, "_version", 0) for t in tensors], val)
    return val


def _normalizer_affine(norm, cout: int):
    """``(scale, offset)`` lists of the output normaliser's affine map, computed on the host and cached per normaliser
    object (+ the versions of its tensor parameters): the per-step call then neither copies nor synchronises, which also
    keeps it legal inside CUDA-graph capture (GraphedMeasurementNorm)."""
    params = getattr(norm, "params", None)
    tensors = tuple(p for p in (params if isinstance(params, (tuple, list)) else ()) if isinstance(p, torch.Tensor))

    def make():
        ya, yb = _affine_of_normalizer(norm, True, cout, torch.device("cpu"), torch.float32)
        return ya.tolist(), yb.tolist()

    return _cached("yaffine", (norm,) + tensors, (cout,), make)


def _kept_rows(mask1d: torch.Tensor) -> torch.Tensor:
    return torch.nonzero(mask1d != 0, as_tuple=False).reshape(-1)


class _MeasurementNormFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lat2d, coords_c, frame_stride, module, y_meas, mask, mask_kind, ya, yb, want_grad, want_field,
                kept=None, rows=None):
        # kept: indices of the rows with a non-zero per-point weight (zero-row skip), or None;
        # rows: (coords_k, meas_k, mask_k, extra_sq) when only the kept rows are to be decoded at all, else None
        lib = _native.load()
        d = module._cdims()
        cin, L, H, nl, cout = module._dims_tuple
        T, P = lat2d.shape[0], coords_c.shape[-2]
        prec = module._precision_code()
        dev = lat2d.device
        packed = module._ensure_packed()
        shift = torch.empty((T, (nl + 1) * H), dtype=torch.float32, device=dev)
        partials = torch.empty(_native.LOSS_PARTIALS, dtype=torch.float32, device=dev)
        norm = torch.empty(2, dtype=torch.float32, device=dev)
        # Rows whose per-point weight is zero have an identically zero seed: with a sparse per-point mask the dense pass
        # below runs WITHOUT the backward stash (every point is still decoded and enters the norm) and only the kept
        # rows are decoded a second time with the stash for the backward -- the stash traffic (2.8 KB / 12 KB per
        # point-frame at case1 / case4) shrinks by P / #kept while the gradient stays exact.
        if rows is not None:
            # Nobody looks at the decoded field, and a row with a zero weight has r = y_meas whatever the decoder
            # returns: decode (with the stash), score and back-propagate the kept rows only and add the dropped rows'
            # measurement energy to the sum of squares (cnf_sensor_loss.d_extra_sq) -- same norm, same gradient.
            return _MeasurementNormFunction._kept_rows_only(ctx, lib, d, module, prec, packed, lat2d, shift, rows, ya, yb,
                                                            partials, norm)
        field = torch.empty((T, P, cout), dtype=torch.float32, device=dev) if (
            want_field or prec == _native.PREC_FP32) else None
        gy = torch.empty((T, P, cout), dtype=torch.float32, device=dev)
        stash, stash_n = None, 0
        if want_grad and kept is None:
            stash_n = _native.stash_bytes(d, prec, T, P)
            stash = torch.empty(stash_n, dtype=torch.uint8, device=dev)
        loss = _native.CnfSensorLoss()
        loss.d_y_meas = y_meas.data_ptr()
        loss.d_mask = mask.data_ptr() if mask is not None else None
        loss.mask_kind = int(mask_kind)
        for o in range(4):
            loss.y_scale[o] = float(ya[o]) if o < cout else 1.0
            loss.y_offset[o] = float(yb[o]) if o < cout else 0.0
        loss.d_gy, loss.d_partials, loss.d_norm = gy.data_ptr(), partials.data_ptr(), norm.data_ptr()
        glat = None
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _native.check(lib.cnf_film_shift(d, packed.data_ptr(), lat2d.data_ptr(), T, shift.data_ptr(), stream),
                          "cnf_film_shift")
            _native.check(lib.cnf_forward_loss(d, packed.data_ptr(), prec, coords_c.data_ptr(), frame_stride,
                                               shift.data_ptr(), field.data_ptr() if field is not None else None, T, P,
                                               stash.data_ptr() if stash is not None else None, stash_n,
                                               ctypes.byref(loss), stream), "cnf_forward_loss")
            if want_grad:  # the gradient is a by-product of the step: run the chain backward now and free the stash
                gshift = torch.empty((T, (nl + 1) * H), dtype=torch.float32, device=dev)
                glat = torch.empty((T, L), dtype=torch.float32, device=dev)
                Pb, gy_b = P, gy
                if kept is not None:  # second, stash-writing decode of the kept rows only (same FiLM shifts)
                    Pb = int(kept.numel())
                    coords_k = coords_c.index_select(0, kept).contiguous()
                    gy_b = gy.index_select(1, kept).contiguous()
                    stash_n = _native.stash_bytes(d, prec, T, Pb)
                    stash = torch.empty(stash_n, dtype=torch.uint8, device=dev)
                    scratch = torch.empty((T, Pb, cout), dtype=torch.float32, device=dev)
                    _native.check(lib.cnf_forward(d, packed.data_ptr(), prec, coords_k.data_ptr(), 0, shift.data_ptr(),
                                                  scratch.data_ptr(), T, Pb, stash.data_ptr(), stash_n, stream),
                                  "cnf_forward")
                _native.check(lib.cnf_backward(d, packed.data_ptr(), prec, gy_b.data_ptr(), stash.data_ptr(), stash_n,
                                               gshift.data_ptr(), T, Pb, stream), "cnf_backward")
                _native.check(lib.cnf_film_shift_backward_scaled(d, packed.data_ptr(), gshift.data_ptr(), T,
                                                                 norm[1:].data_ptr(), glat.data_ptr(), stream),
                              "cnf_film_shift_backward_scaled")
        ctx.glat = glat
        ctx.mark_non_differentiable(*( [field] if field is not None else []))
        out_norm = norm[0]
        if field is not None:
            return out_norm, field
        return out_norm, torch.empty(0, device=dev)

    @staticmethod
    def _kept_rows_only(ctx, lib, d, module, prec, packed, lat2d, shift, rows, ya, yb, partials, norm):
        cin, L, H, nl, cout = module._dims_tuple
        coords_k, meas_k, mask_k, extra = rows
        T, dev, Pk = lat2d.shape[0], lat2d.device, int(coords_k.shape[0])
        gy = torch.empty((T, Pk, cout), dtype=torch.float32, device=dev)
        stash_n = _native.stash_bytes(d, prec, T, Pk)
        stash = torch.empty(stash_n, dtype=torch.uint8, device=dev)
        loss = _native.CnfSensorLoss()
        loss.d_y_meas, loss.d_mask, loss.mask_kind = meas_k.data_ptr(), mask_k.data_ptr(), 1
        for o in range(4):
            loss.y_scale[o] = float(ya[o]) if o < cout else 1.0
            loss.y_offset[o] = float(yb[o]) if o < cout else 0.0
        loss.d_gy, loss.d_partials, loss.d_norm = gy.data_ptr(), partials.data_ptr(), norm.data_ptr()
        loss.d_extra_sq = extra.data_ptr()
        gshift = torch.empty((T, (nl + 1) * H), dtype=torch.float32, device=dev)
        glat = torch.empty((T, L), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _native.check(lib.cnf_film_shift(d, packed.data_ptr(), lat2d.data_ptr(), T, shift.data_ptr(), stream),
                          "cnf_film_shift")
            _native.check(lib.cnf_forward_loss(d, packed.data_ptr(), prec, coords_k.data_ptr(), 0, shift.data_ptr(), None,
                                               T, Pk, stash.data_ptr(), stash_n, ctypes.byref(loss), stream),
                          "cnf_forward_loss")
            _native.check(lib.cnf_backward(d, packed.data_ptr(), prec, gy.data_ptr(), stash.data_ptr(), stash_n,
                                           gshift.data_ptr(), T, Pk, stream), "cnf_backward")
            _native.check(lib.cnf_film_shift_backward_scaled(d, packed.data_ptr(), gshift.data_ptr(), T,
                                                             norm[1:].data_ptr(), glat.data_ptr(), stream),
                          "cnf_film_shift_backward_scaled")
        ctx.glat = glat
        return norm[0], torch.empty(0, device=dev)

    @staticmethod
    def backward(ctx, gnorm, _gfield):
        if ctx.glat is None:
            raise RuntimeError("measurement_norm was evaluated without gradient tracking of the latents")
        return ctx.glat * gnorm, None, None, None, None, None, None, None, None, None, None, None, None


def measurement_norm(model: SIRENAutodecoder_film, coords: torch.Tensor, latents: torch.Tensor,
                     measurement: torch.Tensor, mask: Optional[torch.Tensor] = None, y_normalizer=None,
                     mask_measurement: bool = False, return_field: bool = False, zero_row_skip: bool = True,
                     skip_masked_decode: bool = True):
    """``torch.linalg.norm(measurement - mask * y_normalizer.denormalize(model(coords, latents)))`` as one fused
    CUDA pass, differentiable with respect to ``latents``.

    ``coords`` / ``latents`` follow ``model.forward``'s broadcasting; ``measurement`` broadcasts to the decoded field
    ``(..., cout)``; ``mask`` (optional) is per point ``(P,)`` / ``(P, 1)``, per point and channel ``(P, cout)`` or full.
    ``mask_measurement=True`` evaluates ``mask * (measurement - y)`` instead by masking the
    measurement first.  ``y_normalizer`` is the reference's ``Normalizer_ts`` (any of its affine methods) or ``None``.
    With ``return_field=True`` returns ``(norm, y_phys)`` where ``y_phys`` is the decoded (denormalised, unmasked) field.
    ``zero_row_skip`` (default on): with a sparse per-point mask the backward only visits the rows whose weight is
    non-zero (their seed is identically zero otherwise).  ``skip_masked_decode`` (default on, effective together with
    ``zero_row_skip`` when the field is not requested): those rows are not decoded either -- their residual is the
    measurement itself, so only its energy enters the norm; with ``skip_masked_decode=False`` (or ``return_field=True``)
    every point is decoded and scored by the kernel.  Norm and gradient are the same in all three modes up to fp32
    summation order.
    """
    dev = model._check_inputs(coords, latents)
    grad_on = model._check_grad_mode(coords)
    cin, L, H, nl, cout = model._dims_tuple
    coords_c, stride, lat2d, T, P, out_lead = canonicalize(coords, latents)
    if T * P == 0:
        raise ValueError("empty decode")
    if y_normalizer is not None:
        ya, yb = _normalizer_affine(y_normalizer, cout)
    else:
        ya, yb = [1.0] * cout, [0.0] * cout
    shape_key = (T, P, cout, tuple(out_lead), str(dev))
    mk, kind = (None, 0)
    if mask is not None:
        mk, kind = _cached("mask", (mask,), shape_key, lambda: _mask_kind(mask.to(dev), T, P, cout))

    def prepare_measurement():
        y = measurement.to(device=dev, dtype=torch.float32).expand(out_lead + (cout,)).reshape(T, P, cout)
        if mask is not None and mask_measurement:
            y = y * mk.reshape((1, P, 1) if kind == 1 else (1, P, cout) if kind == 2 else (T, P, cout))
        return y.contiguous()

    y_meas = _cached("meas", (measurement,) + ((mask,) if mask is not None else ()),
                     shape_key + (bool(mask_measurement),), prepare_measurement)
    want_grad = grad_on and lat2d.requires_grad
    # zero-row skip: the rows a sparse per-point mask keeps (frame-shared coordinates, tensor-core precisions)
    kept, rows = None, None
    if (want_grad and zero_row_skip and kind == 1 and stride == 0 and model._precision_code() != _native.PREC_FP32):
        idx = _cached("kept", (mask,), shape_key, lambda: _kept_rows(mk))
        if 0 < idx.numel() <= ZERO_ROW_SKIP_MAX_FRACTION * P:
            kept = idx
    if kept is not None and skip_masked_decode and not return_field:
        coords_k = _cached("coords_k", (coords, mask), shape_key, lambda: coords_c.index_select(0, kept).contiguous())
        mask_k = _cached("mask_k", (mask,), shape_key, lambda: mk.index_select(0, kept).contiguous())

        def gather_measurement():
            meas_k = y_meas.index_select(1, kept).contiguous()
            # energy of the dropped rows = total - kept, in double (zero when the measurement is already masked)
            extra = (torch.linalg.vector_norm(y_meas, dtype=torch.float64).square()
                     - torch.linalg.vector_norm(meas_k, dtype=torch.float64).square())
            return meas_k, extra.clamp_min(0).to(torch.float32).reshape(1)

        meas_k, extra = _cached("meas_k", (measurement, mask), shape_key + (bool(mask_measurement),), gather_measurement)
        rows = (coords_k, meas_k, mask_k, extra)
    norm, field = _MeasurementNormFunction.apply(lat2d, coords_c, stride, model, y_meas, mk, kind, ya, yb, want_grad,
                                                 return_field, kept, rows)
    if return_field:
        ya_t = torch.tensor(ya, device=dev)
        yb_t = torch.tensor(yb, device=dev)
        return norm, (field * ya_t + yb_t).reshape(out_lead + (cout,))
    return norm


class _GraphReplay(torch.autograd.Function):
    @staticmethod
    def forward(ctx, latents, owner):
        owner._lat.detach().copy_(latents.reshape(owner._lat.shape))
        owner._graph.replay()
        ctx.grad = owner._grad.clone().reshape(latents.shape)
        return owner._norm.clone()

    @staticmethod
    def backward(ctx, gnorm):
        return ctx.grad * gnorm, None


class GraphedMeasurementNorm:
    """``measurement_norm`` and its latent gradient captured ONCE in a CUDA graph for a guided-sampling loop whose
    shapes, coordinates, measurement, mask and decoder weights do not change from step to step (the reference's DPS
    loop: condition_methods.py:28-33 called a thousand times).  ``graphed(latents)`` copies the latents into the graph's
    static input, replays K1 -> forward + loss -> finalize -> backward -> K4 and returns the norm as a scalar that is
    differentiable with respect to ``latents`` (so ``torch.autograd.grad(norm, x_prev)`` through the U-Net works as
    with ``measurement_norm``); one replay replaces ~25 Python-level launches and allocations per step.

    ``latents`` fixes the shape (and must be a CUDA tensor); keyword arguments are those of ``measurement_norm``
    (``return_field`` is not supported).  Re-create the object after changing the decoder's weights or ``w0``.
    """

    def __init__(self, model: SIRENAutodecoder_film, coords: torch.Tensor, latents: torch.Tensor,
                 measurement: torch.Tensor, **kwargs):
        if kwargs.get("return_field"):
            raise ValueError("GraphedMeasurementNorm does not return the field")
        if not latents.is_cuda:
            raise ValueError("GraphedMeasurementNorm needs CUDA tensors")
        if any(p.requires_grad for p in model.parameters()):
            raise ValueError("freeze the decoder first (model.disable_gradient()): only the latents are differentiated")
        self._lat = latents.detach().clone().requires_grad_(True)
        self._keep = (model, coords, measurement, kwargs)  # the identity-keyed caches must stay valid

        def step():
            norm = measurement_norm(model, coords, self._lat, measurement, **kwargs)
            (grad,) = torch.autograd.grad(norm, self._lat)
            return norm, grad

        dev = latents.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # warm-up outside capture: weight packing, caches, allocator
            for _ in range(3):
                step()
        torch.cuda.current_stream(dev).wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            norm, grad = step()
        self._norm, self._grad = norm.detach(), grad

    def __call__(self, latents: torch.Tensor) -> torch.Tensor:
        if latents.numel() != self._lat.numel() or latents.device != self._lat.device:
            raise ValueError(f"latents must have {self._lat.numel()} elements on {self._lat.device}")
        return _GraphReplay.apply(latents.to(torch.float32), self)


def sensor_rows(coords: torch.Tensor, mask: torch.Tensor, *fields: torch.Tensor):
    """Gather the rows a per-point ``mask (P,)`` keeps: returns ``(coords[idx], idx, *[f[..., idx, :] for f in fields])``.

    For the dense-grid operators (the mask multiplies the decoded full grid, measurements.py:91-97) every masked-out
    row contributes ``measurement**2`` to the norm and nothing to the gradient, so the guided step only needs the
    decoder on the kept rows: ``measurement_norm(model, coords_s, latents, meas_s)`` on the gathered rows gives the
    same latent gradient direction at ``len(idx)/P`` of the cost.
    """
    idx = torch.nonzero(mask.reshape(-1) != 0, as_tuple=False).reshape(-1)
    return (coords.reshape(-1, coords.shape[-1])[idx], idx) + tuple(f[..., idx, :] for f in fields)
