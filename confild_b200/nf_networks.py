"""Drop-in host side of the CNF decoder: ``SIRENAutodecoder_film``.

Mirrors the reference's ``ConditionalNeuralField/cnf/nf_networks.py:443-508`` (class, constructor
signature, ``forward(coords, latents)``, ``disable_gradient``, ``net1``/``net2``/``nl`` attribute
names and state-dict keys) so that the reference's callers -- ``trainer.infer``
(scripts/train.py:265-279), ``pass_through_model_batch``/``decoder``
(cnf/inference_function.py:22-76), ``CNF_inference.predict`` (:219-259) and the DPS measurement
operators (guided_diffusion/measurements.py:77,120,161,207) -- can use it unchanged.

The arithmetic itself runs in hand-written sm_100a CUDA kernels behind the C ABI declared in
``include/confild_cnf.h`` (``libconfild_cnf.so``, loaded with ctypes).  PyTorch only owns the
buffers and the stream.  There is no CPU, eager-PyTorch or Triton fallback: without a CUDA device
and the built library ``forward`` raises.
"""
from __future__ import annotations

import math
import os
import warnings
from typing import Optional, Tuple

import torch
from torch import nn

from . import _native

DEFAULT_W0 = 30.0  # reference: cnf/initialization.py:5

#: what each ``precision`` string means (bench.py prints it next to every number)
PRECISION_NOTES = {
    "auto": "policy: f16f8 where the tensor-core kernels cover the shape, fp32 otherwise",
    "bf16x3": "tcgen05 bf16 hi/lo split, 3 MMAs per product, fp32 accumulate",
    "fp16": "tcgen05 single fp16 MMA per product, fp32 accumulate",
    "f16f8": "tcgen05 fp16 product + two fp8 (e5m2/e4m3, kind::f8f6f4) correction products = 2 MMA-equivalents per "
             "product, fp32 accumulate",
    "fp32": "CUDA-core fp32 FMA",
}


class Sine(nn.Module):
    """Holder of the sine frequency ``w0`` (reference: cnf/components.py:19-25).

    As in the reference a single shared instance serves every decoder, and ``w0`` is read at call
    time, so code that mutates ``model.nl.w0`` keeps working.
    """

    def __init__(self, w0: float = DEFAULT_W0):
        self.w0 = w0
        super().__init__()

    def forward(self, input):  # noqa: A002 - reference argument name
        return torch.sin(self.w0 * input)


_SHARED_SINE = Sine()


class BatchLinear(nn.Linear):
    """Parameter container with the reference's layer class name (cnf/components.py:55-76).

    Inside the decoder the weights are consumed by the CUDA kernels; this class only provides
    ``weight``/``bias`` with ``nn.Linear``'s default initialisation and state-dict layout.
    """


def sine_init(m: nn.Module, w0: float = DEFAULT_W0) -> None:
    """U(+-sqrt(6/fan_in)/w0) (reference: cnf/initialization.py:117-125)."""
    with torch.no_grad():
        if hasattr(m, "weight"):
            bound = math.sqrt(6 / m.weight.size(-1)) / w0
            m.weight.uniform_(-bound, bound)


def first_layer_sine_init(m: nn.Module) -> None:
    """U(+-1/fan_in) (reference: cnf/initialization.py:127-132)."""
    with torch.no_grad():
        if hasattr(m, "weight"):
            bound = 1 / m.weight.size(-1)
            m.weight.uniform_(-bound, bound)


def _prod(xs) -> int:
    r = 1
    for x in xs:
        r *= int(x)
    return r


def canonicalize(coords: torch.Tensor, latents: torch.Tensor):
    """Map PyTorch-broadcast ``(coords, latents)`` onto the kernel's (frames, points) problem.

    Returns ``(coords_c, frame_stride, lat2d, T, P, out_lead)`` where ``lat2d`` is ``(T, L)``,
    ``coords_c`` is ``(P, cin)`` shared by all frames (``frame_stride == 0``) or ``(T, P, cin)``
    (``frame_stride == P*cin``) and ``out_lead`` the broadcast leading shape of the result.
    Frames are the leading dims over which the latents vary, points the remaining ones
    (reference shapes: SURVEY.md 3.1-3.4).
    """
    cin, L = coords.shape[-1], latents.shape[-1]
    lead_c, lead_l = tuple(coords.shape[:-1]), tuple(latents.shape[:-1])
    out_lead = tuple(torch.broadcast_shapes(lead_c, lead_l))
    nd = len(out_lead)
    lc = (1,) * (nd - len(lead_c)) + lead_c
    ll = (1,) * (nd - len(lead_l)) + lead_l
    s = nd
    while s > 0 and ll[s - 1] == 1:
        s -= 1
    T, P = _prod(out_lead[:s]), _prod(out_lead[s:])
    lat2d = latents.reshape(ll + (L,)).expand(out_lead[:s] + (1,) * (nd - s) + (L,)).reshape(T, L)
    if all(lc[i] == 1 for i in range(s)):
        coords_c = coords.reshape(lc + (cin,)).expand((1,) * s + out_lead[s:] + (cin,)).reshape(P, cin)
        stride = 0
    else:
        coords_c = coords.reshape(lc + (cin,)).expand(out_lead + (cin,)).reshape(T, P, cin)
        stride = P * cin
    return coords_c.contiguous(), stride, lat2d.contiguous(), T, P, out_lead


class _CNFDecodeFunction(torch.autograd.Function):
    """forward: K1 (FiLM shift GEMM) + K2 (layer chain); backward: K3 (chain backward) + K4."""

    @staticmethod
    def forward(ctx, lat2d, coords_c, frame_stride, module, want_grad):
        T, L = lat2d.shape
        P = coords_c.shape[-2]
        out, stash = module._launch_forward(coords_c, frame_stride, lat2d, T, P, want_grad)
        ctx.module = module
        ctx.T, ctx.P = T, P
        ctx.stash = stash
        ctx.precision = module._precision_code()
        ctx.packed = module._packed  # keep the weights of THIS forward alive for backward
        return out

    @staticmethod
    def backward(ctx, gout):
        if ctx.stash is None:
            raise RuntimeError("CNF decode was run without the backward stash (latents did not require grad)")
        glat = ctx.module._launch_backward(gout.contiguous(), ctx.stash, ctx.packed, ctx.precision, ctx.T, ctx.P)
        ctx.stash = None
        return glat, None, None, None, None


class SIRENAutodecoder_film(nn.Module):
    """FiLM-modulated SIREN auto-decoder, same interface as the reference class
    (cnf/nf_networks.py:443-501):

        x_0 = coords;  x_{i+1} = sin(w0 * (net1[i](x_i) + net2[i](latents))), i = 0..nl;  y = net1[nl+1](x)

    Extra, optional keyword (not in the reference): ``precision`` in {"auto", "f16f8", "bf16x3", "fp16",
    "fp32"} selects the operand format of the hidden-layer GEMMs (default "auto", or the environment
    variable CONFILD_PRECISION).  "auto" is the stated policy: "f16f8" -- the fastest mode whose measured
    forward error (2e-5..1.1e-4 on the four recipe shapes) is >= 9x inside the 1e-3 contract -- wherever the
    tensor-core kernels cover the shape, "fp32" (CUDA cores) otherwise.  ``resolved_precision`` tells which
    one runs; see DESIGN.md section 3 for the measured error of each mode.
    """

    def __init__(self, in_coord_features, in_latent_features, out_features, num_hidden_layers, hidden_features,
                 outermost_linear=False, nonlinearity="sine", weight_init=None, bias_init=None,
                 premap_mode=None, **kwargs):
        super().__init__()
        precision = kwargs.pop("precision", None)
        if premap_mode is not None:
            raise NotImplementedError(
                "premap_mode is not None: no recipe of the reference uses a coordinate pre-map on this path; "
                "use the reference module")
        if nonlinearity != "sine":
            raise NotImplementedError(f"nonlinearity={nonlinearity!r}: only 'sine' has CUDA kernels")
        self.premap_mode = premap_mode
        self.first_layer_init = None
        self.nl = _SHARED_SINE
        self.weight_init = weight_init if weight_init is not None else sine_init

        self.net1 = nn.ModuleList(
            [BatchLinear(in_coord_features, hidden_features)]
            + [BatchLinear(hidden_features, hidden_features) for _ in range(num_hidden_layers)]
            + [BatchLinear(hidden_features, out_features)])
        self.net2 = nn.ModuleList(
            [BatchLinear(in_latent_features, hidden_features, bias=False) for _ in range(num_hidden_layers + 1)])
        if self.weight_init is not None:
            self.net1.apply(self.weight_init)
            self.net2.apply(self.weight_init)
        self.net1[0].apply(first_layer_sine_init)
        self.net2[0].apply(first_layer_sine_init)
        if bias_init is not None:
            self.net2.apply(bias_init)

        self.precision = precision or os.environ.get("CONFILD_PRECISION", "auto")
        self._dims_tuple = (int(in_coord_features), int(in_latent_features), int(hidden_features),
                            int(num_hidden_layers), int(out_features))
        self._packed: Optional[torch.Tensor] = None
        self._packed_key = None
        self._timing = None
        self._warned_no_weight_grad = False

    # ------------------------------------------------------------------ reference API
    def disable_gradient(self):
        for param in self.parameters():
            param.requires_grad = False

    def _check_inputs(self, coords, latents) -> torch.device:
        """Type / shape / dtype / device checks shared by every entry point (forward, decode_into, the fused loss):
        raw device pointers cross the C ABI, so nothing unchecked may reach it."""
        if not (isinstance(coords, torch.Tensor) and isinstance(latents, torch.Tensor)):
            raise TypeError("coords and latents must be tensors")
        cin, L, H, nl, cout = self._dims_tuple
        if coords.shape[-1] != cin:
            raise ValueError(f"coords last dim is {coords.shape[-1]}, expected in_coord_features={cin}")
        if latents.shape[-1] != L:
            raise ValueError(f"latents last dim is {latents.shape[-1]}, expected in_latent_features={L}")
        if coords.dtype != torch.float32 or latents.dtype != torch.float32:
            raise TypeError(f"fp32 inputs required (got coords {coords.dtype}, latents {latents.dtype})")
        dev = self.net1[0].weight.device
        if dev.type != "cuda":
            raise RuntimeError(
                "SIRENAutodecoder_film (confild_b200) runs only on a CUDA device: move the module with "
                ".to('cuda'); there is no CPU fallback")
        if coords.device != dev or latents.device != dev:
            raise RuntimeError(f"coords ({coords.device}) / latents ({latents.device}) must be on {dev}")
        return dev

    def _check_grad_mode(self, coords) -> bool:
        """Returns whether autograd is recording.  Only dL/dlatents exists on this path: coordinates that require
        grad and training mode raise; parameters that merely keep ``requires_grad=True`` (the DPS operators only call
        ``.eval()``, measurements.py:209) are tolerated with a one-time warning, since they receive no gradient."""
        grad_on = torch.is_grad_enabled()
        if grad_on and coords.requires_grad:
            raise NotImplementedError("gradient with respect to coords is not implemented on this path")
        if grad_on and any(p.requires_grad for p in self.parameters()):
            if self.training:
                raise NotImplementedError(
                    "weight gradients are not implemented (decode path only): call .eval() or disable_gradient() "
                    "for decoding / DPS, and use the reference module for training")
            if not self._warned_no_weight_grad:
                self._warned_no_weight_grad = True
                warnings.warn(
                    "confild_b200.SIRENAutodecoder_film: parameters have requires_grad=True but this decode path "
                    "only propagates gradients to the latents; parameter .grad stays None (call disable_gradient() "
                    "to silence this)", stacklevel=3)
        return grad_on

    def forward(self, coords, latents):
        self._check_inputs(coords, latents)
        cin, L, H, nl, cout = self._dims_tuple
        grad_on = self._check_grad_mode(coords)
        coords_c, stride, lat2d, T, P, out_lead = canonicalize(coords, latents)
        if T * P == 0:
            return coords.new_zeros(out_lead + (cout,))
        want_grad = grad_on and lat2d.requires_grad
        out = _CNFDecodeFunction.apply(lat2d, coords_c, stride, self, want_grad)
        return out.reshape(out_lead + (cout,))

    # ------------------------------------------------------------------ native plumbing
    @property
    def resolved_precision(self) -> str:
        """The precision that actually runs: ``self.precision`` unless it is "auto" (see the class docstring)."""
        if self.precision != "auto":
            return self.precision
        cin, L, H, nl, cout = self._dims_tuple
        ok = _native.tc_supported(self._cdims()) and nl <= 64
        return "f16f8" if ok else "fp32"

    def _precision_code(self) -> int:
        name = self.resolved_precision
        try:
            code = _native.PRECISIONS[name]
        except KeyError:
            raise ValueError(f"precision must be 'auto' or one of {sorted(_native.PRECISIONS)}, got {self.precision!r}")
        if code != _native.PREC_FP32 and not _native.tc_supported(self._cdims()):
            raise NotImplementedError(
                f"precision={self.precision!r} needs hidden_features in {{128,256,384}}, >=1 hidden layer and "
                f"<=4 coordinate/output features (got dims {self._dims_tuple}); set precision='fp32'")
        return code

    def _cdims(self) -> "_native.CnfDims":
        cin, L, H, nl, cout = self._dims_tuple
        return _native.dims(cin, L, H, nl, cout)

    def _ensure_packed(self) -> torch.Tensor:
        """(Re)pack the parameters into the kernels' device layout whenever they changed
        (load_state_dict, .to(), in-place edits, a new w0)."""
        params = list(self.parameters())
        w0 = float(self.nl.w0)
        key = (tuple((p.data_ptr(), p._version) for p in params), w0, str(params[0].device))
        if self._packed is not None and key == self._packed_key:
            return self._packed
        lib = _native.load()
        d = self._cdims()
        dev = params[0].device
        for p in params:
            if p.dtype != torch.float32 or p.device != dev:
                raise TypeError("all parameters must be fp32 on one CUDA device")
        flat = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
        if flat.numel() != _native.param_count(d):
            raise RuntimeError("parameter count does not match the module dimensions")
        nbytes = _native.packed_bytes(d)
        packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _native.check(lib.cnf_pack_weights(d, flat.data_ptr(), w0, packed.data_ptr(), nbytes, stream),
                          "cnf_pack_weights")
        self._packed, self._packed_key = packed, key
        return packed

    def _launch_forward(self, coords_c, frame_stride, lat2d, T, P, want_grad
                        ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        lib = _native.load()
        d = self._cdims()
        cin, L, H, nl, cout = self._dims_tuple
        prec = self._precision_code()
        dev = lat2d.device
        packed = self._ensure_packed()
        shift = torch.empty((T, (nl + 1) * H), dtype=torch.float32, device=dev)
        out = torch.empty((T, P, cout), dtype=torch.float32, device=dev)
        stash, stash_ptr, stash_n = None, None, 0
        if want_grad:
            stash_n = _native.stash_bytes(d, prec, T, P)
            stash = torch.empty(stash_n, dtype=torch.uint8, device=dev)
            stash_ptr = stash.data_ptr()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _native.check(lib.cnf_film_shift(d, packed.data_ptr(), lat2d.data_ptr(), T, shift.data_ptr(), stream),
                          "cnf_film_shift")
            timing = getattr(self, "_timing", None)  # bench.py: CUDA events around the chain kernel alone
            if timing is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            _native.check(lib.cnf_forward(d, packed.data_ptr(), prec, coords_c.data_ptr(), frame_stride,
                                          shift.data_ptr(), out.data_ptr(), T, P, stash_ptr, stash_n, stream),
                          "cnf_forward")
            if timing is not None:
                e1.record()
                timing.append((e0, e1))
        return out, stash

    @torch.no_grad()
    def decode_into(self, coords, latents, out_ptrs, T_expected: Optional[int] = None) -> Tuple[int, int]:
        """Decode and store this call's ``(T, P, cout)`` block at every device address in ``out_ptrs`` (the fused
        all-gather: the caller passes its own gathered buffer and the peers' NVLink-mapped buffers, each offset to
        this rank's frame range; see ``distributed.FusedGatherDecoder``).  No autograd.  Returns ``(T, P)``."""
        import ctypes

        self._check_inputs(coords, latents)
        lib = _native.load()
        d = self._cdims()
        cin, L, H, nl, cout = self._dims_tuple
        coords_c, stride, lat2d, T, P, _ = canonicalize(coords, latents)
        if T_expected is not None and T != T_expected:
            raise ValueError(f"expected {T_expected} frames, got {T}")
        if not 1 <= len(out_ptrs) <= 8:
            raise ValueError(f"1..8 output targets, got {len(out_ptrs)}")
        if any(int(p) == 0 or int(p) % 4 for p in out_ptrs):
            raise ValueError("output targets must be non-NULL, 4-byte aligned device addresses")
        prec = self._precision_code()
        dev = lat2d.device
        packed = self._ensure_packed()
        shift = torch.empty((T, (nl + 1) * H), dtype=torch.float32, device=dev)
        ptrs = (ctypes.c_void_p * len(out_ptrs))(*[int(p) for p in out_ptrs])
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _native.check(lib.cnf_film_shift(d, packed.data_ptr(), lat2d.data_ptr(), T, shift.data_ptr(), stream),
                          "cnf_film_shift")
            timing = getattr(self, "_timing", None)
            if timing is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            _native.check(lib.cnf_forward_gather(d, packed.data_ptr(), prec, coords_c.data_ptr(), stride,
                                                 shift.data_ptr(), ptrs, len(out_ptrs), T, P, stream),
                          "cnf_forward_gather")
            if timing is not None:
                e1.record()
                timing.append((e0, e1))
        return T, P

    def _launch_backward(self, gout, stash, packed, prec, T, P) -> torch.Tensor:
        lib = _native.load()
        d = self._cdims()
        cin, L, H, nl, cout = self._dims_tuple
        dev = gout.device
        if gout.dtype != torch.float32 or tuple(gout.shape) != (T, P, cout):
            raise RuntimeError(f"unexpected output gradient {tuple(gout.shape)} {gout.dtype}")
        gshift = torch.empty((T, (nl + 1) * H), dtype=torch.float32, device=dev)
        glat = torch.empty((T, L), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _native.check(lib.cnf_backward(d, packed.data_ptr(), prec, gout.data_ptr(), stash.data_ptr(),
                                           stash.numel(), gshift.data_ptr(), T, P, stream), "cnf_backward")
            _native.check(lib.cnf_film_shift_backward(d, packed.data_ptr(), gshift.data_ptr(), T, glat.data_ptr(),
                                                      stream), "cnf_film_shift_backward")
        return glat


class SIRENAutodecoder_film_extra_in(SIRENAutodecoder_film):
    """Variant with a scalar extra coordinate channel prepended (cnf/nf_networks.py:503-508)."""

    def forward(self, coord, latents):
        coord = torch.concat([torch.ones_like(coord[0][..., :1]) * coord[1], coord[0]], dim=-1)
        return super().forward(coord, latents)
