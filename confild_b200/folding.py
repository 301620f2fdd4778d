"""Fold the affine steps around the decoder into its parameters.

The reference wraps every decode in per-channel affine maps that each cost extra element-wise passes over
``(T, P, .)`` tensors:

* ``x_normalizer.normalize(coords)``   (cnf/utils/normalize.py:100-103, called at inference_function.py:42, train.py:272)
* ``y_normalizer.denormalize(out)``    (normalize.py:112-120, called at inference_function.py:41-43, train.py:279)
* the DPS operators' latent ``_unnorm``  ``(z + 1) * (max - min) / 2 + min``  (guided_diffusion/measurements.py:88-89, 219-220)

All three are affine, so they fold exactly into the first layer, the head and the FiLM matrices: the CUDA kernels then
read physical coordinates / normalised latents and write physical fields with no extra pass.  ``fold_normalizers``
returns a NEW module (same class, folded parameters); the original is left untouched.
"""
from __future__ import annotations

import copy
from typing import Optional, Tuple

import torch


def _affine_of_normalizer(norm, inverse: bool, n: int, device, dtype) -> Tuple[torch.Tensor, torch.Tensor]:
    """(a, c) with  normalize(x) = a*x + c   (or, for ``inverse``, denormalize(y) = a*y + c) per channel.

    Supports the reference's methods '-11', '01', 'ms' and 'none' (normalize.py:100-120); ``params`` are
    (max, min) or (mean, std), any shape broadcastable to ``n`` channels.
    """
    method = getattr(norm, "method", "-11")
    one = torch.ones(n, device=device, dtype=dtype)
    if method == "none" or norm is None:
        return one, torch.zeros_like(one)
    p0, p1 = (torch.as_tensor(p, dtype=dtype, device=device).reshape(-1) * one for p in norm.params)
    if method == "-11":       # x_n = (x - min)/(max - min)*2 - 1
        a, c = 2.0 / (p0 - p1), -2.0 * p1 / (p0 - p1) - 1.0
    elif method == "01":      # x_n = (x - min)/(max - min)
        a, c = 1.0 / (p0 - p1), -p1 / (p0 - p1)
    elif method == "ms":      # x_n = (x - mean)/std
        a, c = 1.0 / p1, -p0 / p1
    else:
        raise ValueError(f"unknown normalisation method {method!r}")
    if inverse:               # y = (y_n - c)/a
        return 1.0 / a, -c / a
    return a, c


@torch.no_grad()
def fold_normalizers(model, x_normalizer=None, y_normalizer=None,
                     latent_affine: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
    """Return a copy of ``model`` whose ``forward(coords_physical, latents)`` equals

        y_normalizer.denormalize( model( x_normalizer.normalize(coords_physical), a_z * latents + c_z ) )

    ``latent_affine = (a_z, c_z)`` (per latent channel, e.g. ``((max-min)/2, (max+min)/2)`` for the DPS ``_unnorm``).
    Gradients with respect to the (un-affined) latents follow by the chain rule inside the folded FiLM matrices.
    """
    folded = copy.deepcopy(model)
    w0l, b0l = folded.net1[0].weight, folded.net1[0].bias
    dev, dt = w0l.device, w0l.dtype
    if x_normalizer is not None:
        a, c = _affine_of_normalizer(x_normalizer, False, w0l.shape[1], dev, dt)
        b0l.add_(w0l @ c)          # W (a x + c) + b = (W diag a) x + (b + W c)
        w0l.mul_(a[None, :])
    if y_normalizer is not None:
        wo, bo = folded.net1[-1].weight, folded.net1[-1].bias
        a, c = _affine_of_normalizer(y_normalizer, True, wo.shape[0], dev, dt)
        wo.mul_(a[:, None])        # a (W h + b) + c
        bo.mul_(a).add_(c)
    if latent_affine is not None:
        a_z, c_z = (torch.as_tensor(v, dtype=dt, device=dev).reshape(-1) for v in latent_affine)
        for i, film in enumerate(folded.net2):  # V (a z + c) = (V diag a) z + V c ; the constant joins the layer bias
            folded.net1[i].bias.add_(film.weight @ (c_z * torch.ones(film.weight.shape[1], device=dev, dtype=dt)))
            film.weight.mul_(a_z[None, :])
    if hasattr(folded, "_packed"):
        folded._packed, folded._packed_key = None, None
    return folded
