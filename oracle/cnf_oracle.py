"""CPU oracle for the CNF decode hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain PyTorch (CPU, fp32 or fp64), the arithmetic of the
reference's FiLM-modulated SIREN auto-decoder so that the CUDA path can be
checked on a box where ``/root/reference`` does not exist.  Nothing under
``confild_b200/`` imports it; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may.

What it follows (paths relative to the reference checkout):

* forward            ConditionalNeuralField/cnf/nf_networks.py:480-495
* linear + bias      ConditionalNeuralField/cnf/components.py:64-76
                     (``matmul(input, W^T)`` then in-place ``+= bias.unsqueeze(-2)``)
* sine activation    ConditionalNeuralField/cnf/components.py:19-25 (``sin(w0 * x)``)
* w0 = 30            ConditionalNeuralField/cnf/initialization.py:5
* weight init        ConditionalNeuralField/cnf/initialization.py:117-132 and the
                     constructor order in nf_networks.py:465-476
* DPS gradient       ConditionalDiffusionGeneration/src/guided_diffusion/
                     condition_methods.py:28-33 (``autograd.grad(norm, latents)``)

Pinning: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md section 4), so the pins are generated from the live reference module
imported in the build container: ``tests/golden/make_golden.py`` writes the
fixtures and ``tests/test_oracle.py`` re-checks this restatement bit-for-bit
against them (and against the live module whenever ``/root/reference`` exists).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Callable, Dict, Optional, Tuple

import torch

DEFAULT_W0 = 30.0  # initialization.py:5

#: (cin, L, cout, nl, H) of the reference's training recipes
#: (ConditionalNeuralField/training_recipes/case{1,2,3,4}.yml)
CASE_SHAPES = {
    "case1": (2, 128, 3, 10, 128),
    "case2": (2, 256, 4, 10, 256),
    "case3": (2, 256, 2, 17, 256),
    "case4": (3, 384, 3, 15, 384),
}


def init_params(cin: int, L: int, cout: int, nl: int, H: int, seed: Optional[int] = 0,
                w0: float = DEFAULT_W0) -> "OrderedDict[str, torch.Tensor]":
    """Random-init parameters exactly as the reference constructor draws them.

    The RNG is consumed in the constructor's order (nf_networks.py:465-476):
    ``nn.Linear`` default init for every net1 layer then every net2 layer,
    then ``sine_init`` over net1 and net2 (weights only), then the first-layer
    re-draw for ``net1[0]`` and ``net2[0]``.  Biases keep the nn.Linear default.
    Returns a state dict with the reference's key names.
    """
    if seed is not None:
        torch.manual_seed(seed)
    net1 = [torch.nn.Linear(cin, H)] + [torch.nn.Linear(H, H) for _ in range(nl)] + [torch.nn.Linear(H, cout)]
    net2 = [torch.nn.Linear(L, H, bias=False) for _ in range(nl + 1)]
    with torch.no_grad():
        for lin in net1 + net2:  # sine_init, initialization.py:117-125
            bound = math.sqrt(6 / lin.weight.size(-1)) / w0
            lin.weight.uniform_(-bound, bound)
        for lin in (net1[0], net2[0]):  # first_layer_sine_init, initialization.py:127-132
            bound = 1 / lin.weight.size(-1)
            lin.weight.uniform_(-bound, bound)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for i, lin in enumerate(net1):
        sd[f"net1.{i}.weight"] = lin.weight.detach().clone()
        sd[f"net1.{i}.bias"] = lin.bias.detach().clone()
    for i, lin in enumerate(net2):
        sd[f"net2.{i}.weight"] = lin.weight.detach().clone()
    return sd


def dims_of(sd: Dict[str, torch.Tensor]) -> Tuple[int, int, int, int, int]:
    """(cin, L, cout, nl, H) recovered from a state dict."""
    n1 = sum(1 for k in sd if k.startswith("net1.") and k.endswith(".weight"))
    nl = n1 - 2
    H, cin = sd["net1.0.weight"].shape
    L = sd["net2.0.weight"].shape[1]
    cout = sd[f"net1.{nl + 1}.weight"].shape[0]
    return int(cin), int(L), int(cout), int(nl), int(H)


def _linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    # components.py:71-75: matmul against the transposed weight, then in-place bias add
    out = torch.matmul(x, weight.transpose(-1, -2))
    if bias is not None:
        out += bias.unsqueeze(-2)
    return out


def forward(sd: Dict[str, torch.Tensor], coords: torch.Tensor, latents: torch.Tensor,
            w0: float = DEFAULT_W0) -> torch.Tensor:
    """Decode: same op order as nf_networks.py:491-494 (PyTorch broadcasting)."""
    nl = dims_of(sd)[3]
    x = coords
    for i in range(nl + 1):
        x = _linear(x, sd[f"net1.{i}.weight"], sd[f"net1.{i}.bias"]) + _linear(latents, sd[f"net2.{i}.weight"], None)
        x = torch.sin(w0 * x)
    return _linear(x, sd[f"net1.{nl + 1}.weight"], sd[f"net1.{nl + 1}.bias"])


def to_dtype(sd: Dict[str, torch.Tensor], dtype: torch.dtype) -> "OrderedDict[str, torch.Tensor]":
    return OrderedDict((k, v.to(dtype)) for k, v in sd.items())


def sensor_loss(y: torch.Tensor, y_meas: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
    """DPS measurement distance: Frobenius norm of the (masked) residual
    (condition_methods.py:30-31; mask multiply as in measurements.py:91-97)."""
    diff = y_meas - y
    if mask is not None:
        diff = diff * mask
    return torch.linalg.norm(diff)


def grad_latents(sd: Dict[str, torch.Tensor], coords: torch.Tensor, latents: torch.Tensor,
                 loss_fn: Callable[[torch.Tensor], torch.Tensor], w0: float = DEFAULT_W0
                 ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(loss, y, dloss/dlatents) via autograd on the restated forward, the way
    condition_methods.py:32 asks for it."""
    lat = latents.detach().clone().requires_grad_(True)
    y = forward(sd, coords, lat, w0)
    loss = loss_fn(y)
    (g,) = torch.autograd.grad(loss, lat)
    return loss.detach(), y.detach(), g


def grad_latents_from_gout(sd: Dict[str, torch.Tensor], coords: torch.Tensor, latents: torch.Tensor,
                           gout: torch.Tensor, w0: float = DEFAULT_W0) -> torch.Tensor:
    """Vector-Jacobian product dL/dlatents for a given dL/dy (what the CUDA
    backward entry point computes)."""
    lat = latents.detach().clone().requires_grad_(True)
    y = forward(sd, coords, lat, w0)
    (g,) = torch.autograd.grad(y, lat, grad_outputs=gout)
    return g


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b||_2 / ||b||_2 in fp64 (b is the reference)."""
    a64, b64 = a.detach().double().cpu(), b.detach().double().cpu()
    return float(torch.linalg.norm(a64 - b64) / torch.linalg.norm(b64).clamp_min(1e-300))


def synthetic_inputs(cin: int, L: int, T: int, P: int, sigma: float = 0.1,
                     coord_seed: int = 1, latent_seed: int = 2) -> Tuple[torch.Tensor, torch.Tensor]:
    """coords ~ U(-1,1) of shape (P,cin) (generator seed 1), latents ~ N(0,sigma^2) of
    shape (T,L) (generator seed 2): SURVEY.md section 8(d) 'Inputs'."""
    gc = torch.Generator().manual_seed(coord_seed)
    gl = torch.Generator().manual_seed(latent_seed)
    coords = torch.rand(P, cin, generator=gc) * 2 - 1
    latents = torch.randn(T, L, generator=gl) * sigma
    return coords, latents


def load_reference_module():
    """Import the live reference class when the checkout is present (build
    container only); returns None on the GPU box."""
    import os
    import sys
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "ConditionalNeuralField")):
        return None
    for p in (ref, os.path.join(ref, "ConditionalNeuralField")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from ConditionalNeuralField.cnf.nf_networks import SIRENAutodecoder_film  # type: ignore
    return SIRENAutodecoder_film
