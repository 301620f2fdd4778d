"""Generate the golden fixtures in this directory from the LIVE reference module.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference ships no golden vectors for the CNF decode path, so these are made
by importing ``ConditionalNeuralField.cnf.nf_networks.SIRENAutodecoder_film``
(nf_networks.py:443-501) unmodified, constructing it under ``torch.manual_seed``
and calling its own ``forward`` / ``torch.autograd.grad`` on CPU in fp32
(condition_methods.py:28-33 for the gradient).  Nothing here uses the oracle
restatement or the CUDA path.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
for p in (REF, os.path.join(REF, "ConditionalNeuralField")):
    sys.path.insert(0, p)
from ConditionalNeuralField.cnf.nf_networks import SIRENAutodecoder_film, SIRENAutodecoder_film_extra_in  # noqa: E402


def sha_state(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().contiguous().numpy().tobytes())
    return h.hexdigest()


def make(name, dims, T, P, layout, seed=0, sigma=0.1, sensors=None, store_weights=False):
    cin, L, cout, nl, H = dims
    torch.manual_seed(seed)
    model = SIRENAutodecoder_film(cin, L, cout, nl, H).eval()
    gc = torch.Generator().manual_seed(1)
    gl = torch.Generator().manual_seed(2)
    gm = torch.Generator().manual_seed(3)
    lat = torch.randn(T, L, generator=gl) * sigma
    if layout == "shared":            # pass_through_model_batch: coords (1,P,cin), latents (T,1,L)
        coords = torch.rand(P, cin, generator=gc) * 2 - 1
        c_in, l_in = coords[None], lat[:, None]
    elif layout == "grid":            # CNF_inference.predict: coords (h,w,cin), latents (T,1,1,L)
        h, w = P
        coords = torch.rand(h, w, cin, generator=gc) * 2 - 1
        c_in, l_in = coords, lat[:, None, None]
    elif layout == "perframe":        # training loop: coords (T,P,cin), latents (T,1,L)
        coords = torch.rand(T, P, cin, generator=gc) * 2 - 1
        c_in, l_in = coords, lat[:, None]
    else:
        raise ValueError(layout)
    l_in = l_in.clone().requires_grad_(True)
    y = model(c_in, l_in)
    out = {
        "dims": np.array(dims, dtype=np.int64), "seed": np.array(seed), "layout": np.array(layout),
        "coords": coords.numpy(), "latents": lat.numpy(), "y": y.detach().numpy(),
        "weights_sha256": np.array(sha_state(model.state_dict())),
    }
    if sensors is not None:
        n_pts = int(np.prod(y.shape[1:-1]))
        idx = torch.randperm(n_pts, generator=gm)[:sensors]
        mask = torch.zeros(n_pts)
        mask[idx] = 1.0
        mask = mask.reshape(*y.shape[1:-1], 1)
        y_meas = torch.randn(y.shape, generator=gm) * 0.05
        norm = torch.linalg.norm((y_meas - y) * mask)
        (g,) = torch.autograd.grad(norm, l_in)
        out.update(mask=mask.numpy(), y_meas=y_meas.numpy(), loss=norm.detach().numpy(),
                   dlatents=g.reshape(T, L).numpy())
    if store_weights:
        for k, v in model.state_dict().items():
            out["w:" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: getattr(v, "shape", None) for k, v in out.items() if not k.startswith("w:")})


def make_extra_in(name, dims, T, P, per_point, seed=0, sigma=0.1):
    """SIRENAutodecoder_film_extra_in (nf_networks.py:503-508): ``coord = (coords, extra)``; ``dims[0]`` counts the
    extra channel.  ``extra`` is a scalar, or a per-point ``(1, P, 1)`` tensor."""
    cin, L, cout, nl, H = dims
    torch.manual_seed(seed)
    model = SIRENAutodecoder_film_extra_in(cin, L, cout, nl, H).eval()
    gc = torch.Generator().manual_seed(1)
    gl = torch.Generator().manual_seed(2)
    lat = torch.randn(T, L, generator=gl) * sigma
    coords = torch.rand(P, cin - 1, generator=gc) * 2 - 1
    extra = (torch.rand(1, P, 1, generator=gc) * 2 - 1) if per_point else torch.tensor(0.37)
    l_in = lat[:, None].clone().requires_grad_(True)
    y = model((coords[None], extra), l_in)
    gout = torch.randn(y.shape, generator=torch.Generator().manual_seed(7))
    (g,) = torch.autograd.grad(y, l_in, grad_outputs=gout)
    out = {"dims": np.array(dims, dtype=np.int64), "seed": np.array(seed), "layout": np.array("extra_in"),
           "coords": coords.numpy(), "extra": extra.numpy(), "latents": lat.numpy(), "y": y.detach().numpy(),
           "gout": gout.numpy(), "dlatents": g.reshape(T, L).numpy(),
           "weights_sha256": np.array(sha_state(model.state_dict()))}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    torch.set_num_threads(1)  # one thread: fixed reduction order inside the CPU GEMMs
    only = set(sys.argv[1:])  # optional fixture names: (re)generate just those
    if only:
        _make, _make_extra = make, make_extra_in
        make = lambda name, *a, **k: _make(name, *a, **k) if name in only else None  # noqa: E731
        make_extra_in = lambda name, *a, **k: _make_extra(name, *a, **k) if name in only else None  # noqa: E731
    make("tiny_shared", (2, 32, 3, 2, 64), T=3, P=70, layout="shared", sensors=20, store_weights=True)
    make("case1_shared", (2, 128, 3, 10, 128), T=5, P=300, layout="shared", sensors=100)
    make("case1_grid", (2, 128, 3, 10, 128), T=2, P=(9, 15), layout="grid")
    make("case1_perframe", (2, 128, 3, 10, 128), T=3, P=131, layout="perframe", sensors=50)
    make("case2_shared", (2, 256, 4, 10, 256), T=2, P=140, layout="shared", sensors=30)
    make("case4_shared", (3, 384, 3, 15, 384), T=2, P=130, layout="shared", sensors=40)
    make("case1_sigma1", (2, 128, 3, 10, 128), T=4, P=129, layout="shared", sigma=1.0, sensors=64)
    # round 2: the deepest chain (17 hidden layers) with its DPS gradient, and the _extra_in variant
    make("case3_shared", (2, 256, 2, 17, 256), T=3, P=150, layout="shared", sensors=60)
    make_extra_in("case1_extra_in_scalar", (3, 128, 3, 10, 128), T=3, P=140, per_point=False)
    make_extra_in("case1_extra_in_points", (3, 128, 3, 10, 128), T=2, P=131, per_point=True)
