"""Shared helpers for the tests: golden fixture loading and input reconstruction."""
import os

import numpy as np
import torch

from oracle import cnf_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_NAMES = ["tiny_shared", "case1_shared", "case1_grid", "case1_perframe", "case2_shared", "case4_shared",
                "case1_sigma1", "case3_shared"]
EXTRA_IN_NAMES = ["case1_extra_in_scalar", "case1_extra_in_points"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["dims"] = tuple(int(v) for v in g["dims"])
    g["layout"] = str(g["layout"])
    g["seed"] = int(g["seed"])
    return g


def golden_inputs(g):
    """(state_dict, coords_in, latents_in) shaped the way the fixture's generator fed the reference."""
    cin, L, cout, nl, H = g["dims"]
    sd = O.init_params(cin, L, cout, nl, H, seed=g["seed"])
    coords = torch.from_numpy(g["coords"])
    lat = torch.from_numpy(g["latents"])
    if g["layout"] == "shared":
        return sd, coords[None], lat[:, None]
    if g["layout"] == "grid":
        return sd, coords, lat[:, None, None]
    if g["layout"] == "perframe":
        return sd, coords, lat[:, None]
    raise ValueError(g["layout"])


def extra_in_inputs(g):
    """(state_dict, (coords, extra), latents, concatenated coords) of an ``_extra_in`` fixture: the reference prepends
    the extra channel to the coordinates (nf_networks.py:503-508)."""
    cin, L, cout, nl, H = g["dims"]
    sd = O.init_params(cin, L, cout, nl, H, seed=g["seed"])
    coords = torch.from_numpy(g["coords"])[None]
    extra = torch.from_numpy(g["extra"])
    lat = torch.from_numpy(g["latents"])[:, None]
    cat = torch.concat([torch.ones_like(coords[..., :1]) * extra, coords], dim=-1)
    return sd, (coords, extra), lat, cat


def sha_state(sd):
    import hashlib

    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().contiguous().numpy().tobytes())
    return h.hexdigest()
