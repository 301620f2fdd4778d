"""Randomised fuzz of cnf_group_norm_nhwc_bf16 (both paths) against F.group_norm in fp32: python tests/tools/fuzz_gn.py [n] [seed]"""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
from confild_b200 import _native
from confild_b200.latent_sampler import group_norm_nhwc
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
for it in range(n_cases):
    G = rng.choice([1, 2, 4, 8, 16, 32, 64])
    C = G * rng.choice([1, 2, 3, 4, 5, 8, 12, 16, 28, 32])
    if C % 8 or C > 2048:
        continue
    N, H, W = rng.randint(1, 5), rng.choice([1, 2, 3, 8, 17, 32, 64]), rng.choice([1, 2, 5, 8, 16, 33, 64])
    silu, with_add = rng.random() < 0.5, rng.random() < 0.5
    g = torch.Generator(device="cuda").manual_seed(it)
    x = (torch.randn(N, C, H, W, device="cuda", generator=g) * rng.choice([0.1, 1.0, 5.0]) + rng.choice([0.0, 2.0])).to(torch.bfloat16)
    x = x.contiguous(memory_format=torch.channels_last)
    w, b = torch.randn(C, device="cuda", generator=g), torch.randn(C, device="cuda", generator=g)
    add = torch.randn(N, C, device="cuda", generator=g) if with_add else None
    xin = x.float() + (add[:, :, None, None] if with_add else 0.0)
    if H * W * (C // G) == 1:
        continue
    want = F.group_norm(xin, G, w, b, 1e-5)
    if silu: want = F.silu(want)
    tol = 2.0 ** -7 * want.abs() + 3e-2
    for knob in (0, 1 << 20, _native.KNOB_DEFAULTS["CNF_GN_CLUSTER"]):
        _native.set_knob("CNF_GN_CLUSTER", knob)
        got = group_norm_nhwc(x, w, b, G, 1e-5, add=add, silu=silu)
        if not bool(((got.float() - want).abs() <= tol).all()):
            print(f"FAIL N={N} C={C} G={G} H={H} W={W} silu={silu} add={with_add} knob={knob}: max err {float((got.float() - want).abs().max()):.3e}")
            sys.exit(1)
    print(f"ok   N={N} C={C} G={G} H={H} W={W} silu={silu} add={with_add}", flush=True)
_native.set_knob("CNF_GN_CLUSTER", _native.KNOB_DEFAULTS["CNF_GN_CLUSTER"])
print("all ok")
