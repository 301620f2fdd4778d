"""Randomised LARGE shapes (many tiles per CTA, odd tile counts, ragged last tiles): tensor-core precisions against the fp32
CUDA-core path of the same module (itself checked against the oracle elsewhere): python tests/tools/fuzz_large.py [n] [seed]"""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import confild_b200 as cb
from oracle import cnf_oracle as O
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
FWD = {"bf16x3": 1e-4, "f16f8": 1e-3, "fp16": 5e-3}
for it in range(n_cases):
    H = rng.choice([128, 128, 256, 384])
    cin, cout = rng.randint(1, 4), rng.randint(1, 4)
    L, nl = rng.choice([16, 128]), rng.randint(1, 6)
    prec = rng.choice(["bf16x3", "f16f8", "fp16", "f16f8"])
    T = rng.choice([7, 64, 301, 1000])
    P = rng.choice([129, 513, 1000, 4097, 20000])
    if T * P > 6_000_000: T = max(1, 6_000_000 // P)
    dims = (cin, L, cout, nl, H)
    sd = O.init_params(*dims, seed=it)
    coords, lat = O.synthetic_inputs(cin, L, T, P)
    ms = {}
    for p in (prec, "fp32"):
        m = cb.SIRENAutodecoder_film(cin, L, cout, nl, H, precision=p)
        m.load_state_dict(sd); ms[p] = m.eval().cuda(); ms[p].disable_gradient()
    c = coords.cuda()[None]
    gout = torch.randn(T, P, cout, device="cuda", generator=torch.Generator(device="cuda").manual_seed(it))
    res = {}
    for p, m in ms.items():
        l = lat.cuda()[:, None].requires_grad_(True)
        y = m(c, l)
        (g,) = torch.autograd.grad(y, l, grad_outputs=gout)
        res[p] = (y.detach(), g)
    ef = float((res[prec][0] - res["fp32"][0]).norm() / res["fp32"][0].norm())
    eg = float((res[prec][1] - res["fp32"][1]).norm() / res["fp32"][1].norm())
    tiles = T * ((P + 127) // 128)
    ok = ef <= FWD[prec] and eg <= 1e-2
    print(f"{'ok  ' if ok else 'FAIL'} dims={dims} T={T} P={P} tiles={tiles} {prec}: fwd {ef:.2e} grad {eg:.2e}", flush=True)
    if not ok: sys.exit(1)
    del ms, res
    torch.cuda.empty_cache()
print("all ok")
