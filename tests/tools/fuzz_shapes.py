"""Randomised shape fuzz of the CUDA path against the oracle (forward and dL/dlatent): python tests/tools/fuzz_shapes.py [n] [seed]"""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import confild_b200 as cb
from oracle import cnf_oracle as O
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
FWD = {"bf16x3": 1e-4, "f16f8": 1e-3, "fp16": 4e-3, "fp32": 2e-5}
worst = {}
for it in range(n_cases):
    H = rng.choice([128, 128, 256, 384])
    cin, cout = rng.randint(1, 4), rng.randint(1, 4)
    L = rng.choice([4, 16, 64, 128])
    nl = rng.randint(1, 5)
    prec = rng.choice(["bf16x3", "f16f8", "fp16", "f16f8"])
    T = rng.choice([1, 2, 3, 5, 17, 40])
    P = rng.choice([1, 7, 10, 127, 128, 129, 300, 1000, 2049])
    dims = (cin, L, cout, nl, H)
    sd = O.init_params(*dims, seed=it)
    coords, lat = O.synthetic_inputs(cin, L, T, P)
    gout = torch.randn(T, P, cout, generator=torch.Generator().manual_seed(it))
    want = O.forward(sd, coords[None], lat[:, None])
    gwant = O.grad_latents_from_gout(sd, coords[None], lat[:, None], gout)
    m = cb.SIRENAutodecoder_film(cin, L, cout, nl, H, precision=prec)
    m.load_state_dict(sd); m = m.eval().cuda(); m.disable_gradient()
    l = lat.cuda()[:, None].requires_grad_(True)
    y = m(coords.cuda()[None], l)
    (g,) = torch.autograd.grad(y, l, grad_outputs=gout.cuda())
    with torch.no_grad():
        y2 = m(coords.cuda()[None], lat.cuda()[:, None])
    ef, eg = O.rel_l2(y, want), O.rel_l2(g, gwant)
    ok = ef <= FWD[prec] and eg <= 1e-2 and torch.equal(y2, y.detach())
    worst[prec] = max(worst.get(prec, 0.0), ef)
    print(f"{'ok  ' if ok else 'FAIL'} dims={dims} T={T} P={P} {prec}: fwd {ef:.2e} grad {eg:.2e}", flush=True)
    if not ok:
        sys.exit(1)
print("worst forward error per precision:", {k: f"{v:.2e}" for k, v in worst.items()})
