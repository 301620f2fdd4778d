"""Randomised fuzz of the host-side API around the kernels (default precision = auto): any width (tensor-core shapes or the
fp32 fallback), zero hidden layers, the decoder() / pass_through_model_batch drivers with normalisers, fold_normalizers,
the _extra_in variant -- against the oracle / the unfused composition."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import confild_b200 as cb
from oracle import cnf_oracle as O
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)


class Norm:
    def __init__(self, method, p0, p1): self.method, self.params = method, (torch.tensor(p0), torch.tensor(p1))
    def normalize(self, x):
        a, b = (p.to(x.device) for p in self.params)
        if self.method == "-11": return (x - b) / (a - b) * 2 - 1
        if self.method == "01": return (x - b) / (a - b)
        return (x - a) / b
    def denormalize(self, y):
        a, b = (p.to(y.device) for p in self.params)
        if self.method == "-11": return (y + 1) / 2 * (a - b) + b
        if self.method == "01": return y * (a - b) + b
        return y * b + a


for it in range(n_cases):
    H = rng.choice([32, 64, 100, 128, 128, 192, 256, 384])
    cin, cout = rng.randint(1, 4), rng.randint(1, 4)
    L, nl = rng.choice([4, 16, 64]), rng.randint(0, 3)
    T, P = rng.choice([1, 4, 9, 33]), rng.choice([3, 64, 129, 500])
    dims = (cin, L, cout, nl, H)
    sd = O.init_params(*dims, seed=it)
    coords, lat = O.synthetic_inputs(cin, L, T, P)
    m = cb.SIRENAutodecoder_film(cin, L, cout, nl, H)
    m.load_state_dict(sd); m = m.eval().cuda(); m.disable_gradient()
    prec = m.resolved_precision
    tol = {"fp32": 2e-5, "f16f8": 1e-3}[prec]
    want = O.forward(sd, coords[None], lat[:, None])
    gout = torch.randn(T, P, cout, generator=torch.Generator().manual_seed(it))
    gwant = O.grad_latents_from_gout(sd, coords[None], lat[:, None], gout)
    l = lat.cuda()[:, None].requires_grad_(True)
    y = m(coords.cuda()[None], l)
    (g,) = torch.autograd.grad(y, l, grad_outputs=gout.cuda())
    ef, eg = O.rel_l2(y, want), O.rel_l2(g, gwant)
    # drivers with normalisers
    meth = rng.choice(["-11", "01", "ms"])
    xn = Norm(meth, [1.5] * cin, [-0.5 if meth != "ms" else 0.7] * cin)
    yn = Norm(rng.choice(["-11", "01", "ms"]), [2.0, 1.5, 1.0, 0.5][:cout], [-1.0, 0.8, 0.25, -0.5][:cout] if True else None)
    if yn.method == "ms": yn.params = (yn.params[0], yn.params[1].abs() + 0.1)
    phys = coords * 0.9 + 0.2
    want_d = yn.denormalize(O.forward(sd, xn.normalize(phys)[None], lat[:, None]))
    got_d = cb.decoder(phys, lat, m, xn, yn, rng.choice([1, 4, 64]), "cuda")
    got_p = cb.pass_through_model_batch(phys, lat.cuda(), m, xn, yn, rng.choice([2, 64]), "cuda")
    folded = cb.fold_normalizers(m, xn, yn)
    got_f = folded(phys.cuda()[None], lat.cuda()[:, None])
    ed, ep, eff = O.rel_l2(got_d, want_d), O.rel_l2(got_p, want_d), O.rel_l2(got_f, want_d)
    ok = ef <= tol and eg <= 1e-2 and max(ed, ep, eff) <= 3 * tol
    print(f"{'ok  ' if ok else 'FAIL'} dims={dims} T={T} P={P} {prec}: fwd {ef:.1e} grad {eg:.1e} decoder {ed:.1e} pass_through {ep:.1e} folded {eff:.1e}", flush=True)
    if not ok: sys.exit(1)
print("all ok")
