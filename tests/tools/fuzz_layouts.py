"""Non-contiguous / broadcast input layouts of forward(coords, latents): results must equal the contiguous call bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import confild_b200 as cb
from oracle import cnf_oracle as O
for case, T, P in (("case1", 6, 300), ("case4", 4, 200)):
    dims = O.CASE_SHAPES[case]; sd = O.init_params(*dims, seed=0)
    cin, L, cout, nl, H = dims
    coords, lat = O.synthetic_inputs(cin, L, T, P)
    m = cb.SIRENAutodecoder_film(cin, L, cout, nl, H); m.load_state_dict(sd); m = m.eval().cuda(); m.disable_gradient()
    c, l = coords.cuda(), lat.cuda()
    with torch.no_grad():
        base = m(c[None], l[:, None])
        variants = {
            "coords strided rows": (torch.stack([c, c + 1], 1).reshape(2 * P, cin)[::2][None], l[:, None]),
            "coords strided cols": (torch.cat([c, c], 1)[:, :cin][None], l[:, None]),
            "coords per frame (expanded)": (c[None].expand(T, P, cin), l[:, None]),
            "coords per frame (materialised)": (c[None].expand(T, P, cin).contiguous(), l[:, None]),
            "latents strided": (c[None], torch.stack([l, l * 2], 1).reshape(2 * T, L)[::2][:, None]),
            "latents transposed storage": (c[None], l.t().contiguous().t()[:, None]),
            "latents 4-d": (c.reshape(1, 1, P, cin), l.reshape(T, 1, 1, L)),
            "grid coords": (c.reshape(1, P // 10, 10, cin), l.reshape(T, 1, 1, L)),
        }
        for name, (cc, ll) in variants.items():
            y = m(cc, ll)
            same = torch.equal(y.reshape(base.shape), base)
            print(f"{'ok  ' if same else 'FAIL'} {case}: {name} -> {tuple(y.shape)}", flush=True)
            if not same: sys.exit(1)
    # gradient through a strided latent view
    big = torch.stack([l, l * 2], 1).reshape(2 * T, L).clone().requires_grad_(True)
    y = m(c[None], big[::2][:, None])
    (g,) = torch.autograd.grad(y.sum(), big)
    l2 = l.clone().requires_grad_(True)
    (g2,) = torch.autograd.grad(m(c[None], l2[:, None]).sum(), l2)
    ok = float((g[::2] - g2).norm() / g2.norm()) <= 1e-5 and float(g[1::2].abs().max()) == 0.0
    print(f"{'ok  ' if ok else 'FAIL'} {case}: gradient through a strided latent view", flush=True)
    if not ok: sys.exit(1)
print("all ok")
