"""Randomised fuzz of measurement_norm (all its paths: dense, zero-row skip, masked rows not decoded, CUDA graph) against
the unfused formulation through the drop-in module + PyTorch autograd: python tests/tools/fuzz_dps.py [n] [seed]"""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import confild_b200 as cb
from oracle import cnf_oracle as O
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)


class Norm11:
    method = "-11"
    def __init__(self, hi, lo): self.params = (torch.tensor(hi), torch.tensor(lo))
    def denormalize(self, y):
        hi, lo = (p.to(y.device) for p in self.params)
        return (y + 1) / 2 * (hi - lo) + lo


for it in range(n_cases):
    H = rng.choice([128, 128, 256, 384])
    cin, cout = rng.randint(1, 3), rng.randint(1, 4)
    L, nl = rng.choice([8, 32, 128]), rng.randint(1, 4)
    prec = rng.choice(["bf16x3", "f16f8"])
    T, P = rng.choice([1, 3, 16, 40]), rng.choice([5, 10, 130, 700, 3000])
    dims = (cin, L, cout, nl, H)
    sd = O.init_params(*dims, seed=it)
    coords, lat = O.synthetic_inputs(cin, L, T, P)
    m = cb.SIRENAutodecoder_film(cin, L, cout, nl, H, precision=prec)
    m.load_state_dict(sd); m = m.eval().cuda(); m.disable_gradient()
    g = torch.Generator().manual_seed(it)
    kind = rng.choice(["none", "sparse", "sparse", "dense1", "pc", "full"])
    if kind == "none": mask = None
    elif kind == "sparse":
        mask = torch.zeros(P); mask[torch.randperm(P, generator=g)[:max(1, P // rng.choice([3, 8, 20]))]] = rng.choice([1.0, 0.5])
    elif kind == "dense1": mask = torch.rand(P, generator=g) + 0.1
    elif kind == "pc": mask = torch.rand(P, cout, generator=g)
    else: mask = torch.rand(T, P, cout, generator=g)
    yn = rng.choice([None, Norm11([2.0, 1.5, 1.0, 0.5][:cout], [-1.0, -1.5, -0.25, -0.5][:cout])])
    masked_meas = rng.choice([False, True])
    ym = (torch.randn(T, P, cout, generator=g) * 0.3).cuda()
    c = coords.cuda()[None]
    mk = None if mask is None else mask.cuda()

    def ref():
        l = lat.cuda()[:, None].requires_grad_(True)
        y = m(c, l)
        yp = y if yn is None else yn.denormalize(y)
        if mk is None: r = ym - yp
        else:
            mm = mk.reshape(1, P, 1) if mk.numel() == P else mk
            r = (ym - yp) * mm if masked_meas else ym - mm * yp
        n = torch.linalg.norm(r)
        return float(n), torch.autograd.grad(n, l)[0]

    n_ref, g_ref = ref()
    variants = {"default": {}, "no_fwd_skip": {"skip_masked_decode": False}, "dense": {"zero_row_skip": False},
                "field": {"return_field": True}}
    for name, kw in variants.items():
        l = lat.cuda()[:, None].requires_grad_(True)
        out = cb.measurement_norm(m, c, l, ym, mask=mk, y_normalizer=yn, mask_measurement=masked_meas, **kw)
        n = out[0] if isinstance(out, tuple) else out
        (gg,) = torch.autograd.grad(n, l)
        en = abs(float(n) - n_ref) / max(n_ref, 1e-12)
        eg = float((gg - g_ref).norm() / g_ref.norm().clamp_min(1e-20))
        if not (en <= 1e-5 and eg <= 2e-3):
            print(f"FAIL {name} dims={dims} T={T} P={P} {prec} mask={kind} yn={yn is not None} mm={masked_meas}: norm {en:.2e} grad {eg:.2e}")
            sys.exit(1)
    graphed = cb.GraphedMeasurementNorm(m, c, lat.cuda()[:, None], ym, mask=mk, y_normalizer=yn, mask_measurement=masked_meas)
    l = lat.cuda()[:, None].requires_grad_(True)
    n = graphed(l); (gg,) = torch.autograd.grad(n, l)
    en = abs(float(n) - n_ref) / max(n_ref, 1e-12); eg = float((gg - g_ref).norm() / g_ref.norm().clamp_min(1e-20))
    ok = en <= 1e-5 and eg <= 2e-3
    print(f"{'ok  ' if ok else 'FAIL'} dims={dims} T={T} P={P} {prec} mask={kind} yn={yn is not None} mm={masked_meas}: norm {en:.1e} grad {eg:.1e}", flush=True)
    if not ok: sys.exit(1)
print("all ok")
