"""Would an 8-bit fixed-point cos stash keep dL/dlatent inside the 1e-2 contract?  float64 emulation of the DPS backward with
the stashed cosines rounded to fp16 (what the kernels store) and to int8 (round(127 cos)/127), several shapes.
Result: int8 1.4e-3 with thousands of rows, 7.5e-3 ... 8.7e-3 with 10 sensor rows -- no margin; not built."""
import sys, torch, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cnf_oracle as O
torch.manual_seed(0)
def run(case, T, P, sensors):
    cin, L, cout, nl, H = O.CASE_SHAPES[case]
    sd = O.init_params(cin, L, cout, nl, H, seed=0)
    coords, lat = O.synthetic_inputs(cin, L, T, P)
    w0 = 30.0
    # manual forward/backward in float64 with a hook to quantise cos
    W = [sd[f"net1.{i}.weight"].double() for i in range(nl + 2)]
    B = [sd[f"net1.{i}.bias"].double() for i in range(nl + 2)]
    V = [sd[f"net2.{i}.weight"].double() for i in range(nl + 1)]
    x = coords.double()[None].expand(T, P, cin)
    z = lat.double()
    def fwd_bwd(q):
        h = x; cosl = []
        for l in range(nl + 1):
            a = w0 * (h @ W[l].t() + B[l] + (z @ V[l].t())[:, None, :])
            cosl.append(q(torch.cos(a))); h = torch.sin(a)
        y = h @ W[nl + 1].t() + B[nl + 1]
        mask = torch.zeros(P, dtype=torch.float64); mask[torch.randperm(P, generator=torch.Generator().manual_seed(1))[:sensors]] = 1
        ym = torch.randn(T, P, cout, generator=torch.Generator().manual_seed(2)).double() * 0.3
        r = (ym - y) * mask[None, :, None]
        nrm = r.norm(); gy = -(r * mask[None, :, None]) / nrm
        d = gy @ W[nl + 1]            # dL/dh_last
        glat = torch.zeros_like(z)
        for l in range(nl, -1, -1):
            da = d * cosl[l] * w0      # dL/d(pre-activation argument / w0 part)
            glat += da.sum(1) @ V[l]
            if l > 0: d = da @ W[l]
        return glat
    g_ref = fwd_bwd(lambda c: c)
    def rel(g): return float((g - g_ref).norm() / g_ref.norm())
    g16 = fwd_bwd(lambda c: c.half().double())
    g8 = fwd_bwd(lambda c: torch.round(c * 127) / 127)
    g8s = fwd_bwd(lambda c: (torch.floor(c * 127 + torch.rand_like(c))) / 127)  # stochastic rounding
    print(f"{case} T={T} P={P} sensors={sensors}: fp16 stash {rel(g16):.2e}  int8 {rel(g8):.2e}  int8 stochastic {rel(g8s):.2e}")
run("case1", 4, 2000, 2000)
run("case1", 4, 2000, 100)
run("case1", 8, 10, 10)
run("case4", 2, 1000, 1000)
run("case4", 8, 10, 10)
run("case3", 2, 1000, 1000)
