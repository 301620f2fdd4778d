"""Times the two halves of a DPS step separately: forward with cos stash, backward to dL/dlatent.
    python tests/tools/dps_split.py case1 64 16384"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import confild_b200 as cb
from oracle import cnf_oracle as O
case, T, P = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dims = O.CASE_SHAPES[case]; sd = O.init_params(*dims, seed=0)
coords, lat = O.synthetic_inputs(dims[0], dims[1], T, P)
m = cb.SIRENAutodecoder_film(*dims[:2], dims[2], dims[3], dims[4], precision="bf16x3"); m.load_state_dict(sd); m = m.eval().cuda()
c = coords.cuda()[None]
gout = torch.randn(T, P, dims[2], device="cuda")
ev = lambda: torch.cuda.Event(enable_timing=True)
tf = tb = 0.0
N = 8
for it in range(N + 3):
    l = lat.cuda()[:, None].requires_grad_(True)
    e0, e1, e2 = ev(), ev(), ev()
    e0.record()
    y = m(c, l)
    e1.record()
    (g,) = torch.autograd.grad(y, l, grad_outputs=gout)
    e2.record()
    torch.cuda.synchronize()
    if it >= 3:
        tf += e0.elapsed_time(e1); tb += e1.elapsed_time(e2)
print(f"{case} T={T} P={P}: forward+stash {tf / N:.3f} ms, backward {tb / N:.3f} ms", flush=True)
