"""One DPS-style step (forward with stash, sensor loss, backward to latents) for profiling."""
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
mask = torch.zeros(P, 1, device="cuda"); mask[torch.randperm(P, device="cuda")[:1000]] = 1.0
y_meas = torch.randn(T, P, dims[2], device="cuda") * 0.05
for it in range(4):
    l = lat.cuda()[:, None].requires_grad_(True)
    y = m(c, l)
    loss = torch.linalg.norm((y_meas - y) * mask)
    (g,) = torch.autograd.grad(loss, l)
torch.cuda.synchronize()
print("ok", float(loss))
