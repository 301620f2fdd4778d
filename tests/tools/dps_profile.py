"""Kernel table (torch profiler) of the DPS step of BASELINE config 4: python tests/tools/dps_profile.py case1 [skip|dense]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import confild_b200 as cb
from oracle import cnf_oracle as O
from torch.profiler import profile, ProfilerActivity
case = sys.argv[1] if len(sys.argv) > 1 else "case1"
mode = sys.argv[2] if len(sys.argv) > 2 else "skip"
dims = O.CASE_SHAPES[case]; sd = O.init_params(*dims, seed=0)
T, P, S = 64, 16384, 1000
coords, lat = O.synthetic_inputs(dims[0], dims[1], T, P)
m = cb.SIRENAutodecoder_film(*dims[:2], dims[2], dims[3], dims[4]); m.load_state_dict(sd); m = m.eval().cuda(); m.disable_gradient()
c, l0 = coords.cuda(), lat.cuda()
mask = torch.zeros(P, device="cuda"); mask[torch.randperm(P, device="cuda")[:S]] = 1.0
ym = torch.randn(T, P, dims[2], device="cuda") * 0.05 * mask[None, :, None]
def step():
    l = l0[:, None].detach().requires_grad_(True)
    n = cb.measurement_norm(m, c[None], l, ym, mask=mask, zero_row_skip=(mode == "skip"))
    return torch.autograd.grad(n, l)[0]
for _ in range(5): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=20, max_name_column_width=70))
