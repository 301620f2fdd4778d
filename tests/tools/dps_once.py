"""A few DPS steps of BASELINE config 4 (profiling target):
    python tests/tools/dps_once.py case1 [skip|dense]
64 latents x 16,384 points, 1,000 random sensors, fused measurement norm + gradient; `dense` stashes every row,
`skip` (default policy) only the sensor rows.  Prints the CUDA-event time per step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import confild_b200 as cb
from oracle import cnf_oracle as O
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
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): g = step()
e1.record(); torch.cuda.synchronize()
import time
ts = []
for _ in range(12):
    torch.cuda.synchronize(); t0 = time.perf_counter(); step(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print("wall ms per step:", " ".join(f"{t:.2f}" for t in ts))
graphed = cb.GraphedMeasurementNorm(m, c[None], l0[:, None], ym, mask=mask, zero_row_skip=(mode == "skip"))
def gstep():
    l = l0[:, None].detach().requires_grad_(True)
    return torch.autograd.grad(graphed(l), l)[0]
for _ in range(3): gstep()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50): gg = gstep()
torch.cuda.synchronize()
print(f"graphed step: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms wall; rel diff of the gradient vs eager {float((gg - g).norm() / g.norm()):.2e}")
print(f"{case} DPS step ({mode}, {m.resolved_precision}): {e0.elapsed_time(e1) / 5:.3f} ms; |g| = {float(g.norm()):.4e}")
