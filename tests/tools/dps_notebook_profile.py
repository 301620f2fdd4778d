"""Where the DPS step at the notebook's literal shape (case4, 384 frames x 10 sensor points) spends its time:
torch-profiler kernel table + CUDA-event time of the eager step and of a CUDA-graph replay of the same step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import confild_b200 as cb
from oracle import cnf_oracle as O
dims = O.CASE_SHAPES["case4"]; sd = O.init_params(*dims, seed=0)
T, P = 384, 10
coords, lat = O.synthetic_inputs(dims[0], dims[1], T, P)
m = cb.SIRENAutodecoder_film(*dims[:2], dims[2], dims[3], dims[4]); m.load_state_dict(sd); m = m.eval().cuda(); m.disable_gradient()
c, l0 = coords.cuda(), lat.cuda()
ym = torch.randn(T, P, dims[2], device="cuda") * 0.05
lat_static = l0[:, None].clone().requires_grad_(True)
def step():
    n = cb.measurement_norm(m, c[None], lat_static, ym)
    return n, torch.autograd.grad(n, lat_static)[0]
for _ in range(5): step()
torch.cuda.synchronize()
def timeit(fn, reps=50):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print(f"eager step: {timeit(step):.3f} ms")
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    n_static, g_static = step()
print(f"graph replay: {timeit(g.replay):.3f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=60))
