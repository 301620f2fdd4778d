"""2+ GPU probe (torchrun): fused all-gather decode vs decode + NCCL all_gather_into_tensor."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist
import confild_b200 as cb
from oracle import cnf_oracle as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=dev)
T, P = int(sys.argv[1]), int(sys.argv[2])
dims = O.CASE_SHAPES["case1"]; sd = O.init_params(*dims, seed=0)
m = cb.SIRENAutodecoder_film(2, 128, 3, 10, 128, precision="bf16x3"); m.load_state_dict(sd); m = m.eval().to(dev)
coords, lat = O.synthetic_inputs(2, 128, T, P, latent_seed=2 + rank)
c, l = coords.to(dev)[None], lat.to(dev)[:, None]
gathered = torch.empty((world * T, P, 3), device=dev)

def step_nccl():
    with torch.no_grad():
        y = m(c, l)
        dist.all_gather_into_tensor(gathered, y)
    return gathered

fused = cb.FusedGatherDecoder(m, T, P)
def step_fused():
    return fused(c, l)

ref = step_nccl().clone()
got = step_fused()
torch.cuda.synchronize(); dist.barrier()
ok = torch.equal(ref, got)
def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
t_n, t_f = timeit(step_nccl), timeit(step_fused)
with torch.no_grad():
    t_local = timeit(lambda: m(c, l))
if rank == 0:
    print(f"world={world} T={T} P={P}: identical={ok}  local decode {t_local:.2f} ms | decode+NCCL all-gather {t_n:.2f} ms | fused gather {t_f:.2f} ms", flush=True)
dist.barrier(); dist.destroy_process_group()
