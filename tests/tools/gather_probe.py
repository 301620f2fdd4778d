"""2+ GPU probe (torchrun): fused all-gather decode vs decode + NCCL all_gather_into_tensor.

    torchrun --nproc-per-node N --master-addr 127.0.0.1 tests/tools/gather_probe.py T P [steps]

Multi-step with DIFFERENT latents every step and deliberately skewed ranks (odd ranks run extra kernels before consuming
a result), so that a write-after-read race between one rank's next call and a slow peer's read of the previous result
would show up as a mismatch.  Exits non-zero on any mismatch.  Used by tests/test_multigpu.py.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist

import confild_b200 as cb
from oracle import cnf_oracle as O  # checker only

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
T, P = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
dims = O.CASE_SHAPES["case1"]
sd = O.init_params(*dims, seed=0)
m = cb.SIRENAutodecoder_film(2, 128, 3, 10, 128, precision="bf16x3")
m.load_state_dict(sd)
m = m.eval().to(dev)
coords, _ = O.synthetic_inputs(2, 128, T, P)
c = coords.to(dev)[None]
lats = [O.synthetic_inputs(2, 128, T, P, latent_seed=100 * s + rank)[1].to(dev)[:, None] for s in range(steps)]
gathered = torch.empty((world * T, P, 3), device=dev)
spin = torch.empty(64 << 20, device=dev)

bad = 0
for buffers in (2, 1):
    fused = cb.FusedGatherDecoder(m, T, P, buffers=buffers)
    for s in range(steps):
        with torch.no_grad():
            dist.all_gather_into_tensor(gathered, m(c, lats[s]))
        want = gathered.clone()
        got = fused(c, lats[s])
        if rank % 2 == 1:  # slow consumer: delay this rank's read of the result
            for _ in range(20):
                spin.add_(1.0)
        if not torch.equal(got, want):
            bad += 1
    del fused
t = torch.tensor([bad], device=dev)
dist.all_reduce(t)
bad = int(t.item())


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    tt = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return float(tt)


def step_nccl():
    with torch.no_grad():
        dist.all_gather_into_tensor(gathered, m(c, lats[0]))


f2 = cb.FusedGatherDecoder(m, T, P, buffers=2)
f1 = cb.FusedGatherDecoder(m, T, P, buffers=1)
t_n, t_f2, t_f1 = timeit(step_nccl), timeit(lambda: f2(c, lats[0])), timeit(lambda: f1(c, lats[0]))
with torch.no_grad():
    t_local = timeit(lambda: m(c, lats[0]))
if rank == 0:
    print(f"world={world} T={T} P={P} steps={steps}: mismatches={bad}  local decode {t_local:.2f} ms | decode+NCCL "
          f"all-gather {t_n:.2f} ms | fused gather double-buffered {t_f2:.2f} ms | single buffer + leading barrier "
          f"{t_f1:.2f} ms", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if bad else 0)
