"""A few plain decode calls of one recipe shape (profiling / A-B timing target):
    python tests/tools/decode_once.py case4 16 16384 bf16x3 [check]
Prints the CUDA-event time per call; with `check`, also the rel-L2 error against the CPU oracle on the first 2 frames."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import confild_b200 as cb
from oracle import cnf_oracle as O
case, T, P, prec = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
dims = O.CASE_SHAPES[case]; sd = O.init_params(*dims, seed=0)
coords, lat = O.synthetic_inputs(dims[0], dims[1], T, P)
m = cb.SIRENAutodecoder_film(*dims[:2], dims[2], dims[3], dims[4], precision=prec); m.load_state_dict(sd); m = m.eval().cuda()
c, l = coords.cuda()[None], lat.cuda()[:, None]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    for _ in range(3):
        y = m(c, l)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        y = m(c, l)
    e1.record()
    y_ref = y.clone()
    same = all(torch.equal(m(c, l), y_ref) for _ in range(6))
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
msg = f"{case} T={T} P={P} {prec}: {ms:.3f} ms -> {T * P / ms / 1e6:.4f} G pf/s"
if len(sys.argv) > 5 and sys.argv[5] == "check":
    n = min(P, 4096)
    want = O.forward(sd, coords[None, :n], lat[:2, None])
    msg += f"  rel_l2 vs oracle {O.rel_l2(y[:2, :n].cpu(), want):.3e}"
print(msg + f"  bitwise reproducible over 6 reruns: {same}", flush=True)
