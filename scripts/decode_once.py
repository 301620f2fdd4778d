"""A few plain decode calls of one recipe shape (profiling target): python scripts/decode_once.py case4 16 16384 bf16x3"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import confild_b200 as cb
from oracle import cnf_oracle as O
case, T, P, prec = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
dims = O.CASE_SHAPES[case]; sd = O.init_params(*dims, seed=0)
coords, lat = O.synthetic_inputs(dims[0], dims[1], T, P)
m = cb.SIRENAutodecoder_film(*dims[:2], dims[2], dims[3], dims[4], precision=prec); m.load_state_dict(sd); m = m.eval().cuda()
c, l = coords.cuda()[None], lat.cuda()[:, None]
with torch.no_grad():
    for _ in range(4):
        y = m(c, l)
torch.cuda.synchronize()
print("ok", float(y.abs().sum()))
