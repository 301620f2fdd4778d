import os, sys, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import confild_b200 as cb
from oracle import cnf_oracle as O
mode = sys.argv[1]
dims = O.CASE_SHAPES["case1"]; sd = O.init_params(*dims, seed=0)
coords, lat = O.synthetic_inputs(dims[0], dims[1], 3, 300)
m = cb.SIRENAutodecoder_film(2,128,3,10,128, precision="fp32" if "simt" in mode else "bf16x3"); m.load_state_dict(sd); m = m.eval().cuda()
if "grad" in mode:
    l = lat.cuda()[:, None].requires_grad_(True)
    y = m(coords.cuda()[None], l)
    (g,) = torch.autograd.grad(y.sum(), l)
else:
    with torch.no_grad():
        y = m(coords.cuda()[None], lat.cuda()[:, None])
torch.cuda.synchronize()
print(mode, float(y.sum()), flush=True)
if "clean" in mode:
    del m, y
    gc.collect(); torch.cuda.synchronize()
