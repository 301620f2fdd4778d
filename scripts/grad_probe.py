import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import confild_b200 as cb
from oracle import cnf_oracle as O
T, P = int(sys.argv[1]), int(sys.argv[2])
dims = O.CASE_SHAPES["case1"]; sd = O.init_params(*dims, seed=0)
coords, lat = O.synthetic_inputs(dims[0], dims[1], T, P)
m = cb.SIRENAutodecoder_film(2,128,3,10,128, precision="bf16x3"); m.load_state_dict(sd); m = m.eval().cuda()
l = lat.cuda()[:, None].requires_grad_(True)
y = m(coords.cuda()[None], l)
torch.cuda.synchronize(); print("fwd ok", float(y.detach().abs().sum()))
gout = torch.randn(y.shape, generator=torch.Generator().manual_seed(7))
(g,) = torch.autograd.grad(y, l, grad_outputs=gout.cuda())
torch.cuda.synchronize(); print("bwd ok")
gw = O.grad_latents_from_gout(sd, coords[None], lat[:, None], gout)
print("fwd err", O.rel_l2(y, O.forward(sd, coords[None], lat[:, None])), "dlat err", O.rel_l2(g, gw))
