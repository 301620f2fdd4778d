import sys, torch
sys.path.insert(0, ".")
import confild_b200 as cb
from oracle import cnf_oracle as O
dims = O.CASE_SHAPES["case1"]; sd = O.init_params(*dims, seed=0)
coords, lat = O.synthetic_inputs(dims[0], dims[1], 5, 700)
outs = {}
for d in ("cuda:0", "cuda:1"):
    m = cb.SIRENAutodecoder_film(*dims[:2], dims[2], dims[3], dims[4]); m.load_state_dict(sd); m = m.eval().to(d); m.disable_gradient()
    torch.cuda.set_device(0)  # current device stays 0: the module must launch on ITS device
    l = lat.to(d)[:, None].requires_grad_(True)
    y = m(coords.to(d)[None], l)
    n = cb.measurement_norm(m, coords.to(d)[None], l, torch.zeros(5, 700, 3, device=d))
    (g,) = torch.autograd.grad(n, l)
    outs[d] = (y.detach().cpu(), g.cpu(), float(n))
    x = torch.randn(2, 64, 8, 8, device=d).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.ones(64, device=d)
    from confild_b200.latent_sampler import group_norm_nhwc
    outs[d] += (group_norm_nhwc(x, w, w, 32).float().cpu().abs().mean().item(),)
print(torch.equal(outs["cuda:0"][0], outs["cuda:1"][0]), float((outs["cuda:0"][1] - outs["cuda:1"][1]).norm() / outs["cuda:0"][1].norm()), outs["cuda:0"][2], outs["cuda:1"][2])
print("ok")
