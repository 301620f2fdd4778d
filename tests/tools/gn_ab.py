"""A/B of cnf_group_norm_nhwc_bf16's two paths on the GroupNorm sites of the case1 U-Net (2 samples): all sites replayed
from one CUDA graph per setting of CNF_GN_CLUSTER (threshold in KiB per sample; 0 = two kernels everywhere)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import confild_b200 as cb
from confild_b200 import _native
from confild_b200.latent_sampler import group_norm_nhwc
dev = torch.device("cuda:0")
unet = cb.LatentUNet(image_size=128, num_channels=128, num_res_blocks=2, num_heads=4, num_head_channels=64,
                     attention_resolutions="32,16,8").eval().to(dev)
sites = []
def hook(mod, inp, out):
    sites.append(tuple(inp[0].shape))
hs = [m.register_forward_hook(hook) for m in unet.modules() if isinstance(m, torch.nn.GroupNorm)]
with torch.no_grad():
    unet(torch.randn(2, 1, 128, 128, device=dev), torch.tensor([5, 500], device=dev))
for h in hs: h.remove()
shapes = [(s[0], s[1], s[2], s[3]) if len(s) == 4 else (s[0], s[1], s[2], 1) for s in sites]
print(len(shapes), "sites; bytes per sample (KiB):", sorted({s[1] * s[2] * s[3] * 2 // 1024 for s in shapes}))
xs = [torch.randn(*s, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last) for s in shapes]
ws = [torch.randn(s[1], device=dev) for s in shapes]
def run_all():
    for x, w in zip(xs, ws):
        group_norm_nhwc(x, w, w, 32, 1e-5, silu=True)
def graph_for(k):
    _native.set_knob("CNF_GN_CLUSTER", k)
    run_all(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run_all()
    return g
ks = [0, 64, 128, 256, 512, 1024, 2048]
graphs = {k: graph_for(k) for k in ks}
res = {k: [] for k in ks}
for rep in range(7):
    for k in ks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20): graphs[k].replay()
        e1.record(); torch.cuda.synchronize()
        res[k].append(e0.elapsed_time(e1) / 20)
for k in ks:
    print(f"threshold {k:5d} KiB: {min(res[k]) * 1e3:8.1f} us per U-Net step (min of 7), median {sorted(res[k])[3] * 1e3:8.1f}")
