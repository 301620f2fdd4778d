import sys, time, torch
sys.path.insert(0, ".")
import confild_b200 as cb
dev = torch.device("cuda:0")
torch.manual_seed(0)
unet = cb.LatentUNet(image_size=128, num_channels=128, num_res_blocks=2, num_heads=4, num_head_channels=64,
                     attention_resolutions="32,16,8").eval().to(dev)
def run(tag, steps=400, **kw):
    cb.sample_latents(unet, (2, 1, 128, 128), steps=3, device=dev, **kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    cb.sample_latents(unet, (2, 1, 128, 128), steps=steps, device=dev, **kw)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{tag}: {dt / steps * 1e3:.3f} ms/step", flush=True)
run("fast (forward_inference)")
run("autocast forward", fast_unet=False)
run("fast (forward_inference), again")
# where does the time go: profile one eager autocast step by kernel
from torch.profiler import profile, ProfilerActivity
x = torch.randn(2, 1, 128, 128, device=dev); t = torch.full((2,), 500, device=dev)
with torch.no_grad():
    for _ in range(3): unet.forward_inference(x, t)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        unet.forward_inference(x, t)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
