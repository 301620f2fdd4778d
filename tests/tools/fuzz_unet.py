"""Randomised U-Net configurations: LatentUNet.forward_inference against forward in fp32 and under bf16 autocast."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import confild_b200 as cb
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
for it in range(n_cases):
    cfg = dict(image_size=64, num_channels=rng.choice([32, 64, 96]), num_res_blocks=rng.choice([1, 2]),
               attention_resolutions=rng.choice(["16", "32,16", "32,16,8", "8"]), in_channels=rng.choice([1, 2]),
               out_channels=rng.choice([1, 2]))
    if rng.random() < 0.5: cfg.update(num_heads=rng.choice([1, 2, 4]), num_head_channels=-1)
    else: cfg.update(num_head_channels=rng.choice([16, 32]))
    if rng.random() < 0.3: cfg["channel_mult"] = rng.choice(["1,2,2", "1,1,2,2", "1,2"])
    torch.manual_seed(it)
    try:
        m = cb.LatentUNet(**cfg).eval()
    except ValueError as e:
        print("skip", cfg, e); continue
    g = torch.Generator().manual_seed(it)
    sd = {k: torch.randn(v.shape, generator=g) * 0.05 for k, v in m.state_dict().items()}
    m.load_state_dict(sd); m = m.cuda()
    S = rng.choice([32, 64]); N = rng.choice([1, 2, 3])
    x = torch.randn(N, cfg["in_channels"], S, S, device="cuda"); t = torch.randint(0, 1000, (N,), device="cuda")
    with torch.no_grad():
        want = m(x, t)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            auto = m(x.contiguous(memory_format=torch.channels_last), t).float()
        got = m.forward_inference(x, t)
    ea = float((auto - want).norm() / want.norm()); ef = float((got - want).norm() / want.norm())
    ok = got.shape == want.shape and ef <= max(5e-2, 1.5 * ea)
    print(f"{'ok  ' if ok else 'FAIL'} {cfg} S={S} N={N}: autocast {ea:.2e} fast {ef:.2e}", flush=True)
    if not ok: sys.exit(1)
print("all ok")
