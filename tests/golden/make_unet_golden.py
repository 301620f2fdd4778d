"""Golden fixtures for the latent sampler (f4) from the LIVE reference U-Net / diffusion (build container only):

    python tests/golden/make_unet_golden.py

Imports ``UnconditionalDiffusionTraining_and_Generation/src/{unet,gaussian_diffusion,script_util}.py`` unmodified and
writes
* ``unet_layouts.json``  -- state-dict keys and shapes of ``create_model`` for the case1 and case4 recipe arguments
  (script_util.py:130-187, training_recipes/case{1,4}.yml) and for the tiny test configuration;
* ``unet_tiny.npz``      -- for the tiny configuration with weights ``0.05 * randn(generator seed 1)`` in state-dict
  order (reproducible without the reference): input, timesteps, the reference U-Net's output, the reference sampler's
  ``p_sample`` result for fixed noise at four timesteps, and the schedule arrays of
  ``create_gaussian_diffusion(steps=1000, noise_schedule="cosine")``.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/UnconditionalDiffusionTraining_and_Generation")
from src.script_util import create_gaussian_diffusion, create_model  # noqa: E402

TINY = dict(image_size=64, num_channels=32, num_res_blocks=1, num_heads=4, num_head_channels=32,
            attention_resolutions="32,16,8")
RECIPES = {
    "tiny": TINY,
    "case1": dict(image_size=128, num_channels=128, num_res_blocks=2, num_heads=4, num_head_channels=64,
                  attention_resolutions="32,16,8"),
    "case4": dict(image_size=384, num_channels=128, num_res_blocks=2, num_heads=4, num_head_channels=64,
                  attention_resolutions="32,16,8", channel_mult="1, 1, 2, 2, 4, 4"),
}


def seeded_weights(sd):
    g = torch.Generator().manual_seed(1)
    return {k: torch.randn(v.shape, generator=g) * 0.05 for k, v in sd.items()}


if __name__ == "__main__":
    torch.set_num_threads(1)
    layouts = {}
    for name, cfg in RECIPES.items():
        with torch.device("meta"):
            m = create_model(**cfg)
        layouts[name] = {"config": cfg, "state_dict": [[k, list(v.shape)] for k, v in m.state_dict().items()]}
    with open(os.path.join(HERE, "unet_layouts.json"), "w") as f:
        json.dump(layouts, f)
    torch.manual_seed(0)
    ref = create_model(**TINY).eval()
    init_sha = float(sum(v.double().abs().sum() for v in ref.state_dict().values()))
    ref.load_state_dict(seeded_weights(ref.state_dict()))
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 1, 16, 16, generator=g)
    t = torch.tensor([7, 900])
    with torch.no_grad():
        y = ref(x, t)
    diff = create_gaussian_diffusion(steps=1000, noise_schedule="cosine")
    out = {"x": x.numpy(), "t": t.numpy(), "y": y.numpy(), "init_abs_sum_seed0": np.array(init_sha),
           "betas": np.asarray(diff.betas), "sqrt_recip_alphas_cumprod": np.asarray(diff.sqrt_recip_alphas_cumprod),
           "posterior_mean_coef1": np.asarray(diff.posterior_mean_coef1),
           "posterior_mean_coef2": np.asarray(diff.posterior_mean_coef2)}
    for ti in (999, 500, 1, 0):
        tt = torch.tensor([ti, ti])
        torch.manual_seed(5)
        out[f"p_sample_{ti}"] = diff.p_sample(ref, x, tt)["sample"].detach().numpy()
    torch.manual_seed(5)
    out["noise_seed5"] = torch.randn_like(x).numpy()
    np.savez_compressed(os.path.join(HERE, "unet_tiny.npz"), **out)
    print({k: getattr(v, "shape", None) for k, v in out.items()}, "y norm", float(y.norm()))
