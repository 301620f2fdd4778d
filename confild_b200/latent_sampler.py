"""Latent-space DDPM sampler for unconditional generation (BASELINE config 5).

The reference generates a latent image ``(B, 1, T, L)`` with 1,000 ancestral DDPM steps of an OpenAI guided-diffusion
U-Net in PyTorch eager fp32 (``UnconditionalDiffusionTraining_and_Generation/src/unet.py:396-663`` called from
``src/gaussian_diffusion.py:395-535``), then decodes it frame by frame (``scripts/inference.py:55-79``).  This module
is the B200-side replacement of that outer loop around the CNF decoder:

* ``LatentUNet`` -- a U-Net with the reference's **state-dict layout** (``time_embed.*``, ``input_blocks.{i}.{j}.*``,
  ``middle_block.*``, ``output_blocks.*``, ``out.*``; same parameter shapes, same construction order, so the same seed
  gives bit-identical initial weights and ``ema_*.pt`` checkpoints load unchanged), written for inference: bf16
  autocast for the convolutions / linears with fp32 GroupNorm (as the reference's GroupNorm32), fused
  ``scaled_dot_product_attention`` for the legacy head-major QKV layout (``unet.py:328-357``) and channels-last
  activations.  The recipe configuration only: ``dims=2``, no class conditioning, ``use_scale_shift_norm=False``,
  ``resblock_updown=False``, ``learn_sigma=False`` (``script_util.py:130-187`` with the recipes' arguments).
* ``DDPMSchedule`` / ``sample_latents`` -- the reference's sampler as used by ``create_gaussian_diffusion(steps, cosine)``
  (``script_util.py:388-426``): epsilon prediction, FIXED_LARGE variance, ``clip_denoised`` -- with the coefficient
  tables on the device and ONE sampling step (U-Net + update + noise) captured in a CUDA graph that is replayed
  ``steps`` times, the timestep living in a device counter.
* ``generate_fields`` -- sampler -> latent de-normalisation -> batched CNF decode through ``confild_b200.decoder``: the
  batched replacement of the ``B*T`` one-frame Python iterations of ``scripts/inference.py:71-79``.

This is host-side PyTorch around the hot path (the U-Net is outside the graded kernel set); its GEMM-shaped kernels are
library kernels (cuDNN / cuBLAS / flash attention).  ``LatentUNet.forward_inference`` is the no-grad fast path the
sampler uses on CUDA: bf16 channels-last activations end to end, weights cast once, and every GroupNorm(+SiLU) site --
including the residual block's timestep-embedding add in front of the second one -- as one call of the library's
``cnf_group_norm_nhwc_bf16`` (``csrc/unet_ops.cu``) instead of the eager ``float() / native_group_norm / to(bf16) /
SiLU`` chain and the NCHW<->NHWC conversions cuDNN then runs around every convolution (a profile of the autocast path
shows the convolutions at ~10 % of the step; normalisation, layout conversions and element-wise passes take the rest).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import _native


def group_norm_nhwc(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, num_groups: int, eps: float = 1e-5,
                    add: Optional[torch.Tensor] = None, silu: bool = False) -> torch.Tensor:
    """``act(GroupNorm(x + add[:, :, None, None]))`` for a CUDA bf16 channels-last ``x (N, C, H, W)`` with fp32
    statistics (``cnf_group_norm_nhwc_bf16``); ``weight`` / ``bias`` fp32 ``(C,)``, ``add`` fp32 ``(N, C)`` (rows may be
    a column slice of a wider matrix: only the last dimension must be contiguous) or None.
    Inference only (no autograd).  Reference semantics: src/nn.py:17-19 (GroupNorm32) followed by nn.SiLU."""
    if not (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4):
        raise ValueError("group_norm_nhwc expects a 4-d CUDA bfloat16 tensor")
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    N, C, H, W = x.shape
    if weight.dtype != torch.float32 or bias.dtype != torch.float32 or weight.numel() != C or bias.numel() != C:
        raise ValueError("group_norm_nhwc: weight and bias must be float32 of shape (C,)")
    if add is not None:
        if add.dtype != torch.float32 or tuple(add.shape) != (N, C):
            raise ValueError("group_norm_nhwc: add must be float32 of shape (N, C)")
        if add.stride(1) != 1 or (N > 1 and add.stride(0) < C):
            add = add.contiguous()
    lib = _native.load()
    y = torch.empty_like(x)  # preserves the channels-last strides
    scratch = torch.empty(max(1, lib.cnf_group_norm_scratch_bytes(N) // 4), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _native.check(lib.cnf_group_norm_nhwc_bf16(x.data_ptr(), None if add is None else add.data_ptr(),
                                                   0 if add is None else (add.stride(0) if N > 1 else C),
                                                   weight.contiguous().data_ptr(), bias.contiguous().data_ptr(),
                                                   y.data_ptr(), scratch.data_ptr(), N, H * W, C, int(num_groups),
                                                   float(eps), int(bool(silu)), stream), "cnf_group_norm_nhwc_bf16")
    return y


# ------------------------------------------------------------------------------------------ U-Net
class _GroupNorm32(nn.GroupNorm):
    """GroupNorm computed in fp32 whatever the activation dtype (reference: src/nn.py:17-19)."""

    def forward(self, x):
        return F.group_norm(x.float(), self.num_groups, self.weight, self.bias, self.eps).to(x.dtype)


def _zeroed(module: nn.Module) -> nn.Module:
    for p in module.parameters():
        p.detach().zero_()
    return module


def sinusoidal_embedding(timesteps: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    """[cos | sin] of t * max_period^(-i/half) (reference: src/nn.py:118-135)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


class _Res(nn.Module):
    """Residual block, parameter names ``in_layers.{0,2}``, ``emb_layers.1``, ``out_layers.{0,3}``,
    ``skip_connection`` (reference: unet.py:143-256, the non-updown, additive-embedding variant)."""

    def __init__(self, cin: int, emb: int, cout: int, dropout: float):
        super().__init__()
        self.in_layers = nn.Sequential(_GroupNorm32(32, cin), nn.SiLU(), nn.Conv2d(cin, cout, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb, cout))
        self.out_layers = nn.Sequential(_GroupNorm32(32, cout), nn.SiLU(), nn.Dropout(p=dropout),
                                        _zeroed(nn.Conv2d(cout, cout, 3, padding=1)))
        self.skip_connection = nn.Identity() if cin == cout else nn.Conv2d(cin, cout, 1)

    def forward(self, x, emb):
        h = self.in_layers(x)
        h = h + self.emb_layers(emb).to(h.dtype)[:, :, None, None]
        return self.skip_connection(x) + self.out_layers(h)


class _Attn(nn.Module):
    """Self-attention over the spatial positions, parameter names ``norm``, ``qkv``, ``proj_out``; the 1x1 ``qkv``
    convolution emits, per head, ``[q | k | v]`` channel groups (the legacy order, unet.py:328-357)."""

    def __init__(self, channels: int, num_heads: int, num_head_channels: int):
        super().__init__()
        if num_head_channels == -1:
            self.num_heads = num_heads
        else:
            if channels % num_head_channels:
                raise ValueError(f"{channels} channels are not divisible by num_head_channels {num_head_channels}")
            self.num_heads = channels // num_head_channels
        self.norm = _GroupNorm32(32, channels)
        self.qkv = nn.Conv1d(channels, channels * 3, 1)
        self.proj_out = _zeroed(nn.Conv1d(channels, channels, 1))

    def forward(self, x):
        b, c, hh, ww = x.shape
        t = hh * ww
        xf = x.reshape(b, c, t)
        qkv = self.qkv(self.norm(xf)).reshape(b, self.num_heads, 3, c // self.num_heads, t)
        q, k, v = (qkv[:, :, i].transpose(-1, -2) for i in range(3))  # (b, heads, t, ch)
        a = F.scaled_dot_product_attention(q, k, v)  # softmax(q k^T / sqrt(ch)) v, as q*ch^-1/4 . k*ch^-1/4 there
        a = a.transpose(-1, -2).reshape(b, c, t)
        return (xf + self.proj_out(a)).reshape(b, c, hh, ww)


class _Down(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.op = nn.Conv2d(ch, ch, 3, stride=2, padding=1)

    def forward(self, x):
        return self.op(x)


class _Up(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2, mode="nearest"))


class _Stage(nn.Sequential):
    """Children indexed 0, 1, ... like the reference's TimestepEmbedSequential (unet.py:66-78)."""

    def forward(self, x, emb):
        for layer in self:
            x = layer(x, emb) if isinstance(layer, _Res) else layer(x)
        return x


def default_channel_mult(image_size: int) -> Tuple[float, ...]:
    """script_util.py:149-160."""
    table = {512: (0.5, 1, 1, 2, 2, 4, 4), 256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4)}
    if image_size not in table:
        raise ValueError(f"unsupported image size: {image_size}")
    return table[image_size]


class LatentUNet(nn.Module):
    """State-dict-compatible inference U-Net (see the module docstring).  Constructor arguments follow
    ``create_model`` (script_util.py:130-187): ``attention_resolutions`` is the recipe string ("32,16,8")."""

    def __init__(self, image_size: int, num_channels: int, num_res_blocks: int, channel_mult: Optional[str] = None,
                 attention_resolutions: str = "16", num_heads: int = 1, num_head_channels: int = -1,
                 in_channels: int = 1, out_channels: int = 1, dropout: float = 0.0):
        super().__init__()
        mult = default_channel_mult(image_size) if channel_mult is None else tuple(int(m) for m in channel_mult.split(","))
        attn_ds = tuple(image_size // int(r) for r in attention_resolutions.split(","))
        self.model_channels = mc = num_channels
        emb = 4 * mc
        # construction order = the reference constructor's (unet.py:467-616), so a shared seed gives identical weights
        self.time_embed = nn.Sequential(nn.Linear(mc, emb), nn.SiLU(), nn.Linear(emb, emb))
        ch = int(mult[0] * mc)
        self.input_blocks = nn.ModuleList([_Stage(nn.Conv2d(in_channels, ch, 3, padding=1))])
        skip_chans = [ch]
        ds = 1
        for level, m in enumerate(mult):
            for _ in range(num_res_blocks):
                layers = [_Res(ch, emb, int(m * mc), dropout)]
                ch = int(m * mc)
                if ds in attn_ds:
                    layers.append(_Attn(ch, num_heads, num_head_channels))
                self.input_blocks.append(_Stage(*layers))
                skip_chans.append(ch)
            if level != len(mult) - 1:
                self.input_blocks.append(_Stage(_Down(ch)))
                skip_chans.append(ch)
                ds *= 2
        self.middle_block = _Stage(_Res(ch, emb, ch, dropout), _Attn(ch, num_heads, num_head_channels),
                                   _Res(ch, emb, ch, dropout))
        self.output_blocks = nn.ModuleList([])
        for level, m in list(enumerate(mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [_Res(ch + skip_chans.pop(), emb, int(mc * m), dropout)]
                ch = int(mc * m)
                if ds in attn_ds:
                    layers.append(_Attn(ch, num_heads, num_head_channels))
                if level and i == num_res_blocks:
                    layers.append(_Up(ch))
                    ds //= 2
                self.output_blocks.append(_Stage(*layers))
        self.out = nn.Sequential(_GroupNorm32(32, ch), nn.SiLU(),
                                 _zeroed(nn.Conv2d(int(mult[0] * mc), out_channels, 3, padding=1)))

    # ------------------------------------------------------------------ inference fast path (CUDA, no autograd)
    @torch.no_grad()
    def prepare_inference(self) -> "LatentUNet":
        """(Re)build the bf16 / channels-last weight copies ``forward_inference`` uses (call again after changing or
        loading weights; ``sample_latents`` does so on every call)."""
        bf = torch.bfloat16
        fast = {}
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                fast[m] = (m.weight.detach().to(bf).contiguous(memory_format=torch.channels_last), m.bias.detach().to(bf))
            elif isinstance(m, nn.Conv1d):  # 1x1: a linear layer over the channels
                fast[m] = (m.weight.detach()[:, :, 0].to(bf).contiguous(), m.bias.detach().to(bf))
            elif isinstance(m, nn.Linear):
                fast[m] = (m.weight.detach().to(bf).contiguous(), m.bias.detach().to(bf))
            elif isinstance(m, nn.GroupNorm):
                fast[m] = (m.weight.detach().float().contiguous(), m.bias.detach().float().contiguous())
        # every residual block projects the same SiLU(time embedding): one GEMM for all of them (fp32 result, sliced)
        res = [m for m in self.modules() if isinstance(m, _Res)]
        fast["emb_w"] = torch.cat([m.emb_layers[1].weight.detach() for m in res], 0).to(bf).contiguous()
        fast["emb_b"] = torch.cat([m.emb_layers[1].bias.detach() for m in res], 0).to(bf).contiguous()
        off = 0
        for m in res:
            n = m.emb_layers[1].out_features
            fast[("emb_slice", m)] = (off, off + n)
            off += n
        self._fast = fast
        return self

    @torch.no_grad()
    def forward_inference(self, x, timesteps):
        """Same function as ``forward`` under bf16 autocast, for CUDA inference: see the module docstring."""
        fast = getattr(self, "_fast", None)
        if fast is None:
            fast = self.prepare_inference()._fast
        bf, cl = torch.bfloat16, torch.channels_last

        def lin(m, v):
            w, b = fast[m]
            return F.linear(v, w, b)

        def conv(m, v):
            w, b = fast[m]
            return F.conv2d(v, w, b, m.stride, m.padding).contiguous(memory_format=cl)

        def gn(m, v, silu, add=None):
            w, b = fast[m]
            return group_norm_nhwc(v, w, b, m.num_groups, m.eps, add=add, silu=silu)

        emb = lin(self.time_embed[2], F.silu(lin(self.time_embed[0], sinusoidal_embedding(timesteps, self.model_channels).to(bf))))
        # every residual block starts its embedding branch with the same SiLU -> Linear: one GEMM, fp32, sliced per block
        emb_all = F.linear(F.silu(emb), fast["emb_w"], fast["emb_b"]).float()

        def res(m, v):
            h = conv(m.in_layers[2], gn(m.in_layers[0], v, True))
            lo, hi = fast[("emb_slice", m)]
            e = emb_all[:, lo:hi]  # added inside the second normalisation, in fp32 (a strided view: no copy)
            h = conv(m.out_layers[3], gn(m.out_layers[0], h, True, add=e))
            return (v if isinstance(m.skip_connection, nn.Identity) else conv(m.skip_connection, v)) + h

        def attn(m, v):
            n, c, hh, ww = v.shape
            t = hh * ww
            heads = m.num_heads
            xn = gn(m.norm, v, False).permute(0, 2, 3, 1).reshape(n, t, c)  # a view: the memory is (n, t, c) already
            qkv = lin(m.qkv, xn).reshape(n, t, heads, 3, c // heads)       # legacy order: per head [q | k | v]
            q, k, vv = (qkv[:, :, :, i].transpose(1, 2) for i in range(3))  # (n, heads, t, ch)
            a = F.scaled_dot_product_attention(q, k, vv).transpose(1, 2).reshape(n, t, c)
            return v + lin(m.proj_out, a).reshape(n, hh, ww, c).permute(0, 3, 1, 2)

        def stage(blk, v):
            for layer in blk:
                if isinstance(layer, _Res):
                    v = res(layer, v)
                elif isinstance(layer, _Attn):
                    v = attn(layer, v)
                elif isinstance(layer, _Down):
                    v = conv(layer.op, v)
                elif isinstance(layer, _Up):
                    v = conv(layer.conv, F.interpolate(v, scale_factor=2, mode="nearest"))
                else:
                    v = conv(layer, v)
            return v

        h = x.to(bf).contiguous(memory_format=cl)
        hs = []
        for blk in self.input_blocks:
            h = stage(blk, h)
            hs.append(h)
        h = stage(self.middle_block, h)
        for blk in self.output_blocks:
            h = stage(blk, torch.cat([h, hs.pop()], dim=1).contiguous(memory_format=cl))
        return conv(self.out[2], gn(self.out[0], h, True)).to(x.dtype).contiguous()

    def forward(self, x, timesteps):
        """``x (N, C, T, L)``, ``timesteps (N,)`` -> predicted noise ``(N, C, T, L)`` (reference: unet.py:634-663)."""
        emb = self.time_embed(sinusoidal_embedding(timesteps, self.model_channels))
        hs = []
        h = x
        for blk in self.input_blocks:
            h = blk(h, emb)
            hs.append(h)
        h = self.middle_block(h, emb)
        for blk in self.output_blocks:
            h = blk(torch.cat([h, hs.pop()], dim=1), emb)
        return self.out(h.to(x.dtype))


# ------------------------------------------------------------------------------------------ DDPM sampler
class DDPMSchedule:
    """Coefficient tables of the reference sampler (gaussian_diffusion.py:18-66, 120-160, 232-335): cosine or linear
    betas, epsilon prediction, FIXED_LARGE variance, computed in float64 like the reference."""

    def __init__(self, steps: int = 1000, noise_schedule: str = "cosine"):
        if noise_schedule == "cosine":
            def abar(t):
                return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
            betas = np.array([min(1 - abar((i + 1) / steps) / abar(i / steps), 0.999) for i in range(steps)], dtype=np.float64)
        elif noise_schedule == "linear":
            scale = 1000 / steps
            betas = np.linspace(scale * 0.0001, scale * 0.02, steps, dtype=np.float64)
        else:
            raise NotImplementedError(f"unknown beta schedule: {noise_schedule}")
        alphas = 1.0 - betas
        acp = np.cumprod(alphas, axis=0)
        acp_prev = np.append(1.0, acp[:-1])
        post_var = betas * (1.0 - acp_prev) / (1.0 - acp)
        self.steps = steps
        self.betas = betas
        self.sqrt_recip_acp = np.sqrt(1.0 / acp)
        self.sqrt_recipm1_acp = np.sqrt(1.0 / acp - 1)
        self.post_coef1 = betas * np.sqrt(acp_prev) / (1.0 - acp)
        self.post_coef2 = (1.0 - acp_prev) * np.sqrt(alphas) / (1.0 - acp)
        self.sigma = np.sqrt(np.append(post_var[1], betas[1:]))  # FIXED_LARGE: exp(0.5 * log variance)

    def table(self, device) -> torch.Tensor:
        """(steps, 5) fp32: sqrt_recip, sqrt_recipm1, coef1, coef2, sigma (sigma[0] = 0: no noise at t = 0)."""
        sig = self.sigma.copy()
        sig[0] = 0.0
        t = np.stack([self.sqrt_recip_acp, self.sqrt_recipm1_acp, self.post_coef1, self.post_coef2, sig], axis=1)
        return torch.tensor(t, dtype=torch.float32, device=device)


def ddpm_step(eps: torch.Tensor, x: torch.Tensor, coef: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    """x_{t-1} from x_t and the predicted noise: x0 = clip(a x_t - b eps); mean = c1 x0 + c2 x_t; + sigma * noise
    (gaussian_diffusion.py:293-314, 328-335, 395-439).  ``coef`` = the schedule table row(s) of t, shape (5,) or (N, 5)."""
    c = coef.reshape(-1, 5)[:, :, None, None, None]
    x0 = (c[:, 0] * x - c[:, 1] * eps).clamp(-1, 1)
    return c[:, 2] * x0 + c[:, 3] * x + c[:, 4] * noise


@torch.no_grad()
def sample_latents(model: nn.Module, shape: Sequence[int], steps: int = 1000, noise_schedule: str = "cosine",
                   device=None, autocast_dtype: Optional[torch.dtype] = torch.bfloat16, use_cuda_graph: bool = True,
                   generator: Optional[torch.Generator] = None, fast_unet: bool = True) -> torch.Tensor:
    """Ancestral sampling ``p_sample_loop(model, shape)`` (gaussian_diffusion.py:441-535) -> ``(B, C, T, L)`` in [-1, 1].

    On a CUDA device one step (U-Net under ``autocast_dtype``, update, fresh noise) is captured in a CUDA graph and
    replayed ``steps`` times; the timestep is a device counter, so the host only enqueues replays.  With
    ``fast_unet`` (default) and bf16 the U-Net runs ``LatentUNet.forward_inference`` instead of ``forward`` under autocast.  ``generator`` makes
    the eager path reproducible (graph capture uses the default CUDA generator, which is graph-safe).
    """
    dev = torch.device(device) if device is not None else next(model.parameters()).device
    sched = DDPMSchedule(steps, noise_schedule)
    table = sched.table(dev)
    B = int(shape[0])
    x = torch.randn(*shape, device=dev, generator=generator)
    cuda = dev.type == "cuda"

    fast = (cuda and fast_unet and autocast_dtype == torch.bfloat16 and hasattr(model, "forward_inference"))
    if fast:
        model.prepare_inference()  # bf16 channels-last weight copies, rebuilt here so that they are never stale

    def unet(xx, tt):
        if fast:
            return model.forward_inference(xx, tt)
        if cuda and autocast_dtype is not None:
            with torch.autocast("cuda", dtype=autocast_dtype):
                return model(xx.contiguous(memory_format=torch.channels_last), tt).float()
        return model(xx, tt)

    if not (cuda and use_cuda_graph):
        for i in reversed(range(steps)):
            t = torch.full((B,), i, device=dev, dtype=torch.long)
            noise = torch.randn(x.shape, device=dev, generator=generator)
            x = ddpm_step(unet(x, t), x, table[i], noise)
        return x

    t_dev = torch.full((B,), steps - 1, device=dev, dtype=torch.long)
    x_static = x.clone()

    def one_step():
        eps = unet(x_static, t_dev)
        noise = torch.randn_like(x_static)
        x_static.copy_(ddpm_step(eps, x_static, table.index_select(0, t_dev), noise))  # (B, 5) rows: no host sync
        t_dev.sub_(1)

    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):  # warm-up outside capture (cuDNN autotune, allocator), then restore the state
        for _ in range(2):
            one_step()
        x_static.copy_(x)
        t_dev.fill_(steps - 1)
    torch.cuda.current_stream(dev).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        one_step()
    x_static.copy_(x)
    t_dev.fill_(steps - 1)
    for _ in range(steps):
        graph.replay()
    return x_static.clone()


# ------------------------------------------------------------------------------------------ end-to-end generation
@torch.no_grad()
def generate_fields(unet: nn.Module, cnf_model, coords: torch.Tensor, latent_max, latent_min, x_normalizer,
                    y_normalizer, n_samples: int, time_length: int, latent_length: int, device, steps: int = 1000,
                    noise_schedule: str = "cosine", out: Optional[torch.Tensor] = None, **sampler_kwargs):
    """Unconditional generation end to end (scripts/inference.py:55-79): sample ``(n_samples, 1, T, L)`` latents,
    de-normalise them ``(z + 1)(max - min)/2 + min`` (:59-61) and decode ALL ``n_samples * T`` frames with one batched
    call of the CUDA decoder (instead of one ``trainer.infer`` per frame, :71-77).  Returns ``(fields, latents)``:
    ``fields (n_samples * T, P, cout)`` on the host (pinned), ``latents (n_samples, T, L)`` on the device."""
    from .inference_function import decoder

    z = sample_latents(unet, (n_samples, 1, time_length, latent_length), steps=steps, noise_schedule=noise_schedule,
                       device=device, **sampler_kwargs)[:, 0]
    hi = torch.as_tensor(latent_max, dtype=torch.float32, device=z.device)
    lo = torch.as_tensor(latent_min, dtype=torch.float32, device=z.device)
    latents = (z + 1) * (hi - lo) / 2.0 + lo
    fields = decoder(coords, latents.reshape(n_samples * time_length, latent_length), cnf_model, x_normalizer,
                     y_normalizer, 16, device, out=out)
    return fields, latents
