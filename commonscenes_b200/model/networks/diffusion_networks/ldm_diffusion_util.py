"""Host-side helpers of the denoiser, mirroring the names of the reference's
model/networks/diffusion_networks/ldm_diffusion_util.py so callers can switch imports:

  make_beta_schedule (:43-65)   make_ddim_timesteps (:68-83)   make_ddim_sampling_parameters (:86-96)
  extract_into_tensor (:118-122)   timestep_embedding (:174-194)   normalization/GroupNorm32 (:225-239)
  conv_nd / linear (:241-260)   zero_module (:197-203)   noise_like (:289-292)

The schedule helpers are float64 numpy host code run once; timestep_embedding runs on the GPU through the
C ABI (cs_timestep_embedding).  The nn.Module factories only create PARAMETER HOLDERS: their forward()
is never used on the hot path (the kernels are called by the owning block).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .... import ops


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    if schedule == "linear":
        # torch.linspace (not numpy) so the float64 table is bit-identical to the reference's
        betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2).numpy()
    elif schedule == "cosine":
        ts = np.arange(n_timestep + 1, dtype=np.float64) / n_timestep + cosine_s
        alphas = np.cos(ts / (1 + cosine_s) * np.pi / 2) ** 2
        alphas = alphas / alphas[0]
        betas = np.clip(1 - alphas[1:] / alphas[:-1], 0, 0.999)
    elif schedule == "sqrt_linear":
        betas = np.linspace(linear_start, linear_end, n_timestep, dtype=np.float64)
    elif schedule == "sqrt":
        betas = np.linspace(linear_start, linear_end, n_timestep, dtype=np.float64) ** 0.5
    else:
        raise ValueError(f"schedule '{schedule}' unknown.")
    return betas


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=True):
    if ddim_discr_method == "uniform":
        c = num_ddpm_timesteps // num_ddim_timesteps
        ddim_timesteps = np.asarray(list(range(0, num_ddpm_timesteps, c)))
    elif ddim_discr_method == "quad":
        ddim_timesteps = ((np.linspace(0, np.sqrt(num_ddpm_timesteps * .8), num_ddim_timesteps)) ** 2).astype(int)
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
    steps_out = ddim_timesteps + 1   # same +1 convention as the reference (so S=1000 overruns, like the reference)
    if verbose:
        print(f"Selected timesteps for ddim sampler: {steps_out}")
    return steps_out


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=True):
    """alphacums: fp32 numpy array / CPU tensor of alphas_cumprod.  Returns fp32 numpy (sigmas, alphas, alphas_prev)."""
    ac = np.asarray(alphacums, dtype=np.float32)
    alphas = ac[ddim_timesteps]
    alphas_prev = np.asarray([ac[0]] + ac[ddim_timesteps[:-1]].tolist(), dtype=np.float32)
    sigmas = (eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))).astype(np.float32)
    if verbose:
        print(f"Selected alphas for ddim sampler: a_t: {alphas}; a_(t-1): {alphas_prev}")
    return sigmas, alphas, alphas_prev


def extract_into_tensor(a, t, x_shape):
    b, *_ = t.shape
    out = a.gather(-1, t)
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))


def timestep_embedding(timesteps, dim, max_period=10000, repeat_only=False):
    """[N] int64 CUDA -> [N, dim] fp32 sinusoidal embedding (cs_timestep_embedding)."""
    if repeat_only:
        return timesteps[:, None].float().repeat(1, dim)
    if dim % 2:
        raise NotImplementedError("odd embedding dims are not used by the reference configs")
    return ops.timestep_embedding(timesteps.to(torch.int64).contiguous(), dim, float(max_period))


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


class GroupNorm32(nn.GroupNorm):
    """Parameter holder for GroupNorm(32, C); the arithmetic is cs_groupnorm_* (fp32 statistics)."""


def normalization(channels):
    return GroupNorm32(32, channels)


def conv_nd(dims, *args, **kwargs):
    if dims == 1:                      # the 1x1 qkv / proj_out projections of AttentionBlock (openai_model_3d.py:340,348)
        return nn.Conv1d(*args, **kwargs)
    if dims in (3, 4):
        return nn.Conv3d(*args, **kwargs)
    raise ValueError(f"unsupported dimensions: {dims} (the shape branch is 3-D)")


def linear(*args, **kwargs):
    return nn.Linear(*args, **kwargs)


def noise_like(shape, device, repeat=False):
    if repeat:
        return torch.randn((1, *shape[1:]), device=device).repeat(shape[0], *((1,) * (len(shape) - 1)))
    return torch.randn(shape, device=device)
