"""DDPMSampler — ancestral (1000-step) sampling with classifier-free guidance on the B200 kernels (BASELINE cfg5).

The reference ships only DDIM and PLMS samplers, and its DDIM cannot take S = 1000 (make_ddim_timesteps adds 1 to every
timestep and then indexes alphas_cumprod[1000]: ldm_diffusion_util.py:79,87 — SURVEY.md §0).  It does register the DDPM
posterior buffers (posterior_variance, posterior_mean_coef1/2: sdfusion_txt2shape_model.py:214-224) without using them.
This sampler is the standard ancestral step over those buffers,

    x0 = sqrt(1/abar_t) x - sqrt(1/abar_t - 1) eps ;  x_{t-1} = coef1_t x0 + coef2_t x + [t > 0] sqrt(beta~_t) z ,

executed by the same fused kernel as DDIM (cs_ddim_step): the posterior mean equals DDIM's update with
a_t = abar_t, a_prev = abar_{t-1} (1 at t = 0) and sigma_t^2 = beta~_t (DDIM with eta = 1 over all T steps).  The UNet
evaluation is the CUDA-graph replay of DDIMSampler, guidance is e = e_uc + s (e_c - e_uc) on a [uncond; cond] batch.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from ..... import ops
from .ddim import DDIMSampler


class DDPMSampler(DDIMSampler):
    def make_schedule(self, timesteps: Optional[int] = None, verbose: bool = False, **unused):
        """Per-step scalars from the model's fp32 alphas_cumprod (float64 arithmetic on the host, like register_schedule)."""
        ac = self.model.alphas_cumprod.detach().float().cpu().numpy().astype(np.float64)
        T = ac.shape[0]
        assert T == self.ddpm_num_timesteps, "alphas have to be defined for each timestep"
        ac_prev = np.append(1.0, ac[:-1])
        betas = 1.0 - ac / ac_prev
        post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
        n = T if timesteps is None else int(timesteps)
        assert 1 <= n <= T
        self.ddim_timesteps = np.arange(n)                       # t = 0 .. n-1; the loop runs them in reverse
        self.ddim_alphas, self.ddim_alphas_prev = ac[:n], ac_prev[:n]
        self.ddim_sigmas = np.sqrt(post_var[:n])                 # exactly 0 at t = 0: the last step adds no noise
        self.ddim_sqrt_one_minus_alphas = np.sqrt(1.0 - ac[:n])

    @torch.no_grad()
    def sample(self, batch_size, shape, conditioning=None, x_T=None, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, timesteps: Optional[int] = None, temperature=1., callback=None,
               img_callback=None, log_every_t=100, generator: Optional[torch.Generator] = None, verbose=False, **kwargs):
        """-> (samples, intermediates{'x_inter','pred_x0'}) like DDIMSampler.sample.  `timesteps` < T starts the chain at
        t = timesteps - 1 (tests); `generator` seeds the per-step noise."""
        self.make_schedule(timesteps)
        device = self.model.betas.device
        size = (batch_size, *shape)
        img = torch.randn(size, device=device, generator=generator) if x_T is None else x_T.float().contiguous()
        guided = unconditional_conditioning is not None and unconditional_guidance_scale != 1.
        ca_vecs, concat = self._conditioning(conditioning, unconditional_conditioning, guided)
        t_dev = torch.empty(2 * batch_size if guided else batch_size, dtype=torch.int64, device=device)
        intermediates = {"x_inter": [img], "pred_x0": [img]}
        total = self.ddim_timesteps.shape[0]
        with self.frozen_weights():
            for i, t in enumerate(self.ddim_timesteps[::-1]):
                t = int(t)
                t_dev.fill_(t)
                img, pred_x0 = self.p_sample(img, t_dev, ca_vecs, t, guided, float(unconditional_guidance_scale), temperature,
                                             generator=generator, concat=concat)
                if callback:
                    callback(i)
                if img_callback:
                    img_callback(pred_x0, i)
                if t % log_every_t == 0 or t == total - 1:
                    intermediates["x_inter"].append(img)
                    intermediates["pred_x0"].append(pred_x0)
        return img, intermediates

    def p_sample(self, x, t_dev, ca_vecs, t: int, guided: bool, scale: float, temperature=1., noise=None, generator=None,
                 concat=None):
        """One ancestral step at timestep t: UNet evaluation + fused CFG / x0 / posterior sample."""
        eps = self._eps(x, t_dev, ca_vecs, concat=concat)
        sigma = float(self.ddim_sigmas[t])
        if sigma > 0 and noise is None:
            noise = torch.randn(x.shape, device=x.device, generator=generator) * temperature
        return ops.ddim_step(x, eps, guided=guided, scale=scale, a_t=float(self.ddim_alphas[t]), a_prev=float(self.ddim_alphas_prev[t]),
                             sigma=sigma, sqrt_one_minus_at=float(self.ddim_sqrt_one_minus_alphas[t]),
                             noise=noise if sigma > 0 else None)
