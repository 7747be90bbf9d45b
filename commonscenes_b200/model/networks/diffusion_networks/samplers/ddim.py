"""DDIMSampler — the denoising time loop on the B200 kernels.

Drop-in for the reference's model/networks/diffusion_networks/samplers/ddim.py (make_schedule :28-57,
sample :60-123, ddim_sampling :126-179, p_sample_ddim :182-244).  What changes underneath:

  * the schedule tables are built once per (S, eta) on the host in the reference's float64/float32 recipe and
    never pushed to the device — the per-step scalars travel as kernel arguments;
  * classifier-free guidance does not materialise cat([x, x]): the UNet's stem reads x for both halves, and the
    context-only cross-attention vectors for [uncond; cond] are computed once per trajectory;
  * e = e_uc + s (e_c - e_uc), pred_x0 and x_prev are one kernel (cs_ddim_step) instead of ~12 torch ops and
    four torch.full allocations per step; randn is only drawn when sigma > 0;
  * the UNet forward of a step is captured in a CUDA graph on first use and replayed (~500 launches -> 1).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from ..... import ops
from ..ldm_diffusion_util import make_ddim_sampling_parameters, make_ddim_timesteps


class DDIMSampler(object):
    def __init__(self, model, schedule="linear", use_cuda_graph: bool = True, **kwargs):
        super().__init__()
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule
        self.use_cuda_graph = use_cuda_graph
        self.max_graphs = int(kwargs.get("max_graphs", 4))
        self._graphs = {}                               # (x shape, batch, weight generation) -> captured evaluation, LRU order
        self._frozen = False

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        self.ddim_timesteps = make_ddim_timesteps(ddim_discr_method=ddim_discretize, num_ddim_timesteps=ddim_num_steps,
                                                  num_ddpm_timesteps=self.ddpm_num_timesteps, verbose=verbose)
        ac = self.model.alphas_cumprod.detach().float().cpu().numpy()
        assert ac.shape[0] == self.ddpm_num_timesteps, "alphas have to be defined for each timestep"
        self.ddim_sigmas, self.ddim_alphas, self.ddim_alphas_prev = make_ddim_sampling_parameters(
            alphacums=ac, ddim_timesteps=self.ddim_timesteps, eta=ddim_eta, verbose=verbose)
        self.ddim_sqrt_one_minus_alphas = np.sqrt(1. - self.ddim_alphas)

    # ------------------------------------------------------------------------------------------
    def _unet(self):
        return self._df().diffusion_net

    def _df(self):
        return self.model.df_module if hasattr(self.model, "df_module") else self.model.df

    def _is_concat(self):
        return getattr(self._df(), "conditioning_key", None) == "concat"

    def _conditioning(self, cond, uncond, guided):
        """-> (ca_vecs, concat): the per-trajectory conditioning in kernel form.  Cross-attention: the context-only vectors
        of all blocks for [uncond; cond]; concat (network.py:25-27): the (2B | B, Cc, D, H, W) volume appended to x."""
        ctx = torch.cat([uncond, cond]) if guided else cond                          # [uncond; cond] (ddim.py:206-209)
        if self._is_concat():
            return None, ctx.float().contiguous()
        return self._unet().context_vectors(ctx), None

    def frozen_weights(self):
        """Context manager for a sampling trajectory: the UNet's packed-weight cache is validated ONCE (walking the 496
        parameters' versions costs ~0.4 ms of host time, which a 1000-step chain on a small per-rank batch cannot hide);
        weights must not change inside the block."""
        sampler = self

        class _Frozen:
            def __enter__(self_inner):
                sampler._unet()._ensure_packed()
                self_inner.prev, sampler._frozen = sampler._frozen, True

            def __exit__(self_inner, *exc):
                sampler._frozen = self_inner.prev
                return False
        return _Frozen()

    def _eps(self, x, t_dev, ca_vecs, concat=None):
        """One UNet evaluation for the whole (guided) batch, optionally through a cached CUDA graph."""
        unet = self._unet()
        if concat is not None:          # x_in = cat([x (repeated for [uncond; cond]), c_concat], dim=1)
            r = concat.shape[0] // x.shape[0]
            x = torch.cat([x.repeat(r, 1, 1, 1, 1) if r > 1 else x, concat], dim=1)
        # [uncond; cond] share x and t (t_dev is one value, or cat([t] * 2) in p_sample_ddim): shared conditioning-free prefix
        shared = concat is None and x.shape[0] < t_dev.shape[0]
        call = (lambda a, b, c: unet(a, b)) if concat is not None else (lambda a, b, c: unet(a, b, context_vecs=c, shared_prefix=shared))
        if not self.use_cuda_graph:
            return call(x, t_dev, ca_vecs)
        if not self._frozen:            # inside frozen_weights() the packing was checked once for the whole trajectory
            unet._ensure_packed()
        key = (tuple(x.shape), t_dev.shape[0], unet._pack_generation)
        g = self._graphs.get(key)
        if g is None:
            sx, st, sc = x.clone(), t_dev.clone(), (None if ca_vecs is None else ca_vecs.clone())
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):                      # warm-up: lazy packing, attribute setup, allocator pools
                    call(sx, st, sc)
            torch.cuda.current_stream().wait_stream(s)
            graph = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(graph):
                out = call(sx, st, sc)
            g = {"graph": graph, "x": sx, "t": st, "c": sc, "out": out, "launches": ops.launch_count() - n0}
            # a few resident graphs, least recently used first out: the ragged last mini-batch of a scene, or a trainer
            # alternating validation batch sizes, must not re-capture (two warm-up evaluations + capture) on every call;
            # graphs of an older weight packing can never be hit again and go first
            for k in [k for k in self._graphs if k[2] != unet._pack_generation]:
                del self._graphs[k]
            while len(self._graphs) >= self.max_graphs:
                del self._graphs[next(iter(self._graphs))]
            self._graphs[key] = g
        else:
            self._graphs[key] = self._graphs.pop(key)   # mark as most recently used
        g["x"].copy_(x); g["t"].copy_(t_dev)
        if ca_vecs is not None:
            g["c"].copy_(ca_vecs)
        g["graph"].replay()
        self.kernels_per_eval = g["launches"]
        return g["out"]

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0., score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, mm_cls_free=False, **kwargs):
        if mask is not None or score_corrector is not None or mm_cls_free or quantize_x0 or noise_dropout > 0.:
            raise NotImplementedError("mask / score_corrector / mm_cls_free / quantize_x0 / noise_dropout are not used by "
                                      "the shape branch (sdfusion_txt2shape_model.py:500-508)")
        if conditioning is not None and not isinstance(conditioning, dict) and conditioning.shape[0] != batch_size:
            print(f"Warning: Got {conditioning.shape[0]} conditionings but batch-size is {batch_size}")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        size = (batch_size, *shape)
        return self.ddim_sampling(conditioning, size, callback=callback, img_callback=img_callback, x_T=x_T,
                                  log_every_t=log_every_t, temperature=temperature,
                                  unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning)

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, x_T=None, callback=None, timesteps=None, img_callback=None, log_every_t=100,
                      temperature=1., unconditional_guidance_scale=1., unconditional_conditioning=None, **kwargs):
        device = self.model.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T.float().contiguous()
        if timesteps is None:
            timesteps = self.ddim_timesteps
        else:
            subset_end = int(min(timesteps / self.ddim_timesteps.shape[0], 1) * self.ddim_timesteps.shape[0]) - 1
            timesteps = self.ddim_timesteps[:subset_end]
        intermediates = {"x_inter": [img], "pred_x0": [img]}
        time_range = np.flip(timesteps)
        total_steps = timesteps.shape[0]

        guided = unconditional_conditioning is not None and unconditional_guidance_scale != 1.
        ca_vecs, concat = self._conditioning(cond, unconditional_conditioning, guided)     # once per trajectory
        t_dev = torch.empty(2 * b if guided else b, dtype=torch.int64, device=device)
        with self.frozen_weights():
            for i, step in enumerate(time_range):
                index = total_steps - i - 1
                t_dev.fill_(int(step))
                img, pred_x0 = self._step(img, t_dev, ca_vecs, index, guided, float(unconditional_guidance_scale), temperature,
                                          concat=concat)
                if callback:
                    callback(i)
                if img_callback:
                    img_callback(pred_x0, i)
                if index % log_every_t == 0 or index == total_steps - 1:
                    intermediates["x_inter"].append(img)
                    intermediates["pred_x0"].append(pred_x0)
        return img, intermediates

    def _step(self, x, t_dev, ca_vecs, index, guided, scale, temperature=1., want_pred_x0=True, concat=None):
        """UNet evaluation (graph replay) + fused CFG / pred_x0 / x_prev update (cs_ddim_step)."""
        eps = self._eps(x, t_dev, ca_vecs, concat=concat)
        sigma = float(self.ddim_sigmas[index])
        noise = torch.randn_like(x) * temperature if sigma > 0 else None
        return ops.ddim_step(x, eps, guided=guided, scale=scale, a_t=float(self.ddim_alphas[index]),
                             a_prev=float(self.ddim_alphas_prev[index]), sigma=sigma,
                             sqrt_one_minus_at=float(self.ddim_sqrt_one_minus_alphas[index]), noise=noise,
                             want_pred_x0=want_pred_x0)

    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None, mm_cls_free=False):
        """One reverse step (ddim.py:182-244); returns (x_prev, pred_x0).  `t` is the (b,) timestep tensor."""
        if use_original_steps or quantize_denoised or score_corrector is not None or mm_cls_free or noise_dropout > 0.:
            raise NotImplementedError("options unused by the shape branch")
        guided = unconditional_conditioning is not None and unconditional_guidance_scale != 1.
        ca_vecs, concat = self._conditioning(c, unconditional_conditioning, guided)
        tt = (torch.cat([t] * 2) if guided else t).to(torch.int64).contiguous()
        return self._step(x.float().contiguous(), tt, ca_vecs, index, guided, float(unconditional_guidance_scale), temperature,
                          concat=concat)
