"""SDFusionText2ShapeModel — the shape-branch wrapper (schedule, q_sample / p_losses, rel2shape) on the B200 path.

API mirror of the reference's model/sdfusion_txt2shape_model.py:51-703 for everything its callers use
(Sg2ScVAEModel: VAEGAN_V2FULL.py:94,153,318,379,519-521,615,644,692-694; train_3dfront.py:350-351,399,428,443;
VAE.py:140-157): attributes df / df_module / vqvae / vqvae_module / trainable_params / z_shape / num_timesteps /
betas / alphas_cumprod(_prev) / device / ddim_steps / uc_scale, methods set_input, set_requires_grad, switch_train /
switch_eval, q_sample, apply_model, p_losses, forward, update_loss, get_current_errors, gen_shape_after_foward,
rel2shape, save, load_ckpt, state-dict layout {'vqvae','df','global_step'[,'opt']}.  Mesh rendering / tensorboard
visuals (pytorch3d, mcubes) are outside the hot path and not provided.

Differences underneath: rel2shape runs ALL objects in one batch (the reference's mini-batches of 7 exist for A100
memory, :493; objects are independent so results are identical), the x_T noise may be seeded, rel2shape can also run the
1000-step ancestral sampler (samplers/ddpm.py), and `loss_df` from forward() back-propagates through explicit CUDA backward
kernels (autograd bridge in UNet3DModel.forward; native step in commonscenes_b200/train.py).
"""
from __future__ import annotations

import os
import time
from collections import OrderedDict
from functools import partial

import numpy as np
import torch

from .. import ops
from .base_model import BaseModel
from .model_utils import load_vqvae
from .networks.diffusion_networks.ldm_diffusion_util import extract_into_tensor, make_beta_schedule
from .networks.diffusion_networks.network import DiffusionUNet
from .networks.diffusion_networks.samplers.ddim import DDIMSampler

# config/sdfusion-txt2shape.yaml (values restated; a yaml path in opt.network.df_cfg overrides them)
DF_MODEL_PARAMS = dict(linear_start=0.00085, linear_end=0.012, conditioning_key="crossattn", timesteps=1000, scale_factor=0.18215)
UNET_PARAMS = dict(image_size=16, in_channels=3, out_channels=3, model_channels=224, num_res_blocks=2,
                   attention_resolutions=[4, 2], channel_mult=[1, 2, 3], num_heads=8, dims=3, use_spatial_transformer=True,
                   transformer_depth=1, context_dim=1280, use_checkpoint=True, legacy=False)
# config/sdfusion-txt2shape_concat.yaml: conditioning concatenated as a 4th latent channel
DF_MODEL_PARAMS_CONCAT = dict(DF_MODEL_PARAMS, conditioning_key="concat")
UNET_PARAMS_CONCAT = dict(image_size=16, in_channels=4, out_channels=3, model_channels=224, num_res_blocks=2,
                          attention_resolutions=[4, 2], channel_mult=[1, 2, 3], num_heads=8, dims=4, use_spatial_transformer=False,
                          transformer_depth=1, context_dim=None, use_checkpoint=True, legacy=False)
# config/vqvae_snet.yaml
VQ_CONF = dict(model=dict(params=dict(embed_dim=3, n_embed=8192, ddconfig=dict(
    double_z=False, z_channels=3, resolution=64, in_channels=1, out_ch=1, ch=64, ch_mult=[1, 2, 4], num_res_blocks=1,
    attn_resolutions=[], dropout=0.0))))


class _Cfg(dict):
    """dict with attribute access (stands in for OmegaConf nodes, which also work)."""
    __getattr__ = dict.get

    @staticmethod
    def wrap(x):
        if isinstance(x, dict):
            return _Cfg({k: _Cfg.wrap(v) for k, v in x.items()})
        return x


def _load_yaml(path):
    import yaml
    with open(path) as f:
        return _Cfg.wrap(yaml.safe_load(f))


def default_opt(device="cuda", vq_ckpt=None, df_cfg=None, vq_cfg=None, ckpt_dir=None, conditioning_key="crossattn"):
    """The fields of config/v2_full.yaml (conditioning_key="concat": config/v2_full_concat.yaml) that the shape branch reads.
    `df_cfg` (a yaml path, e.g. the reference's config/sdfusion-txt2shape[_concat].yaml) overrides the built-in values."""
    return _Cfg.wrap(dict(hyper=dict(batch_size=4 if conditioning_key == "crossattn" else 32, isTrain=True, device=device, distributed=0),
                          network=dict(df_cfg=df_cfg, vq_cfg=vq_cfg, vq_ckpt=vq_ckpt, ddim_steps=100, ddim_eta=0.0, uc_scale=3.0,
                                       conditioning_key=conditioning_key),
                          misc=dict(debug=0, seed=111, local_rank=0), ckpt_dir=ckpt_dir))


def diffusion_schedule(timesteps=1000, linear_start=0.00085, linear_end=0.012, v_posterior=0.0):
    """register_schedule (sdfusion_txt2shape_model.py:184-236): float64 host math, fp32 tables."""
    betas = make_beta_schedule("linear", timesteps, linear_start=linear_start, linear_end=linear_end)
    alphas = 1. - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1., ac[:-1])
    f = partial(torch.tensor, dtype=torch.float32)
    post_var = (1 - v_posterior) * betas * (1. - ac_prev) / (1. - ac) + v_posterior * betas
    s = dict(betas=f(betas), alphas_cumprod=f(ac), alphas_cumprod_prev=f(ac_prev), sqrt_alphas_cumprod=f(np.sqrt(ac)),
             sqrt_one_minus_alphas_cumprod=f(np.sqrt(1. - ac)), log_one_minus_alphas_cumprod=f(np.log(1. - ac)),
             sqrt_recip_alphas_cumprod=f(np.sqrt(1. / ac)), sqrt_recipm1_alphas_cumprod=f(np.sqrt(1. / ac - 1)),
             posterior_variance=f(post_var), posterior_log_variance_clipped=f(np.log(np.maximum(post_var, 1e-20))),
             posterior_mean_coef1=f(betas * np.sqrt(ac_prev) / (1. - ac)),
             posterior_mean_coef2=f((1. - ac_prev) * np.sqrt(alphas) / (1. - ac)))
    lvlb = s["betas"] ** 2 / (2 * s["posterior_variance"] * f(alphas) * (1 - s["alphas_cumprod"]))
    lvlb[0] = lvlb[1]
    s["lvlb_weights"] = lvlb
    return s


class SDFusionText2ShapeModel(BaseModel):
    def __init__(self, opt=None):
        super().__init__()
        opt = default_opt() if opt is None else opt
        BaseModel.initialize(self, opt)
        self.isTrain = opt.hyper.isTrain
        self.model_name = self.name()
        self.device = opt.hyper.device

        if opt.network.df_cfg:
            df_conf = _load_yaml(opt.network.df_cfg)
        elif opt.network.conditioning_key == "concat":
            df_conf = _Cfg.wrap(dict(model=dict(params=DF_MODEL_PARAMS_CONCAT), unet=dict(params=UNET_PARAMS_CONCAT)))
        else:
            df_conf = _Cfg.wrap(dict(model=dict(params=DF_MODEL_PARAMS), unet=dict(params=UNET_PARAMS)))
        vq_conf = _load_yaml(opt.network.vq_cfg) if opt.network.vq_cfg else _Cfg.wrap(VQ_CONF)
        ddconfig = vq_conf.model.params.ddconfig
        z_sp = ddconfig.resolution // (2 ** (len(ddconfig.ch_mult) - 1))
        self.z_shape = (ddconfig.z_channels, z_sp, z_sp, z_sp)

        unet_params = dict(df_conf.unet.params)
        unet_params.setdefault("use_spatial_transformer", df_conf.model.params.conditioning_key != "concat")
        self.df = DiffusionUNet(unet_params, vq_conf=vq_conf, conditioning_key=df_conf.model.params.conditioning_key)
        self.df.to(self.device)
        self.init_diffusion_params(uc_scale=3., df_model_params=df_conf.model.params)
        self.ddim_sampler = DDIMSampler(self)
        self.vqvae = load_vqvae(vq_conf, vq_ckpt=opt.network.vq_ckpt, opt=opt)
        self.trainable_params = [p for p in self.df.parameters() if p.requires_grad]
        self.df_module = self.df
        self.vqvae_module = self.vqvae
        self.ddim_steps = 7 if (opt.misc and opt.misc.debug == 1) else 100

    def name(self):
        return "SDFusion-Text2Shape-Model"

    # ---- diffusion parameters ------------------------------------------------------------------
    def init_diffusion_params(self, uc_scale=3., df_model_params=None, opt=None):
        p = df_model_params if df_model_params is not None else _Cfg(DF_MODEL_PARAMS)
        self.parameterization = "eps"
        self.learn_logvar = False
        self.v_posterior = 0.
        self.original_elbo_weight = 0.
        self.l_simple_weight = 1.
        self.register_schedule(timesteps=p.timesteps, linear_start=p.linear_start, linear_end=p.linear_end)
        self.logvar = torch.full(fill_value=0., size=(self.num_timesteps,))
        self.uc_scale = uc_scale

    def register_schedule(self, given_betas=None, beta_schedule="linear", timesteps=1000, linear_start=1e-4, linear_end=2e-2,
                          cosine_s=8e-3):
        if given_betas is not None or beta_schedule != "linear":
            raise NotImplementedError("only the linear schedule of config/sdfusion-txt2shape.yaml is used")
        s = diffusion_schedule(timesteps, linear_start, linear_end, self.v_posterior)
        self.num_timesteps = int(timesteps)
        self.linear_start, self.linear_end = linear_start, linear_end
        for k, v in s.items():
            setattr(self, k, v.to(self.device))

    # ---- inputs / modes ------------------------------------------------------------------------
    def set_input(self, input=None, max_sample=None):
        self.x = input["sdf"]
        self.rel = input["rel"]
        self.uc_rel = input["uc"]
        if self.df.conditioning_key == "concat":           # reference :246-248: the 4096-d vector becomes a latent channel
            B = self.x.shape[0]
            self.rel = self.rel.view(B, -1, *self.z_shape[1:])
            self.uc_rel = self.uc_rel.view(B, -1, *self.z_shape[1:])
        if max_sample is not None:
            self.x, self.rel, self.uc_rel = self.x[:max_sample], self.rel[:max_sample], self.uc_rel[:max_sample]
        self.tocuda(var_names=["x"])

    def switch_train(self):
        self.df.train()

    def switch_eval(self):
        self.df.eval()
        self.vqvae.eval()

    # ---- diffusion forward ---------------------------------------------------------------------
    def q_sample(self, x_start, t, noise=None):
        noise = torch.randn_like(x_start) if noise is None else noise
        return ops.q_sample(x_start.float().contiguous(), noise.float().contiguous(), t.to(torch.int64).contiguous(),
                            self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod)

    def apply_model(self, x_noisy, t, cond, return_ids=False):
        if not isinstance(cond, dict):
            key = "c_concat" if self.df_module.conditioning_key == "concat" else "c_crossattn"      # reference :281-283
            cond = {key: cond if isinstance(cond, list) else [cond]}
        out = self.df(x_noisy, t, **cond)
        return out[0] if isinstance(out, tuple) and not return_ids else out

    def get_loss(self, pred, target, loss_type="l2", mean=True):
        if loss_type == "l1":
            loss = (target - pred).abs()
            return loss.mean() if mean else loss
        if loss_type == "l2":
            return torch.nn.functional.mse_loss(target, pred, reduction="mean" if mean else "none")
        raise NotImplementedError(f"unknown loss type '{loss_type}'")

    def p_losses(self, x_start, cond, t, noise=None):
        noise = torch.randn_like(x_start) if noise is None else noise
        x_noisy = self.q_sample(x_start=x_start, t=t, noise=noise)
        model_output = self.apply_model(x_noisy, t, cond)
        target = noise
        loss_dict = {}
        loss_simple = self.get_loss(model_output, target, mean=False).mean([1, 2, 3, 4])
        loss_dict["loss_simple"] = loss_simple.mean()
        logvar_t = self.logvar[t.cpu()].to(self.device)
        loss = self.l_simple_weight * (loss_simple / torch.exp(logvar_t) + logvar_t).mean()
        loss_vlb = (self.lvlb_weights[t] * loss_simple).mean()
        loss_dict["loss_vlb"] = loss_vlb
        loss = loss + self.original_elbo_weight * loss_vlb
        loss_dict["loss_total"] = loss.clone().detach().mean()
        return x_noisy, target, loss, loss_dict

    def forward(self):
        """One training forward (reference :348-365): frozen VQ-VAE encode, random t, p_losses.  `loss_df` carries a grad_fn
        when the denoiser's parameters (or the conditioning) require grad."""
        self.switch_train()
        with torch.no_grad():
            z = self.vqvae(self.x, forward_no_quant=True, encode_only=True)
        t = torch.randint(0, self.num_timesteps, (z.shape[0],), device=self.device).long()
        z_noisy, target, loss, loss_dict = self.p_losses(z, self.rel, t)
        self.loss_df = loss
        self.loss_dict = loss_dict

    def backward(self, retain_graph=False):
        """reference :568-580 (loss-dict reduction across ranks is the caller's `reduce_loss_dict`; values are unchanged
        on a single process)."""
        self.update_loss()
        self.loss.backward(retain_graph=retain_graph)

    def optimize_parameters(self, total_steps=0):
        """reference :591-600"""
        self.set_requires_grad([self.df], requires_grad=True)
        self.forward()
        self.optimizer.zero_grad()
        self.backward()
        self.optimizer.step()

    def update_loss(self):
        self.loss = self.loss_df
        self.loss_total = self.loss_dict["loss_total"]
        self.loss_simple = self.loss_dict["loss_simple"]
        self.loss_vlb = self.loss_dict["loss_vlb"]

    def get_current_errors(self):
        return OrderedDict([("total", self.loss_total.mean().data), ("simple", self.loss_simple.mean().data),
                            ("vlb", self.loss_vlb.mean().data)])

    # ---- sampling ------------------------------------------------------------------------------
    @torch.no_grad()
    def gen_shape_after_foward(self, num_obj, ddim_steps=None, uc_scale=None, ddim_eta=0.):
        self.switch_eval()
        ddim_steps = self.ddim_steps if ddim_steps is None else ddim_steps
        uc_scale = self.uc_scale if uc_scale is None else uc_scale
        B = self.rel[:num_obj].shape[0]
        samples, _ = self.ddim_sampler.sample(S=ddim_steps, batch_size=B, shape=self.z_shape, conditioning=self.rel[:num_obj],
                                              verbose=False, unconditional_guidance_scale=uc_scale,
                                              unconditional_conditioning=self.uc_rel[:num_obj], eta=ddim_eta)
        self.gen_df = self.vqvae_module.decode_no_quant(samples)
        self.switch_train()

    @torch.no_grad()
    def rel2shape(self, data, ddim_steps=100, ddim_eta=0.0, uc_scale=None, seed=None, return_latent=False, sampler="ddim",
                  ddpm_timesteps=None):
        """Scene-graph conditioning -> (O, 1, R, R, R) SDFs.  One shared x_T for all objects (reference :487-491).
        sampler="ddpm": ancestral sampling (BASELINE cfg5; samplers/ddpm.py) instead of DDIM: `ddim_steps` is ignored and
        the chain runs over all num_timesteps, or starts at t = ddpm_timesteps - 1 when `ddpm_timesteps` is given."""
        self.switch_eval()
        self.set_input(data)
        ddim_steps = self.ddim_steps if ddim_steps is None else ddim_steps
        uc_scale = self.uc_scale if uc_scale is None else uc_scale
        B = self.rel.shape[0]
        gen = torch.Generator(device=self.device)
        gen.manual_seed(int(time.time()) if seed is None else int(seed))
        noise = torch.randn((1, *self.z_shape), device=self.device, generator=gen).repeat(B, 1, 1, 1, 1)
        if sampler == "ddpm":
            if getattr(self, "ddpm_sampler", None) is None:
                from .networks.diffusion_networks.samplers.ddpm import DDPMSampler
                self.ddpm_sampler = DDPMSampler(self)
            samples, _ = self.ddpm_sampler.sample(batch_size=B, shape=self.z_shape, conditioning=self.rel, x_T=noise,
                                                  unconditional_guidance_scale=uc_scale, unconditional_conditioning=self.uc_rel,
                                                  timesteps=ddpm_timesteps, generator=gen)
        elif sampler == "ddim":
            samples, _ = self.ddim_sampler.sample(S=ddim_steps, batch_size=B, shape=self.z_shape, conditioning=self.rel, x_T=noise,
                                                  verbose=False, unconditional_guidance_scale=uc_scale,
                                                  unconditional_conditioning=self.uc_rel, eta=ddim_eta)
        else:
            raise ValueError(f"unknown sampler '{sampler}' (ddim | ddpm)")
        self.gen_df = self.vqvae_module.decode_no_quant(samples)
        return (self.gen_df, samples) if return_latent else self.gen_df

    # ---- checkpoints ---------------------------------------------------------------------------
    def save(self, label, global_step, save_opt=False):
        state_dict = {"vqvae": self.vqvae_module.state_dict(), "df": self.df_module.state_dict(), "global_step": global_step}
        if save_opt:
            state_dict["opt"] = self.optimizer.state_dict()
        path = os.path.join(self.opt.ckpt_dir or ".", "df_%s.pth" % label)
        torch.save(state_dict, path)
        return path

    def load_ckpt(self, ckpt, load_opt=False):
        state_dict = torch.load(ckpt, map_location=lambda storage, loc: storage) if isinstance(ckpt, str) else ckpt
        self.vqvae.load_state_dict(state_dict["vqvae"])
        self.df.load_state_dict(state_dict["df"])
        if load_opt:
            self.optimizer.load_state_dict(state_dict["opt"])
