"""Oracle (test infrastructure): CPU fp32 restatement of the reference denoiser.

Covers SURVEY.md §8 rows a1-a14: the UNet3DModel forward, the diffusion schedule, q_sample / p_losses,
the DDIM sampler with classifier-free guidance.  Every function names the reference lines it restates.
Pinned (oracle/validate_against_reference.py): against the reference modules AND against the real SDFusionText2ShapeModel
class built on the CPU (oracle/reference_diffusion_model.py): schedule, q_sample, p_losses, forward(), rel2shape.
All functions are pure: weights come in as a state dict with the reference's key names
(`diffusion_net.*`), nothing is cached, nothing touches CUDA.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# config/sdfusion-txt2shape.yaml:13-38 (the cross-attention denoiser of v2_full)
UNET_FULL = dict(image_size=16, in_channels=3, out_channels=3, model_channels=224, num_res_blocks=2,
                 attention_resolutions=(4, 2), channel_mult=(1, 2, 3), num_heads=8, dims=3,
                 transformer_depth=1, context_dim=1280)
# a small configuration with the same topology, for CPU-fast tests
UNET_TINY = dict(image_size=8, in_channels=3, out_channels=3, model_channels=32, num_res_blocks=1,
                 attention_resolutions=(4, 2), channel_mult=(1, 2, 3), num_heads=4, dims=3,
                 transformer_depth=1, context_dim=64)
# config/sdfusion-txt2shape_concat.yaml:13-38: conditioning concatenated on the channel axis
# (in_channels 4), AttentionBlock self-attention instead of the spatial transformer, `dims: 4` = Conv3d with isotropic
# stride-2 resampling (ldm_diffusion_util.py:251-252, openai_model_3d.py:150-155, 188)
UNET_CONCAT_FULL = dict(image_size=16, in_channels=4, out_channels=3, model_channels=224, num_res_blocks=2,
                        attention_resolutions=(4, 2), channel_mult=(1, 2, 3), num_heads=8, dims=4,
                        transformer_depth=1, context_dim=None, use_spatial_transformer=False)
UNET_CONCAT_TINY = dict(image_size=8, in_channels=4, out_channels=3, model_channels=32, num_res_blocks=1,
                        attention_resolutions=(4, 2), channel_mult=(1, 2, 3), num_heads=4, dims=4,
                        transformer_depth=1, context_dim=None, use_spatial_transformer=False)
# config/sdfusion-txt2shape.yaml:3-7
DIFFUSION = dict(timesteps=1000, linear_start=0.00085, linear_end=0.012)


# ------------------------------------------------------------------------------------------------
# network layout (openai_model_3d.py:558-728)
# ------------------------------------------------------------------------------------------------
def unet_layout(cfg: dict) -> Dict[str, list]:
    """Block structure of UNet3DModel as lists of layer descriptors.

    ("conv", cin, cout) | ("res", cin, cout) | ("st", ch, heads, d_head) | ("attn", ch, heads) | ("down", ch) | ("up", ch)
    use_spatial_transformer=False (the concat variant) puts an AttentionBlock where the transformer would be
    (openai_model_3d.py:593-598, 649-654, 690-697).
    """
    mc, mult, nrb = cfg["model_channels"], cfg["channel_mult"], cfg["num_res_blocks"]
    heads, attn_res = cfg["num_heads"], tuple(cfg["attention_resolutions"])
    use_st = cfg.get("use_spatial_transformer", True)
    att = (lambda c: ("st", c, heads, c // heads)) if use_st else (lambda c: ("attn", c, heads))
    inp: List[list] = [[("conv", cfg["in_channels"], mc)]]
    chans = [mc]
    ch, ds = mc, 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            layers = [("res", ch, m * mc)]
            ch = m * mc
            if ds in attn_res:
                layers.append(att(ch))                          # legacy=False: dim_head = ch // num_heads (:584-591)
            inp.append(layers)
            chans.append(ch)
        if level != len(mult) - 1:
            inp.append([("down", ch)])
            chans.append(ch)
            ds *= 2
    mid = [("res", ch, ch), att(ch), ("res", ch, ch)]
    out: List[list] = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, mc * m)]
            ch = mc * m
            if ds in attn_res:
                layers.append(att(ch))
            if level and i == nrb:
                layers.append(("up", ch))
                ds //= 2
            out.append(layers)
    return {"input": inp, "middle": mid, "output": out, "final_ch": ch}


def unet_param_shapes(cfg: dict, prefix: str = "diffusion_net.") -> Dict[str, Tuple[int, ...]]:
    """Every parameter of DiffusionUNet (network.py:11-22) by state-dict key -> shape."""
    mc, ted, ctx = cfg["model_channels"], cfg["model_channels"] * 4, cfg["context_dim"]
    shapes: Dict[str, Tuple[int, ...]] = {}

    def lin(name, i, o, bias=True):
        shapes[name + ".weight"] = (o, i)
        if bias:
            shapes[name + ".bias"] = (o,)

    def conv(name, i, o, k):
        shapes[name + ".weight"] = (o, i, k, k, k)
        shapes[name + ".bias"] = (o,)

    def norm(name, c):
        shapes[name + ".weight"] = (c,)
        shapes[name + ".bias"] = (c,)

    def layer(name, d):
        if d[0] == "conv":
            conv(name, d[1], d[2], 3)
        elif d[0] == "res":                                   # ResBlock (openai_model_3d.py:240-280)
            _, cin, cout = d
            norm(name + ".in_layers.0", cin); conv(name + ".in_layers.2", cin, cout, 3)
            lin(name + ".emb_layers.1", ted, cout)
            norm(name + ".out_layers.0", cout); conv(name + ".out_layers.3", cout, cout, 3)
            if cin != cout:
                conv(name + ".skip_connection", cin, cout, 1)
        elif d[0] == "st":                                    # SpatialTransformer3D (attention.py:306-328)
            _, ch, heads, dh = d
            inner = heads * dh
            norm(name + ".norm", ch); conv(name + ".proj_in", ch, inner, 1)
            tb = name + ".transformer_blocks.0"
            for a, cd in (("attn1", inner), ("attn2", ctx)):
                lin(f"{tb}.{a}.to_q", inner, inner, bias=False)
                lin(f"{tb}.{a}.to_k", cd, inner, bias=False)
                lin(f"{tb}.{a}.to_v", cd, inner, bias=False)
                lin(f"{tb}.{a}.to_out.0", inner, inner)
            lin(f"{tb}.ff.net.0.proj", inner, inner * 8)
            lin(f"{tb}.ff.net.2", inner * 4, inner)
            for n in ("norm1", "norm2", "norm3"):
                norm(f"{tb}.{n}", inner)
            conv(name + ".proj_out", inner, ch, 1)
        elif d[0] == "attn":                                  # AttentionBlock (openai_model_3d.py:324-352): Conv1d k=1
            _, ch, heads = d
            norm(name + ".norm", ch)
            shapes[name + ".qkv.weight"] = (3 * ch, ch, 1); shapes[name + ".qkv.bias"] = (3 * ch,)
            shapes[name + ".proj_out.weight"] = (ch, ch, 1); shapes[name + ".proj_out.bias"] = (ch,)
        elif d[0] == "down":
            conv(name + ".op", d[1], d[1], 3)
        elif d[0] == "up":
            conv(name + ".conv", d[1], d[1], 3)

    lay = unet_layout(cfg)
    lin(prefix + "time_embed.0", mc, ted); lin(prefix + "time_embed.2", ted, ted)
    for bi, block in enumerate(lay["input"]):
        for li, d in enumerate(block):
            layer(f"{prefix}input_blocks.{bi}.{li}", d)
    for li, d in enumerate(lay["middle"]):
        layer(f"{prefix}middle_block.{li}", d)
    for bi, block in enumerate(lay["output"]):
        for li, d in enumerate(block):
            layer(f"{prefix}output_blocks.{bi}.{li}", d)
    norm(prefix + "out.0", lay["final_ch"]); conv(prefix + "out.2", mc, cfg["out_channels"], 3)
    return shapes


# ------------------------------------------------------------------------------------------------
# building blocks
# ------------------------------------------------------------------------------------------------
def timestep_embedding(t: Tensor, dim: int, max_period: float = 10000.0) -> Tensor:
    """ldm_diffusion_util.py:174-194."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _gn(sd, name, x, eps):
    return F.group_norm(x.float(), 32, sd[name + ".weight"], sd[name + ".bias"], eps)


def _conv(sd, name, x, stride=1, padding=1):
    return F.conv3d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding)


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def res_block(sd, name: str, x: Tensor, emb: Tensor) -> Tensor:
    """ResBlock._forward, use_scale_shift_norm=False, no up/down (openai_model_3d.py:294-314)."""
    h = _conv(sd, name + ".in_layers.2", F.silu(_gn(sd, name + ".in_layers.0", x, 1e-5)))
    e = _lin(sd, name + ".emb_layers.1", F.silu(emb))
    h = h + e[:, :, None, None, None]
    h = _conv(sd, name + ".out_layers.3", F.silu(_gn(sd, name + ".out_layers.0", h, 1e-5)))   # dropout p=0
    skip = x if (name + ".skip_connection.weight") not in sd else _conv(sd, name + ".skip_connection", x, padding=0)
    return skip + h


def cross_attention(sd, name: str, x: Tensor, context: Optional[Tensor], heads: int) -> Tensor:
    """CrossAttention.forward without mask (attention.py:172-219); scale = dim_head ** -0.5 (:160)."""
    ctx = x if context is None else context
    q, k, v = _lin(sd, name + ".to_q", x), _lin(sd, name + ".to_k", ctx), _lin(sd, name + ".to_v", ctx)
    b, n, inner = q.shape
    d = inner // heads

    def split(t):
        return t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3).reshape(b * heads, t.shape[1], d)
    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("bid,bjd->bij", q, k) * (d ** -0.5)
    out = torch.einsum("bij,bjd->bid", sim.softmax(dim=-1), v)
    out = out.reshape(b, heads, n, d).permute(0, 2, 1, 3).reshape(b, n, inner)
    return _lin(sd, name + ".to_out.0", out)


def transformer_block(sd, name: str, x: Tensor, context: Tensor, heads: int) -> Tensor:
    """BasicTransformerBlock._forward (attention.py:237-245) with GEGLU feed-forward (:39-66)."""
    def ln(n, t):
        return F.layer_norm(t, (t.shape[-1],), sd[f"{name}.{n}.weight"], sd[f"{name}.{n}.bias"], 1e-5)
    x = cross_attention(sd, name + ".attn1", ln("norm1", x), None, heads) + x
    x = cross_attention(sd, name + ".attn2", ln("norm2", x), context, heads) + x
    a, gate = _lin(sd, name + ".ff.net.0.proj", ln("norm3", x)).chunk(2, dim=-1)
    x = _lin(sd, name + ".ff.net.2", a * F.gelu(gate)) + x
    return x


def spatial_transformer(sd, name: str, x: Tensor, context: Tensor, heads: int) -> Tensor:
    """SpatialTransformer3D.forward (attention.py:335-351): GN eps 1e-6, tokens in (d h w) order."""
    b, c, d, h, w = x.shape
    t = _conv(sd, name + ".proj_in", _gn(sd, name + ".norm", x, 1e-6), padding=0)
    inner = t.shape[1]
    t = t.reshape(b, inner, d * h * w).permute(0, 2, 1)
    t = transformer_block(sd, name + ".transformer_blocks.0", t, context, heads)
    t = t.permute(0, 2, 1).reshape(b, inner, d, h, w)
    return _conv(sd, name + ".proj_out", t, padding=0) + x


def attention_block(sd, name: str, x: Tensor, heads: int) -> Tensor:
    """AttentionBlock._forward (openai_model_3d.py:358-364) with QKVAttentionLegacy (:386-411): GroupNorm32 (eps 1e-5),
    1x1 Conv1d qkv, heads split BEFORE the q/k/v split, scale ch^-1/4 on q and on k, softmax in fp32, 1x1 proj_out, + x."""
    b, c = x.shape[:2]
    xf = x.reshape(b, c, -1)
    qkv = F.conv1d(_gn(sd, name + ".norm", xf, 1e-5), sd[name + ".qkv.weight"], sd[name + ".qkv.bias"])
    bs, width, length = qkv.shape
    ch = width // (3 * heads)
    q, k, v = qkv.reshape(bs * heads, ch * 3, length).split(ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    weight = torch.softmax(torch.einsum("bct,bcs->bts", q * scale, k * scale).float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", weight, v).reshape(bs, -1, length)
    h = F.conv1d(a, sd[name + ".proj_out.weight"], sd[name + ".proj_out.bias"])
    return (xf + h).reshape(x.shape)


def _run_layers(sd, prefix: str, block: Sequence[tuple], h: Tensor, emb: Tensor, context: Tensor, dims: int = 3) -> Tensor:
    """TimestepEmbedSequential.forward (openai_model_3d.py:119-127).  dims == 3 resamples H, W only; dims == 4 (the concat
    variant's Conv3d alias) takes the isotropic branches (:154-155, :188)."""
    for li, d in enumerate(block):
        name = f"{prefix}.{li}"
        if d[0] == "conv":
            h = _conv(sd, name, h)
        elif d[0] == "res":
            h = res_block(sd, name, h, emb)
        elif d[0] == "st":
            h = spatial_transformer(sd, name, h, context, d[2])
        elif d[0] == "attn":
            h = attention_block(sd, name, h, d[2])
        elif d[0] == "down":                                  # Downsample, dims=3: stride (1,2,2) (:188-192)
            h = _conv(sd, name + ".op", h, stride=(1, 2, 2) if dims == 3 else 2)
        elif d[0] == "up":                                    # Upsample, dims=3: (D, 2H, 2W) nearest (:150-157)
            if dims == 3:
                h = F.interpolate(h, (h.shape[2], h.shape[3] * 2, h.shape[4] * 2), mode="nearest")
            else:
                h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = _conv(sd, name + ".conv", h)
    return h


def unet_forward(sd: Dict[str, Tensor], cfg: dict, x: Tensor, t: Tensor, context: Optional[Tensor] = None,
                 prefix: str = "diffusion_net.", c_concat: Optional[Tensor] = None) -> Tensor:
    """DiffusionUNet.forward (network.py:24-30: 'crossattn' passes `context`, 'concat' runs the UNet on
    cat([x, c_concat], dim=1)) -> UNet3DModel.forward (openai_model_3d.py:752-789)."""
    lay = unet_layout(cfg)
    dims = cfg.get("dims", 3)
    if c_concat is not None:
        x = torch.cat([x, c_concat], dim=1)
    emb = timestep_embedding(t, cfg["model_channels"])
    emb = _lin(sd, prefix + "time_embed.2", F.silu(_lin(sd, prefix + "time_embed.0", emb)))
    hs = []
    h = x
    for bi, block in enumerate(lay["input"]):
        h = _run_layers(sd, f"{prefix}input_blocks.{bi}", block, h, emb, context, dims)
        hs.append(h)
    h = _run_layers(sd, f"{prefix}middle_block", lay["middle"], h, emb, context, dims)
    for bi, block in enumerate(lay["output"]):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_layers(sd, f"{prefix}output_blocks.{bi}", block, h, emb, context, dims)
    return _conv(sd, prefix + "out.2", F.silu(_gn(sd, prefix + "out.0", h, 1e-5)))


# ------------------------------------------------------------------------------------------------
# diffusion schedule, losses, sampler
# ------------------------------------------------------------------------------------------------
def register_schedule(timesteps: int = 1000, linear_start: float = 0.00085, linear_end: float = 0.012) -> Dict[str, Tensor]:
    """make_beta_schedule('linear') in float64 (ldm_diffusion_util.py:43-47) and the fp32 buffers of
    SDFusionText2ShapeModel.register_schedule (sdfusion_txt2shape_model.py:184-236), v_posterior = 0."""
    betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=torch.float64) ** 2).numpy()
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    f = lambda a: torch.tensor(a, dtype=torch.float32)
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    s = {
        "betas": f(betas), "alphas_cumprod": f(ac), "alphas_cumprod_prev": f(ac_prev),
        "sqrt_alphas_cumprod": f(np.sqrt(ac)), "sqrt_one_minus_alphas_cumprod": f(np.sqrt(1.0 - ac)),
        "log_one_minus_alphas_cumprod": f(np.log(1.0 - ac)), "sqrt_recip_alphas_cumprod": f(np.sqrt(1.0 / ac)),
        "sqrt_recipm1_alphas_cumprod": f(np.sqrt(1.0 / ac - 1)), "posterior_variance": f(post_var),
        "posterior_log_variance_clipped": f(np.log(np.maximum(post_var, 1e-20))),
        "posterior_mean_coef1": f(betas * np.sqrt(ac_prev) / (1.0 - ac)),
        "posterior_mean_coef2": f((1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)),
    }
    lvlb = s["betas"] ** 2 / (2 * s["posterior_variance"] * f(alphas) * (1 - s["alphas_cumprod"]))
    lvlb[0] = lvlb[1]
    s["lvlb_weights"] = lvlb
    return s


def q_sample(sched, x0: Tensor, t: Tensor, noise: Tensor) -> Tensor:
    """sdfusion_txt2shape_model.py:268-272."""
    shp = (-1,) + (1,) * (x0.dim() - 1)
    return sched["sqrt_alphas_cumprod"][t].reshape(shp) * x0 + sched["sqrt_one_minus_alphas_cumprod"][t].reshape(shp) * noise


def p_losses(sd, cfg, sched, x0: Tensor, cond: Tensor, t: Tensor, noise: Tensor, concat: bool = False):
    """sdfusion_txt2shape_model.py:311-345 (eps parameterisation, logvar = 0, l_simple_weight = 1,
    original_elbo_weight = 0).  concat=True: `cond` is the (B, 1, D, H, W) volume apply_model passes as c_concat (:281-283)."""
    x_noisy = q_sample(sched, x0, t, noise)
    out = unet_forward(sd, cfg, x_noisy, t, c_concat=cond) if concat else unet_forward(sd, cfg, x_noisy, t, cond)
    loss_simple = ((out - noise) ** 2).mean(dim=(1, 2, 3, 4))
    loss = loss_simple.mean()
    loss_vlb = (sched["lvlb_weights"][t] * loss_simple).mean()
    return x_noisy, noise, loss, {"loss_simple": loss_simple.mean(), "loss_vlb": loss_vlb, "loss_total": loss.detach()}


def ddim_schedule(sched, S: int, eta: float = 0.0):
    """make_ddim_timesteps('uniform') + make_ddim_sampling_parameters (ldm_diffusion_util.py:68-96) and
    DDIMSampler.make_schedule (samplers/ddim.py:28-57)."""
    T = sched["alphas_cumprod"].shape[0]
    c = T // S
    steps = np.asarray(list(range(0, T, c))) + 1
    ac = sched["alphas_cumprod"].numpy()   # the sampler indexes the fp32 buffer (ddim.py:41-46)
    alphas = ac[steps]
    alphas_prev = np.asarray([ac[0]] + ac[steps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return {"timesteps": steps, "alphas": alphas, "alphas_prev": alphas_prev, "sigmas": sigmas,
            "sqrt_one_minus_alphas": np.sqrt(1.0 - alphas)}


def p_sample_ddim(sd, cfg, dd, x: Tensor, c: Tensor, step: int, index: int, scale: float, uc: Optional[Tensor],
                  noise: Optional[Tensor] = None, concat: bool = False):
    """DDIMSampler.p_sample_ddim (samplers/ddim.py:182-244): CFG batch is [uncond; cond] (:206-210).  concat=True: c / uc are
    (B, 1, D, H, W) volumes that apply_model routes to c_concat (sdfusion_txt2shape_model.py:281-283)."""
    b = x.shape[0]
    t = torch.full((b,), int(step), dtype=torch.long)
    net = (lambda xx, tt, cc: unet_forward(sd, cfg, xx, tt, c_concat=cc)) if concat else (lambda xx, tt, cc: unet_forward(sd, cfg, xx, tt, cc))
    if uc is None or scale == 1.0:
        e_t = net(x, t, c)
    else:
        e_uc, e_c = net(torch.cat([x] * 2), torch.cat([t] * 2), torch.cat([uc, c])).chunk(2)
        e_t = e_uc + scale * (e_c - e_uc)
    a_t = torch.full((b, 1, 1, 1, 1), float(dd["alphas"][index]))
    a_prev = torch.full((b, 1, 1, 1, 1), float(dd["alphas_prev"][index]))
    sigma_t = torch.full((b, 1, 1, 1, 1), float(dd["sigmas"][index]))
    s1m = torch.full((b, 1, 1, 1, 1), float(dd["sqrt_one_minus_alphas"][index]))
    pred_x0 = (x - s1m * e_t) / a_t.sqrt()
    dir_xt = (1.0 - a_prev - sigma_t ** 2).sqrt() * e_t
    nz = sigma_t * (noise if noise is not None else torch.zeros_like(x))
    return a_prev.sqrt() * pred_x0 + dir_xt + nz, pred_x0, e_t


def ddim_sample(sd, cfg, sched, cond: Tensor, uc: Optional[Tensor], x_T: Tensor, S: int = 100, eta: float = 0.0,
                scale: float = 3.0, max_steps: Optional[int] = None):
    """DDIMSampler.ddim_sampling (samplers/ddim.py:126-179), eta = 0 path.  `max_steps` truncates the loop
    (tests only); returns (x, list of (x_prev, pred_x0, e_t) per step)."""
    dd = ddim_schedule(sched, S, eta)
    steps = np.flip(dd["timesteps"])
    total = steps.shape[0]
    x = x_T
    trace = []
    for i, step in enumerate(steps):
        if max_steps is not None and i >= max_steps:
            break
        index = total - i - 1
        x, pred_x0, e_t = p_sample_ddim(sd, cfg, dd, x, cond, int(step), index, scale, uc)
        trace.append((x, pred_x0, e_t))
    return x, trace


def p_sample_ddpm(sd, cfg, sched, x: Tensor, c: Tensor, t_int: int, scale: float, uc: Optional[Tensor], noise: Tensor):
    """One ancestral DDPM step (BASELINE cfg5) in the posterior-mean form, from the buffers the reference registers in
    register_schedule (sdfusion_txt2shape_model.py:214-224: sqrt_recip(m1)_alphas_cumprod, posterior_mean_coef1/2,
    posterior_log_variance_clipped) with DDIMSampler's guidance rule (samplers/ddim.py:206-210).
    PARITY UNPINNED for the step itself: the reference ships no ancestral sampler to run against -- this is
    the textbook DDPM posterior over the reference's own (pinned) schedule tables."""
    b = x.shape[0]
    t = torch.full((b,), int(t_int), dtype=torch.long)
    if uc is None or scale == 1.0:
        e_t = unet_forward(sd, cfg, x, t, c)
    else:
        e_uc, e_c = unet_forward(sd, cfg, torch.cat([x] * 2), torch.cat([t] * 2), torch.cat([uc, c])).chunk(2)
        e_t = e_uc + scale * (e_c - e_uc)
    x0 = sched["sqrt_recip_alphas_cumprod"][t_int] * x - sched["sqrt_recipm1_alphas_cumprod"][t_int] * e_t
    mean = sched["posterior_mean_coef1"][t_int] * x0 + sched["posterior_mean_coef2"][t_int] * x
    nonzero = 0.0 if t_int == 0 else 1.0
    return mean + nonzero * torch.exp(0.5 * sched["posterior_log_variance_clipped"][t_int]) * noise, x0, e_t
