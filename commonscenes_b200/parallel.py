"""Object sharding across the GPUs of one box.

Every object's denoising trajectory depends only on its own (x_T, c, uc): the UNet and VQ-VAE have no cross-sample
ops, so sampling shards by contiguous blocks of objects with NO collective inside the DDIM loop; the only
communication is one all_gather of the decoded SDFs (1 MiB fp32 per object) at the end.  The tiny GCN that produces the
conditioning is evaluated replicated on the full graph on every rank (its BatchNorm statistics span the whole graph
batch), then each rank slices its objects.  One process per GPU, torch.distributed (NCCL on GPUs; gloo in CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def partition(num_objects: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous block partition: the first (num_objects % world_size) ranks get one extra object."""
    base, extra = divmod(num_objects, world_size)
    bounds, start = [], 0
    for r in range(world_size):
        n = base + (1 if r < extra else 0)
        bounds.append((start, start + n))
        start += n
    return bounds


def shard(t: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    lo, hi = partition(t.shape[0], world_size)[rank]
    return t[lo:hi]


def gather_objects(local: torch.Tensor, num_objects: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """all_gather of per-rank object blocks (ragged: padded to the largest block) back into object order."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    bounds = partition(num_objects, world)
    max_n = max(hi - lo for lo, hi in bounds)
    padded = local.new_zeros((max_n,) + tuple(local.shape[1:]))
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    return torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, bounds)], dim=0)


@torch.no_grad()
def rel2shape_sharded(diff_model, data: dict, ddim_steps: int = 100, ddim_eta: float = 0.0, uc_scale: float = 3.0,
                      seed: Optional[int] = None, group: Optional[dist.ProcessGroup] = None, **sampler_kw) -> torch.Tensor:
    """SDFusionText2ShapeModel.rel2shape with the objects of `data` ('sdf', 'rel', 'uc': one row per object) split
    across the ranks of `group`; every rank returns the full (O, 1, R, R, R) result.  All ranks must pass the same
    `data` and `seed` (the reference shares one x_T across objects, sdfusion_txt2shape_model.py:487-491)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return diff_model.rel2shape(data, ddim_steps=ddim_steps, ddim_eta=ddim_eta, uc_scale=uc_scale, seed=seed, **sampler_kw)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = data["rel"].shape[0]
    if seed is None:        # the reference seeds x_T from the clock (:487): draw ONE seed on rank 0 so every rank shares x_T
        import time
        box = torch.tensor([int(time.time())], dtype=torch.int64, device=data["rel"].device)
        dist.broadcast(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        seed = int(box.item())
    lo, hi = partition(n, world)[rank]
    if hi > lo:
        local = {k: v[lo:hi] for k, v in data.items()}
        out = diff_model.rel2shape(local, ddim_steps=ddim_steps, ddim_eta=ddim_eta, uc_scale=uc_scale, seed=seed, **sampler_kw)
    else:   # more ranks than objects: this rank only takes part in the gather
        r = diff_model.z_shape[-1] * 4
        out = torch.zeros((0, 1, r, r, r), dtype=torch.float32, device=data["rel"].device)
    return gather_objects(out, n, group)


# ----------------------------------------------------------------------------------------------------------------------
# CFG-pair split (SURVEY.md §8e "Sampling partition", cfg5): a scene has fewer objects than a box has GPUs (10 objects
# on 8 ranks leave a 2:1 imbalance), but the unconditional and the conditional evaluation of one object are independent
# too.  The 2·O UNet forwards of a guided step -- ordered [uncond_0..uncond_{O-1}; cond_0..cond_{O-1}] like the
# reference's batch (samplers/ddim.py:206-209) -- are block-partitioned over the ranks (20 forwards -> 3,3,3,3,2,2,2,2).
# This path HAS an exchange step: the guidance e = e_uc + s (e_c - e_uc) needs both halves of an object, so every step
# ends with one all_gather of the eps rows (48 KiB fp32 each); after it every rank applies the (tiny, fused) update to all
# O latents, so x never travels.  The VQ-VAE decode at the end is object-sharded and gathered as in rel2shape_sharded.
# ----------------------------------------------------------------------------------------------------------------------
def pair_units(num_objects: int, world_size: int) -> List[Tuple[int, int]]:
    """[lo, hi) ranges over the 2 * num_objects forwards of a guided step, one per rank."""
    return partition(2 * num_objects, world_size)


def exchange_eps(local_eps: torch.Tensor, num_objects: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """all_gather of each rank's eps rows back into the (2 * O, ...) [uncond; cond] batch (ragged blocks are padded)."""
    return gather_objects(local_eps, 2 * num_objects, group)


@torch.no_grad()
def rel2shape_pair_sharded(diff_model, data: dict, ddim_steps: int = 100, ddim_eta: float = 0.0, uc_scale: float = 3.0,
                           seed: Optional[int] = None, group: Optional[dist.ProcessGroup] = None, sampler: str = "ddim",
                           ddpm_timesteps: Optional[int] = None, return_latent: bool = False):
    """SDFusionText2ShapeModel.rel2shape with the 2 * O forwards of every guided step split across the ranks of `group`
    (see above).  Same arguments and result as rel2shape_sharded; all ranks must pass the same `data` (and `seed`)."""
    from . import ops
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return diff_model.rel2shape(data, ddim_steps=ddim_steps, ddim_eta=ddim_eta, uc_scale=uc_scale, seed=seed,
                                    sampler=sampler, ddpm_timesteps=ddpm_timesteps, return_latent=return_latent)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    m = diff_model
    m.switch_eval()
    m.set_input(data)
    dev = m.rel.device
    n = m.rel.shape[0]
    if seed is None:
        import time
        box = torch.tensor([int(time.time())], dtype=torch.int64, device=dev)
        dist.broadcast(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        seed = int(box.item())
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))                           # same stream on every rank: x_T and the per-step noise agree
    x = torch.randn((1, *m.z_shape), device=dev, generator=gen).repeat(n, 1, 1, 1, 1)       # one shared x_T (reference :487-491)
    if sampler == "ddpm":
        if getattr(m, "ddpm_sampler", None) is None:
            from .model.networks.diffusion_networks.samplers.ddpm import DDPMSampler
            m.ddpm_sampler = DDPMSampler(m)
        s = m.ddpm_sampler
        s.make_schedule(ddpm_timesteps)
        order = [(int(t), int(t)) for t in s.ddim_timesteps[::-1]]                        # (timestep, table index)
    elif sampler == "ddim":
        s = m.ddim_sampler
        s.make_schedule(ddim_num_steps=ddim_steps, ddim_eta=ddim_eta, verbose=False)
        total = s.ddim_timesteps.shape[0]
        order = [(int(t), total - i - 1) for i, t in enumerate(s.ddim_timesteps[::-1])]
    else:
        raise ValueError(f"unknown sampler '{sampler}' (ddim | ddpm)")
    lo, hi = pair_units(n, world)[rank]
    units = torch.arange(lo, hi, device=dev)
    ca_all, concat = s._conditioning(m.rel, m.uc_rel, True)                                 # rows [uncond; cond]
    if concat is not None:
        raise NotImplementedError("the CFG-pair split is built for the cross-attention conditioning of v2_full")
    ca = ca_all[lo:hi].contiguous()
    obj_of_unit = units % n
    t_dev = torch.empty(hi - lo, dtype=torch.int64, device=dev)
    # the exchange step: every rank's eps rows, padded to the largest block, land in ONE pre-allocated buffer through a
    # single all_gather_into_tensor; `take` maps the (world * max_n) padded rows back to the (2 n) [uncond; cond] order
    bounds = pair_units(n, world)
    max_n = max(b_hi - b_lo for b_lo, b_hi in bounds)
    row_shape = tuple(x.shape[1:])
    send = torch.zeros((max_n,) + row_shape, dtype=torch.float32, device=dev)
    recv = torch.empty((world * max_n,) + row_shape, dtype=torch.float32, device=dev)
    take = torch.tensor([r * max_n + i for r, (b_lo, b_hi) in enumerate(bounds) for i in range(b_hi - b_lo)], dtype=torch.int64, device=dev)
    with s.frozen_weights():
        for step, index in order:
            if hi > lo:
                t_dev.fill_(step)
                send[:hi - lo].copy_(s._eps(x.index_select(0, obj_of_unit), t_dev, ca))
            dist.all_gather_into_tensor(recv, send, group=group)                                 # the path's collective
            eps = recv.index_select(0, take)
            sigma = float(s.ddim_sigmas[index])
            noise = torch.randn(x.shape, device=dev, generator=gen) if sigma > 0 else None
            x, _ = ops.ddim_step(x, eps, guided=True, scale=float(uc_scale), a_t=float(s.ddim_alphas[index]),
                                 a_prev=float(s.ddim_alphas_prev[index]), sigma=sigma,
                                 sqrt_one_minus_at=float(s.ddim_sqrt_one_minus_alphas[index]), noise=noise, want_pred_x0=False)
    olo, ohi = partition(n, world)[rank]
    if ohi > olo:
        dec = m.vqvae_module.decode_no_quant(x[olo:ohi].contiguous())
    else:
        r = m.z_shape[-1] * 4
        dec = torch.zeros((0, 1, r, r, r), dtype=torch.float32, device=dev)
    out = gather_objects(dec, n, group)
    return (out, x) if return_latent else out
