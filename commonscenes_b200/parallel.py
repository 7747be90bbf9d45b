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
