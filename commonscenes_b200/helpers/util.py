"""Mirror of the reference's `helpers/util.py:31-45 sample_points`: every cloud re-sampled to `num` points -- a random subset
when it has at least `num` points, draws with replacement otherwise.  The index vectors come from torch's default (CPU)
generator through the same two calls as the reference (same seed -> same indices, pinned in
oracle/validate_against_reference.py); the gather runs on the device the points live on."""
from typing import List, Sequence

import torch


def _resample_indices(n_points: int, num: int) -> torch.Tensor:
    # randperm(n)[:num] for a large enough cloud, randint(n, (num,)) otherwise: the reference's RNG consumption
    return torch.randperm(n_points)[:num] if n_points >= num else torch.randint(n_points, size=(num,))


def sample_points(points_list: Sequence[torch.Tensor], num: int) -> List[torch.Tensor]:
    return [cloud.index_select(0, _resample_indices(cloud.shape[0], num).to(cloud.device)) for cloud in points_list]
