"""Mirror of the reference's `helpers/util.py:31-45 sample_points`: every cloud re-sampled to `num` points -- a random subset
when it has at least `num` points, draws with replacement otherwise.  The index vectors come from torch's default (CPU)
generator exactly as in the reference (same seed -> same indices); the gather runs where the points live."""
import torch


def sample_points(points_list, num):
    resampled_point_clouds = []
    for point_cloud in points_list:
        n_points = point_cloud.size(0)
        if n_points >= num:
            random_indices = torch.randperm(n_points)[:num]
        else:
            random_indices = torch.randint(n_points, size=(num,))
        resampled_point_clouds.append(point_cloud[random_indices.to(point_cloud.device)])
    return resampled_point_clouds
