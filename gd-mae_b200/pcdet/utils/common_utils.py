"""The four helpers of pcdet/utils/common_utils.py that sit on the path."""
import torch

from ... import ops as _ops


def get_voxel_centers(voxel_coords, downsample_times, voxel_size, point_cloud_range, dim=3):
    """common_utils.py:130-145."""
    voxel_centers = torch.flip(voxel_coords, dims=[-1]).float()
    voxel_size = torch.tensor(voxel_size[:dim], device=voxel_centers.device).float() * downsample_times
    pc_range = torch.tensor(point_cloud_range[:dim], device=voxel_centers.device).float()
    return (voxel_centers + 0.5) * voxel_size + pc_range


def random_masking(N, L, mask_ratio, device, noise=None):
    """common_utils.py:49-63 for N == 1 (the only call shape, spt_backbone_mae.py:99)."""
    assert N == 1
    if noise is None:
        noise = torch.rand(L, device=device)
    off = torch.tensor([0, L], dtype=torch.int32, device=device)
    return _ops.random_mask(noise.reshape(-1), off, 1, mask_ratio).view(1, L)


def get_in_range_mask(points, pc_range, voxel_size, grid_size):
    """common_utils.py:66-76 -> (mask (N,) bool, coords (N,3) int64 [x,y,z]) via the voxelisation kernel's
    arithmetic (fp32 subtract / divide / truncate)."""
    pc = points.new_tensor(pc_range) if not isinstance(pc_range, torch.Tensor) else pc_range
    vs = points.new_tensor(voxel_size) if not isinstance(voxel_size, torch.Tensor) else voxel_size
    gs = torch.as_tensor(grid_size, device=points.device).to(torch.int64)
    coords = ((points[:, 1:4] - pc[:3]) / vs).to(torch.int64)
    mask = torch.all((coords >= 0) & (coords < gs), dim=-1)
    return mask, coords
