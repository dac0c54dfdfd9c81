"""Mirror of pcdet/ops/sst_ops/sst_ops_utils.py:5-27 over the B200 kernels.

``sst_ops_cuda`` keeps the reference's pybind names and calling convention
(pcdet/ops/sst_ops/src/sst_ops_api.cpp:6-9): outputs are written in place into
caller-allocated tensors and the wrappers return 1.  Unlike the reference (exit(-1) on a CPU
tensor, sst_ops.cpp:7-19) a wrong device raises.  Order inside a group is deterministic
(ascending element index) where the reference's atomics leave it to arrival order."""
import torch

from .... import ops as _ops


class _SstOpsCuda:
    @staticmethod
    def ingroup_inds_wrapper(group_inds_tensor, out_inds_tensor):
        _ops.ingroup_inds(group_inds_tensor, out_inds_tensor)
        return 1

    @staticmethod
    def group_inner_inds_wrapper(inverse_inds_tensor, group_inds_tensor):
        M, K = group_inds_tensor.shape
        _ops.group_inner_inds(inverse_inds_tensor, M, K, out=group_inds_tensor)
        return 1


sst_ops_cuda = _SstOpsCuda()


def get_inner_win_inds(group_inds):
    """
    Args:
        group_inds: (N,)
    """
    out_inds = torch.zeros_like(group_inds) - 1
    sst_ops_cuda.ingroup_inds_wrapper(group_inds.contiguous(), out_inds)
    return out_inds


def group_inner_inds(points, inverse_inds, K):
    """
    Args:
        points: (N, C)
        inverse_inds: (N, )
    Return:
        group_points: (valid_voxel_num + 1, K, C)
    """
    valid_voxel_num = inverse_inds.max().item()
    group_inds = torch.full((valid_voxel_num + 1, K), -1, dtype=torch.long, device=points.device)
    sst_ops_cuda.group_inner_inds_wrapper(inverse_inds.contiguous(), group_inds)
    return points[group_inds]
