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
    """(N,) group id of every element -> (N,) its running index inside the group, in ascending element order
    (sst_ops_utils.py:5-12; the reference's atomic counter leaves the order to arrival)."""
    running = torch.full_like(group_inds, -1)
    sst_ops_cuda.ingroup_inds_wrapper(group_inds.contiguous(), running)
    return running


def group_inner_inds(points, inverse_inds, K):
    """points (N, C), inverse_inds (N,) pillar of every point -> (n_pillars, K, C): the first K points of every pillar,
    cyclically repeated when a pillar holds fewer (sst_ops_utils.py:15-27).  The pillar count is read from the device
    (one host sync, as in the reference); SPTBackboneMAE uses ops.group_points_centered on the CSR instead."""
    n_groups = int(inverse_inds.max()) + 1
    slots = torch.full((n_groups, K), -1, dtype=torch.long, device=points.device)
    sst_ops_cuda.group_inner_inds_wrapper(inverse_inds.contiguous(), slots)
    return points[slots]
