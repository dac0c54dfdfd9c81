"""The torch_scatter surface used by the path (dyn_vfe.py:81,109), over the B200 kernels.
``index`` must be a grouping of rows (dim=0).  Arbitrary (unsorted) indices are supported by
building a CSR per call (torch argsort/bincount: API-surface plumbing, not the hot path); DynVFE
itself reuses the CSR that voxelisation already produced."""
import torch

from . import ops as _ops


def _csr(index, M):
    sorted_idx = torch.argsort(index, stable=True).int()
    counts = torch.bincount(index, minlength=M)
    off = torch.zeros(M + 1, dtype=torch.int32, device=index.device)
    off[1:] = counts.cumsum(0).int()
    return off, sorted_idx


def scatter(src, index, dim=0, reduce="mean", dim_size=None):
    assert dim == 0 and reduce == "mean" and src.dim() == 2
    M = int(index.max()) + 1 if dim_size is None else dim_size
    off, pts = _csr(index, M)
    return _ops.segment_mean(src.contiguous().float(), 0, src.shape[1], off, pts, M)


def scatter_max(src, index, dim=0, dim_size=None):
    assert dim == 0 and src.dim() == 2 and src.shape[1] % 4 == 0
    M = int(index.max()) + 1 if dim_size is None else dim_size
    off, pts = _csr(index, M)
    out = _ops.SegmentMax.apply(src, off, pts, M)
    return out, None
