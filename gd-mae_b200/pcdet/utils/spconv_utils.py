"""Mirror of pcdet/utils/spconv_utils.py:1-56 and of the spconv.pytorch surface the path uses
(SparseConvTensor, SubMConv2d, SparseConv2d, SparseSequential), built on the B200 structure
kernels (rank grids + 3x3 neighbour maps) and gather -> GEMM.

Weight layout is spconv 2.x "KRSC": (C_out, kH, kW, C_in), which makes the
sparse conv ONE GEMM of the gathered (N, 9*C_in) rows with weight.view(C_out, 9*C_in)^T.
Output rows of the strided conv are ordered lexicographically by (b, y, x)."""
import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops as _ops


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, _struct=None):
        self.features = features
        self.indices = indices.int().contiguous() if indices.dtype != torch.int32 else indices.contiguous()
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self._struct = {} if _struct is None else _struct  # structure cache shared by replace_feature()

    def replace_feature(self, new_features):
        return SparseConvTensor(new_features, self.indices, self.spatial_shape, self.batch_size, self._struct)

    # ---- structure (index) side, cached
    def rank_grid(self):
        if "rank_grid" not in self._struct:
            H, W = self.spatial_shape
            self._struct["rank_grid"] = _ops.build_rank_grid(self.indices, self.batch_size, H, W)
        return self._struct["rank_grid"]

    def subm_map(self):
        if "subm_map" not in self._struct:
            H, W = self.spatial_shape
            self._struct["subm_map"] = _ops.subm_neighbor_map(self.indices, self.rank_grid(), self.batch_size, H, W)
        return self._struct["subm_map"]

    def window_tables(self):
        if "win" not in self._struct:
            H, W = self.spatial_shape
            self._struct["win"] = [_ops.window_table(self.indices, self.batch_size, H, W, s) for s in (0, 1)]
        return self._struct["win"]

    def down(self):
        """Output site set + maps of SparseConv2d(3, stride 2, pad 1); planned ahead by
        plan_pyramid() or computed here with one host sync for the site count."""
        if "down" not in self._struct:
            plan_pyramid(self, 1)
        return self._struct["down"]

    def dense(self, channels_first=True):
        """(B, C, H, W), zeros at empty cells.  API surface only: the MAE decoder never
        materialises this (see ops.DenseFill)."""
        H, W = self.spatial_shape
        idx = self.indices.long()
        out = self.features.new_zeros((self.batch_size, H, W, self.features.shape[1]))
        out[idx[:, 0], idx[:, 1], idx[:, 2]] = self.features
        return out.permute(0, 3, 1, 2).contiguous() if channels_first else out


_PINNED = {}


def plan_pyramid_launch(sp, n_levels):
    """First half of plan_pyramid: launches the site-set kernels of ``n_levels`` successive stride-2 sparse convs and an
    asynchronous copy of their counts into pinned memory; returns a handle for plan_pyramid_finish.  GPU work enqueued
    between the two calls runs while the host waits for the counts and builds the structures that depend on them."""
    B = sp.batch_size
    H, W = sp.spatial_shape
    idx, n_rows = sp.indices, sp.indices.shape[0]
    pending = []
    for _ in range(n_levels):
        out_idx, grid, count, Ho, Wo = _ops.down_sites(idx, n_rows, B, H, W)
        pending.append((out_idx, grid, count, H, W, Ho, Wo))
        idx, n_rows, H, W = out_idx, out_idx.shape[0], Ho, Wo
    if not pending:
        return sp, pending, None, None
    host = _PINNED.get(n_levels)
    if host is None:
        host = _PINNED[n_levels] = torch.empty((n_levels,), dtype=torch.int32).pin_memory()
    host.copy_(torch.cat([p[2] for p in pending]), non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    return sp, pending, host, ev


def plan_pyramid_finish(handle):
    sp, pending, host, ev = handle
    if not pending:
        return
    ev.synchronize()                       # the one sync: waits for the count copy only, not for work enqueued after it
    counts = host.tolist()
    B = sp.batch_size
    cur = sp
    for (out_idx, grid, _, H, W, Ho, Wo), n in zip(pending, counts):
        out_idx = out_idx[:n]
        nbr_down, nbr_up = _ops.down_neighbor_maps(cur.indices, cur.rank_grid(), H, W, out_idx, grid)
        nxt_struct = {"rank_grid": grid}
        cur._struct["down"] = SimpleNamespace(indices=out_idx, spatial_shape=[Ho, Wo], nbr_down=nbr_down, nbr_up=nbr_up,
                                              struct=nxt_struct)
        cur = SparseConvTensor(None, out_idx, [Ho, Wo], B, nxt_struct)


def plan_pyramid(sp, n_levels):
    """Site sets of ``n_levels`` successive stride-2 sparse convs starting at ``sp`` with ONE host
    sync for all their counts; attaches the result to each level's structure cache."""
    plan_pyramid_finish(plan_pyramid_launch(sp, n_levels))


def prebuild_structures(sp, n_levels, tensor_core_units=False):
    """Builds every index structure the ``n_levels`` + 1 pyramid levels below ``sp`` will ask for - submanifold neighbour
    maps, both window tables (and the work units of the tensor-core SRA kernels) - in one go.  They depend on the site sets
    only, so a caller that has GPU work in flight (SPTBackboneMAE right after it enqueued the VFE feature pass) gets all the
    host-side table building done while the GPU is busy, instead of in front of each block with the queue empty."""
    cur = sp
    for lvl in range(n_levels + 1):
        cur.subm_map()
        for t in cur.window_tables():
            if tensor_core_units:
                t.bin_units()
        if lvl < n_levels:
            d = cur.down()
            cur = SparseConvTensor(None, d.indices, d.spatial_shape, cur.batch_size, d.struct)


class SparseModule(nn.Module):
    pass


class SparseConvolution(SparseModule):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=False, indice_key=None, subm=False):
        super().__init__()
        assert kernel_size == 3 and not bias, "the path uses 3x3 convs without bias (spconv_utils.py:37-56)"
        assert subm or (stride == 2 and padding == 1)
        self.in_channels, self.out_channels, self.subm, self.indice_key = in_channels, out_channels, subm, indice_key
        self.weight = nn.Parameter(torch.empty(out_channels, 3, 3, in_channels))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))

    def forward(self, x):
        from ... import fused as _fused
        if self.subm:
            nbr = x.subm_map()
            return x.replace_feature(_fused.SparseConvFunction.apply(x.features, self.weight, nbr, nbr, True))
        d = x.down()
        y = _fused.SparseConvFunction.apply(x.features, self.weight, d.nbr_down, d.nbr_up, False)
        return SparseConvTensor(y, d.indices, d.spatial_shape, x.batch_size, d.struct)

    def forward_bn_relu(self, x, bn):
        """this convolution followed by BatchNorm1d (training statistics) + ReLU as one autograd node (bf16 configuration)"""
        from ... import fused as _fused
        if self.subm:
            nbr = x.subm_map()
            return x.replace_feature(_fused.sparse_conv_bn_relu(self, bn, x.features, nbr, nbr, True))
        d = x.down()
        y = _fused.sparse_conv_bn_relu(self, bn, x.features, d.nbr_down, d.nbr_up, False)
        return SparseConvTensor(y, d.indices, d.spatial_shape, x.batch_size, d.struct)


class SubMConv2d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=False, indice_key=None):
        super().__init__(in_channels, out_channels, kernel_size, 1, 0, bias, indice_key, subm=True)


class SparseConv2d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=False, indice_key=None):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, bias, indice_key, subm=False)


class SparseSequential(nn.Sequential):
    def forward(self, x):
        from ... import fused as _fused
        mods = list(self)
        i = 0
        while i < len(mods):
            m = mods[i]
            if (isinstance(m, SparseConvolution) and i + 2 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm1d)
                    and isinstance(mods[i + 2], nn.ReLU)
                    and _fused.sparse_conv_bn_relu_ok(m, mods[i + 1], x.features, self.training and mods[i + 1].training)):
                # conv + BatchNorm1d + ReLU (post_act_block) as one node: bf16 gradient hand-over inside (fused.py)
                x = m.forward_bn_relu(x, mods[i + 1])
                i += 2
            elif isinstance(m, SparseModule):
                x = m(x)
            elif (isinstance(m, nn.BatchNorm1d) and i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU) and x.features.is_cuda
                  and m.track_running_stats):
                # BatchNorm1d(train: batch statistics) + ReLU as one fused kernel pair (csrc/batchnorm.cu)
                x = x.replace_feature(_fused.batchnorm_relu(m, x.features, self.training and m.training)[0])
                i += 1
            else:
                x = x.replace_feature(m(x.features))
            i += 1
        return x


spconv = SimpleNamespace(SparseConvTensor=SparseConvTensor, SubMConv2d=SubMConv2d, SparseConv2d=SparseConv2d,
                         SparseSequential=SparseSequential, SparseModule=SparseModule,
                         conv=SimpleNamespace(SparseConvolution=SparseConvolution))


def find_all_spconv_keys(model, prefix=""):
    """state_dict keys of every sparse-convolution weight below ``model`` (pcdet/utils/spconv_utils.py:11-26): the keys
    whose layout differs between spconv versions and may need adapting when a checkpoint is loaded."""
    found = set()
    for name, child in model.named_children():
        new_prefix = f"{prefix}.{name}" if prefix != "" else name
        if isinstance(child, SparseConvolution):
            found.add(f"{new_prefix}.weight")
        found.update(find_all_spconv_keys(child, prefix=new_prefix))
    return found


def replace_feature(out, new_features):
    return out.replace_feature(new_features)


def post_act_block(in_channels, out_channels, kernel_size, indice_key=None, stride=1, padding=0,
                   conv_type='subm', norm_fn=None, dim=3):
    """spconv_utils.py:37-56 for dim=2, conv_type in {'subm', 'spconv'}."""
    assert dim == 2
    if conv_type == 'subm':
        conv = SubMConv2d(in_channels, out_channels, kernel_size, bias=False, indice_key=indice_key)
    elif conv_type == 'spconv':
        conv = SparseConv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=False,
                            indice_key=indice_key)
    else:
        raise NotImplementedError
    return SparseSequential(conv, norm_fn(out_channels), nn.ReLU())
