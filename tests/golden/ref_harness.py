"""Import the UNMODIFIED reference Python from /root/reference in this container.

Test infrastructure, used only by ``make_golden.py`` (never on the GPU box, never by the
product).  The reference's hot-path files import unchanged once the packages that are
absent from the image are replaced by stand-ins:

* ``torch_scatter``, ``spconv.pytorch``, ``pytorch3d.loss`` - third-party, not vendored in
  the reference; the stand-ins restate their published semantics with plain torch ops
  and are written INDEPENDENTLY of oracle/gdmae_oracle.py (dense-conv based spconv,
  cdist based chamfer) so that agreement between the two is a real cross-check.
* ``pcdet.ops.sst_ops.sst_ops_cuda`` - the reference op has no CPU path
  (sst_ops.cpp:7-19 exits on CPU tensors); the stand-in implements sst_ops_gpu.cu:14-39
  with a sequential loop, i.e. the arrival order "ascending element index".
* package ``__init__`` files are bypassed with namespace stubs because ``pcdet/__init__.py``
  needs a generated version.py and ``pcdet.models`` imports every CUDA extension.
"""
import importlib
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

import os

# the reference checkout in the build container; on the GPU box the byte-identical copies under baseline/_ref/
# (tools/install_reference.py) - bench.py's reference arm runs from there
_LOCAL = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "baseline", "_ref")
REF = os.environ.get("GDMAE_REF_ROOT") or ("/root/reference" if os.path.isdir("/root/reference") else _LOCAL)


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(d):
    if isinstance(d, dict):
        return AttrDict({k: to_attr(v) for k, v in d.items()})
    if isinstance(d, list):
        return [to_attr(v) for v in d]
    return d


# ------------------------------------------------------------------ torch_scatter stand-in
def _scatter(src, index, dim=0, reduce="sum", dim_size=None):
    assert dim == 0
    M = int(index.max()) + 1 if dim_size is None else dim_size
    out = torch.zeros((M,) + tuple(src.shape[1:]), dtype=src.dtype)
    out = out.index_add(0, index, src)
    if reduce in ("sum", "add"):
        return out
    assert reduce == "mean"
    cnt = torch.zeros(M, dtype=src.dtype).index_add(0, index, torch.ones(index.shape[0], dtype=src.dtype))
    cnt = cnt.clamp(min=1)
    return out / cnt.view((-1,) + (1,) * (src.dim() - 1))


def _scatter_max(src, index, dim=0, dim_size=None):
    assert dim == 0
    M = int(index.max()) + 1 if dim_size is None else dim_size
    if src.dim() == 1:
        out = torch.full((M,), torch.iinfo(src.dtype).min if not src.is_floating_point() else float("-inf"), dtype=src.dtype)
        out = out.scatter_reduce(0, index, src, reduce="amax", include_self=True)
        return out, None
    idx = index.unsqueeze(1).expand_as(src)
    out = torch.zeros((M, src.shape[1]), dtype=src.dtype).scatter_reduce(0, idx, src, reduce="amax", include_self=False)
    return out, None


# ------------------------------------------------------------------ spconv stand-in (dense-conv based)
class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size):
        self.features, self.indices = features, indices
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = batch_size

    def replace_feature(self, f):
        return SparseConvTensor(f, self.indices, self.spatial_shape, self.batch_size)

    def dense(self):
        H, W = self.spatial_shape
        idx = self.indices.long()
        out = self.features.new_zeros((self.batch_size, H, W, self.features.shape[1]))
        out[idx[:, 0], idx[:, 1], idx[:, 2]] = self.features
        return out.permute(0, 3, 1, 2).contiguous()


class SparseModule(nn.Module):
    pass


class SparseConvolution(SparseModule):
    def __init__(self, cin, cout, k, stride=1, padding=0, bias=False, indice_key=None, subm=False):
        super().__init__()
        assert not bias
        self.k, self.stride, self.padding, self.subm = k, stride, padding, subm
        self.weight = nn.Parameter(torch.empty(cout, k, k, cin))  # spconv 2.x KRSC
        nn.init.kaiming_uniform_(self.weight, a=np.sqrt(5))

    def forward(self, x):
        dense = x.dense()
        w = self.weight.permute(0, 3, 1, 2)
        if self.subm:
            y = F.conv2d(dense, w, padding=self.k // 2)
            idx = x.indices.long()
            feats = y.permute(0, 2, 3, 1)[idx[:, 0], idx[:, 1], idx[:, 2]]
            return SparseConvTensor(feats, x.indices, x.spatial_shape, x.batch_size)
        y = F.conv2d(dense, w, stride=self.stride, padding=self.padding)
        H, W = x.spatial_shape
        occ = torch.zeros((x.batch_size, 1, H, W))
        idx = x.indices.long()
        occ[idx[:, 0], 0, idx[:, 1], idx[:, 2]] = 1
        occ = F.max_pool2d(occ, self.k, stride=self.stride, padding=self.padding)[:, 0]
        oidx = torch.nonzero(occ > 0)
        feats = y.permute(0, 2, 3, 1)[oidx[:, 0], oidx[:, 1], oidx[:, 2]]
        return SparseConvTensor(feats, oidx.int(), list(y.shape[2:]), x.batch_size)


class SubMConv2d(SparseConvolution):
    def __init__(self, cin, cout, k, stride=1, padding=0, bias=False, indice_key=None):
        super().__init__(cin, cout, k, 1, 0, bias, indice_key, subm=True)


class SparseConv2d(SparseConvolution):
    def __init__(self, cin, cout, k, stride=1, padding=0, bias=False, indice_key=None):
        super().__init__(cin, cout, k, stride, padding, bias, indice_key, subm=False)


class SparseSequential(nn.Sequential):
    def forward(self, x):
        for m in self:
            if isinstance(m, SparseModule):
                x = m(x)
            else:
                x = x.replace_feature(m(x.features))
        return x


# ------------------------------------------------------------------ pytorch3d stand-in
def _chamfer_distance(x, y, weights=None):
    d = torch.cdist(x, y) ** 2  # (N,P1,P2)
    cx, cy = d.min(2)[0], d.min(1)[0]
    if weights is not None:
        if weights.sum() == 0:
            return x.sum() * 0.0, None
        cx, cy = cx * weights.view(-1, 1), cy * weights.view(-1, 1)
    cx, cy = cx.sum(1) / x.shape[1], cy.sum(1) / y.shape[1]
    div = weights.sum() if weights is not None else x.shape[0]
    return cx.sum() / div + cy.sum() / div, None


# ------------------------------------------------------------------ sst_ops_cuda stand-in
def _ingroup_inds_wrapper(group_inds, out_inds):
    g = group_inds.numpy()
    cnt = {}
    o = out_inds.numpy()
    for i in range(g.shape[0]):
        c = cnt.get(g[i], 0)
        o[i] = c
        cnt[g[i]] = c + 1
    return 1


def _group_inner_inds_wrapper(inverse_inds, group_inds):
    inv = inverse_inds.numpy()
    g = group_inds.numpy()
    M, K = g.shape
    cnt = np.zeros(M, dtype=np.int64)
    for i in range(inv.shape[0]):
        c = cnt[inv[i]]
        if c < K:
            g[inv[i], c] = i
        cnt[inv[i]] = c + 1
    for m in range(M):
        c = cnt[m]
        if c == 0:
            continue
        for i in range(c, K):
            g[m, i] = g[m, i % c]
    return 1


def _stub(name, path=None):
    m = types.ModuleType(name)
    if path is not None:
        m.__path__ = [path]
    sys.modules[name] = m
    return m


def install():
    """Seed sys.modules so that the reference's hot-path files import unchanged."""
    if "pcdet" in sys.modules and getattr(sys.modules["pcdet"], "_gdmae_harness", False):
        return
    for pkg in ["pcdet", "pcdet.models", "pcdet.models.backbones_3d", "pcdet.models.backbones_3d.vfe",
                "pcdet.models.model_utils", "pcdet.utils", "pcdet.ops", "pcdet.ops.sst_ops"]:
        _stub(pkg, REF + "/" + pkg.replace(".", "/"))
    sys.modules["pcdet"]._gdmae_harness = True
    ts = _stub("torch_scatter")
    ts.scatter, ts.scatter_max = _scatter, _scatter_max
    sp = _stub("spconv", "")
    spp = _stub("spconv.pytorch", "")
    spc = _stub("spconv.pytorch.conv")
    for m in (spp, sp):
        m.SparseConvTensor, m.SubMConv2d, m.SparseConv2d = SparseConvTensor, SubMConv2d, SparseConv2d
        m.SparseSequential, m.SparseModule, m.conv = SparseSequential, SparseModule, spc
    spc.SparseConvolution = SparseConvolution
    p3 = _stub("pytorch3d", "")
    p3l = _stub("pytorch3d.loss")
    p3l.chamfer_distance = _chamfer_distance
    p3.loss = p3l
    so = _stub("pcdet.ops.sst_ops.sst_ops_cuda")
    so.ingroup_inds_wrapper, so.group_inner_inds_wrapper = _ingroup_inds_wrapper, _group_inner_inds_wrapper
    _stub("SharedArray")


def ref_modules():
    install()
    names = dict(
        common_utils="pcdet.utils.common_utils", dyn_vfe="pcdet.models.backbones_3d.vfe.dyn_vfe",
        spt_backbone="pcdet.models.backbones_3d.spt_backbone", spt_backbone_mae="pcdet.models.backbones_3d.spt_backbone_mae",
        sst_utils="pcdet.models.model_utils.sst_utils", sst_basic_block="pcdet.models.model_utils.sst_basic_block",
        cosine_msa="pcdet.models.model_utils.cosine_msa", sst_ops_utils="pcdet.ops.sst_ops.sst_ops_utils")
    return AttrDict({k: importlib.import_module(v) for k, v in names.items()})


def ref_data_augmentor():
    """The reference's pcdet/datasets/augmentor/data_augmentor.py, unmodified.  ``database_sampler`` (gt_sampling, not on the
    SSL path, imports SharedArray and the CUDA iou ops) is replaced by an empty stand-in; package __init__s are bypassed."""
    install()
    for pkg in ["pcdet.datasets", "pcdet.datasets.augmentor"]:
        if pkg not in sys.modules:
            _stub(pkg, REF + "/" + pkg.replace(".", "/"))
    if "pcdet.datasets.augmentor.database_sampler" not in sys.modules:
        ds = _stub("pcdet.datasets.augmentor.database_sampler")
        ds.DataBaseSampler = object
        sys.modules["pcdet.datasets.augmentor"].database_sampler = ds
    return importlib.import_module("pcdet.datasets.augmentor.data_augmentor")


def load_model_cfg(rel_yaml):
    import yaml
    with open(REF + "/" + rel_yaml) as f:
        return to_attr(yaml.safe_load(f))


class RefMAE(nn.Module):
    """The reference's DynVFE + SPTBackboneMAE wired the way Detector3DTemplate.build_vfe /
    build_backbone_3d do (detector3d_template.py:70-100), with GDMAE's module names
    (``vfe``, ``backbone_3d``) so that state_dict keys match Appendix A."""

    def __init__(self, model_cfg, n_feat, voxel_size, pc_range, grid_size):
        super().__init__()
        R = ref_modules()
        self.vfe = R.dyn_vfe.DynVFE(model_cfg=model_cfg.VFE, num_point_features=n_feat, voxel_size=voxel_size,
                                    point_cloud_range=pc_range, grid_size=grid_size)
        self.backbone_3d = R.spt_backbone_mae.SPTBackboneMAE(
            model_cfg=model_cfg.BACKBONE_3D, input_channels=self.vfe.get_output_feature_dim(), grid_size=grid_size,
            voxel_size=voxel_size, point_cloud_range=pc_range)

    def forward(self, batch_dict):
        batch_dict = self.vfe(batch_dict)
        batch_dict = self.backbone_3d(batch_dict)
        loss, _ = self.backbone_3d.get_loss()
        return loss, batch_dict


# ------------------------------------------------------------------ finetune path
def install_finetune():
    """Additional stand-ins so that the reference's CenterPoint-path files import unchanged on this CPU-only container:
    * ``pcdet.ops.iou3d_nms.iou3d_nms_cuda``: the GPU entry points are replaced by the C oracle (oracle/iou3d_oracle.c, itself
      pinned bit-exact to the reference's own iou3d_cpu.cpp) working on CPU tensors;
    * ``pcdet.ops.roiaware_pool3d.roiaware_pool3d_cuda``: imported by box_utils, not used on this path - empty module."""
    install()
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from oracle import iou3d_oracle as IO
    for pkg in ["pcdet.models.backbones_2d", "pcdet.models.dense_heads", "pcdet.ops.iou3d_nms", "pcdet.ops.roiaware_pool3d"]:
        if pkg not in sys.modules:
            _stub(pkg, REF + "/" + pkg.replace(".", "/"))
    if "pcdet.ops.iou3d_nms.iou3d_nms_cuda" not in sys.modules:
        m = _stub("pcdet.ops.iou3d_nms.iou3d_nms_cuda")

        def boxes_overlap_bev_gpu(a, b, out):
            out.copy_(torch.from_numpy(IO.boxes_overlap_bev(a.detach().numpy(), b.detach().numpy())))
            return 1

        def boxes_iou_bev_gpu(a, b, out):
            out.copy_(torch.from_numpy(IO.boxes_iou_bev(a.detach().numpy(), b.detach().numpy())))
            return 1

        m.boxes_overlap_bev_gpu, m.boxes_iou_bev_gpu = boxes_overlap_bev_gpu, boxes_iou_bev_gpu
        sys.modules["pcdet.ops.iou3d_nms"].iou3d_nms_cuda = m
        r = _stub("pcdet.ops.roiaware_pool3d.roiaware_pool3d_cuda")
        sys.modules["pcdet.ops.roiaware_pool3d"].roiaware_pool3d_cuda = r


class _CpuAsCuda:
    """``tensor.cuda()`` / ``torch.cuda.FloatTensor`` are hard-coded in center_head.py:66 and iou3d_nms_utils.py:41,66: inside
    this context they return CPU tensors (the container has no GPU)."""

    def __enter__(self):
        self._cuda = torch.Tensor.cuda
        torch.Tensor.cuda = lambda t, *a, **k: t
        self._ft = getattr(torch.cuda, "FloatTensor", None)
        torch.cuda.FloatTensor = lambda size: torch.empty(size, dtype=torch.float32)
        return self

    def __exit__(self, *exc):
        torch.Tensor.cuda = self._cuda
        if self._ft is not None:
            torch.cuda.FloatTensor = self._ft
        return False


class RefCenterPoint(nn.Module):
    """The reference's DynVFE + SPTBackbone + SSTBEVBackbone + CenterHead wired the way Detector3DTemplate.build_networks
    does for tools/cfgs/waymo_models/gd_mae_iou.yaml (detector3d_template.py:70-157), with CenterPoint's module names."""

    def __init__(self, model_cfg, n_feat, voxel_size, pc_range, grid_size, class_names):
        super().__init__()
        install_finetune()
        R = ref_modules()
        bev = importlib.import_module("pcdet.models.backbones_2d.sst_bev_backbone")
        ch = importlib.import_module("pcdet.models.dense_heads.center_head")
        self.vfe = R.dyn_vfe.DynVFE(model_cfg=model_cfg.VFE, num_point_features=n_feat, voxel_size=voxel_size,
                                    point_cloud_range=pc_range, grid_size=grid_size)
        self.backbone_3d = R.spt_backbone.SPTBackbone(model_cfg=model_cfg.BACKBONE_3D, input_channels=self.vfe.get_output_feature_dim(),
                                                      grid_size=grid_size, voxel_size=voxel_size, point_cloud_range=pc_range)
        self.backbone_2d = bev.SSTBEVBackbone(model_cfg=model_cfg.BACKBONE_2D, input_channels=None)
        with _CpuAsCuda():
            self.dense_head = ch.CenterHead(model_cfg=model_cfg.DENSE_HEAD, input_channels=self.backbone_2d.num_bev_features,
                                            num_class=len(class_names), class_names=class_names, grid_size=grid_size,
                                            point_cloud_range=pc_range, predict_boxes_when_training=False, voxel_size=voxel_size)

    def forward(self, batch_dict):
        with _CpuAsCuda():
            for m in (self.vfe, self.backbone_3d, self.backbone_2d, self.dense_head):
                batch_dict = m(batch_dict)
            loss, tb_dict = self.dense_head.get_loss()
        return loss, tb_dict, batch_dict
