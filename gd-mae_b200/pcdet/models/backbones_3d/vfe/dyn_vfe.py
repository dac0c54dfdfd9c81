"""Mirror of pcdet/models/backbones_3d/vfe/dyn_vfe.py:10-124 (DynVFE) over the B200 kernels.

Same constructor signature, parameter names (``dvfe_mlps.0.{0,1,3,4}``) and batch_dict keys.
Supported configuration = the one every GD-MAE config uses: TYPE mean, WITH_DISTANCE False,
USE_ABSLOTE_XYZ / USE_CLUSTER_XYZ True, one MLP group, no aggregation MLP."""
from functools import partial

import torch
import torch.nn as nn

from ..... import fused as _fused
from ..... import ops as _ops
from ...model_utils.network_utils import make_fc_layers


class VFETemplate(nn.Module):
    def __init__(self, model_cfg, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg

    def get_output_feature_dim(self):
        raise NotImplementedError


class DynVFE(VFETemplate):
    def __init__(self, model_cfg, num_point_features, voxel_size, point_cloud_range, grid_size, **kwargs):
        super().__init__(model_cfg=model_cfg)
        self.sample_type = model_cfg.get('TYPE', 'mean')
        mlps = model_cfg.get('MLPS', None)
        if (self.sample_type != 'mean' or mlps is None or len(mlps) != 1 or model_cfg.get('AGGREGATION_MLPS', None)
                or model_cfg.WITH_DISTANCE or not model_cfg.USE_ABSLOTE_XYZ or not model_cfg.USE_CLUSTER_XYZ):
            raise NotImplementedError("gd-mae_b200 DynVFE implements the GD-MAE configuration (TYPE mean, one MLP group, "
                                      "absolute + cluster xyz, no distance)")
        input_channels = num_point_features + 6
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        self.dvfe_mlps = nn.ModuleList([make_fc_layers(mlps[0], input_channels, norm_fn=norm_fn)])
        self.num_point_features = mlps[0][-1]
        self.voxel_size = [float(v) for v in voxel_size]
        self.point_cloud_range = [float(v) for v in point_cloud_range]
        self.grid_size = [int(g) for g in grid_size]
        self.fused_mlp = True   # one autograd node for the MLP + scatter_max (fused.VfeMlpFunction); False = op by op

    def get_output_feature_dim(self):
        return self.num_point_features

    def index_pass(self, batch_dict):
        """Dynamic voxelisation only (one host sync for the counts): everything of forward() that does not touch the
        parameters.  GDMAE.prefetch_index runs it for the NEXT batch on a side stream while this step's backward is queued."""
        points = batch_dict['points']
        ps = _ops.dynamic_voxelize(points, self.point_cloud_range, self.voxel_size, self.grid_size, batch_dict['batch_size'])
        batch_dict['points'] = ps.points
        batch_dict['point_coords'] = ps.point_coords
        batch_dict['point_inverse_indices'] = ps.inverse
        batch_dict['voxel_coords'] = ps.voxel_coords
        batch_dict['pillar_set'] = ps  # extra key: CSR / batch offsets reused by SPTBackboneMAE
        # the point features (pillar mean, offsets to centre / cluster) are parameter-free too: built here, in pillar order
        n_feat = points.shape[1] - 1
        mean = _ops.segment_mean(ps.points, 1, n_feat, ps.seg_offsets, ps.seg_points, ps.n_pillars)
        batch_dict['vfe_point_features'] = _ops.vfe_point_features(ps, mean, self.point_cloud_range, self.voxel_size)
        return batch_dict

    def forward(self, batch_dict, **kwargs):
        if batch_dict.get('pillar_set', None) is None:
            self.index_pass(batch_dict)
        ps = batch_dict['pillar_set']
        n_feat = batch_dict['points'].shape[1] - 1
        if batch_dict.pop('defer_vfe_features', False):
            # GDMAE.forward asks for this when the next module is SPTBackboneMAE: the backbone first launches its own index kernels
            # (mask, visible sites, pyramid site sets - they need voxel_coords only), then calls this closure, then reads the
            # site counts back.  The feature pass (~0.8 ms of GPU work) thus covers the host work that follows the count
            # read, during which the GPU queue would otherwise run dry.
            batch_dict['pillar_features'] = batch_dict['voxel_features'] = None
            batch_dict['deferred_vfe'] = lambda: self._features(batch_dict, ps, n_feat)
            return batch_dict
        return self._features(batch_dict, ps, n_feat)

    def _features(self, batch_dict, ps, n_feat):
        x = batch_dict.pop('vfe_point_features', None)
        if x is None:
            mean = _ops.segment_mean(ps.points, 1, n_feat, ps.seg_offsets, ps.seg_points, ps.n_pillars)
            x = _ops.vfe_point_features(ps, mean, self.point_cloud_range, self.voxel_size)
        mlp = self.dvfe_mlps[0]
        if self.fused_mlp and self.training and _fused.vfe_mlp_supported(mlp, x) and ps.n_pillars > 0:
            x = _fused.vfe_mlp(mlp, x, ps)            # Linear-BN-ReLU x2 + scatter_max as one node (csrc/vfe_mlp.cu)
        else:
            with torch.autocast("cuda", enabled=False):  # absolute coordinates (|x| up to 75 m) stay fp32
                for k in range(0, len(mlp), 3):          # Linear (cuBLAS) -> fused BN1d(batch statistics)+ReLU, twice
                    x = _fused.batchnorm_relu(mlp[k + 1], mlp[k](x), self.training and mlp[k + 1].training)[0]
            x = _ops.SegmentMax.apply(x, ps.seg_offsets, ps.seg_points, ps.n_pillars)
        batch_dict['pillar_features'] = x
        batch_dict['voxel_features'] = x
        batch_dict.pop('deferred_vfe', None)
        return batch_dict
