"""Mirror of pcdet/models/backbones_3d/spt_backbone_mae.py:11-153 (SPTBackboneMAE) over the B200
kernels: random masking, 3 SST blocks on the visible pillars, the generative BEV decoder, gather
at all pillars, 16-point prediction against 64 grouped ground-truth points, chamfer loss.

Same constructor signature, state_dict names (sst_blocks.*, decoder_deblocks.{i}.{0,1},
decoder_conv_out.{0,1}, decoder_pred) and batch_dict keys.  Parity-harness inputs: if
``batch_dict['voxel_mae_mask']`` (M,) or ``batch_dict['voxel_mae_noise']`` (M,) is present it is
used instead of drawing ``torch.rand`` (the RNG stream differs between CPU and GPU)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .... import fused as _fused
from .... import ops as _ops
from ...utils.spconv_utils import plan_pyramid_launch, plan_pyramid_finish, prebuild_structures, spconv, plan_pyramid
from .spt_backbone import SSTBlockV1


class SPTBackboneMAE(nn.Module):
    def __init__(self, model_cfg, input_channels, grid_size, voxel_size, point_cloud_range, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.grid_size = [int(g) for g in grid_size]
        self.voxel_size = [float(v) for v in voxel_size]
        self.point_cloud_range = [float(v) for v in point_cloud_range]
        self.sparse_shape = [self.grid_size[1], self.grid_size[0]]
        in_channels = input_channels

        self.mask_cfg = self.model_cfg.get('MASK_CONFIG', None)
        self.mask_ratio = self.mask_cfg.RATIO if self.mask_cfg is not None else 0.0

        self.sst_blocks = nn.ModuleList()
        for sst_block_cfg in model_cfg.SST_BLOCK_LIST:
            self.sst_blocks.append(SSTBlockV1(sst_block_cfg, in_channels, sst_block_cfg.NAME))
            in_channels = sst_block_cfg.ENCODER.D_MODEL

        in_channels = 0
        self.decoder_deblocks = nn.ModuleList()
        self.fuse_strides = []
        for src in model_cfg.FEATURES_SOURCE:
            conv_cfg = model_cfg.FUSE_LAYER[src]
            self.decoder_deblocks.append(nn.Sequential(
                nn.ConvTranspose2d(conv_cfg.NUM_FILTER, conv_cfg.NUM_UPSAMPLE_FILTER, conv_cfg.UPSAMPLE_STRIDE,
                                   stride=conv_cfg.UPSAMPLE_STRIDE, bias=False),
                nn.BatchNorm2d(conv_cfg.NUM_UPSAMPLE_FILTER, eps=1e-3, momentum=0.01),
                nn.ReLU(inplace=True)))
            in_channels += conv_cfg.NUM_UPSAMPLE_FILTER
            self.fuse_strides.append(int(conv_cfg.UPSAMPLE_STRIDE))
        n_src = len(self.decoder_deblocks)
        if n_src != 3 or len({m[0].out_channels for m in self.decoder_deblocks}) != 1:
            raise NotImplementedError("the B200 decoder fill handles the three equal-width pyramid sources of the GD-MAE configs")
        self.decoder_conv_out = nn.Sequential(
            nn.Conv2d(in_channels, in_channels // n_src, 3, padding=1, bias=False),
            nn.BatchNorm2d(in_channels // n_src, eps=1e-3, momentum=0.01),
            nn.ReLU(inplace=True))
        in_channels = in_channels // n_src
        self.decoder_pred = nn.Linear(in_channels, self.mask_cfg.NUM_PRD_POINTS * 3, bias=True)
        self.forward_ret_dict = {}
        self.num_point_features = in_channels
        # dtype of the dense 384-channel BEV map and of the cuDNN decoder conv (fp32 = parity mode,
        # torch.bfloat16 = the bf16 configuration of BASELINE.json; BN statistics stay fp32 either way)
        self.decoder_dtype = torch.float32
        # True: batch_dict['spatial_features'] holds the dense BN+ReLU map like the reference (finetune heads read it);
        # False: the MAE step evaluates BN+ReLU only at the pillar cells its head gathers (ops.DecoderTail)
        self.dense_spatial_features = True

    # ------------------------------------------------------------------ loss (spt_backbone_mae.py:83-89)
    def get_loss(self, tb_dict=None):
        tb_dict = {} if tb_dict is None else tb_dict
        r = self.forward_ret_dict
        loss, _ = _ops.chamfer_distance(r['pred_points'], r['gt_points'], weights=r['mask'])
        return loss, tb_dict

    # ------------------------------------------------------------------ target_assigner (:57-81)
    def target_assigner(self, batch_dict):
        voxel_features = batch_dict['voxel_features']
        ps = batch_dict.get('pillar_set', None)
        K = self.mask_cfg.NUM_GT_POINTS
        if ps is None:
            raise _ops.L.GdmaeError("SPTBackboneMAE needs batch_dict['pillar_set'] written by gd-mae_b200's DynVFE")
        norm_gt_points = batch_dict.get('mae_gt_points', None)      # prefetched by index_pass, else built here
        if norm_gt_points is None:
            norm_gt_points = _ops.group_points_centered(ps, self.point_cloud_range, self.voxel_size, K)
        if voxel_features.is_cuda and self.training:
            pred_points = _ops.TallLinear.apply(voxel_features.contiguous(), self.decoder_pred.weight, self.decoder_pred.bias)
        else:
            pred_points = self.decoder_pred(voxel_features)
        pred_points = pred_points.view(voxel_features.shape[0], -1, 3)
        return {'pred_points': pred_points, 'gt_points': norm_gt_points, 'mask': batch_dict['voxel_mae_mask']}

    # ------------------------------------------------------------------ decoder (:123-132)
    def _deblock_rows(self, i, sp, n_cells_total):
        """ConvTranspose2d(k=s) + BatchNorm2d(train) + ReLU of one pyramid level, evaluated on the
        sparse rows: every active site yields its k x k block, every other cell is exactly 0 before
        BN, so the batch statistics follow from the sparse rows and the zero count."""
        deconv, bn = self.decoder_deblocks[i][0], self.decoder_deblocks[i][1]
        k = self.fuse_strides[i]
        c_out = deconv.out_channels
        if (_fused.GEMM_DTYPE == torch.bfloat16 and self.training and deconv.bias is None and deconv.in_channels % 64 == 0
                and (k * k * c_out) % 64 == 0):
            # bf16 configuration: GEMM on the own tcgen05 kernel, BatchNorm backward hands bf16 gradients to the backward GEMMs
            return _fused.deblock_rows(deconv, bn, k, sp.features, n_cells_total)
        w = deconv.weight.permute(0, 2, 3, 1).reshape(deconv.in_channels, k * k * c_out)  # (C_in, [a, b, c_out])
        u = (sp.features @ w).view(-1, c_out)                                                 # (N*k*k, c_out)
        # fused BN(batch statistics over all B*Y*X cells, zeros included)+ReLU on the sparse rows; bg = value of an empty cell
        return _fused.batchnorm_relu(bn, u, self.training, relu=True, count=n_cells_total)

    def _mask_and_sites(self, batch_dict):
        """random masking per frame (spt_backbone_mae.py:96-100), the visible site set and the launch of the pyramid plan"""
        all_voxel_coords = batch_dict['voxel_coords']
        batch_size = batch_dict['batch_size']
        ps = batch_dict.get('pillar_set', None)
        if ps is None:
            raise _ops.L.GdmaeError("SPTBackboneMAE needs batch_dict['pillar_set'] written by gd-mae_b200's DynVFE")
        if self.grid_size[2] != 1:
            raise AssertionError("pillar grids only: z extent must be 1 (spt_backbone_mae.py:94)")
        Y, X = self.sparse_shape
        M = all_voxel_coords.shape[0]
        if 'voxel_mae_mask' in batch_dict and batch_dict['voxel_mae_mask'] is not None:
            voxel_mae_mask = batch_dict['voxel_mae_mask'].float().contiguous()
            n_visible = int((voxel_mae_mask == 0).sum().item())
        else:
            noise = batch_dict.get('voxel_mae_noise', None)
            if noise is None:
                noise = torch.rand(M, device=all_voxel_coords.device)
            voxel_mae_mask = _ops.random_mask(noise, ps.batch_offsets_dev, batch_size, self.mask_ratio)
            n_visible = sum(int((ps.batch_offsets[b + 1] - ps.batch_offsets[b]) * (1 - self.mask_ratio))
                            for b in range(batch_size))
        batch_dict['voxel_mae_mask'] = voxel_mae_mask
        vis_idx, indices, rank_grid, _ = _ops.visible_sites(all_voxel_coords, voxel_mae_mask, n_visible, batch_size, Y, X)
        sp = spconv.SparseConvTensor(None, indices, self.sparse_shape, batch_size, {"rank_grid": rank_grid})
        n_down = sum(1 for b in self.sst_blocks if b.conv_down is not None)
        plan = plan_pyramid_launch(sp, n_down)  # all site sets of the pyramid, counts copied asynchronously
        return vis_idx, sp, n_down, plan

    def index_pass(self, batch_dict):
        """Everything of forward() that depends on voxel_coords only - mask, visible sites, pyramid site sets, neighbour
        maps, window tables, SRA work units - stored under batch_dict['mae_index'].  GDMAE.prefetch_index runs it for the
        NEXT batch on a side stream; forward() then has no host sync and no table building left."""
        vis_idx, sp, n_down, plan = self._mask_and_sites(batch_dict)
        plan_pyramid_finish(plan)
        prebuild_structures(sp, n_down, tensor_core_units=bool(_ops.SRA_TENSOR_CORES))
        batch_dict['mae_index'] = (vis_idx, sp)
        # the reconstruction targets (first K points of every pillar, centred) depend on the points only
        batch_dict['mae_gt_points'] = _ops.group_points_centered(batch_dict['pillar_set'], self.point_cloud_range, self.voxel_size,
                                                                 self.mask_cfg.NUM_GT_POINTS)
        return batch_dict

    def forward(self, batch_dict):
        batch_size = batch_dict['batch_size']
        Y, X = self.sparse_shape
        deferred = batch_dict.get('deferred_vfe', None)
        if batch_dict.get('mae_index', None) is not None:
            vis_idx, input_sp_tensor = batch_dict['mae_index']     # prefetched one step ahead
            if deferred is not None:
                deferred()
        else:
            vis_idx, input_sp_tensor, n_down, plan = self._mask_and_sites(batch_dict)
            # the index kernels above need voxel_coords only: a DynVFE that deferred its feature pass runs it now, so that
            # its kernels cover the count read and the host-side structure building that follows
            if deferred is not None:
                deferred()
            plan_pyramid_finish(plan)                            # the one host sync of the backbone
            if self.training:
                # window tables, neighbour maps and SRA work units of all three scales now, while the VFE kernels run
                prebuild_structures(input_sp_tensor, n_down, tensor_core_units=bool(_ops.SRA_TENSOR_CORES))
        all_voxel_coords = batch_dict['voxel_coords']
        ps = batch_dict['pillar_set']
        M = all_voxel_coords.shape[0]
        voxel_mae_mask = batch_dict['voxel_mae_mask']
        all_voxel_features = batch_dict['voxel_features']
        input_sp_tensor = input_sp_tensor.replace_feature(all_voxel_features.index_select(0, vis_idx))

        x = input_sp_tensor
        x_hidden = []
        for sst_block in self.sst_blocks:
            x = sst_block(x)
            x_hidden.append(x)

        batch_dict.update({'encoded_spconv_tensor': x_hidden[-1],
                           'encoded_spconv_tensor_stride': self.sparse_shape[0] // x_hidden[-1].spatial_shape[0]})
        multi_scale_3d_features, multi_scale_3d_strides = {}, {}
        for i in range(len(x_hidden)):
            multi_scale_3d_features[f'x_conv{i + 1}'] = x_hidden[i]
            multi_scale_3d_strides[f'x_conv{i + 1}'] = self.sparse_shape[0] // x_hidden[i].spatial_shape[0]

        srcs = [multi_scale_3d_features[s] for s in self.model_cfg.FEATURES_SOURCE]
        n_cells_total = batch_size * Y * X
        rows, bgs = zip(*[self._deblock_rows(i, sp, n_cells_total) for i, sp in enumerate(srcs)])
        dt = self.decoder_dtype
        fused = _ops.DenseFill.apply(rows[0], rows[1], rows[2], bgs[0], bgs[1], bgs[2], [sp.rank_grid() for sp in srcs],
                                     [sp.indices for sp in srcs], self.fuse_strides, batch_size, Y, X, dt)  # (B,Y,X,384) NHWC
        conv, bn = self.decoder_conv_out[0], self.decoder_conv_out[1]
        own_wgrad = (dt == torch.bfloat16 and self.training and conv.in_channels == 384 and conv.out_channels == 128
                     and conv.kernel_size == (3, 3) and conv.bias is None)
        if own_wgrad:
            # cuDNN forward / input gradient, own tcgen05 weight gradient (ops.DecoderConv3x3)
            y = _ops.DecoderConv3x3.apply(fused, conv.weight, _fused._gw(conv.weight)).permute(0, 3, 1, 2)
        else:
            w = _fused.cast_param(conv.weight, dt)
            y = F.conv2d(fused.permute(0, 3, 1, 2), w.contiguous(memory_format=torch.channels_last), padding=1)  # cuDNN, NHWC
        fuse_tail = self.training and not self.dense_spatial_features and ps.cell2pillar.numel() == batch_size * Y * X
        if fuse_tail:
            # BN statistics over the whole map, BN + ReLU values only at the pillar cells the head gathers (ops.DecoderTail)
            all_pyramid_voxel_features = _ops.DecoderTail.apply(y.permute(0, 2, 3, 1), bn.weight, bn.bias, bn.running_mean,
                                                                bn.running_var, bn.momentum, bn.eps, True, all_voxel_coords,
                                                                ps.cell2pillar)
            spatial_features = None          # not materialised: nothing on the MAE path reads it
        else:
            y = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, self.training, bn.momentum, bn.eps)
            spatial_features = F.relu(y)                                                                     # (B, C, Y, X)
        if self.training:
            bn.num_batches_tracked += 1
        spatial_features_stride = multi_scale_3d_strides[self.model_cfg.FEATURES_SOURCE[0]] // self.fuse_strides[0]

        batch_dict['multi_scale_3d_features'] = multi_scale_3d_features
        batch_dict['multi_scale_3d_strides'] = multi_scale_3d_strides
        batch_dict['spatial_features'] = spatial_features
        batch_dict['spatial_features_stride'] = spatial_features_stride
        all_voxel_shuffle_inds = torch.arange(M, device=all_voxel_coords.device, dtype=torch.long)
        if not fuse_tail:
            assert spatial_features.shape[0] == batch_size and spatial_features.shape[2] == Y and spatial_features.shape[3] == X
            all_pyramid_voxel_features = _ops.GatherNHWC.apply(spatial_features.permute(0, 2, 3, 1), all_voxel_coords)
        batch_dict.update({'voxel_features': all_pyramid_voxel_features, 'voxel_coords': all_voxel_coords,
                           'voxel_shuffle_inds': all_voxel_shuffle_inds})
        self.forward_ret_dict = self.target_assigner(batch_dict)
        return batch_dict
