"""Mirror of pcdet/models/backbones_3d/spt_backbone.py:11-264 (SSTInputLayer, SSTBlockV1) over the
B200 kernels.  Same class names, constructor signatures and parameter names
(conv_down.{0,1}, encoder_blocks.{e}.encoder_list.{l}.*, conv_out.{0,1})."""
from functools import partial

import numpy as np
import torch
import torch.nn as nn

from ..model_utils.sst_basic_block import BasicShiftBlockV2
from ...utils.spconv_utils import replace_feature, post_act_block


def pos_embed_table(feat_dim, pos_temperature, window_shape=(8, 8, 1), normalize_pos=False):
    """SSTInputLayer.get_pos_embed (spt_backbone.py:137-172) evaluated on the 64 in-window cells:
    row yy*8+xx holds [sin/cos(x terms) | sin/cos(y terms)].  Computed once on the CPU in fp32."""
    win_x, win_y = window_shape[:2]
    yy, xx = torch.meshgrid(torch.arange(win_y), torch.arange(win_x), indexing="ij")
    y, x = yy.flatten() - win_y / 2, xx.flatten() - win_x / 2
    if normalize_pos:
        x = x / win_x * 2 * 3.1415
        y = y / win_y * 2 * 3.1415
    pos_length = feat_dim // 2
    inv_freq = torch.arange(pos_length, dtype=torch.float32)
    inv_freq = pos_temperature ** (2 * (torch.div(inv_freq, 2, rounding_mode='floor')) / pos_length)
    embed_x = x[:, None] / inv_freq[None, :]
    embed_y = y[:, None] / inv_freq[None, :]
    embed_x = torch.stack([embed_x[:, ::2].sin(), embed_x[:, 1::2].cos()], dim=-1).flatten(1)
    embed_y = torch.stack([embed_y[:, ::2].sin(), embed_y[:, 1::2].cos()], dim=-1).flatten(1)
    return torch.cat([embed_x, embed_y], dim=-1).float()


class SSTInputLayer(nn.Module):
    """Window partition of the active sites for both shifts (spt_backbone.py:11-194).  With 8x8x1
    windows and the GD-MAE DROP_INFO (level 2 keeps 64 tokens) no voxel is ever dropped, so the
    layer reduces to building the two window tables; SHUFFLE_VOXELS must be False."""

    def __init__(self, model_cfg, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.window_shape = list(model_cfg.WINDOW_SHAPE)
        self.shuffle_voxels = model_cfg.SHUFFLE_VOXELS
        drop_info = model_cfg.DROP_INFO['train' if self.training else 'test']
        self.drop_info = {int(k): v for k, v in drop_info.items()}
        self.pos_temperature = model_cfg.POS_TEMPERATURE
        self.normalize_pos = model_cfg.NORMALIZE_POS
        assert self.window_shape[2] == 1
        expect = {0: (16, 0, 16), 1: (32, 16, 32), 2: (64, 32)}
        ok = self.window_shape == [8, 8, 1] and not self.shuffle_voxels and all(
            k in self.drop_info and self.drop_info[k]['max_tokens'] == v[0]
            and list(self.drop_info[k]['drop_range'])[:len(v) - 1] == list(v[1:]) for k, v in expect.items())
        if not ok or len(self.drop_info) != 3:
            raise NotImplementedError("gd-mae_b200 implements the 8x8x1 window / 16-32-64 drop levels of the GD-MAE configs")
        self._pos_tables = {}

    def pos_table(self, feat_dim, device):
        key = (feat_dim, str(device))
        if key not in self._pos_tables:
            self._pos_tables[key] = pos_embed_table(feat_dim, self.pos_temperature, self.window_shape,
                                                    self.normalize_pos).to(device)
        return self._pos_tables[key]

    def forward(self, input_dict):
        sp = input_dict['sp_tensor']
        tables = sp.window_tables()
        pos = self.pos_table(input_dict['voxel_features'].shape[1], input_dict['voxel_features'].device)
        voxel_info = dict(input_dict)
        for i in range(2):
            voxel_info[f'flat2win_inds_shift{i}'] = tables[i]
            voxel_info[f'pos_dict_shift{i}'] = pos
            voxel_info[f'key_mask_shift{i}'] = None
        return voxel_info


class SSTBlockV1(nn.Module):
    def __init__(self, model_cfg, input_channels, indice_key, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        encoder_cfg = model_cfg.ENCODER
        d_model = encoder_cfg.D_MODEL
        stride = encoder_cfg.STRIDE
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        if stride > 1:
            self.conv_down = post_act_block(input_channels, d_model, 3, norm_fn=norm_fn, stride=stride, padding=1,
                                            indice_key=f'{indice_key}_spconv', conv_type='spconv', dim=2)
        else:
            self.conv_down = None
        self.sst_input_layer = SSTInputLayer(model_cfg.PREPROCESS)
        self.encoder_blocks = nn.ModuleList([
            BasicShiftBlockV2(d_model, encoder_cfg.NHEAD, encoder_cfg.DIM_FEEDFORWARD, encoder_cfg.DROPOUT,
                              encoder_cfg.ACTIVATION, batch_first=False, layer_cfg=encoder_cfg.LAYER_CFG)
            for _ in range(encoder_cfg.NUM_BLOCKS)])
        self.conv_out = post_act_block(d_model, d_model, 3, norm_fn=norm_fn, indice_key=f'{indice_key}_subm', dim=2)

    def decouple_sp_tensor(self, sp_tensor):
        voxel_features = sp_tensor.features
        voxel_coords = sp_tensor.indices.long()
        voxel_coords = torch.cat([voxel_coords[:, 0:1], torch.zeros_like(voxel_coords[:, 0:1]), voxel_coords[:, 1:]], dim=-1)
        grid_size = sp_tensor.spatial_shape
        return voxel_features, voxel_coords, [grid_size[1], grid_size[0], 1]

    def encoder_forward(self, sp_tensor):
        voxel_info = self.sst_input_layer({'sp_tensor': sp_tensor, 'voxel_features': sp_tensor.features})
        ind_dict_list = [voxel_info[f'flat2win_inds_shift{i}'] for i in range(2)]
        pos_embed_list = [voxel_info[f'pos_dict_shift{i}'] for i in range(2)]
        output = sp_tensor.features
        for block in self.encoder_blocks:
            output = block(output, pos_embed_list, ind_dict_list, None)
        return output

    def forward(self, sp_tensor):
        if self.conv_down is not None:
            sp_tensor = self.conv_down(sp_tensor)
        encoded = self.encoder_forward(sp_tensor)
        sp_tensor = replace_feature(sp_tensor, sp_tensor.features + encoded)
        return self.conv_out(sp_tensor)


class SPTBackbone(nn.Module):
    """Mirror of pcdet/models/backbones_3d/spt_backbone.py:267-347: the Sparse Pyramid Transformer of the finetune /
    detection configs (tools/cfgs/*/gd_mae_iou.yaml) - the three SST blocks on ALL pillars (no masking: 35 k / 25.6 k /
    11 k tokens per Waymo frame, windows of all three drop levels), three deblocks, concat, 3x3 conv -> the dense
    ``spatial_features`` map the BEV backbone and CenterHead read.  Same state_dict names (``sst_blocks.*``,
    ``deblocks.{i}.{0,1}``, ``conv_out.{0,1}``) and batch_dict keys.  Every kernel is the one the MAE path uses; the deblocks
    run on the sparse rows and ``ops.DenseFill`` writes the 384-channel map once (see SPTBackboneMAE)."""

    def __init__(self, model_cfg, input_channels, grid_size, voxel_size, point_cloud_range, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.grid_size = [int(g) for g in grid_size]
        self.voxel_size = [float(v) for v in voxel_size]
        self.point_cloud_range = [float(v) for v in point_cloud_range]
        self.sparse_shape = [self.grid_size[1], self.grid_size[0]]
        in_channels = input_channels
        self.sst_blocks = nn.ModuleList()
        for cfg in model_cfg.SST_BLOCK_LIST:
            self.sst_blocks.append(SSTBlockV1(cfg, in_channels, cfg.NAME))
            in_channels = cfg.ENCODER.D_MODEL
        in_channels = 0
        self.deblocks = nn.ModuleList()
        self.fuse_strides = []
        for src in model_cfg.FEATURES_SOURCE:
            c = model_cfg.FUSE_LAYER[src]
            self.deblocks.append(nn.Sequential(
                nn.ConvTranspose2d(c.NUM_FILTER, c.NUM_UPSAMPLE_FILTER, c.UPSAMPLE_STRIDE, stride=c.UPSAMPLE_STRIDE, bias=False),
                nn.BatchNorm2d(c.NUM_UPSAMPLE_FILTER, eps=1e-3, momentum=0.01), nn.ReLU(inplace=True)))
            in_channels += c.NUM_UPSAMPLE_FILTER
            self.fuse_strides.append(int(c.UPSAMPLE_STRIDE))
        n_src = len(self.deblocks)
        if n_src != 3 or len({m[0].out_channels for m in self.deblocks}) != 1:
            raise NotImplementedError("the B200 decoder fill handles the three equal-width pyramid sources of the GD-MAE configs")
        self.conv_out = nn.Sequential(nn.Conv2d(in_channels, in_channels // n_src, 3, padding=1, bias=False),
                                      nn.BatchNorm2d(in_channels // n_src, eps=1e-3, momentum=0.01), nn.ReLU(inplace=True))
        self.num_point_features = in_channels // n_src
        self.decoder_dtype = torch.float32     # torch.bfloat16: dense map and 3x3 conv in bf16 (config.set_precision)

    def _deblock_rows(self, i, sp, n_cells_total):
        from .... import fused as _fused
        deconv, bn = self.deblocks[i][0], self.deblocks[i][1]
        k = self.fuse_strides[i]
        w = deconv.weight.permute(0, 2, 3, 1).reshape(deconv.in_channels, k * k * deconv.out_channels)
        u = (sp.features @ w).view(-1, deconv.out_channels)
        return _fused.batchnorm_relu(bn, u, self.training, relu=True, count=n_cells_total)

    def forward(self, batch_dict):
        import torch.nn.functional as F
        from .... import fused as _fused
        from .... import ops as _ops
        from ...utils.spconv_utils import spconv
        voxel_features, voxel_coords = batch_dict['voxel_features'], batch_dict['voxel_coords']
        batch_size = batch_dict['batch_size']
        if self.grid_size[2] != 1:
            raise AssertionError("pillar grids only: z extent must be 1 (spt_backbone.py:308)")
        Y, X = self.sparse_shape
        indices = voxel_coords[:, [0, 2, 3]].contiguous().int()
        x = spconv.SparseConvTensor(voxel_features, indices, self.sparse_shape, batch_size)
        x_hidden = []
        for blk in self.sst_blocks:
            x = blk(x)
            x_hidden.append(x)
        batch_dict.update({'encoded_spconv_tensor': x_hidden[-1], 'encoded_spconv_tensor_stride': 2 ** len(x_hidden)})
        feats = {f'x_conv{i + 1}': x_hidden[i] for i in range(len(x_hidden))}
        strides = {f'x_conv{i + 1}': 2 ** (i + 1) for i in range(len(x_hidden))}
        srcs = [feats[s] for s in self.model_cfg.FEATURES_SOURCE]
        n_cells_total = batch_size * Y * X
        rows, bgs = zip(*[self._deblock_rows(i, sp, n_cells_total) for i, sp in enumerate(srcs)])
        dt = self.decoder_dtype
        fused_map = _ops.DenseFill.apply(rows[0], rows[1], rows[2], bgs[0], bgs[1], bgs[2], [sp.rank_grid() for sp in srcs],
                                         [sp.indices for sp in srcs], self.fuse_strides, batch_size, Y, X, dt)     # (B, Y, X, 384)
        conv, bn = self.conv_out[0], self.conv_out[1]
        w = _fused.cast_param(conv.weight, dt)
        y = F.conv2d(fused_map.permute(0, 3, 1, 2), w.contiguous(memory_format=torch.channels_last), padding=1)
        y = F.batch_norm(y.float(), bn.running_mean, bn.running_var, bn.weight, bn.bias, self.training, bn.momentum, bn.eps)
        if self.training:
            bn.num_batches_tracked += 1
        spatial_features = F.relu(y)
        assert spatial_features.shape[0] == batch_size and spatial_features.shape[2] == Y and spatial_features.shape[3] == X
        batch_dict['multi_scale_3d_features'] = feats
        batch_dict['multi_scale_3d_strides'] = strides
        batch_dict['spatial_features'] = spatial_features
        batch_dict['spatial_features_stride'] = strides[self.model_cfg.FEATURES_SOURCE[0]] // self.fuse_strides[0]
        return batch_dict
