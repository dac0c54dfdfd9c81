"""Mirror of pcdet/models/dense_heads/center_head.py:11-392 (SeparateHead, CenterHead) for the GD-MAE finetune configs
(tools/cfgs/*/gd_mae_iou.yaml: one head over all classes, heads center / center_z / dim / rot / iou + hm).

Same constructor signatures and state_dict names (``shared_conv.{0,1}``, ``heads_list.{h}.{name}.{i}...``).  What changed:
* assign_targets: ONE kernel launch per head on the device (ops.center_assign_targets) instead of the reference's Python
  loop over frames and boxes on CPU tensors with a numpy gaussian per box (center_head.py:105-231);
* the heat-map focal loss takes the logits and runs as one fused kernel (pcdet/utils/loss_utils.FocalLossCenterNet);
* tb_dict holds tensors (no .item() host syncs inside get_loss, center_head.py:258-259,282).
The dense 3x3 convolutions are library calls (cuDNN through torch)."""
import copy

import numpy as np
import torch
import torch.nn as nn
from torch.nn.init import kaiming_normal_

from .... import ops as _ops
from ...utils import loss_utils
from ..model_utils import centernet_utils, model_nms_utils


class SeparateHead(nn.Module):
    def __init__(self, input_channels, sep_head_dict, init_bias=-2.19, use_bias=False):
        super().__init__()
        self.sep_head_dict = sep_head_dict
        for cur_name, spec in self.sep_head_dict.items():
            layers = [nn.Sequential(nn.Conv2d(input_channels, input_channels, kernel_size=3, stride=1, padding=1, bias=use_bias),
                                    nn.BatchNorm2d(input_channels), nn.ReLU()) for _ in range(spec['num_conv'] - 1)]
            layers.append(nn.Conv2d(input_channels, spec['out_channels'], kernel_size=3, stride=1, padding=1, bias=True))
            fc = nn.Sequential(*layers)
            if 'hm' in cur_name:
                fc[-1].bias.data.fill_(init_bias)
            else:
                for m in fc.modules():
                    if isinstance(m, nn.Conv2d):
                        kaiming_normal_(m.weight.data)
                        if m.bias is not None:
                            nn.init.constant_(m.bias, 0)
            self.__setattr__(cur_name, fc)

    def forward(self, x):
        return {name: self.__getattr__(name)(x) for name in self.sep_head_dict}


class CenterHead(nn.Module):
    def __init__(self, model_cfg, input_channels, num_class, class_names, grid_size, point_cloud_range, voxel_size,
                 predict_boxes_when_training=True, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_class = num_class
        self.grid_size = grid_size
        self.point_cloud_range = [float(v) for v in point_cloud_range]
        self.voxel_size = [float(v) for v in voxel_size]
        self.feature_map_stride = self.model_cfg.TARGET_ASSIGNER_CONFIG.get('FEATURE_MAP_STRIDE', None)
        self.class_names = list(class_names)
        self.class_names_each_head = [[x for x in names if x in self.class_names] for names in self.model_cfg.CLASS_NAMES_EACH_HEAD]
        # class id (0-based, global) of every class of a head, and the 1-based global id -> 1-based in-head id table the
        # target kernel reads (0 = class not in this head)
        self._id_mapping = [np.array([self.class_names.index(x) for x in names]) for names in self.class_names_each_head]
        self._class_map = []
        for names in self.class_names_each_head:
            cmap = np.zeros(len(self.class_names) + 1, dtype=np.int32)
            for j, x in enumerate(names):
                cmap[self.class_names.index(x) + 1] = j + 1
            self._class_map.append(cmap)
        self._dev_tables = {}
        assert sum(len(x) for x in self.class_names_each_head) == len(self.class_names), \
            f'class_names_each_head={self.class_names_each_head}'
        use_bias = self.model_cfg.get('USE_BIAS_BEFORE_NORM', False)
        self.shared_conv = nn.Sequential(
            nn.Conv2d(input_channels, self.model_cfg.SHARED_CONV_CHANNEL, 3, stride=1, padding=1, bias=use_bias),
            nn.BatchNorm2d(self.model_cfg.SHARED_CONV_CHANNEL), nn.ReLU())
        self.heads_list = nn.ModuleList()
        self.separate_head_cfg = self.model_cfg.SEPARATE_HEAD_CFG
        for names in self.class_names_each_head:
            head_dict = copy.deepcopy(dict(self.separate_head_cfg.HEAD_DICT))
            head_dict['hm'] = dict(out_channels=len(names), num_conv=self.model_cfg.NUM_HM_CONV)
            self.heads_list.append(SeparateHead(self.model_cfg.SHARED_CONV_CHANNEL, head_dict, init_bias=-2.19, use_bias=use_bias))
        self.with_iou = 'iou' in self.separate_head_cfg.HEAD_DICT
        self.predict_boxes_when_training = predict_boxes_when_training
        self.forward_ret_dict = {}
        self.build_losses()

    def build_losses(self):
        self.add_module('hm_loss_func', loss_utils.FocalLossCenterNet())
        self.add_module('reg_loss_func', loss_utils.RegLossCenterNet())
        if self.with_iou:
            self.add_module('iou_loss_func', loss_utils.IoULossCenterNet())

    def _tables(self, device):
        key = str(device)
        if key not in self._dev_tables:
            self._dev_tables[key] = ([torch.from_numpy(m).to(device) for m in self._id_mapping],
                                     [torch.from_numpy(m).to(device) for m in self._class_map])
        return self._dev_tables[key]

    def assign_targets(self, gt_boxes, feature_map_size=None, **kwargs):
        """gt_boxes (B, M, 8) on the device, feature_map_size [H, W] -> the reference's ret_dict of per-head lists"""
        cfg = self.model_cfg.TARGET_ASSIGNER_CONFIG
        ret_dict = {'heatmaps': [], 'target_boxes': [], 'iou_boxes': [], 'inds': [], 'masks': []}
        _, class_maps = self._tables(gt_boxes.device)
        for h, names in enumerate(self.class_names_each_head):
            heat, tgt, iou_boxes, inds, mask = _ops.center_assign_targets(
                gt_boxes, class_maps[h], len(names), feature_map_size, self.point_cloud_range, self.voxel_size,
                cfg.FEATURE_MAP_STRIDE, num_max_objs=cfg.NUM_MAX_OBJS, gaussian_overlap=cfg.GAUSSIAN_OVERLAP, min_radius=cfg.MIN_RADIUS)
            for k, v in zip(('heatmaps', 'target_boxes', 'iou_boxes', 'inds', 'masks'), (heat, tgt, iou_boxes, inds, mask)):
                ret_dict[k].append(v)
        return ret_dict

    def sigmoid(self, x):
        return torch.clamp(x.sigmoid(), min=1e-4, max=1 - 1e-4)

    def get_loss(self):
        pred_dicts, target_dicts = self.forward_ret_dict['pred_dicts'], self.forward_ret_dict['target_dicts']
        weights = self.model_cfg.LOSS_CONFIG.LOSS_WEIGHTS
        tb_dict, loss = {}, 0
        for idx, pred_dict in enumerate(pred_dicts):
            hm_loss = self.hm_loss_func(pred_dict['hm'], target_dicts['heatmaps'][idx], from_logits=True) * weights['cls_weight']
            pred_boxes = torch.cat([pred_dict[name] for name in self.separate_head_cfg.HEAD_ORDER], dim=1)
            reg_loss = self.reg_loss_func(pred_boxes, target_dicts['masks'][idx], target_dicts['inds'][idx], target_dicts['target_boxes'][idx])
            loc_loss = (reg_loss * reg_loss.new_tensor(weights['code_weights'])).sum() * weights['loc_weight']
            loss = loss + hm_loss + loc_loss
            tb_dict['hm_loss_head_%d' % idx], tb_dict['loc_loss_head_%d' % idx] = hm_loss.detach(), loc_loss.detach()
            if self.with_iou:
                batch_dim = pred_dict['dim'].exp()
                batch_rot = torch.atan2(pred_dict['rot'][:, 1:2], pred_dict['rot'][:, 0:1])
                B, _, H, W = batch_dim.shape
                ys, xs = torch.meshgrid([torch.arange(H, device=batch_dim.device), torch.arange(W, device=batch_dim.device)], indexing='ij')
                xs = xs.to(batch_dim).view(1, 1, H, W) + pred_dict['center'][:, 0:1]
                ys = ys.to(batch_dim).view(1, 1, H, W) + pred_dict['center'][:, 1:2]
                xs = xs * self.feature_map_stride * self.voxel_size[0] + self.point_cloud_range[0]
                ys = ys * self.feature_map_stride * self.voxel_size[1] + self.point_cloud_range[1]
                batch_box_preds = torch.cat([xs, ys, pred_dict['center_z'], batch_dim, batch_rot], dim=1)      # (B, 7, H, W)
                iou_loss = self.iou_loss_func(pred_dict['iou'], target_dicts['masks'][idx], target_dicts['inds'][idx],
                                              batch_box_preds.detach(), target_dicts['iou_boxes'][idx]) * weights['iou_weight']
                loss = loss + iou_loss
                tb_dict['iou_loss_head_%d' % idx] = iou_loss.detach()
        return loss, tb_dict

    def generate_predicted_boxes(self, batch_size, pred_dicts):
        """center_head.py:286-348: top-K decode per head, score / range filter, (multi-class) rotated NMS"""
        cfg = self.model_cfg.POST_PROCESSING
        device = pred_dicts[0]['hm'].device
        limit = torch.tensor(cfg.POST_CENTER_LIMIT_RANGE, device=device).float()
        id_maps, _ = self._tables(device)
        ret = [{'pred_boxes': [], 'pred_scores': [], 'pred_labels': []} for _ in range(batch_size)]
        for idx, pred_dict in enumerate(pred_dicts):
            batch_hm = pred_dict['hm'].sigmoid()
            batch_iou = torch.clamp((pred_dict['iou'] + 1) * 0.5, min=0, max=1) if 'iou' in pred_dict else torch.ones_like(batch_hm[:, 0:1])
            finals = centernet_utils.decode_bbox_from_heatmap(
                heatmap=batch_hm, rot_cos=pred_dict['rot'][:, 0:1], rot_sin=pred_dict['rot'][:, 1:2], center=pred_dict['center'],
                center_z=pred_dict['center_z'], dim=pred_dict['dim'].exp(),
                vel=pred_dict['vel'] if 'vel' in self.separate_head_cfg.HEAD_ORDER else None, iou=batch_iou,
                point_cloud_range=self.point_cloud_range, voxel_size=self.voxel_size, feature_map_stride=self.feature_map_stride,
                K=cfg.MAX_OBJ_PER_SAMPLE, circle_nms=(cfg.NMS_CONFIG.NMS_TYPE == 'circle_nms'), score_thresh=cfg.SCORE_THRESH,
                post_center_limit_range=limit)
            for k, fd in enumerate(finals):
                fd['pred_labels'] = id_maps[idx][fd['pred_labels'].long()]
                if cfg.NMS_CONFIG.NMS_TYPE == 'nms_gpu':
                    selected, selected_scores = model_nms_utils.class_agnostic_nms(fd['pred_scores'], fd['pred_boxes'], cfg.NMS_CONFIG, None)
                elif cfg.NMS_CONFIG.NMS_TYPE == 'multi_class_nms':
                    selected, selected_scores = model_nms_utils.multi_class_agnostic_nms(fd['pred_scores'], fd['pred_ious'], fd['pred_labels'],
                                                                                        fd['pred_boxes'], cfg.NMS_CONFIG)
                else:
                    raise NotImplementedError
                ret[k]['pred_boxes'].append(fd['pred_boxes'][selected])
                ret[k]['pred_scores'].append(selected_scores)
                ret[k]['pred_labels'].append(fd['pred_labels'][selected])
        for k in range(batch_size):
            ret[k]['pred_boxes'] = torch.cat(ret[k]['pred_boxes'], dim=0)
            ret[k]['pred_scores'] = torch.cat(ret[k]['pred_scores'], dim=0)
            ret[k]['pred_labels'] = torch.cat(ret[k]['pred_labels'], dim=0) + 1
        return ret

    @staticmethod
    def reorder_rois_for_refining(batch_size, pred_dicts):
        num_max_rois = max(1, max(len(d['pred_boxes']) for d in pred_dicts))
        ref = pred_dicts[0]['pred_boxes']
        rois = ref.new_zeros((batch_size, num_max_rois, ref.shape[-1]))
        roi_scores = ref.new_zeros((batch_size, num_max_rois))
        roi_labels = ref.new_zeros((batch_size, num_max_rois)).long()
        for b in range(batch_size):
            n = len(pred_dicts[b]['pred_boxes'])
            rois[b, :n], roi_scores[b, :n], roi_labels[b, :n] = pred_dicts[b]['pred_boxes'], pred_dicts[b]['pred_scores'], pred_dicts[b]['pred_labels']
        return rois, roi_scores, roi_labels

    def forward(self, data_dict):
        spatial_features_2d = data_dict['spatial_features_2d']
        x = self.shared_conv(spatial_features_2d)
        pred_dicts = [head(x) for head in self.heads_list]
        if self.training:
            self.forward_ret_dict['target_dicts'] = self.assign_targets(
                data_dict['gt_boxes'], feature_map_size=spatial_features_2d.size()[2:],
                feature_map_stride=data_dict.get('spatial_features_2d_strides', None))
        self.forward_ret_dict['pred_dicts'] = pred_dicts
        if not self.training or self.predict_boxes_when_training:
            boxes = self.generate_predicted_boxes(data_dict['batch_size'], pred_dicts)
            data_dict['cls_preds_normalized'] = True
            if self.predict_boxes_when_training:
                data_dict['rois'], data_dict['roi_scores'], data_dict['roi_labels'] = self.reorder_rois_for_refining(data_dict['batch_size'], boxes)
                data_dict['has_class_labels'] = True
            else:
                data_dict['final_box_dicts'] = boxes
        return data_dict
