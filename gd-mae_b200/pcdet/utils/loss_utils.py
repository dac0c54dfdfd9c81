"""Mirror of the CenterNet losses of pcdet/utils/loss_utils.py:273-418 (FocalLossCenterNet, RegLossCenterNet,
IoULossCenterNet).  The heat-map focal loss is ONE fused kernel over the logits (ops.CenterFocalLoss: clamped sigmoid, both
terms, the positive count and the gradient in a single pass over the (B, C, H, W) map; the reference builds ~15 map-sized
temporaries); the two sparse losses gather <= 500 cells per frame and stay index arithmetic in torch."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..models.model_utils.centernet_utils import _transpose_and_gather_feat
from ..ops.iou3d_nms import iou3d_nms_utils


class FocalLossCenterNet(nn.Module):
    """forward(out, target): ``out`` are the heat-map LOGITS when ``from_logits`` (the CenterHead mirror passes them and
    the clamped sigmoid of center_head.py:233-235 is applied inside the kernel), else probabilities (reference signature)."""

    def forward(self, out, target, mask=None, from_logits=False):
        from ... import ops as _ops
        if from_logits and mask is None and out.is_cuda:
            return _ops.CenterFocalLoss.apply(out, target)
        return neg_loss_cornernet(torch.clamp(out.sigmoid(), min=1e-4, max=1 - 1e-4) if from_logits else out, target, mask=mask)


def neg_loss_cornernet(pred, gt, mask=None):
    """loss_utils.py:273-309 on probabilities (kept for callers that pass a mask or probabilities)"""
    pos_inds, neg_inds = gt.eq(1).float(), gt.lt(1).float()
    pos_loss = torch.log(pred) * torch.pow(1 - pred, 2) * pos_inds
    neg_loss = torch.log(1 - pred) * torch.pow(pred, 2) * torch.pow(1 - gt, 4) * neg_inds
    if mask is not None:
        m = mask[:, None, :, :].float()
        pos_loss, neg_loss, num_pos = pos_loss * m, neg_loss * m, (pos_inds * m).sum()
    else:
        num_pos = pos_inds.sum()
    pos_loss, neg_loss = pos_loss.sum(), neg_loss.sum()
    return -neg_loss if num_pos == 0 else -(pos_loss + neg_loss) / num_pos


def _reg_loss(regr, gt_regr, mask):
    """L1 over the masked objects, per regression channel, divided by max(#objects, 1) (loss_utils.py:323-352)"""
    num = mask.float().sum()
    m = mask.unsqueeze(2).expand_as(gt_regr).float() * (~torch.isnan(gt_regr)).float()
    loss = torch.abs(regr * m - gt_regr * m).sum(dim=(0, 1))
    return loss / torch.clamp_min(num, min=1.0)


class RegLossCenterNet(nn.Module):
    def forward(self, output, mask, ind=None, target=None):
        pred = output if ind is None else _transpose_and_gather_feat(output, ind)
        return _reg_loss(pred, target, mask)


class IoULossCenterNet(nn.Module):
    """L1 between the predicted IoU and 2 * IoU3D(decoded box, gt box) - 1 at the object cells (loss_utils.py:398-418).  The
    reference takes the diagonal of an (n, n) IoU matrix; the same values are computed here on the n pairs only."""

    def forward(self, iou_pred, mask, ind, box_pred, box_gt):
        mask = mask.bool()
        pred = _transpose_and_gather_feat(iou_pred, ind)[mask]
        pred_box = _transpose_and_gather_feat(box_pred, ind)
        a, b = pred_box[mask].contiguous(), box_gt[mask].contiguous()
        target = torch.diagonal(iou3d_nms_utils.boxes_iou3d_gpu(a, b)).unsqueeze(-1) if a.shape[0] > 0 else a.new_zeros((0, 1))
        target = 2 * target - 1
        loss = F.l1_loss(pred, target, reduction='sum')
        return loss / (mask.sum() + 1e-4)
