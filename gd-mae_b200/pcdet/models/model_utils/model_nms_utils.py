"""Mirror of pcdet/models/model_utils/model_nms_utils.py:6-46 over the B200 NMS kernels (pcdet/ops/iou3d_nms)."""
import torch

from ...ops.iou3d_nms import iou3d_nms_utils


def class_agnostic_nms(box_scores, box_preds, nms_config, score_thresh=None):
    src_box_scores = box_scores
    if score_thresh is not None:
        scores_mask = box_scores >= score_thresh
        box_scores, box_preds = box_scores[scores_mask], box_preds[scores_mask]
    selected = []
    if box_scores.shape[0] > 0:
        top_scores, indices = torch.topk(box_scores, k=min(nms_config.NMS_PRE_MAXSIZE, box_scores.shape[0]))
        keep_idx, _ = getattr(iou3d_nms_utils, nms_config.NMS_TYPE)(box_preds[indices][:, 0:7], top_scores, nms_config.NMS_THRESH,
                                                                    **nms_config)
        selected = indices[keep_idx[:nms_config.NMS_POST_MAXSIZE]]
    if score_thresh is not None:
        selected = scores_mask.nonzero().view(-1)[selected]
    return selected, src_box_scores[selected]


def multi_class_agnostic_nms(box_scores, box_ious, box_labels, box_preds, nms_config):
    """per-class rotated NMS on IoU-rectified scores score^(1-a) * iou^a (model_nms_utils.py:28-46)"""
    alpha = box_scores.new_tensor(nms_config.IOU_RECTIFIER)[box_labels.long()]
    rect_scores = torch.pow(box_scores, 1 - alpha) * torch.pow(box_ious, alpha)
    selected = []
    for cls in range(len(nms_config.NMS_THRESH)):
        src_idx = (box_labels == cls).nonzero(as_tuple=True)[0]
        if src_idx.numel() == 0:
            continue
        top_scores, indices = torch.topk(rect_scores[src_idx], k=min(nms_config.NMS_PRE_MAXSIZE[cls], src_idx.numel()))
        keep_idx, _ = iou3d_nms_utils.nms_gpu(box_preds[src_idx][indices][:, 0:7], top_scores, nms_config.NMS_THRESH[cls])
        selected.append(src_idx[indices[keep_idx[:nms_config.NMS_POST_MAXSIZE[cls]]]])
    if len(selected) > 0:
        selected = torch.cat(selected, dim=0)
    return selected, rect_scores[selected]
