"""Mirror of pcdet/models/model_utils/centernet_utils.py (the functions the CenterHead path uses).  Target drawing
(gaussian_radius / gaussian2D / draw_gaussian_to_heatmap, :9-72) runs inside ops.center_assign_targets on the device; the
decode side (:122-220) is index arithmetic on top-k results and stays torch."""
import torch


def gaussian_radius(height, width, min_overlap=0.5):
    """(N), (N) -> (N) smallest of the three CenterNet radii (centernet_utils.py:9-37)"""
    b1 = height + width
    c1 = width * height * (1 - min_overlap) / (1 + min_overlap)
    r1 = (b1 + (b1 ** 2 - 4 * c1).sqrt()) / 2
    b2 = 2 * (height + width)
    c2 = (1 - min_overlap) * width * height
    r2 = (b2 + (b2 ** 2 - 16 * c2).sqrt()) / 2
    a3 = 4 * min_overlap
    b3 = -2 * min_overlap * (height + width)
    c3 = (min_overlap - 1) * width * height
    r3 = (b3 + (b3 ** 2 - 4 * a3 * c3).sqrt()) / 2
    return torch.min(torch.min(r1, r2), r3)


def _gather_feat(feat, ind, mask=None):
    dim = feat.size(2)
    feat = feat.gather(1, ind.unsqueeze(2).expand(ind.size(0), ind.size(1), dim))
    if mask is not None:
        feat = feat[mask.unsqueeze(2).expand_as(feat)].view(-1, dim)
    return feat


def _transpose_and_gather_feat(feat, ind):
    """(B, C, H, W), (B, K) flat cell indices -> (B, K, C)"""
    feat = feat.permute(0, 2, 3, 1).contiguous()
    return _gather_feat(feat.view(feat.size(0), -1, feat.size(3)), ind)


def _topk(scores, K=40):
    """per-class top K, then the overall top K of those (centernet_utils.py:138-154)"""
    batch, num_class, height, width = scores.size()
    topk_scores, topk_inds = torch.topk(scores.flatten(2, 3), K)
    topk_inds = topk_inds % (height * width)
    topk_ys = torch.div(topk_inds, width, rounding_mode='floor').float()
    topk_xs = (topk_inds % width).int().float()
    topk_score, topk_ind = torch.topk(topk_scores.view(batch, -1), K)
    topk_classes = torch.div(topk_ind, K, rounding_mode='floor').int()
    pick = lambda t: _gather_feat(t.view(batch, -1, 1), topk_ind).view(batch, K)  # noqa: E731
    return topk_score, pick(topk_inds), topk_classes, pick(topk_ys), pick(topk_xs)


def decode_bbox_from_heatmap(heatmap, rot_cos, rot_sin, center, center_z, dim, vel=None, iou=None, point_cloud_range=None,
                             voxel_size=None, feature_map_stride=None, K=100, circle_nms=False, score_thresh=None,
                             post_center_limit_range=None):
    """centernet_utils.py:157-220 -> per frame {'pred_boxes', 'pred_scores', 'pred_ious', 'pred_labels'}"""
    assert not circle_nms, 'circle_nms is marked "not checked yet" in the reference (centernet_utils.py:163-166) and is not provided'
    assert post_center_limit_range is not None
    batch_size = heatmap.size(0)
    scores, inds, class_ids, ys, xs = _topk(heatmap, K=K)
    take = lambda t, c: _transpose_and_gather_feat(t, inds).view(batch_size, K, c)  # noqa: E731
    ious, center, rot_sin, rot_cos = take(iou, 1), take(center, 2), take(rot_sin, 1), take(rot_cos, 1)
    center_z, dim = take(center_z, 1), take(dim, 3)
    angle = torch.atan2(rot_sin, rot_cos)
    xs = (xs.view(batch_size, K, 1) + center[:, :, 0:1]) * feature_map_stride * voxel_size[0] + point_cloud_range[0]
    ys = (ys.view(batch_size, K, 1) + center[:, :, 1:2]) * feature_map_stride * voxel_size[1] + point_cloud_range[1]
    parts = [xs, ys, center_z, dim, angle]
    if vel is not None:
        parts.append(take(vel, 2))
    boxes = torch.cat(parts, dim=-1)
    mask = (boxes[..., :3] >= post_center_limit_range[:3]).all(2) & (boxes[..., :3] <= post_center_limit_range[3:]).all(2)
    if score_thresh is not None:
        mask &= scores > score_thresh
    return [{'pred_boxes': boxes[k, mask[k]], 'pred_scores': scores[k, mask[k]], 'pred_ious': ious.view(batch_size, K)[k, mask[k]],
             'pred_labels': class_ids[k, mask[k]]} for k in range(batch_size)]
