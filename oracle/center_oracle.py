"""TEST INFRASTRUCTURE - CPU restatement of the CenterHead pieces that run as own kernels on the finetune path
(SURVEY.md 8f rank 1).  Only tests/ may import this.  Each function cites the reference lines it follows
(paths relative to /root/reference).  Pinned by tests/golden/finetune_tiny.npz, written by the unmodified reference
(tests/golden/make_golden_finetune.py); tests/test_oracle_golden.py holds this file to it."""
import numpy as np
import torch


def gaussian_radius(h, w, min_overlap):
    """pcdet/models/model_utils/centernet_utils.py:9-37 (float32 torch arithmetic)"""
    b1 = h + w
    c1 = w * h * (1 - min_overlap) / (1 + min_overlap)
    r1 = (b1 + (b1 ** 2 - 4 * c1).sqrt()) / 2
    b2 = 2 * (h + w)
    c2 = (1 - min_overlap) * w * h
    r2 = (b2 + (b2 ** 2 - 16 * c2).sqrt()) / 2
    a3 = 4 * min_overlap
    b3 = -2 * min_overlap * (h + w)
    c3 = (min_overlap - 1) * w * h
    r3 = (b3 + (b3 ** 2 - 4 * a3 * c3).sqrt()) / 2
    return torch.min(torch.min(r1, r2), r3)


def assign_targets(gt_boxes, class_map, C, H, W, pc_range, voxel, stride=1, max_objs=500, overlap=0.1, min_radius=2):
    """CenterHead.assign_targets + assign_target_of_single_head for one head (pcdet/models/dense_heads/center_head.py:105-231)
    and draw_gaussian_to_heatmap / gaussian2D (centernet_utils.py:40-72).  gt_boxes (B, M, 8) float32 torch tensor, class id
    last (0 = padding); class_map[c] = 1-based id inside the head or 0."""
    B = gt_boxes.shape[0]
    heat = torch.zeros(B, C, H, W)
    tgt, iou_boxes = torch.zeros(B, max_objs, 8), torch.zeros(B, max_objs, 7)
    inds, mask = torch.zeros(B, max_objs, dtype=torch.int64), torch.zeros(B, max_objs, dtype=torch.int64)
    for b in range(B):
        rows = [g.clone() for g in gt_boxes[b] if class_map[int(g[7])] > 0]        # :196-208 (order kept)
        if not rows:
            continue
        g = torch.stack(rows)
        g[:, 7] = torch.tensor([float(class_map[int(v)]) for v in g[:, 7]])
        cx = torch.clamp((g[:, 0] - pc_range[0]) / voxel[0] / stride, min=0, max=W - 0.5)   # :125-128
        cy = torch.clamp((g[:, 1] - pc_range[1]) / voxel[1] / stride, min=0, max=H - 0.5)
        ci = torch.stack([cx, cy], 1).int()
        dx, dy = g[:, 3] / voxel[0] / stride, g[:, 4] / voxel[1] / stride
        radius = torch.clamp_min(gaussian_radius(dx, dy, overlap).int(), min_radius)           # :137-138
        for k in range(min(max_objs, g.shape[0])):
            if dx[k] <= 0 or dy[k] <= 0:
                continue
            r = int(radius[k])
            x, y = int(ci[k, 0]), int(ci[k, 1])
            d = 2 * r + 1
            yy, xx = np.ogrid[-r:r + 1, -r:r + 1]
            gauss = np.exp(-(xx * xx + yy * yy) / (2 * (d / 6) ** 2))                          # float64, centernet_utils.py:40-47
            gauss[gauss < np.finfo(gauss.dtype).eps * gauss.max()] = 0
            left, right, top, bottom = min(x, r), min(W - x, r + 1), min(y, r), min(H - y, r + 1)
            patch = torch.from_numpy(gauss[r - top:r + bottom, r - left:r + right]).float()
            c = int(g[k, 7]) - 1
            region = heat[b, c, y - top:y + bottom, x - left:x + right]
            if min(patch.shape) > 0 and min(region.shape) > 0:
                torch.max(region, patch, out=region)
            inds[b, k], mask[b, k] = y * W + x, 1
            tgt[b, k, 0], tgt[b, k, 1], tgt[b, k, 2] = cx[k] - ci[k, 0].float(), cy[k] - ci[k, 1].float(), g[k, 2]
            tgt[b, k, 3:6] = g[k, 3:6].log()
            tgt[b, k, 6], tgt[b, k, 7] = torch.cos(g[k, 6]), torch.sin(g[k, 6])
            iou_boxes[b, k] = g[k, :7]
    return heat, tgt, iou_boxes, inds, mask


def focal_loss_from_logits(logits, gt):
    """CenterHead.sigmoid (center_head.py:233-235) + neg_loss_cornernet (pcdet/utils/loss_utils.py:273-309)"""
    pred = torch.clamp(logits.sigmoid(), min=1e-4, max=1 - 1e-4)
    pos_inds, neg_inds = gt.eq(1).float(), gt.lt(1).float()
    pos_loss = (torch.log(pred) * torch.pow(1 - pred, 2) * pos_inds).sum()
    neg_loss = (torch.log(1 - pred) * torch.pow(pred, 2) * torch.pow(1 - gt, 4) * neg_inds).sum()
    num_pos = pos_inds.sum()
    return -neg_loss if num_pos == 0 else -(pos_loss + neg_loss) / num_pos
