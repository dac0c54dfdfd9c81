"""tests/golden/finetune_tiny.npz: one training step (forward + loss + backward) of the reference's finetune detector
(tools/cfgs/waymo_models/gd_mae_iou.yaml: DynVFE + SPTBackbone + SSTBEVBackbone + CenterHead) on the tiny grid, run from the
UNMODIFIED reference Python through ref_harness.py.  Run in the build container only:
    python tests/golden/make_golden_finetune.py
Stores the input points / gt boxes, the seeds of the weights, the assigned targets, every loss term, features and gradient
norms.  Weights are not stored: both sides rebuild them from oracle.init_params (backbone) and oracle.finetune_head_state."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness as RH  # noqa: E402
from oracle import gdmae_oracle as O  # noqa: E402

SUB = 8
CLASS_NAMES = ['Vehicle', 'Pedestrian', 'Cyclist']


def tiny_scene(seed, cfg, n_per_frame=2500, B=2, n_boxes=9):
    r = np.random.RandomState(seed)
    lim = cfg["pc_range"]
    pts, boxes = [], []
    for b in range(B):
        n = n_per_frame + 101 * b
        xy = np.stack([r.uniform(lim[0], lim[3], n), r.uniform(lim[1], lim[4], n)], 1)
        p = np.concatenate([xy, r.uniform(-1.5, 2.0, (n, 1)), r.uniform(0, 1, (n, cfg["n_feat"] - 3))], 1)
        pts.append(np.concatenate([np.full((n, 1), b), p], 1))
        cls = r.randint(1, 4, n_boxes)
        dims = np.array([[4.7, 2.1, 1.7], [0.9, 0.9, 1.7], [1.8, 0.8, 1.7]])[cls - 1] * r.uniform(0.8, 1.2, (n_boxes, 3))
        c = np.stack([r.uniform(lim[0] + 1, lim[3] - 1, n_boxes), r.uniform(lim[1] + 1, lim[4] - 1, n_boxes), r.uniform(-0.5, 1.0, n_boxes)], 1)
        bx = np.concatenate([c, dims, r.uniform(-np.pi, np.pi, (n_boxes, 1)), cls[:, None]], 1)
        bx[0, :2] = [lim[3] - 0.05, lim[4] - 0.05]     # centre in the last cell: the clamp to size - 0.5 and the clipped gaussian
        bx[1, :2] = bx[2, :2] + 0.4                       # two boxes of possibly one class next to each other: overlapping gaussians
        if b == 1:
            bx[-2:, :] = 0                                # collate_batch zero-pads to the longest frame (dataset.py:185-190)
        boxes.append(bx)
    return np.concatenate(pts, 0).astype(np.float32), np.stack(boxes, 0).astype(np.float32)


if __name__ == "__main__":
    cfg = O.make_cfg("tiny")
    mcfg = RH.load_model_cfg("tools/cfgs/waymo_models/gd_mae_iou.yaml").MODEL
    grid = np.array(cfg["grid"], dtype=np.int64)
    torch.manual_seed(31)
    model = RH.RefCenterPoint(mcfg, cfg["n_feat"], cfg["voxel"], np.array(cfg["pc_range"], dtype=np.float32), grid, CLASS_NAMES)
    # backbone weights from the oracle's initialiser (as the MAE goldens), BEV backbone / head weights from
    # O.finetune_head_state (a function of key, shape and seed): nothing but the seeds is stored
    P, Bf = O.init_params(cfg, 7)
    sd = model.state_dict()
    sd.update(O.finetune_head_state({k: v.shape for k, v in sd.items() if k.startswith(("backbone_2d.", "dense_head."))}, seed=3))
    for k, v in list(P.items()) + list(Bf.items()):
        k2 = k.replace("backbone_3d.decoder_deblocks", "backbone_3d.deblocks").replace("backbone_3d.decoder_conv_out", "backbone_3d.conv_out")
        if k2 in sd and sd[k2].shape == v.shape:
            sd[k2] = v.clone()
    model.load_state_dict(sd)
    pts, gt = tiny_scene(5, cfg)
    model.train()
    out = {"points_in": pts, "gt_boxes": gt, "batch_size": np.int64(2)}
    out["param_seed"], out["head_seed"] = np.int64(7), np.int64(3)
    out["state_keys"] = np.array(list(sd.keys()))
    loss, tb, bd = model(dict(points=torch.from_numpy(pts), batch_size=2, gt_boxes=torch.from_numpy(gt)))
    loss.backward()
    out["loss"] = loss.detach().numpy()
    for k, v in tb.items():
        out["tb." + k] = np.float64(v)
    td = model.dense_head.forward_ret_dict["target_dicts"]
    hm = td["heatmaps"][0].numpy()
    nz = np.nonzero(hm.reshape(-1))[0]
    out["heatmap.shape"], out["heatmap.nz_index"], out["heatmap.nz_value"] = np.array(hm.shape), nz.astype(np.int64), hm.reshape(-1)[nz]
    out["target_boxes"], out["iou_boxes"] = td["target_boxes"][0].numpy(), td["iou_boxes"][0].numpy()
    out["inds"], out["masks"] = td["inds"][0].numpy(), td["masks"][0].numpy()
    out["spatial_features.sub"] = bd["spatial_features"].detach()[:, ::16, ::5, ::5].numpy()
    out["spatial_features_2d.sub"] = bd["spatial_features_2d"].detach()[:, ::16, ::5, ::5].numpy()
    for i in range(3):
        sp = bd["multi_scale_3d_features"][f"x_conv{i + 1}"]
        out[f"x_conv{i + 1}.indices"], out[f"x_conv{i + 1}.features.sub"] = sp.indices.numpy(), sp.features.detach()[::SUB].numpy()
    pd = model.dense_head.forward_ret_dict["pred_dicts"][0]
    for k, v in pd.items():
        out["pred." + k + ".sub"] = v.detach()[:, :, ::5, ::5].numpy()
    names, norms = [], []
    for k, p in model.named_parameters():
        names.append(k)
        norms.append(float(p.grad.norm()) if p.grad is not None else 0.0)
    out["grad_keys"], out["grad_norms"] = np.array(names), np.array(norms, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "finetune_tiny.npz"), **out)
    print("loss", float(loss), {k: float(v) for k, v in tb.items()}, "objects", int(out["masks"].sum()),
          "tokens", [out[f"x_conv{i+1}.indices"].shape[0] for i in range(3)], "size MB", os.path.getsize(os.path.join(HERE, "finetune_tiny.npz")) / 1e6)
