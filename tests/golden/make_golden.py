"""Generate the golden fixtures in tests/golden/*.npz by running the UNMODIFIED reference
Python (imported from /root/reference through ref_harness.py) on seeded inputs.

Run in the build container only:  python tests/golden/make_golden.py
The fixtures are committed; nothing at test/bench time reads /root/reference.

Weights are not stored: both sides rebuild them from ``oracle.init_params(cfg, seed)``;
the reference model receives them through ``load_state_dict(strict=True)``, which also
pins the state_dict key/shape schema.
"""
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


def tiny_points(seed, cfg, n_per_frame, B):
    """Small dense scene inside the tiny range, with a few out-of-range / boundary points
    (x just below max, z below range) to exercise the divide-then-truncate quirks."""
    r = np.random.RandomState(seed)
    lim = cfg["pc_range"]
    rows = []
    for b in range(B):
        n = n_per_frame + 37 * b
        xy = r.normal(0, 3.0, (n, 2)).astype(np.float32)
        z = r.uniform(lim[2] - 3.0, lim[5] + 0.5, (n, 1)).astype(np.float32)
        f = r.uniform(0, 1, (n, cfg["n_feat"] - 3)).astype(np.float32)
        p = np.concatenate([xy, z, f], axis=1)
        p[0, 0] = np.float32(lim[3]) - np.float32(1e-6)   # rounds to the max edge
        p[1, 0] = np.float32(lim[0])                        # exactly the min edge -> bin 0
        p[2, 1] = np.float32(lim[4])                        # exactly the max edge -> dropped
        p[3, :2] = 0.0
        p[4, :2] = 0.0                                      # duplicate xy
        rows.append(np.concatenate([np.full((n, 1), b, np.float32), p], axis=1))
    return np.concatenate(rows, 0)


def build_ref(cfg, yaml_rel, seed):
    mcfg = RH.load_model_cfg(yaml_rel).MODEL
    mcfg.BACKBONE_3D.MASK_CONFIG.RATIO = cfg["mask_ratio"]  # same effect as tools/train.py --set
    grid = np.array(cfg["grid"], dtype=np.int64)
    model = RH.RefMAE(mcfg, cfg["n_feat"], cfg["voxel"], np.array(cfg["pc_range"], dtype=np.float32), grid)
    P, Bf = O.init_params(cfg, seed)
    sd = dict(P)
    sd.update(Bf)
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model, P, Bf


SUB = 8


def t2n(t):
    return t.detach().cpu().numpy()


def golden_mae(name, cfg, yaml_rel, pts_np, B, seed):
    torch.manual_seed(666)  # train.py:82 --fix_random_seed value
    model, P, Bf = build_ref(cfg, yaml_rel, seed)
    model.train()
    bd = dict(points=torch.from_numpy(pts_np), batch_size=B)
    loss, bd = model(bd)
    loss.backward()
    out = dict(points_in=pts_np, batch_size=np.int64(B), param_seed=np.int64(seed), loss=t2n(loss))
    # integer outputs in full (they compress well); float feature rows sub-sampled (every SUB-th row)
    for k in ["point_coords", "point_inverse_indices", "voxel_coords", "voxel_mae_mask"]:
        out[k] = t2n(bd[k])
    out["n_points_kept"] = np.int64(bd["points"].shape[0])
    for k in ["pillar_features", "voxel_features"]:
        out[k + ".sub"] = t2n(bd[k][::SUB])
    for i in range(3):
        sp = bd["multi_scale_3d_features"][f"x_conv{i + 1}"]
        out[f"x_conv{i + 1}.features.sub"], out[f"x_conv{i + 1}.indices"] = t2n(sp.features[::SUB]), t2n(sp.indices)
    sf = bd["spatial_features"]
    out["spatial_features.shape"] = np.array(sf.shape)
    out["spatial_features.mean_c"] = t2n(sf.mean(dim=(0, 2, 3)))
    out["spatial_features.sub"] = t2n(sf[:, ::16, ::5, ::5])
    frd = model.backbone_3d.forward_ret_dict
    out["pred_points.sub"], out["gt_points.sub"] = t2n(frd["pred_points"][::SUB]), t2n(frd["gt_points"][::SUB])
    gn = {}
    for k, p in model.named_parameters():
        gn[k] = float(p.grad.norm()) if p.grad is not None else 0.0
    out["grad_keys"] = np.array(list(gn.keys()))
    out["grad_norms"] = np.array(list(gn.values()), dtype=np.float64)
    for k in ["vfe.dvfe_mlps.0.0.weight", "vfe.dvfe_mlps.0.1.weight", "backbone_3d.decoder_pred.weight",
              "backbone_3d.sst_blocks.0.encoder_blocks.0.encoder_list.0.win_attn.self_attn.tau",
              "backbone_3d.sst_blocks.1.encoder_blocks.1.encoder_list.1.norm2.weight",
              "backbone_3d.sst_blocks.2.conv_out.1.bias"]:
        out["grad." + k] = t2n(dict(model.named_parameters())[k].grad)
    sdict = model.state_dict()
    for k in ["vfe.dvfe_mlps.0.1.running_mean", "vfe.dvfe_mlps.0.4.running_var",
              "backbone_3d.decoder_conv_out.1.running_mean", "backbone_3d.sst_blocks.1.conv_down.1.running_var"]:
        out["buf." + k] = t2n(sdict[k])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss", float(loss), "Np", int(out["n_points_kept"]), "M", out["voxel_coords"].shape[0],
          "tokens", [out[f"x_conv{i+1}.indices"].shape[0] for i in range(3)])
    return model, bd


def golden_window(name, model, bd):
    """Per-function KATs of the window bookkeeping (a10-a16) on a dense random occupancy
    (70 % of a 20x30 grid, B=2) so that all three drop levels occur."""
    R = RH.ref_modules()
    import spconv.pytorch as spconv
    out = {}
    for bi in (0, 1):
        blk = model.backbone_3d.sst_blocks[bi]
        g = torch.Generator().manual_seed(100 + bi)
        occ = torch.rand(2, 30, 20, generator=g) < 0.7
        occ[1, :, 10:] &= torch.rand(30, 10, generator=g) < 0.2
        idx = torch.nonzero(occ).int()
        d = 128 if bi == 0 else 256
        sp = spconv.SparseConvTensor(torch.zeros(idx.shape[0], d), idx, [30, 20], 2)
        feats, coords, grid = blk.decouple_sp_tensor(sp)
        info = blk.sst_input_layer(dict(voxel_features=feats, voxel_coords=coords,
                                        voxel_shuffle_inds=torch.arange(coords.shape[0]), grid_size=grid))
        out[f"b{bi}.coords"] = t2n(coords)
        out[f"b{bi}.grid"] = np.array(grid)
        for s in range(2):
            out[f"b{bi}.s{s}.batch_win_inds"] = t2n(info[f"batch_win_inds_shift{s}"])
            out[f"b{bi}.s{s}.coors_in_win"] = t2n(info[f"coors_in_win_shift{s}"])
            out[f"b{bi}.s{s}.drop_level"] = t2n(info[f"voxel_drop_level_shift{s}"])
            f2w = info[f"flat2win_inds_shift{s}"]
            for dl in (0, 1, 2):
                if dl in f2w:
                    out[f"b{bi}.s{s}.l{dl}.flat2win"] = t2n(f2w[dl][0])
                    out[f"b{bi}.s{s}.l{dl}.where"] = t2n(f2w[dl][1][0])
                    out[f"b{bi}.s{s}.l{dl}.key_mask"] = t2n(info[f"key_mask_shift{s}"][dl])
                    pe = info[f"pos_dict_shift{s}"][dl]
                    out[f"b{bi}.s{s}.l{dl}.pos_shape"] = np.array(pe.shape)
                    out[f"b{bi}.s{s}.l{dl}.pos_sub"] = t2n(pe[:4])
        # one attention layer (a17-a19) on random features
        g = torch.Generator().manual_seed(5 + bi)
        x = torch.randn(coords.shape[0], feats.shape[1], generator=g)
        layer = blk.encoder_blocks[0].encoder_list[1]
        y = layer(x, info["pos_dict_shift1"], info["flat2win_inds_shift1"], info["key_mask_shift1"])
        a = layer.win_attn(x, info["pos_dict_shift1"], info["flat2win_inds_shift1"], info["key_mask_shift1"])
        out[f"b{bi}.layer_in_seed"] = np.int64(5 + bi)
        out[f"b{bi}.layer_out.sub"], out[f"b{bi}.attn_out.sub"] = t2n(y[::4]), t2n(a[::4])
    # sst_ops KATs (canonical sequential order)
    g = torch.Generator().manual_seed(11)
    grp = torch.randint(0, 40, (500,), generator=g)
    out["ops.group_inds"] = t2n(grp)
    out["ops.inner"] = t2n(R.sst_ops_utils.get_inner_win_inds(grp))
    inv = torch.randint(0, 25, (900,), generator=g)
    inv[:80] = 3  # a pillar with > K points
    ptsx = torch.randn(900, 3, generator=g)
    out["ops.inverse"], out["ops.points"] = t2n(inv), t2n(ptsx)
    out["ops.grouped"] = t2n(R.sst_ops_utils.group_inner_inds(ptsx, inv, 64))
    # random_masking (a7): same generator state on both sides
    torch.manual_seed(123)
    m = R.common_utils.random_masking(1, 777, 0.85, "cpu")[0]
    torch.manual_seed(123)
    out["mask.noise"] = t2n(torch.rand(1, 777)[0])
    out["mask.mask"] = t2n(m)
    # pos-embed closed form check input: full table via reference on all 64 cells
    il = model.backbone_3d.sst_blocks[0].sst_input_layer
    ciw = torch.stack([torch.zeros(64, dtype=torch.long), torch.arange(64) // 8, torch.arange(64) % 8], 1)
    fake = {0: (torch.arange(64), (torch.arange(64),)), "voxel_drop_level": torch.zeros(64, dtype=torch.long),
            "batching_info": {0: {"max_tokens": 64, "drop_range": (0, 100000)}}}
    for d in (128, 256):
        out[f"pos_table.{d}"] = t2n(il.get_pos_embed(fake, ciw, d)[0][0])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "written", len(out), "arrays")


def golden_kitti_c1(name):
    """BASELINE config C1 (plumbing): KITTI-shape frame, DynVFE + sst_block_x1 forward, no grad."""
    R = RH.ref_modules()
    cfg = O.make_cfg("kitti")
    r = np.random.RandomState(0)
    n = 4000
    az = r.uniform(-0.78, 0.78, n)
    rr = np.minimum(3 + r.exponential(18.0, n), 69.0)
    pts = np.stack([rr * np.cos(az), rr * np.sin(az), r.uniform(-2.5, 0.5, n), r.uniform(0, 1, n)], 1).astype(np.float32)
    pts = np.concatenate([np.zeros((n, 1), np.float32), pts], 1)
    mcfg = RH.load_model_cfg("tools/cfgs/kitti_models/gd_mae.yaml").MODEL
    torch.manual_seed(0)
    vfe = R.dyn_vfe.DynVFE(model_cfg=mcfg.VFE, num_point_features=4, voxel_size=cfg["voxel"],
                           point_cloud_range=np.array(cfg["pc_range"], dtype=np.float32), grid_size=np.array(cfg["grid"]))
    blk = R.spt_backbone.SSTBlockV1(mcfg.BACKBONE_3D.SST_BLOCK_LIST[0], 128, "sst_block_x1")
    P, _ = O.init_params(cfg, 3)
    vfe.load_state_dict({k[len("vfe."):]: v for k, v in P.items() if k.startswith("vfe.")}, strict=False)
    blk.load_state_dict({k[len("backbone_3d.sst_blocks.0."):]: v for k, v in P.items()
                         if k.startswith("backbone_3d.sst_blocks.0.")}, strict=False)
    import spconv.pytorch as spconv
    with torch.no_grad():
        bd = vfe(dict(points=torch.from_numpy(pts), batch_size=1))
        vc = bd["voxel_coords"]
        sp = spconv.SparseConvTensor(bd["voxel_features"], vc[:, [0, 2, 3]].contiguous().int(),
                                     np.array(cfg["grid"])[[1, 0]], 1)
        y = blk(sp)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), points_in=pts, param_seed=np.int64(3),
                        voxel_coords=t2n(vc), inverse=t2n(bd["point_inverse_indices"]),
                        pillar_features_sub=t2n(bd["voxel_features"][::SUB]), block_out_sub=t2n(y.features[::SUB]))
    print(name, "Np", pts.shape[0], "M", vc.shape[0])


def golden_augment(name):
    """SURVEY 8f rank 3 (input side): the reference's DataAugmentor (random_world_flip / rotation / scaling of
    tools/cfgs/waymo_models/gd_mae_ssl.yaml:18-31) applied frame by frame the way a dataset worker does, then
    data_processor.shuffle_points (data_processor.py:92-102), with numpy seeded; stores inputs, drawn parameters, outputs."""
    m = RH.ref_data_augmentor()
    acfg = RH.load_model_cfg("tools/cfgs/waymo_models/gd_mae_ssl.yaml").DATA_CONFIG.DATA_AUGMENTOR
    aug = m.DataAugmentor(None, acfg, ['Vehicle', 'Pedestrian', 'Cyclist'], logger=None)
    r = np.random.RandomState(21)
    out = {}
    np.random.seed(1234)
    for f in range(3):
        n = 1500 + 200 * f
        pts = np.concatenate([r.uniform(-70, 70, (n, 2)), r.uniform(-2, 4, (n, 1)), r.uniform(0, 1, (n, 2))], 1).astype(np.float32)
        out[f"f{f}.points_in"] = pts.copy()
        d = aug.forward({"points": pts.copy()})
        p3 = d["transformation_3d_params"]
        out[f"f{f}.flip_x"] = np.int64("x" in p3["random_world_flip"])
        out[f"f{f}.flip_y"] = np.int64("y" in p3["random_world_flip"])
        out[f"f{f}.rotation"] = np.float64(p3["random_world_rotation"])
        out[f"f{f}.scaling"] = np.float64(p3["random_world_scaling"])
        out[f"f{f}.points_aug"] = np.asarray(d["points"], dtype=np.float32)
        perm = np.random.permutation(n)                      # data_processor.py:98
        out[f"f{f}.perm"] = perm.astype(np.int64)
        out[f"f{f}.points_out"] = out[f"f{f}.points_aug"][perm]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "written", len(out), "arrays")


def golden_optimizer(name):
    """adam_onecycle of the path: the reference's build_optimizer / build_scheduler (tools/train_utils/optimization/
    __init__.py:11-66, fastai_optim.py:104-152 OptimWrapper, learning_schedules_fastai.py:44-77 OneCycle), imported
    unmodified, driven the way train_one_epoch does (train_utils.py:34-53: scheduler.step(it), zero_grad, backward,
    clip_grad_norm_(all parameters, 10), optimizer.step()) for 5 iterations on the tiny MAE model with seeded synthetic
    gradients.  Pins which parameters are updated (leaf modules only: in_proj_* / tau never move), the two weight-decay
    groups (bn_wd=True: BatchNorm too), the lr / momentum schedule and the Adam arithmetic."""
    sys.path.insert(0, RH.REF + "/tools")
    import train_utils.optimization as RO  # noqa: E402  (tools/train_utils has no __init__: namespace package)
    from torch.nn.utils import clip_grad_norm_
    cfg = O.make_cfg("tiny")
    model, P, Bf = build_ref(cfg, "tools/cfgs/waymo_models/gd_mae_ssl.yaml", 4)
    ocfg = RH.load_model_cfg("tools/cfgs/waymo_models/gd_mae_ssl.yaml").OPTIMIZATION
    total_steps = 12
    opt = RO.build_optimizer(model, ocfg)
    sched, _ = RO.build_scheduler(opt, total_iters_each_epoch=total_steps, total_epochs=1, last_epoch=-1, optim_cfg=ocfg)
    names = [k for k, _ in model.named_parameters()]
    watch = ["vfe.dvfe_mlps.0.0.weight", "vfe.dvfe_mlps.0.1.weight", "vfe.dvfe_mlps.0.1.bias",
             "backbone_3d.sst_blocks.0.encoder_blocks.0.encoder_list.0.win_attn.self_attn.in_proj_weight",
             "backbone_3d.sst_blocks.0.encoder_blocks.0.encoder_list.0.win_attn.self_attn.tau",
             "backbone_3d.sst_blocks.0.encoder_blocks.0.encoder_list.0.win_attn.self_attn.out_proj.weight",
             "backbone_3d.sst_blocks.1.encoder_blocks.1.encoder_list.1.norm2.weight",
             "backbone_3d.sst_blocks.1.conv_down.0.weight", "backbone_3d.sst_blocks.2.conv_out.1.bias",
             "backbone_3d.decoder_deblocks.2.0.weight", "backbone_3d.decoder_conv_out.1.weight", "backbone_3d.decoder_pred.bias"]
    out = {"param_seed": np.int64(4), "total_steps": np.int64(total_steps), "n_iters": np.int64(5), "watch": np.array(watch),
           "clip": np.float64(ocfg.GRAD_NORM_CLIP)}
    lrs, moms, norms = [], [], []
    params = dict(model.named_parameters())
    for it in range(5):
        sched.step(it)
        lrs.append(float(opt.lr))
        moms.append(float(opt.mom))
        opt.zero_grad()
        g = torch.Generator().manual_seed(900 + it)
        scale = 30.0 if it == 1 else 1.0          # iteration 1 exceeds the clip norm, the others do not
        for k in names:
            params[k].grad = torch.randn(params[k].shape, generator=g) * (0.002 * scale)
        norms.append(float(clip_grad_norm_(model.parameters(), ocfg.GRAD_NORM_CLIP)))
        opt.step()
        for k in watch:
            flat = params[k].detach().reshape(-1)
            out[f"it{it}.{k}"] = t2n(flat[::max(1, flat.numel() // 1500)]).copy()   # every stride-th element (<= ~1500 per tensor)
    out["lr"], out["mom"], out["total_norm"] = np.array(lrs), np.array(moms), np.array(norms)
    # the whole parameter vector after 5 iterations, as per-tensor sums / L2 norms (compact, covers every tensor)
    out["final_keys"] = np.array(names)
    out["final_sum"] = np.array([float(params[k].double().sum()) for k in names])
    out["final_norm"] = np.array([float(params[k].double().norm()) for k in names])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "lr", lrs, "mom", moms, "norms", norms)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "optimizer":
        golden_optimizer("optimizer_kat")
        sys.exit(0)
    cfg = O.make_cfg("tiny")
    pts = tiny_points(7, cfg, 1500, 2)
    model, bd = golden_mae("mae_tiny_b2", cfg, "tools/cfgs/waymo_models/gd_mae_ssl.yaml", pts, 2, seed=1)
    cfg2 = O.make_cfg("tiny")
    cfg2["mask_ratio"] = 0.3  # dense visible set -> drop levels 1 and 2 occur inside the full step
    r = np.random.RandomState(3)
    n = 5000
    dense = np.concatenate([r.randint(0, 2, (n, 1)).astype(np.float32), r.uniform(-6.4, 6.4, (n, 1)),
                            r.uniform(-7.68, 7.68, (n, 1)), r.uniform(-2, 4, (n, 1)), r.uniform(0, 1, (n, 2))], 1)
    dense = dense[np.argsort(dense[:, 0], kind="stable")].astype(np.float32)
    golden_mae("mae_tiny_dense", cfg2, "tools/cfgs/waymo_models/gd_mae_ssl.yaml", dense, 2, seed=2)
    golden_window("window_kat", model, bd)
    golden_kitti_c1("kitti_c1")
    golden_augment("augment_kat")
    golden_optimizer("optimizer_kat")
