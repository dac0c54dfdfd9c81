"""CPU: the oracle restatement (oracle/gdmae_oracle.py) against fixtures produced by the
UNMODIFIED reference Python (tests/golden/make_golden.py).  Pins rows a1-a27 of SURVEY.md 8a."""
import numpy as np
import pytest
import torch

from oracle import gdmae_oracle as O

SUB = 8


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)


@pytest.mark.parametrize("name,mask_ratio,seed", [("mae_tiny_b2", 0.85, 1), ("mae_tiny_dense", 0.3, 2)])
def test_full_step_matches_reference(golden, name, mask_ratio, seed):
    G = golden(name)
    cfg = O.make_cfg("tiny")
    cfg["mask_ratio"] = mask_ratio
    P, Bf = O.init_params(cfg, seed)
    leaves = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    pts = torch.from_numpy(G["points_in"])
    B = int(G["batch_size"])
    loss, T = O.mae_forward(leaves, pts, B, cfg, mask=torch.from_numpy(G["voxel_mae_mask"]), stats=Bf)
    # integer outputs: bit exact
    assert T["points"].shape[0] == int(G["n_points_kept"])
    for k in ["point_coords", "point_inverse_indices", "voxel_coords"]:
        assert np.array_equal(T[k].numpy(), G[k]), k
    for i in range(3):
        assert np.array_equal(T[f"x_conv{i+1}.indices"].numpy(), G[f"x_conv{i+1}.indices"])
    # features / loss: 1e-3 relative (north_star); in practice ~1e-6
    tol = 1e-4
    assert rel(T["pillar_features"][::SUB].detach(), G["pillar_features.sub"]) < tol
    for i in range(3):
        assert rel(T[f"x_conv{i+1}.features"][::SUB].detach(), G[f"x_conv{i+1}.features.sub"]) < tol
    assert rel(T["voxel_features"][::SUB].detach(), G["voxel_features.sub"]) < tol
    assert rel(T["spatial_features"][:, ::16, ::5, ::5].detach(), G["spatial_features.sub"]) < tol
    assert rel(T["pred_points"][::SUB].detach(), G["pred_points.sub"]) < tol
    assert np.array_equal(T["gt_points"][::SUB].numpy(), G["gt_points.sub"])
    assert abs(float(loss) - float(G["loss"])) / abs(float(G["loss"])) < 1e-5
    loss.backward()
    keys = [str(k) for k in G["grad_keys"]]
    assert keys == list(P.keys()), "state_dict key order/schema"
    for k, gn in zip(keys, G["grad_norms"]):
        mine = float(leaves[k].grad.norm()) if leaves[k].grad is not None else 0.0
        # tau's gradient is a tiny sum of cancelling terms (ill-conditioned in fp32): looser bound
        rtol = 5e-2 if k.endswith(".tau") else 2e-3
        assert abs(mine - gn) <= rtol * max(gn, 1e-6) + 1e-7, (k, mine, gn)
    for k in G.files:
        if k.startswith("grad."):
            assert rel(leaves[k[5:]].grad, G[k]) < (5e-2 if k.endswith(".tau") else 2e-3), k
        if k.startswith("buf."):
            assert rel(Bf[k[4:]], G[k]) < 1e-4, k


def test_window_bookkeeping_kat(golden):
    G = golden("window_kat")
    for bi, d in ((0, 128), (1, 256)):
        coords = torch.from_numpy(G[f"b{bi}.coords"])
        grid = [int(v) for v in G[f"b{bi}.grid"]]
        info = O.window_info(coords, grid, (8, 8, 1), d, 1000.0)
        levels_seen = set()
        for s in range(2):
            assert np.array_equal(info[s]["win"].numpy(), G[f"b{bi}.s{s}.batch_win_inds"])
            assert np.array_equal(info[s]["ciw"].numpy(), G[f"b{bi}.s{s}.coors_in_win"])
            assert np.array_equal(info[s]["lvl"].numpy(), G[f"b{bi}.s{s}.drop_level"])
            f2w = info[s]["flat2win"]
            pos3d = O.flat2window(info[s]["pos"], f2w)
            ones = O.flat2window(torch.ones((coords.shape[0], 1), dtype=torch.bool), f2w)
            for dl in (0, 1, 2):
                key = f"b{bi}.s{s}.l{dl}.flat2win"
                assert (key in G.files) == (dl in f2w)
                if dl not in f2w:
                    continue
                levels_seen.add(dl)
                assert np.array_equal(f2w[dl][0].numpy(), G[key])
                assert np.array_equal(f2w[dl][1].numpy(), G[f"b{bi}.s{s}.l{dl}.where"])
                assert np.array_equal(ones[dl].logical_not().squeeze(2).numpy(), G[f"b{bi}.s{s}.l{dl}.key_mask"])
                assert list(pos3d[dl].shape) == list(G[f"b{bi}.s{s}.l{dl}.pos_shape"])
                assert np.array_equal(pos3d[dl][:4].numpy(), G[f"b{bi}.s{s}.l{dl}.pos_sub"])
        assert levels_seen == {0, 1, 2}
        # one encoder layer on shift-1 windows (a17-a19)
        cfg = O.make_cfg("tiny")
        P, _ = O.init_params(cfg, 1)
        g = torch.Generator().manual_seed(int(G[f"b{bi}.layer_in_seed"]))
        x = torch.randn(coords.shape[0], d, generator=g)
        pre = f"backbone_3d.sst_blocks.{bi}.encoder_blocks.0.encoder_list.1."
        a = O.window_attention(P, pre + "win_attn.self_attn.", x, info[1], 8, 0.01)
        y = O.encoder_layer(P, pre, x, info[1], 8, 0.01)
        assert rel(a[::4], G[f"b{bi}.attn_out.sub"]) < 1e-5
        assert rel(y[::4], G[f"b{bi}.layer_out.sub"]) < 1e-5


def test_sst_ops_and_mask_kat(golden):
    G = golden("window_kat")
    assert np.array_equal(O.get_inner_win_inds(torch.from_numpy(G["ops.group_inds"])).numpy(), G["ops.inner"])
    inv = torch.from_numpy(G["ops.inverse"])
    gi = O.group_inner_inds(inv, int(inv.max()) + 1, 64)
    assert np.array_equal(torch.from_numpy(G["ops.points"])[gi].numpy(), G["ops.grouped"])
    m = O.random_masking_from_noise(torch.from_numpy(G["mask.noise"]), 0.85)
    assert np.array_equal(m.numpy(), G["mask.mask"])
    for d in (128, 256):
        assert np.array_equal(O.pos_embed_table(d, 1000.0).numpy(), G[f"pos_table.{d}"])


def test_kitti_c1_plumbing(golden):
    """BASELINE config 0: KITTI gd_mae.yaml, batch=1, DynVFE + one SRA block forward on CPU."""
    G = golden("kitti_c1")
    cfg = O.make_cfg("kitti")
    assert cfg["grid"] == [216, 248, 1]
    P, _ = O.init_params(cfg, int(G["param_seed"]))
    pts = torch.from_numpy(G["points_in"])
    with torch.no_grad():
        keep, p, pc, vc, inv = O.voxelize(pts, cfg)
        assert np.array_equal(vc.numpy(), G["voxel_coords"]) and np.array_equal(inv.numpy(), G["inverse"])
        pf, _, _ = O.vfe_forward(P, p, pc, inv, vc.shape[0], cfg)
        assert rel(pf[::SUB], G["pillar_features_sub"]) < 1e-5
        y, _, _, _ = O.sst_block(P, "backbone_3d.sst_blocks.0.", pf, vc[:, [0, 2, 3]], 1, 248, 216, cfg["blocks"][0], cfg)
        assert rel(y[::SUB], G["block_out_sub"]) < 1e-4


def test_onecycle_schedule_matches_survey_probe():
    """SURVEY.md section 5 [probe]: lr/mom at 3000 total steps."""
    cfg = O.make_cfg("waymo_ssl")
    for step, lr, mom in [(0, 3.0e-4, 0.95), (600, 1.65e-3, 0.90), (1200, 3.0e-3, 0.85)]:
        l, m = O.onecycle(step, 3000, cfg)
        assert abs(l - lr) / lr < 1e-2 and abs(m - mom) < 1e-3
    l, m = O.onecycle(2999, 3000, cfg)
    assert l < 1e-7 * 5 and abs(m - 0.95) < 1e-3


def test_param_count_matches_survey():
    cfg = O.make_cfg("waymo_ssl")
    P, Bf = O.init_params(cfg, 0)
    assert sum(v.numel() for v in P.values()) == 8091516
    assert sum(v.numel() for k, v in P.items() if O.in_optimizer(k)) == 6314352
    assert len(P) + len(Bf) == 224  # Appendix A: 224 state_dict entries (global_step is added by the detector shell)


def test_world_augmentation_kat(golden):
    """the oracle's world augmentation + shuffle against the reference's DataAugmentor / shuffle_points run
    (tests/golden/make_golden.py::golden_augment): same numpy stream -> same parameters, bit-identical points."""
    K = golden("augment_kat")
    np.random.seed(1234)
    for f in range(3):
        pts = K[f"f{f}.points_in"]
        prm = O.draw_world_aug_params(n_points=pts.shape[0])
        assert prm["flip_x"] == bool(K[f"f{f}.flip_x"]) and prm["flip_y"] == bool(K[f"f{f}.flip_y"])
        assert prm["rotation"] == float(K[f"f{f}.rotation"]) and prm["scaling"] == float(K[f"f{f}.scaling"])
        assert np.array_equal(prm["perm"], K[f"f{f}.perm"])
        out = O.world_augment(pts, prm["flip_x"], prm["flip_y"], prm["rotation"], prm["scaling"], prm["perm"])
        assert np.array_equal(out, K[f"f{f}.points_out"])


def _optimizer_kat_grads(P, it):
    """the seeded synthetic gradients of tests/golden/make_golden.py::golden_optimizer (named_parameters order = P's order)"""
    g = torch.Generator().manual_seed(900 + it)
    scale = 30.0 if it == 1 else 1.0
    return {k: torch.randn(v.shape, generator=g) * (0.002 * scale) for k, v in P.items()}


def test_adam_onecycle_matches_reference_optimizer(golden):
    """O.AdamOneCycle / O.onecycle against the reference's OptimWrapper + OneCycle (fastai_optim.py:104-152,
    learning_schedules_fastai.py:44-77) run for 5 iterations by make_golden.py: schedule values, clip norm, watched tensors
    after every iteration (1e-6 relative), never-optimised in_proj_* / tau bit-identical, every tensor's sum / norm at the end."""
    K = golden("optimizer_kat")
    cfg = O.make_cfg("tiny")
    P, _ = O.init_params(cfg, int(K["param_seed"]))
    assert list(P.keys()) == [str(k) for k in K["final_keys"]]
    P0 = {k: v.clone() for k, v in P.items()}
    opt = O.AdamOneCycle(P, cfg, int(K["total_steps"]))
    watch = [str(k) for k in K["watch"]]
    for it in range(int(K["n_iters"])):
        G = _optimizer_kat_grads(P, it)
        norm, lr, mom = opt.step(P, G, it)
        assert abs(lr - float(K["lr"][it])) <= 1e-12 + 1e-9 * lr and abs(mom - float(K["mom"][it])) <= 1e-12
        assert abs(norm - float(K["total_norm"][it])) / float(K["total_norm"][it]) < 1e-5
        for k in watch:
            flat = P[k].reshape(-1)
            mine = flat[::max(1, flat.numel() // 1500)]
            if O.in_optimizer(k):
                assert rel(mine, K[f"it{it}.{k}"]) < 1e-6, (it, k)
            else:
                assert np.array_equal(mine.numpy(), K[f"it{it}.{k}"]), (it, k)
    for k, s, n in zip(P.keys(), K["final_sum"], K["final_norm"]):
        assert abs(float(P[k].double().norm()) - n) <= 1e-6 * max(n, 1e-6), k
        assert abs(float(P[k].double().sum()) - s) <= 1e-5 * max(n, 1e-6), k
        if not O.in_optimizer(k):
            assert torch.equal(P[k], P0[k]), k


# ------------------------------------------------------------------------------ iou3d_nms oracle
def test_iou3d_oracle_matches_reference_cpu_golden(golden):
    """oracle/iou3d_oracle.c against the known answers written by the REFERENCE's own iou3d_cpu.cpp (compiled from
    /root/reference by oracle/build_oracle.py, tests/golden/make_golden_iou3d.py): bit for bit, including the cases where
    the reference's 1e-2 corner margin pushes the IoU of identical boxes above 1."""
    from oracle import iou3d_oracle as IO
    K = golden("iou3d_kat")
    for name in ("special", "rand_a", "rand_self"):
        iou = IO.boxes_iou_bev(K[name + ".a"], K[name + ".b"])
        assert np.array_equal(iou, K[name + ".iou"]), (name, np.abs(iou - K[name + ".iou"]).max())
    assert float(K["special.iou"].max()) > 1.0          # the margin quirk is part of the pinned behaviour


def test_iou3d_oracle_matches_compiled_reference_when_present():
    """where oracle/_ref exists (this container; it travels to the GPU box with gpurun) the oracle is also held to the
    compiled reference on fresh seeds"""
    from oracle import build_oracle, iou3d_oracle as IO
    ref = build_oracle.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref was not built (no /root/reference here)")
    a, b = IO.random_boxes(90, 11, spread=6.0), IO.random_boxes(70, 12, spread=6.0)
    ans = torch.zeros(a.shape[0], b.shape[0])
    ref.boxes_iou_bev_cpu(torch.from_numpy(a), torch.from_numpy(b), ans)
    assert np.array_equal(IO.boxes_iou_bev(a, b), ans.numpy())


def test_iou3d_oracle_nms_properties():
    from oracle import iou3d_oracle as IO
    boxes = IO.random_boxes(300, 5, spread=10.0)
    scores = np.random.RandomState(5).uniform(0, 1, 300).astype(np.float32)
    for fn in (IO.nms, IO.nms_normal):
        keep = fn(boxes, scores, 0.3)
        assert len(set(keep.tolist())) == len(keep) and 0 < len(keep) < 300
        assert np.all(np.diff(scores[keep]) <= 0)                              # kept in descending score order
    keep = IO.nms(boxes, scores, 0.3)
    iou = IO.boxes_iou_bev(boxes[keep], boxes[keep])
    assert (np.triu(iou, 1) <= 0.3).all()                                        # survivors do not suppress each other
    gone = np.setdiff1d(np.arange(300), keep)
    best = IO.boxes_iou_bev(boxes[gone], boxes[keep])
    assert ((best > 0.3) & (scores[keep][None, :] >= scores[gone][:, None])).any(1).all()   # every removed box has a better survivor
    top = np.argsort(-scores, kind="stable")[:50]                                # pre_maxsize = NMS over the 50 best only
    assert np.array_equal(IO.nms(boxes, scores, 0.3, pre_maxsize=50), top[IO.nms(boxes[top], scores[top], 0.3)])
    i3 = IO.boxes_iou3d(boxes[:40], boxes[:40])
    assert np.allclose(np.diag(i3), 1.0, atol=2e-2) and (i3 >= 0).all()


def test_iou3d_host_entry_point_matches_reference_golden(golden):
    """gdmae_boxes_iou_bev_cpu (HOST pointers, the counterpart of the reference's boxes_iou_bev_cpu) through the Python
    mirror pcdet.ops.iou3d_nms.iou3d_nms_utils.boxes_bev_iou_cpu: same geometry code as the kernels, compiled for the host."""
    import gd_mae_b200  # noqa: F401
    from gd_mae_b200.pcdet.ops.iou3d_nms import iou3d_nms_utils as U
    K = golden("iou3d_kat")
    for name in ("special", "rand_a", "rand_self"):
        iou = U.boxes_bev_iou_cpu(K[name + ".a"], K[name + ".b"])
        assert np.abs(iou - K[name + ".iou"]).max() <= 1e-6, name
    with pytest.raises(Exception):
        U.boxes_iou_bev(torch.zeros(2, 7), torch.zeros(2, 7))          # CPU tensors on the GPU entry point: no fallback


# ------------------------------------------------------------------------------ CenterHead oracle
def test_center_head_oracle_matches_reference_golden(golden):
    """oracle/center_oracle.py against the targets the unmodified reference CenterHead assigned in
    tests/golden/finetune_tiny.npz: heat map, indices and masks bit-exact, regression targets to float rounding."""
    from oracle import center_oracle as CO
    K = golden("finetune_tiny")
    cfg = O.make_cfg("tiny")
    X, Y, _ = cfg["grid"]
    heat, tgt, iou_boxes, inds, mask = CO.assign_targets(torch.from_numpy(K["gt_boxes"]), [0, 1, 2, 3], 3, Y, X, cfg["pc_range"], cfg["voxel"])
    ref_heat = np.zeros(int(np.prod(K["heatmap.shape"])), dtype=np.float32)
    ref_heat[K["heatmap.nz_index"]] = K["heatmap.nz_value"]
    assert np.array_equal(heat.numpy().reshape(-1), ref_heat)
    assert np.array_equal(inds.numpy(), K["inds"]) and np.array_equal(mask.numpy(), K["masks"])
    assert int(mask.sum()) == 16 and np.array_equal(iou_boxes.numpy(), K["iou_boxes"])
    assert np.abs(tgt.numpy() - K["target_boxes"]).max() <= 1e-6
