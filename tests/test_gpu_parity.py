"""GPU parity tests: the CUDA path (through the C ABI, ctypes) against the CPU oracle and the
golden fixtures written by the unmodified reference.  Bit-exact for index outputs; features/loss
within the tolerances written next to each assert (north_star: 1e-3 relative fp32)."""
import numpy as np
import pytest
import torch

from oracle import gdmae_oracle as O

pytestmark = pytest.mark.gpu
SUB = 8


def rel(a, b):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)


@pytest.fixture(scope="module")
def G():
    import gd_mae_b200
    from gd_mae_b200 import ops, config
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    assert gd_mae_b200._lib.lib().gdmae_check_device() == 0
    return type("G", (), dict(ops=ops, config=config, pkg=gd_mae_b200))


def build(G, name, mask_ratio, seed):
    cfg = G.config.builtin_cfg(name)
    cfg.MODEL.BACKBONE_3D.MASK_CONFIG.RATIO = mask_ratio
    model = G.config.build_mae_model(cfg).cuda()
    ocfg = O.make_cfg(name)
    ocfg["mask_ratio"] = mask_ratio
    P, Bf = O.init_params(ocfg, seed)
    sd = dict(P)
    sd.update(Bf)
    missing = model.load_state_dict(sd, strict=False)
    assert missing.missing_keys == ["global_step"] and not missing.unexpected_keys
    return model, cfg, ocfg, P, Bf


# ------------------------------------------------------------------------------ a1-a6 VFE
@pytest.mark.parametrize("case", ["tiny", "waymo"])
def test_voxelize_mean_features_max(G, golden, case):
    if case == "tiny":
        cfg = O.make_cfg("tiny")
        pts = torch.from_numpy(golden("mae_tiny_b2")["points_in"])
        B = 2
    else:
        cfg = O.make_cfg("waymo_ssl")
        pts = torch.from_numpy(O.synth_batch([0, 1], cfg, n=60000))
        B = 2
    keep, opts, ocoords, ovc, oinv = O.voxelize(pts, cfg)
    ps = G.ops.dynamic_voxelize(pts.cuda(), cfg["pc_range"], cfg["voxel"], cfg["grid"], B)
    assert ps.n_points == opts.shape[0] and ps.n_pillars == ovc.shape[0]
    assert torch.equal(ps.points.cpu(), opts)
    assert torch.equal(ps.point_coords.cpu(), ocoords)
    assert torch.equal(ps.voxel_coords.cpu(), ovc)
    assert torch.equal(ps.inverse.cpu(), oinv)
    # CSR: pillar m owns exactly the points with inverse == m, in ascending index order
    off, sp = ps.seg_offsets.cpu().long(), ps.seg_points.cpu().long()
    assert torch.equal(oinv[sp], torch.repeat_interleave(torch.arange(ps.n_pillars), off[1:] - off[:-1]))
    same = oinv[sp][1:] == oinv[sp][:-1]
    assert bool((sp[1:][same] > sp[:-1][same]).all())
    assert ps.batch_offsets == [int((ovc[:, 0] < b).sum()) for b in range(B + 1)]
    # a3 mean: bit exact (same summation order as the sequential CPU scatter)
    M = ps.n_pillars
    omean = O.scatter_mean(opts[:, 1:], oinv, M)
    mean = G.ops.segment_mean(ps.points, 1, opts.shape[1] - 1, ps.seg_offsets, ps.seg_points, M)
    assert torch.equal(mean.cpu(), omean)
    # a4 features: bit exact
    ox = O.vfe_point_features(opts, ocoords, omean, oinv, cfg)
    x = G.ops.vfe_point_features(ps, mean, cfg["pc_range"], cfg["voxel"])
    assert torch.equal(x.cpu(), ox)
    # a6 max + backward
    g = torch.Generator().manual_seed(0)
    h = torch.randn(opts.shape[0], 128, generator=g).relu()
    hc = h.cuda().requires_grad_(True)
    out = G.ops.SegmentMax.apply(hc, ps.seg_offsets, ps.seg_points, M)
    assert torch.equal(out.detach().cpu(), O.scatter_max(h, oinv, M))
    w = torch.randn(M, 128, generator=g)
    (out * w.cuda()).sum().backward()
    h2 = h.clone().requires_grad_(True)
    (O.scatter_max(h2, oinv, M) * w).sum().backward()
    nz = h > 0  # ties at 0 (ReLU) may route to a different equal element; positive maxima are unique
    assert torch.equal(hc.grad.cpu()[nz], h2.grad[nz])


def test_voxelize_edge_cases(G):
    cfg = O.make_cfg("tiny")
    dev = "cuda"
    empty = torch.zeros((0, 6), device=dev)
    ps = G.ops.dynamic_voxelize(empty, cfg["pc_range"], cfg["voxel"], cfg["grid"], 2)
    assert ps.n_points == 0 and ps.n_pillars == 0 and ps.batch_offsets == [0, 0, 0]
    out = torch.tensor([[0, 100.0, 0, 0, 0, 0], [1, 0, -100.0, 0, 0, 0]], device=dev)
    ps = G.ops.dynamic_voxelize(out, cfg["pc_range"], cfg["voxel"], cfg["grid"], 2)
    assert ps.n_points == 0 and ps.n_pillars == 0
    # z below the range truncates toward zero into bin 0 and is KEPT; x == max edge is dropped
    q = torch.tensor([[0, 0.0, 0.0, -5.0, 0, 0], [0, 6.4, 0.0, 0.0, 0, 0], [1, -6.4, -7.68, 0.0, 0, 0]], device=dev)
    ps = G.ops.dynamic_voxelize(q, cfg["pc_range"], cfg["voxel"], cfg["grid"], 2)
    _, opts, _, ovc, oinv = O.voxelize(q.cpu(), cfg)
    assert ps.n_points == opts.shape[0] == 2 and torch.equal(ps.voxel_coords.cpu(), ovc)
    with pytest.raises(Exception):
        G.ops.dynamic_voxelize(torch.tensor([[5, 0.0, 0, 0, 0, 0]], device=dev), cfg["pc_range"], cfg["voxel"], cfg["grid"], 2)
    with pytest.raises(Exception):
        G.ops.dynamic_voxelize(torch.zeros((4, 6)), cfg["pc_range"], cfg["voxel"], cfg["grid"], 2)  # CPU tensor: no fallback


# ------------------------------------------------------------------------------ a7, a11, a24
def test_mask_and_sst_ops(G, golden):
    K = golden("window_kat")
    from gd_mae_b200.pcdet.ops.sst_ops import sst_ops_utils
    from gd_mae_b200.pcdet.utils import common_utils
    noise = torch.from_numpy(K["mask.noise"]).cuda()
    m = common_utils.random_masking(1, noise.shape[0], 0.85, "cuda", noise=noise)[0]
    assert np.array_equal(m.cpu().numpy(), K["mask.mask"])
    # several frames + ties
    g = torch.Generator().manual_seed(3)
    noise = (torch.randint(0, 50, (1000,), generator=g).float() / 50)
    vc = torch.zeros(1000, 4, dtype=torch.long)
    vc[300:, 0] = 1
    vc[650:, 0] = 2
    off = torch.tensor([0, 300, 650, 1000], dtype=torch.int32).cuda()
    got = G.ops.random_mask(noise.cuda(), off, 3, 0.85)
    assert torch.equal(got.cpu(), O.mae_mask(vc, 3, noise, 0.85))
    grp = torch.from_numpy(K["ops.group_inds"]).cuda()
    assert np.array_equal(sst_ops_utils.get_inner_win_inds(grp).cpu().numpy(), K["ops.inner"])
    big = torch.randint(0, 2 ** 40, (5000,), generator=g)
    big[::3] = big[0]
    assert torch.equal(sst_ops_utils.get_inner_win_inds(big.cuda()).cpu(), O.get_inner_win_inds(big))
    inv = torch.from_numpy(K["ops.inverse"]).cuda()
    pts = torch.from_numpy(K["ops.points"]).cuda()
    assert np.array_equal(sst_ops_utils.group_inner_inds(pts, inv, 64).cpu().numpy(), K["ops.grouped"])


# ------------------------------------------------------------------------------ a10-a16 window tables
def test_window_tables_match_reference_kat(G, golden):
    K = golden("window_kat")
    from gd_mae_b200.pcdet.models.model_utils import sst_utils
    for bi in (0, 1):
        coords = torch.from_numpy(K[f"b{bi}.coords"])
        X, Y, _ = [int(v) for v in K[f"b{bi}.grid"]]
        idx = coords[:, [0, 2, 3]].int().cuda()
        B = int(coords[:, 0].max()) + 1
        for s in range(2):
            t = G.ops.window_table(idx, B, Y, X, s)
            win, ciw = sst_utils.get_window_coors(t)
            assert np.array_equal(win.cpu().numpy(), K[f"b{bi}.s{s}.batch_win_inds"])
            assert np.array_equal(ciw.cpu().numpy(), K[f"b{bi}.s{s}.coors_in_win"])
            f2w = sst_utils.get_flat2win_inds_v2(t)
            assert np.array_equal(f2w["voxel_drop_level"].cpu().numpy(), K[f"b{bi}.s{s}.drop_level"])
            for dl in (0, 1, 2):
                key = f"b{bi}.s{s}.l{dl}.flat2win"
                assert (key in K.files) == (dl in f2w)
                if dl in f2w:
                    assert np.array_equal(f2w[dl][0].cpu().numpy(), K[key])
                    assert np.array_equal(f2w[dl][1][0].cpu().numpy(), K[f"b{bi}.s{s}.l{dl}.where"])
                    ones = sst_utils.flat2window_v2(torch.ones((idx.shape[0], 1), dtype=torch.bool, device="cuda"), f2w)
                    assert np.array_equal(ones[dl].logical_not().squeeze(2).cpu().numpy(), K[f"b{bi}.s{s}.l{dl}.key_mask"])
            # round trip: window2flat(flat2window(x)) == x
            x = torch.randn(idx.shape[0], 8, device="cuda")
            assert torch.equal(sst_utils.window2flat_v2(sst_utils.flat2window_v2(x, f2w), f2w), x)
            # CSR consistency
            off, tok = t.win_off.cpu().long(), t.win_tok.cpu().long()
            assert int(off[-1]) == idx.shape[0] and sorted(tok.tolist()) == list(range(idx.shape[0]))
    from gd_mae_b200.pcdet.models.backbones_3d.spt_backbone import pos_embed_table
    for d in (128, 256):
        assert np.array_equal(pos_embed_table(d, 1000).numpy(), K[f"pos_table.{d}"])


# ------------------------------------------------------------------------------ a9/a21 sparse conv structure + features
def test_sparse_conv_structure_and_features(G):
    from gd_mae_b200.pcdet.utils.spconv_utils import spconv, plan_pyramid
    g = torch.Generator().manual_seed(1)
    B, H, W, C = 2, 47, 40, 32
    occ = torch.rand(B, H, W, generator=g) < 0.15
    idx = torch.nonzero(occ)
    feat = torch.randn(idx.shape[0], C, generator=g)
    sp = spconv.SparseConvTensor(feat.cuda().requires_grad_(True), idx.int().cuda(), [H, W], B)
    plan_pyramid(sp, 2)
    o1, Ho, Wo = O.down_sites(idx, B, H, W)
    d = sp.down()
    assert d.spatial_shape == [Ho, Wo] and torch.equal(d.indices.cpu().long(), o1)
    assert torch.equal(d.nbr_down.cpu().long(), O.down_neighbor_map(idx, o1, B, H, W))
    assert torch.equal(sp.subm_map().cpu().long(), O.subm_neighbor_map(idx, B, H, W))
    o2, Ho2, Wo2 = O.down_sites(o1, B, Ho, Wo)
    sp2 = spconv.SparseConvTensor(None, d.indices, d.spatial_shape, B, d.struct)
    assert torch.equal(sp2.down().indices.cpu().long(), o2)
    # features fwd/bwd of both conv types
    for subm in (True, False):
        conv = (spconv.SubMConv2d if subm else spconv.SparseConv2d)(C, 48, 3, stride=1 if subm else 2, padding=1).cuda()
        y = conv(sp)
        wgt = torch.randn(y.features.shape, generator=g)
        sp.features.grad = None
        (y.features * wgt.cuda()).sum().backward()
        f2 = feat.clone().requires_grad_(True)
        w2 = conv.weight.detach().cpu().clone().requires_grad_(True)
        nbr = O.subm_neighbor_map(idx, B, H, W) if subm else O.down_neighbor_map(idx, o1, B, H, W)
        yo = O.sparse_conv(f2, nbr, w2)
        (yo * wgt).sum().backward()
        assert rel(y.features, yo) < 1e-5
        assert rel(sp.features.grad, f2.grad) < 1e-5 and rel(conv.weight.grad, w2.grad) < 1e-5


# ------------------------------------------------------------------------------ a17-a19 SRA layer
@pytest.mark.parametrize("bi,d", [(0, 128), (1, 256)])
def test_sra_encoder_layer_matches_reference_kat(G, golden, bi, d):
    K = golden("window_kat")
    from gd_mae_b200.pcdet.models.model_utils.sst_basic_block import EncoderLayer
    from gd_mae_b200.pcdet.models.backbones_3d.spt_backbone import pos_embed_table
    coords = torch.from_numpy(K[f"b{bi}.coords"])
    X, Y, _ = [int(v) for v in K[f"b{bi}.grid"]]
    idx = coords[:, [0, 2, 3]].int().cuda()
    table = G.ops.window_table(idx, 2, Y, X, 1)
    cfg = O.make_cfg("tiny")
    P, _ = O.init_params(cfg, 1)
    pre = f"backbone_3d.sst_blocks.{bi}.encoder_blocks.0.encoder_list.1."
    layer = EncoderLayer(d, 8, 2 * d, 0.0, "gelu", layer_cfg={"cosine": True, "tau_min": 0.01}).cuda()
    layer.load_state_dict({k[len(pre):]: v for k, v in P.items() if k.startswith(pre)})
    g = torch.Generator().manual_seed(int(K[f"b{bi}.layer_in_seed"]))
    x = torch.randn(coords.shape[0], d, generator=g)
    pos = pos_embed_table(d, 1000).cuda()
    xc = x.cuda().requires_grad_(True)
    a = layer.win_attn(xc, pos, table)
    y = layer(xc, pos, table)
    assert rel(a[::4], K[f"b{bi}.attn_out.sub"]) < 1e-4   # fp32, 1e-3 is the contract
    assert rel(y[::4], K[f"b{bi}.layer_out.sub"]) < 1e-4
    # backward of the whole layer against the oracle's autograd
    info = O.window_info(coords, [X, Y, 1], (8, 8, 1), d, 1000.0)
    leaves = {k: v.clone().requires_grad_(True) for k, v in P.items() if k.startswith(pre)}
    xo = x.clone().requires_grad_(True)
    yo = O.encoder_layer(leaves, pre, xo, info[1], 8, 0.01)
    wgt = torch.randn(yo.shape, generator=g)
    (yo * wgt).sum().backward()
    (y * wgt.cuda()).sum().backward()
    assert rel(xc.grad, xo.grad) < 1e-4
    for k, p in layer.named_parameters():
        tol = 5e-2 if k.endswith("tau") else 1e-3
        assert rel(p.grad, leaves[pre + k].grad) < tol, k


# ------------------------------------------------------------------------------ a24-a26 chamfer
def test_chamfer_fwd_bwd(G):
    g = torch.Generator().manual_seed(2)
    N = 700
    x = torch.randn(N, 16, 3, generator=g)
    y = torch.randn(N, 64, 3, generator=g)
    y[5] = y[5, :7].repeat(10, 1)[:64]  # cyclic duplicates like group_inner_inds
    w = (torch.rand(N, generator=g) < 0.85).float()
    xo = x.clone().requires_grad_(True)
    lo = O.chamfer_distance(xo, y, w)
    lo.backward()
    xc = x.cuda().requires_grad_(True)
    lc, _ = G.ops.chamfer_distance(xc, y.cuda(), w.cuda())
    lc.backward()
    assert abs(float(lc) - float(lo)) / float(lo) < 1e-5
    assert rel(xc.grad, xo.grad) < 1e-5
    same, _ = G.ops.chamfer_distance(y[:, :16].cuda().contiguous(), y[:, :16].cuda().contiguous(), w.cuda())
    assert float(same) == 0.0  # chamfer(x, x) == 0
    zero, _ = G.ops.chamfer_distance(xc, y.cuda(), torch.zeros(N, device="cuda"))
    assert float(zero) == 0.0


# ------------------------------------------------------------------------------ full step vs the reference's golden run
@pytest.mark.parametrize("dense_map", [True, False])
@pytest.mark.parametrize("name,mask_ratio,seed", [("mae_tiny_b2", 0.85, 1), ("mae_tiny_dense", 0.3, 2)])
def test_full_mae_step_matches_reference(G, golden, name, mask_ratio, seed, dense_map):
    """dense_map=False: BN + ReLU of the decoder evaluated at the pillar cells only (ops.DecoderTail) - the
    loss, the head inputs and every gradient must still match the reference's golden step."""
    K = golden(name)
    model, cfg, ocfg, P, Bf = build(G, "tiny", mask_ratio, seed)
    model.backbone_3d.dense_spatial_features = dense_map
    model.train()
    B = int(K["batch_size"])
    bd = dict(points=torch.from_numpy(K["points_in"]).cuda(), batch_size=B,
              voxel_mae_mask=torch.from_numpy(K["voxel_mae_mask"]).cuda())
    ret, tb, _ = model(bd)
    loss = ret["loss"]
    loss.backward()
    assert bd["points"].shape[0] == int(K["n_points_kept"])
    for k in ["point_coords", "point_inverse_indices", "voxel_coords"]:
        assert np.array_equal(bd[k].cpu().numpy(), K[k]), k
    for i in range(3):
        sp = bd["multi_scale_3d_features"][f"x_conv{i + 1}"]
        assert np.array_equal(sp.indices.cpu().numpy(), K[f"x_conv{i + 1}.indices"])
        assert rel(sp.features[::SUB], K[f"x_conv{i + 1}.features.sub"]) < 1e-3
    assert rel(bd["pillar_features"][::SUB], K["pillar_features.sub"]) < 1e-4
    assert rel(bd["voxel_features"][::SUB], K["voxel_features.sub"]) < 1e-3
    sf = bd["spatial_features"]
    if dense_map:
        assert list(sf.shape) == list(K["spatial_features.shape"])
        assert rel(sf[:, ::16, ::5, ::5], K["spatial_features.sub"]) < 1e-3
    else:
        assert sf is None
    frd = model.backbone_3d.forward_ret_dict
    assert np.array_equal(frd["gt_points"][::SUB].cpu().numpy(), K["gt_points.sub"])
    assert rel(frd["pred_points"][::SUB], K["pred_points.sub"]) < 1e-3
    assert abs(float(loss) - float(K["loss"])) / float(K["loss"]) < 1e-4
    grads = {k: p.grad for k, p in model.named_parameters()}
    for k, gn in zip([str(s) for s in K["grad_keys"]], K["grad_norms"]):
        mine = float(grads[k].norm()) if grads[k] is not None else 0.0
        rtol = 5e-2 if k.endswith(".tau") else 5e-3
        assert abs(mine - gn) <= rtol * max(gn, 1e-6) + 1e-7, (k, mine, gn)
    for k in K.files:
        if k.startswith("grad."):
            assert rel(grads[k[5:]], K[k]) < (5e-2 if k.endswith(".tau") else 5e-3), k
        if k.startswith("buf."):
            assert rel(model.state_dict()[k[4:]], K[k]) < 1e-3, k


def test_mask_from_noise_inside_model_and_trainer_steps(G):
    """Three optimizer iterations: CUDA trainer (flat bucket + fused Adam) vs the oracle's step."""
    from gd_mae_b200.trainer import MAETrainer, optimised_parameter_names
    model, cfg, ocfg, P, Bf = build(G, "tiny", 0.85, 5)
    assert optimised_parameter_names(model) == {k for k in P if O.in_optimizer(k)}
    trainer = MAETrainer(model, cfg.OPTIMIZATION, total_steps=20)
    opt = O.AdamOneCycle(P, ocfg, 20)
    P0 = {k: v.clone() for k, v in P.items()}
    r = np.random.RandomState(0)
    for it in range(3):
        n = 1800
        pts = np.concatenate([r.randint(0, 2, (n, 1)), r.normal(0, 3, (n, 2)), r.uniform(-2, 4, (n, 1)), r.uniform(0, 1, (n, 2))], 1)
        pts = torch.from_numpy(pts[np.argsort(pts[:, 0], kind="stable")].astype(np.float32))
        _, _, _, ovc, _ = O.voxelize(pts, ocfg)
        noise = torch.rand(ovc.shape[0], generator=torch.Generator().manual_seed(it))
        lo, _, _ = O.train_step(P, Bf, opt, pts, 2, ocfg, noise, it)
        lc = trainer.step(dict(points=pts.cuda(), batch_size=2, voxel_mae_noise=noise.cuda()))
        assert abs(float(lc) - lo) / lo < 2e-3, (it, float(lc), lo)
    sd = model.state_dict()
    # Adam's first steps move every element by ~lr * sign(g): elements whose gradient is at noise
    # level may flip, so compare the update as a whole (relative L2 over all optimised parameters) ...
    num = sum(float(((sd[k].cpu().double() - P[k].double()) ** 2).sum()) for k in P)
    den = sum(float(((P[k].double() - P0[k].double()) ** 2).sum()) for k in P)
    assert (num / den) ** 0.5 < 5e-2, (num / den) ** 0.5
    for k in P:  # ... and the never-updated attention in-proj / tau exactly (optimizer quirk)
        if not O.in_optimizer(k):
            assert torch.equal(sd[k].cpu(), P0[k]), k


def test_deferred_gradient_handover_fills_the_bucket(G, monkeypatch):
    """MAETrainer's deferred hand-over (AccumulateGrad keeps the produced tensors, one multi-tensor copy moves them into the
    flat bucket) must leave the same gradient bucket as `.grad = bucket view` + one `grad += g` per parameter."""
    from gd_mae_b200.trainer import MAETrainer
    r = np.random.RandomState(4)
    n = 2200
    pts = np.concatenate([r.randint(0, 2, (n, 1)), r.normal(0, 3, (n, 2)), r.uniform(-2, 4, (n, 1)), r.uniform(0, 1, (n, 2))], 1)
    pts = torch.from_numpy(pts[np.argsort(pts[:, 0], kind="stable")].astype(np.float32)).cuda()
    buckets = []
    for defer in ("1", "0"):
        monkeypatch.setenv("GDMAE_DEFER_GRADS", defer)
        model, cfg, ocfg, P, Bf = build(G, "tiny", 0.85, 5)
        _, _, _, ovc, _ = O.voxelize(pts.cpu(), ocfg)
        noise = torch.rand(ovc.shape[0], generator=torch.Generator().manual_seed(1)).cuda()
        tr = MAETrainer(model, cfg.OPTIMIZATION, total_steps=20)
        assert (len(tr._deferred) > 0) == (defer == "1")
        tr.step(dict(points=pts.clone(), batch_size=2, voxel_mae_noise=noise))
        for p_, view in tr._deferred:
            assert p_.grad is not None and p_.grad.data_ptr() == view.data_ptr()      # .grad is the bucket view again
        buckets.append((tr.flat_grads.clone(), [(o, k) for o, k in tr.slices.values()]))
    (g1, sl), (g0, _) = buckets
    assert rel(g1, g0) < 1e-4, rel(g1, g0)                       # float atomics order only
    for off, k in sl:                                            # every tensor of the bucket received its gradient in both runs
        assert bool((g1[off:off + k] != 0).any()) == bool((g0[off:off + k] != 0).any())


def test_fused_adam_onecycle_kernel_matches_oracle(G):
    """The clip + decoupled-wd + Adam kernel in isolation: identical synthetic gradients on both sides."""
    from gd_mae_b200.trainer import MAETrainer
    model, cfg, ocfg, P, Bf = build(G, "tiny", 0.85, 6)
    trainer = MAETrainer(model, cfg.OPTIMIZATION, total_steps=50)
    opt = O.AdamOneCycle(P, ocfg, 50)
    params = dict(model.named_parameters())
    g = torch.Generator().manual_seed(0)
    for it in range(4):
        scale = [3.0, 0.01, 1.0, 30.0][it]  # exercises both sides of the clip threshold
        Gd = {k: torch.randn(v.shape, generator=g) * scale * 1e-3 for k, v in P.items()}
        for k in P:
            params[k].grad.copy_(Gd[k].cuda())
        norm_o, lr_o, mom_o = opt.step(P, Gd, it)
        lr_c, mom_c = trainer.optimizer_step()
        assert (lr_c, mom_c) == (lr_o, mom_o)
        assert abs(float(trainer.sumsq.sqrt()) - norm_o) / norm_o < 1e-5
    sd = model.state_dict()
    for k in P:
        assert rel(sd[k], P[k]) < 2e-5, k


# ------------------------------------------------------------------------------ full-size properties (Waymo shape)
def test_waymo_shape_properties(G):
    cfg = O.make_cfg("waymo_ssl")
    pts = torch.from_numpy(O.synth_batch([11, 12], cfg)).cuda()
    ps = G.ops.dynamic_voxelize(pts, cfg["pc_range"], cfg["voxel"], cfg["grid"], 2)
    vc = ps.voxel_coords
    key = ((vc[:, 0] * 1 + vc[:, 1]) * 468 + vc[:, 2]) * 468 + vc[:, 3]
    assert bool((key[1:] > key[:-1]).all())                       # unique + lexicographically sorted
    assert torch.equal(vc[ps.inverse], ps.point_coords)           # voxel_coords[inverse] == point_coords
    assert int((ps.seg_offsets[1:] - ps.seg_offsets[:-1]).min()) >= 1
    model = G.config.build_mae_model(G.config.builtin_cfg("waymo_ssl")).cuda().train()
    bd = dict(points=pts, batch_size=2)
    ret, _, _ = model(bd)
    ret["loss"].backward()
    assert torch.isfinite(ret["loss"]) and 0.5 < float(ret["loss"]) < 50
    mask = bd["voxel_mae_mask"]
    for b in range(2):
        L = ps.batch_offsets[b + 1] - ps.batch_offsets[b]
        assert int((mask[ps.batch_offsets[b]:ps.batch_offsets[b + 1]] == 0).sum()) == int(L * (1 - 0.85))
    x1 = bd["multi_scale_3d_features"]["x_conv1"]
    for t in x1.window_tables():                                   # nothing is dropped; every window has <= 64 tokens
        cnt = t.win_off[1:] - t.win_off[:-1]
        assert int(cnt.max()) <= 64 and int(cnt.sum()) == x1.indices.shape[0]
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


@pytest.mark.parametrize("d", [128, 256])
def test_sra_tensor_core_kernel_matches_simt(G, d):
    """bf16 tensor-core SRA forward (bf16 q/k/v) vs the fp32 SIMT kernel on the same rounded inputs;
    windows of every size 1..64 and all drop levels, both shifts, ragged last bin."""
    g = torch.Generator().manual_seed(d)
    occ = torch.rand(3, 70, 61, generator=g) < 0.55
    occ[2, :, 30:] &= torch.rand(70, 31, generator=g) < 0.1
    occ[1, 40:, :] &= torch.rand(30, 61, generator=g) < 0.04        # tiny windows (1-3 tokens)
    idx = torch.nonzero(occ).int().contiguous().cuda()  # nonzero() returns a column-major (N,3) tensor
    N = idx.shape[0]
    qkv_b = torch.randn(N, 3 * d, generator=g).cuda().to(torch.bfloat16)
    qkv = qkv_b.float()
    lut = (0.5 * torch.randn(64, 2 * d, generator=g)).cuda()
    tau = torch.tensor([0.7]).cuda()
    bv = torch.randn(d, generator=g).cuda()
    for shift in (0, 1):
        table = G.ops.window_table(idx, 3, 70, 61, shift)
        o_ref, lse_ref = G.ops.sra_fwd(qkv, lut, tau, table, 0.01, 8)
        o_tc, lse_tc = G.ops.sra_fwd(qkv_b, lut, tau, table, 0.01, 8)
        assert torch.isfinite(o_tc).all() and torch.isfinite(lse_tc).all()
        e_o, e_l = rel(o_tc, o_ref), rel(lse_tc, lse_ref)
        print(f"sra tc d={d} shift={shift}: rel err out {e_o:.2e} lse {e_l:.2e}")
        assert e_o < 1e-2, e_o      # bf16 operands (8-bit mantissa): q, k, P and the LUT are rounded
        assert e_l < 5e-3, e_l
        # value bias folded into the output, bf16 output (the bench configuration's call)
        o_ref, _ = G.ops.sra_fwd(qkv, lut, tau, table, 0.01, 8, bv=bv, out_dtype=torch.bfloat16)
        o_tc, _ = G.ops.sra_fwd(qkv_b, lut, tau, table, 0.01, 8, bv=bv, out_dtype=torch.bfloat16)
        assert rel(o_tc.float(), o_ref.float()) < 1.5e-2
        # backward: tensor-core kernel (bf16 qkv / dO / dqkv) vs the fp32 kernel on the same rounded inputs
        do_b = torch.randn(N, d, generator=g).cuda().to(torch.bfloat16)
        o32, lse32 = G.ops.sra_fwd(qkv, lut, tau, table, 0.01, 8, bv=bv)
        dq_ref, dt_ref = G.ops.sra_bwd(qkv, lut, tau, table, 0.01, 8, o32, lse32, do_b.float(), bv=bv)
        dq_tc, dt_tc = G.ops.sra_bwd(qkv_b, lut, tau, table, 0.01, 8, None, lse32, do_b)
        assert torch.isfinite(dq_tc.float()).all()
        for nm, c0, c1 in (("dq", 0, d), ("dk", d, 2 * d), ("dv", 2 * d, 3 * d)):
            e = rel(dq_tc[:, c0:c1].float(), dq_ref[:, c0:c1])
            print(f"sra tc bwd d={d} shift={shift}: {nm} rel err {e:.2e}")
            assert e < 2e-2, (nm, e)
        e_t = abs(float(dt_tc) - float(dt_ref)) / max(abs(float(dt_ref)), 1e-6)
        print(f"   dtau_sum {float(dt_tc):.4f} vs {float(dt_ref):.4f}")
        assert e_t < 3e-2, e_t


def test_bf16_configuration_close_to_reference(G, golden):
    """The bench configuration (bf16 GEMM operands + bf16 decoder map, fp32 accumulation / statistics)
    against the reference's fp32 golden step: looser, stated tolerances (SURVEY.md section 7:
    bf16 inputs alone move decoder features by ~1.6e-2 relative)."""
    K = golden("mae_tiny_dense")
    model, cfg, ocfg, P, Bf = build(G, "tiny", 0.3, 2)
    from gd_mae_b200 import fused
    try:
        G.config.set_precision(model, "bf16", gemm_bf16=True, dense_spatial_features=False)
        model.train()
        bd = dict(points=torch.from_numpy(K["points_in"]).cuda(), batch_size=int(K["batch_size"]),
                  voxel_mae_mask=torch.from_numpy(K["voxel_mae_mask"]).cuda())
        ret, _, _ = model(bd)
        ret["loss"].backward()
        assert np.array_equal(bd["voxel_coords"].cpu().numpy(), K["voxel_coords"])          # indices stay bit exact
        assert abs(float(ret["loss"]) - float(K["loss"])) / float(K["loss"]) < 2e-2
        assert rel(bd["voxel_features"][::SUB], K["voxel_features.sub"]) < 6e-2
        grads = {k: p.grad for k, p in model.named_parameters()}
        tot_ref = float(np.sqrt((K["grad_norms"] ** 2).sum()))
        tot = float(torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())))
        assert abs(tot - tot_ref) / tot_ref < 5e-2
        # the same configuration through the trainer (flat bucket, bf16 parameter mirror, in-place gradients):
        # every parameter the reference's step gives a gradient must get one of the same size
        from gd_mae_b200.trainer import MAETrainer
        tr = MAETrainer(model, cfg.OPTIMIZATION, total_steps=10)
        tr.zero_grad()
        tr.refresh_bf16_mirror()
        assert len(fused.BF16_SHADOW) > 0
        bd = dict(points=torch.from_numpy(K["points_in"]).cuda(), batch_size=int(K["batch_size"]),
                  voxel_mae_mask=torch.from_numpy(K["voxel_mae_mask"]).cuda())
        ret, _, _ = model(bd)
        fused.INPLACE_PARAM_GRADS = True            # what MAETrainer.step sets around its backward
        try:
            ret["loss"].backward()
        finally:
            fused.INPLACE_PARAM_GRADS = False
        grads = {k: p.grad for k, p in model.named_parameters()}
        for k, gn in zip([str(s) for s in K["grad_keys"]], K["grad_norms"]):
            mine = float(grads[k].norm())
            assert abs(mine - gn) <= 0.15 * gn + 1e-4 * tot_ref, (k, mine, gn)
    finally:
        fused.BF16_SHADOW.clear()
        G.config.set_precision(model, "fp32")


# ------------------------------------------------------------------------------ BASELINE config 0 (KITTI) on the GPU
def test_kitti_c1_matches_reference_golden(G, golden):
    """KITTI gd_mae.yaml shape (216x248 grid, 4 point features -> VFE input 10), batch 1: DynVFE + sst_block_x1
    forward through the plugin classes against the unmodified reference's outputs (tests/golden/kitti_c1.npz)."""
    from gd_mae_b200.pcdet.utils.spconv_utils import spconv
    K = golden("kitti_c1")
    model, cfg, ocfg, P, Bf = build(G, "kitti", 0.85, int(K["param_seed"]))
    assert [int(v) for v in cfg.GRID_SIZE] == [216, 248, 1]
    model.train()
    with torch.no_grad():
        bd = model.vfe(dict(points=torch.from_numpy(K["points_in"]).cuda(), batch_size=1))
        assert np.array_equal(bd["voxel_coords"].cpu().numpy(), K["voxel_coords"])
        assert np.array_equal(bd["point_inverse_indices"].cpu().numpy(), K["inverse"])
        assert rel(bd["voxel_features"][::SUB], K["pillar_features_sub"]) < 1e-4
        vc = bd["voxel_coords"]
        sp = spconv.SparseConvTensor(bd["voxel_features"], vc[:, [0, 2, 3]].contiguous().int(), [248, 216], 1)
        y = model.backbone_3d.sst_blocks[0](sp)
        assert rel(y.features[::SUB], K["block_out_sub"]) < 1e-3   # north_star: 1e-3 relative fp32


# ------------------------------------------------------------------------------ BASELINE config 4 (ONCE) dense stress
def test_once_dense_stress_step_matches_oracle(G):
    """ONCE SSL shape (z in [-5,3], voxel z 8.0, 4 point features, 468x468 grid), one frame whose points crowd the
    sensor: hundreds of pillars hold more than NUM_GT_POINTS=64 points (ground-truth truncation in group_inner_inds)
    and, at mask ratio 0.3, windows of all three drop levels (16/32/64 tokens) occur at every scale.  One full
    iteration (fwd + loss + bwd + clip + AdamOneCycle) against the CPU oracle, fp32 configuration."""
    from gd_mae_b200.trainer import MAETrainer
    model, cfg, ocfg, P, Bf = build(G, "once_ssl", 0.3, 11)
    G.config.set_precision(model, "fp32")
    model.backbone_3d.dense_spatial_features = False
    r = np.random.RandomState(5)
    n = 80000
    az = r.uniform(-np.pi, np.pi, n)
    rr = 2 + r.exponential(3.0, n)
    pts = np.stack([np.zeros(n), rr * np.cos(az), rr * np.sin(az), r.uniform(-4.5, 2.5, n), r.uniform(0, 1, n)], 1).astype(np.float32)
    pts = torch.from_numpy(pts)
    _, opts, _, ovc, oinv = O.voxelize(pts, ocfg)
    counts = torch.bincount(oinv)
    assert int((counts > 64).sum()) >= 50, int((counts > 64).sum())
    noise = torch.rand(ovc.shape[0], generator=torch.Generator().manual_seed(4))
    # all three drop levels present among the visible pillars of the first scale
    mask = O.mae_mask(ovc, 1, noise, 0.3)
    vis = ovc[mask == 0]
    info = O.window_info(vis, ocfg["grid"], (8, 8, 1), 128, 1000.0)
    assert sorted(torch.unique(info[0]["lvl"]).tolist()) == [0, 1, 2]
    opt = O.AdamOneCycle(P, ocfg, 10)
    P0 = {k: v.clone() for k, v in P.items()}
    lo, Gr, _ = O.train_step(P, Bf, opt, pts, 1, ocfg, noise, 0)
    trainer = MAETrainer(model, cfg.OPTIMIZATION, total_steps=10)
    lc = float(trainer.step(dict(points=pts.cuda(), batch_size=1, voxel_mae_noise=noise.cuda())))
    assert abs(lc - lo) / abs(lo) < 1e-3, (lc, lo)
    sd = model.state_dict()
    num = sum(float(((sd[k].cpu().double() - P[k].double()) ** 2).sum()) for k in P)
    den = sum(float(((P[k].double() - P0[k].double()) ** 2).sum()) for k in P)
    assert (num / den) ** 0.5 < 5e-2, (num / den) ** 0.5


# ------------------------------------------------------------------------------ work units of the tensor-core SRA kernels
def test_sra_bin_units_cover_every_row(G, golden):
    """gdmae_sra_bin_units: per 64-row bin the packed (query tile, key range) units of the windows that START in the bin.
    Every CSR row must be a query row of exactly one unit; a packed unit holds whole windows totalling <= 16 rows and
    attends exactly to itself; the chunks of a larger window tile it and share its full key range."""
    K = golden("window_kat")
    g = torch.Generator().manual_seed(11)
    cases = [(torch.from_numpy(K["b1.coords"])[:, [0, 2, 3]].int(), 2) + tuple(int(v) for v in K["b1.grid"][:2])[::-1]]
    dense = torch.nonzero(torch.rand(2, 90, 70, generator=g) < 0.45).int()      # windows of every size up to 64
    cases.append((dense, 2, 90, 70))
    for idx, B, Y, X in cases:
        for shifted in (0, 1):
            t = G.ops.window_table(idx.contiguous().cuda(), B, Y, X, shifted)
            N = t.N
            nbins = (N + 63) // 64
            raw = t.bin_units().cpu()
            off = int(G.pkg._lib.lib().gdmae_sra_tok_info_offset(G.pkg._lib.i64(N)))
            blocks = raw[:off].view(-1, 576)                                     # per bin: 64 ints of units | 128 row records
            units = blocks[:, :64]
            info = t.row_info.cpu()
            # tok_info (N) after the blocks: CSR row | cell << 26 per token
            tok_info = raw[off:off + N].long()
            rows_of_tok = torch.empty(N, dtype=torch.long)
            rows_of_tok[info[:, 0].long()] = torch.arange(N)
            assert torch.equal(tok_info & 0x3ffffff, rows_of_tok)
            assert torch.equal((tok_info >> 26) & 63, t.pos_of_token.cpu().long())
            start, end = info[:, 1].long(), info[:, 2].long()                    # window extent [start, end) of every CSR row
            assert units.shape[0] >= nbins
            seen = torch.zeros(N, dtype=torch.int32)
            for b in range(nbins):
                bin0 = 64 * b
                row0 = bin0 if int(start[bin0]) == bin0 else int(end[bin0])      # first window that starts in the bin
                row1 = N if bin0 + 64 >= N else (bin0 + 64 if int(start[bin0 + 64]) == bin0 + 64 else int(end[bin0 + 64]))
                nu = int(units[b, 48])
                assert 0 <= nu <= 48 and int(units[b, 49]) == 0
                assert int(units[b, 50]) == row0 and int(units[b, 51]) == max(row1 - row0, 0)
                nrec = min(128, N - row0)                                        # the block carries the records from row0 on
                assert torch.equal(blocks[b, 64:64 + 4 * nrec].view(-1, 4), info[row0:row0 + nrec])
                for u in range(nu):
                    code = int(units[b, u])
                    q0, qn, k0, kn = code & 127, (code >> 7) & 31, (code >> 12) & 127, (code >> 19) & 127
                    rows = torch.arange(row0 + q0, row0 + q0 + qn)
                    seen[rows] += 1
                    assert 1 <= qn <= 16 and row0 + k0 + kn <= N
                    # the key range covers exactly the windows of the query rows
                    assert int(start[rows].min()) == row0 + k0 and int(end[rows].max()) == row0 + k0 + kn
                    if kn <= 16:
                        assert (q0, qn) == (k0, kn)                              # packed run of whole windows
                    else:
                        assert int(start[rows[0]]) == row0 + k0 and int(end[rows[-1]]) == row0 + k0 + kn and (q0 - k0) % 16 == 0
                        assert qn == min(16, k0 + kn - q0)
            assert bool((seen == 1).all()), (int((seen == 0).sum()), int((seen > 1).sum()))


# ------------------------------------------------------------------------------ window-major epilogues of the own GEMM
def test_tc_gemm_window_major_epilogues(G, golden):
    """gdmae_tc_gemm mode 4 (in-projection: + positional LUT row, per-head L2 norm, log2(e)/tau into q, rows to CSR order,
    1/|q| and 1/|k| into the row records) and mode 5 (rows of a plain bf16 result to CSR order) against the plain GEMM
    followed by the stand-alone re-layout kernel (gdmae_sra_relayout), for both model widths."""
    from gd_mae_b200 import fused, _lib as L
    import ctypes
    g = torch.Generator().manual_seed(5)
    sites = torch.nonzero(torch.rand(2, 61, 45, generator=g) < 0.3).int().contiguous().cuda()
    for d, shifted in ((128, 0), (256, 1)):
        t = G.ops.window_table(sites, 2, 61, 45, shifted)
        N = t.N
        gen = torch.Generator("cuda").manual_seed(d)
        x = torch.randn(N, d, device="cuda", generator=gen).to(torch.bfloat16)
        w = (torch.randn(3 * d, d, device="cuda", generator=gen) * d ** -0.5).to(torch.bfloat16)
        lut = 0.5 * torch.randn(64, 2 * d, device="cuda", generator=gen)
        tau = torch.full((1,), 0.07, device="cuda")
        units = t.bin_units()
        tok_info = units[int(L.lib().gdmae_sra_tok_info_offset(L.i64(N))):]
        lib = L.lib()
        # reference: plain bf16 GEMM, then the re-layout kernel
        qkv = fused.tc_gemm(x, w.t(), out_dtype=torch.bfloat16)
        dflat = fused.tc_gemm(x, w[:d].t(), out_dtype=torch.bfloat16)                   # any (N, d) bf16 result
        ref = torch.zeros(4 * N * d, dtype=torch.bfloat16, device="cuda")
        lrr_ref = torch.zeros(N, 24, device="cuda")
        L.check(lib.gdmae_sra_relayout(L.P(qkv), L.P(lut), L.P(t.row_info), L.P(tau), L.f32(0.01), L.i64(N), d, L.P(dflat), None,
                                       L.P(ref), L.P(lrr_ref), L.stream()), "gdmae_sra_relayout")
        got = torch.zeros(4 * N * d, dtype=torch.bfloat16, device="cuda")
        lrr = torch.zeros(N, 24, device="cuda")
        E = fused.TcEpilogue()
        E.mode, E.tok_info, E.lut, E.tau, E.tau_min, E.lrr = 4, tok_info.data_ptr(), lut.data_ptr(), tau.data_ptr(), 0.01, lrr.data_ptr()
        L.check(lib.gdmae_tc_gemm(0, 1, L.i64(N), L.i64(3 * d), L.i64(d), L.P(x), L.i64(d), L.P(w), L.i64(d), L.P(got), L.i64(3 * d), 1,
                                  L.f32(0.0), 0, ctypes.byref(E), L.stream()), "gdmae_tc_gemm mode 4")
        E5 = fused.TcEpilogue()
        E5.mode, E5.tok_info, E5.plane0 = 5, tok_info.data_ptr(), 3 * (d // 64)
        L.check(lib.gdmae_tc_gemm(0, 1, L.i64(N), L.i64(d), L.i64(d), L.P(x), L.i64(d), L.P(w[:d].contiguous()), L.i64(d), L.P(got), L.i64(d), 1,
                                  L.f32(0.0), 0, ctypes.byref(E5), L.stream()), "gdmae_tc_gemm mode 5")
        torch.cuda.synchronize()
        a, b = got.view(4, N * d).float(), ref.view(4, N * d).float()
        # q^, k^: the fused epilogue normalises the fp32 accumulator, the reference a bf16-rounded copy of it (2 roundings)
        for part, tol in ((0, 2e-2), (1, 2e-2), (2, 0.0)):
            err = float((a[part] - b[part]).abs().max() / b[part].abs().max())
            print(f"window-major epilogue d={d} tensor {part}: rel err {err:.2e}")
            assert err <= tol, (d, part, err)
        assert torch.equal(got.view(4, -1)[3], ref.view(4, -1)[3])                        # mode 5 is a pure row permutation
        assert rel(lrr[:, 8:], lrr_ref[:, 8:]) < 5e-3


def test_conv3x3_wgrad_matches_torch(G):
    """gdmae_conv3x3_wgrad (tcgen05 / TMA weight gradient of the decoder conv, shifted TMA boxes as tap operands) against
    torch's conv2d_weight in fp32 on the same bf16-rounded tensors; map sizes that are not multiples of the 64-pixel box."""
    from gd_mae_b200 import _lib as L
    for B, Y, X in ((2, 37, 70), (1, 9, 130)):
        gen = torch.Generator("cuda").manual_seed(Y)
        x = torch.randn(B, Y, X, 384, device="cuda", generator=gen).to(torch.bfloat16)
        dy = torch.randn(B, Y, X, 128, device="cuda", generator=gen).to(torch.bfloat16)
        dw = torch.empty(128, 3, 3, 384, device="cuda")
        L.check(L.lib().gdmae_conv3x3_wgrad(L.P(dy), L.P(x), B, Y, X, 384, 128, L.P(dw), 0, L.stream()), "gdmae_conv3x3_wgrad")
        ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (128, 384, 3, 3), dy.float().permute(0, 3, 1, 2), padding=1)
        err = rel(dw.permute(0, 3, 1, 2), ref)
        print(f"conv3x3 wgrad B={B} Y={Y} X={X}: rel err {err:.2e}")
        assert err < 2e-3, err
        # accumulate = 1 adds to dW
        L.check(L.lib().gdmae_conv3x3_wgrad(L.P(dy), L.P(x), B, Y, X, 384, 128, L.P(dw), 1, L.stream()), "gdmae_conv3x3_wgrad")
        assert rel(dw.permute(0, 3, 1, 2), 2 * ref) < 2e-3
    to = ctypes_int()
    L.check(L.lib().gdmae_tc_gemm_timeouts(to), "gdmae_tc_gemm_timeouts")
    assert to._obj.value == 0


def test_deblock_rows_and_tall_linear_match_torch(G):
    """fused.DeblockRowsFunction (ConvTranspose2d as a GEMM on the sparse rows + BatchNorm + ReLU in one node, bf16 operands on
    the own tcgen05 GEMM, bf16 gradient hand-over) against the same computation written with torch in fp32, forward and all
    gradients; ops.TallLinear (decoder_pred with the block-wise weight gradient) against F.linear."""
    from gd_mae_b200 import fused
    torch.manual_seed(3)
    old = fused.GEMM_DTYPE
    fused.GEMM_DTYPE = torch.bfloat16
    try:
        for C_in, c_out, k, N in ((128, 128, 1, 3001), (256, 128, 2, 2500), (256, 128, 4, 1777)):
            deconv = torch.nn.ConvTranspose2d(C_in, c_out, k, stride=k, bias=False).cuda()
            bn = torch.nn.BatchNorm2d(c_out, eps=1e-3, momentum=0.01).cuda()
            with torch.no_grad():
                bn.weight.uniform_(0.5, 1.5)
                bn.bias.uniform_(-0.3, 0.3)
            x = torch.randn(N, C_in, device="cuda", requires_grad=True)
            count = float(N * k * k * 3)
            out, bg = fused.deblock_rows(deconv, bn, k, x, count)
            assert out.dtype == torch.bfloat16          # the rows feed the bf16 map (ops.DenseFill) and nothing else
            gout, gbg = torch.randn_like(out), torch.randn_like(bg)
            torch.autograd.backward([out, bg], [gout, gbg])
            got = [out, bg, x.grad.clone(), deconv.weight.grad.clone(), bn.weight.grad.clone(), bn.bias.grad.clone()]
            # torch fp32 reference of the same function on the bf16-rounded GEMM operands (fp32 accumulation, as the tensor
            # cores do): with unrounded operands u differs by ~4e-3 and a handful of ReLU masks flip, which a max-norm
            # comparison of dx sees as whole rows of W
            x2 = x.detach().bfloat16().float().requires_grad_()
            w2 = deconv.weight.detach().bfloat16().float().requires_grad_()
            g2, b2 = bn.weight.detach().clone().requires_grad_(), bn.bias.detach().clone().requires_grad_()
            u = (x2 @ w2.permute(0, 2, 3, 1).reshape(C_in, k * k * c_out)).view(-1, c_out)
            if fused.DEBLOCK_U_DTYPE == torch.bfloat16:
                # the bf16 configuration keeps u as bf16 (statistics and ReLU mask are taken from those values): the same
                # rounding here, straight through for the gradient - otherwise a handful of masks differ near zero
                u = u + (u.bfloat16().float() - u).detach()
            mean = u.sum(0) / count
            var = (u * u).sum(0) / count - mean * mean
            rstd = torch.rsqrt(var + bn.eps)
            ref_out = torch.relu((u - mean) * rstd * g2 + b2)
            ref_bg = torch.relu(b2 - mean * rstd * g2)
            torch.autograd.backward([ref_out, ref_bg], [gout.float(), gbg])
            ref = [ref_out, ref_bg, x2.grad, w2.grad, g2.grad, b2.grad]
            for name, a, b in zip(("out", "bg", "dx", "dW", "dgamma", "dbeta"), got, ref):
                e = rel(a, b)
                print(f"deblock k={k} {name}: rel err {e:.2e}")
                assert e < 2e-2, (k, name, e)          # bf16 GEMM operands (8-bit mantissa) against fp32
    finally:
        fused.GEMM_DTYPE = old
    x = torch.randn(50003, 128, device="cuda", requires_grad=True)
    lin = torch.nn.Linear(128, 48).cuda()
    y = G.ops.TallLinear.apply(x, lin.weight, lin.bias)
    g = torch.randn_like(y)
    y.backward(g)
    got = [y, x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone()]
    x2 = x.detach().clone().requires_grad_()
    lin.zero_grad()
    y2 = torch.nn.functional.linear(x2, lin.weight, lin.bias)
    y2.backward(g)
    for name, a, b in zip(("y", "dx", "dW", "db"), got, [y2, x2.grad, lin.weight.grad, lin.bias.grad]):
        assert rel(a, b) < 1e-5, (name, rel(a, b))


def test_dense_fill_typed_rows_and_one_pass_backward(G):
    """ops.DenseFill (SparseConvTensor.dense() + cat of the three deblock outputs, spt_backbone_mae.py:125-132) with fp32 and
    bf16 sparse rows, fp32 and bf16 map, on an odd-sized grid (boundary sites whose k x k block leaves the map) against the
    same map assembled with torch indexing; the one-pass backward (rows gathered + background sums) against autograd."""
    torch.manual_seed(11)
    B, Y, X, Cs = 2, 21, 19, 128
    strides = [1, 2, 4]
    lat = lambda n, k: n if k == 1 else lat((n - 1) // 2 + 1, k // 2)  # noqa: E731
    bb, yy, xx = torch.meshgrid(torch.arange(B), torch.arange(Y), torch.arange(X), indexing="ij")
    bb, yy, xx = bb.cuda(), yy.cuda(), xx.cuda()
    for row_dtype, out_dtype in ((torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16), (torch.float32, torch.bfloat16),
                                 (torch.bfloat16, torch.float32)):
        grids, rows, bgs = [], [], []
        for k in strides:
            H, W = lat(Y, k), lat(X, k)
            occ = torch.rand(B, H, W, device="cuda") < 0.4
            grid = torch.full((B, H, W), -1, dtype=torch.int32, device="cuda")
            grid[occ] = torch.arange(int(occ.sum()), dtype=torch.int32, device="cuda")
            grids.append(grid.contiguous())
            rows.append(torch.randn(int(occ.sum()) * k * k, Cs, device="cuda").to(row_dtype).requires_grad_())
            bgs.append(torch.randn(Cs, device="cuda").requires_grad_())
        out = G.ops.DenseFill.apply(rows[0], rows[1], rows[2], bgs[0], bgs[1], bgs[2], grids, None, strides, B, Y, X, out_dtype)
        assert out.dtype == out_dtype and out.shape == (B, Y, X, 3 * Cs)
        gout = torch.randn(B, Y, X, 3 * Cs, device="cuda").to(out_dtype)
        out.backward(gout)
        # torch restatement on fp32 leaves
        rows2 = [r.detach().float().requires_grad_() for r in rows]
        bgs2 = [b.detach().clone().requires_grad_() for b in bgs]
        parts = []
        for k, grid, r2, b2 in zip(strides, grids, rows2, bgs2):
            rank = grid[bb, yy // k, xx // k].long()
            row = rank * k * k + (yy % k) * k + (xx % k)
            part = torch.where((rank >= 0)[..., None], r2[row.clamp(min=0)], b2.expand(B, Y, X, Cs))
            parts.append(part)
        ref = torch.cat(parts, dim=-1)
        assert torch.equal(out.float(), ref.to(out_dtype).float()), (row_dtype, out_dtype)
        ref.backward(gout.float())
        for s in range(3):
            want = rows2[s].grad.to(row_dtype)
            assert rows[s].grad.dtype == row_dtype and torch.equal(rows[s].grad, want), (row_dtype, out_dtype, s)   # pure data movement
            e = rel(bgs[s].grad, bgs2[s].grad)
            assert e < 1e-5, (s, e)        # float sums in a different order


def test_typed_batchnorm_relu_matches_fp32_kernels(G):
    """gdmae_batchnorm_relu_fwd_t / _bwd_t (typed tensors, used by the decoder deblocks of the bf16 configuration) against the
    fp32 entry points: identical arithmetic when every tensor is fp32, and only the roundings of the bf16 tensors otherwise
    (bf16 out / dout / dy; y stays fp32 here so that the ReLU masks are the same on both sides)."""
    import ctypes
    from gd_mae_b200 import _lib as L
    lib = L.lib()
    torch.manual_seed(5)
    N, C = 3001, 128
    u = torch.randn(N, C, device="cuda")
    gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.rand(C, device="cuda") - 0.5
    count = float(3 * N)
    ws = torch.empty(lib.gdmae_batchnorm_workspace_bytes(C), dtype=torch.uint8, device="cuda")
    wsz = ctypes.c_size_t(ws.numel())
    code = {torch.float32: 0, torch.bfloat16: 1}
    o0, m0, r0 = torch.empty_like(u), torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    L.check(lib.gdmae_batchnorm_relu_fwd(L.P(u), L.P(gamma), L.P(beta), L.i64(N), C, ctypes.c_double(count), L.f32(1e-3), L.f32(0.01), 1,
                                         L.P(o0), L.P(m0), L.P(r0), None, None, L.P(ws), wsz, L.stream()), "fwd")
    for od in (torch.float32, torch.bfloat16):
        o1 = torch.empty((N, C), dtype=od, device="cuda")
        m1, r1 = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
        L.check(lib.gdmae_batchnorm_relu_fwd_t(L.P(u), 0, L.P(gamma), L.P(beta), L.i64(N), C, ctypes.c_double(count), L.f32(1e-3), L.f32(0.01),
                                               1, L.P(o1), code[od], L.P(m1), L.P(r1), None, None, L.P(ws), wsz, L.stream()), "fwd_t")
        assert rel(m1, m0) < 1e-6 and rel(r1, r0) < 1e-6
        assert torch.equal(o1, o0.to(od)) or rel(o1.float(), o0) < (1e-6 if od == torch.float32 else 4e-3), od
    dout = torch.randn(N, C, device="cuda").bfloat16()          # a bf16-representable upstream gradient for every variant
    y0, g0, b0 = torch.empty_like(u), torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    L.check(lib.gdmae_batchnorm_relu_bwd(L.P(u), L.P(beta), L.P(dout.float().contiguous()), L.P(gamma), L.P(m0), L.P(r0), L.i64(N), C,
                                         ctypes.c_double(count), 1, None, None, L.P(y0), None, L.P(g0), L.P(b0), L.P(ws), wsz, L.stream()), "bwd")
    for dd in (torch.float32, torch.bfloat16):
        for gd in (torch.float32, torch.bfloat16):
            d = dout.to(dd).contiguous()
            y1 = torch.empty((N, C), dtype=gd, device="cuda")
            g1, b1 = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
            L.check(lib.gdmae_batchnorm_relu_bwd_t(L.P(u), 0, L.P(beta), L.P(d), code[dd], L.P(gamma), L.P(m0), L.P(r0), L.i64(N), C,
                                                   ctypes.c_double(count), 1, None, None, L.P(y1), code[gd], L.P(g1), L.P(b1), L.P(ws), wsz,
                                                   L.stream()), "bwd_t")
            assert rel(g1, g0) < 1e-5 and rel(b1, b0) < 1e-5, (dd, gd)
            assert rel(y1.float(), y0) < (1e-6 if gd == torch.float32 else 4e-3), (dd, gd, rel(y1.float(), y0))


@pytest.mark.parametrize("C", [128, 256, 96])
def test_sparse_conv_gathers_match_indexing(G, C):
    """gdmae_gather_rows / gdmae_gather_rows_transposed (im2col over a 9-tap neighbour map and its transposed gather): the
    bf16 fast paths (C = 128, 256) and the generic kernels (C = 96, fp32 output) against plain indexing."""
    torch.manual_seed(C)
    N, Ns, K = 4099, 3500, 9
    x = torch.randn(Ns, C, device="cuda")
    fmap = torch.randint(-1, Ns, (N, K), device="cuda", dtype=torch.int32)
    fmap[torch.rand(N, K, device="cuda") < 0.5] = -1
    safe = fmap.long().clamp(min=0)
    ref = torch.where((fmap >= 0)[..., None], x[safe], torch.zeros((), device="cuda")).reshape(N, K * C)
    for od in (torch.bfloat16, torch.float32):
        col = G.ops.gather_rows(x, fmap, od)
        assert col.dtype == od and torch.equal(col, ref.to(od)), (C, od)
    # transposed: dsrc[i] = sum_k dcol[tmap[i, k], k]   (mirror = 0)
    tmap = torch.randint(-1, N, (Ns, K), device="cuda", dtype=torch.int32)
    tmap[torch.rand(Ns, K, device="cuda") < 0.5] = -1
    dcol = torch.randn(N, K * C, device="cuda").bfloat16()
    d3 = dcol.float().view(N, K, C)
    for mirror in (False, True):
        want = torch.zeros(Ns, C, device="cuda")
        for k in range(K):
            idx = tmap[:, K - 1 - k if mirror else k].long()
            want += torch.where((idx >= 0)[:, None], d3[idx.clamp(min=0), k], torch.zeros((), device="cuda"))
        got = G.ops.gather_rows_transposed(dcol, tmap, Ns, mirror)
        assert rel(got, want) < 1e-6, (C, mirror, rel(got, want))
        got32 = G.ops.gather_rows_transposed(dcol.float(), tmap, Ns, mirror)
        assert rel(got32, want) < 1e-6, (C, mirror)


def ctypes_int():
    import ctypes
    return ctypes.byref(ctypes.c_int(0))


# ------------------------------------------------------------------------------ index pipeline one step ahead
def test_prefetched_index_pipeline_matches_inline(G):
    """GDMAE.prefetch_index (voxelisation, mask, site sets, window tables of the NEXT batch on a side stream, used by
    MAETrainer.step(batch, next_batch)) must give the same step as building them inside forward()."""
    from gd_mae_b200.trainer import MAETrainer
    model, cfg, ocfg, P, Bf = build(G, "tiny", 0.85, 9)
    model.train()
    r = np.random.RandomState(3)
    batches = []
    for it in range(3):
        n = 2500
        pts = np.concatenate([r.randint(0, 2, (n, 1)), r.normal(0, 3, (n, 2)), r.uniform(-2, 4, (n, 1)), r.uniform(0, 1, (n, 2))], 1)
        pts = torch.from_numpy(pts[np.argsort(pts[:, 0], kind="stable")].astype(np.float32)).cuda()
        _, _, _, ovc, _ = O.voxelize(pts.cpu(), ocfg)
        noise = torch.rand(ovc.shape[0], generator=torch.Generator().manual_seed(it)).cuda()
        batches.append((pts, noise))

    def run(prefetch):
        m, *_ = build(G, "tiny", 0.85, 9)
        tr = MAETrainer(m, cfg.OPTIMIZATION, total_steps=20)
        bds = [dict(points=p.clone(), batch_size=2, voxel_mae_noise=nz) for p, nz in batches]
        losses, lagged = [], []
        for i, bd in enumerate(bds):
            nxt = bds[i + 1] if (prefetch and i + 1 < len(bds)) else None
            loss = tr.step(bd, next_batch=nxt)
            lagged.append(tr.loss_to_host(loss))            # pinned-slot read: the loss of the step before
            losses.append(float(loss))
            if prefetch and i > 0:
                assert '_index_event' not in bd and bd.get('mae_index') is not None   # the prefetched structures were consumed
        assert lagged[0] is None and lagged[1:] == losses[:-1] and tr.drain_loss() == losses[-1], (lagged, losses)
        tr.reserve_memory(main_gb=0.25, side_gb=0.05)       # maps pool memory, leaves every live tensor alone
        return losses, {k: v.detach().clone() for k, v in m.state_dict().items()}

    l0, s0 = run(False)
    l1, s1 = run(True)
    for a, b in zip(l0, l1):
        assert abs(a - b) <= 1e-5 * abs(a), (l0, l1)
    # Adam's first steps move every element by ~lr * sign(g): elements whose gradient is at noise level (float atomics
    # order) may flip, so compare the update as a whole, like test_mask_from_noise_inside_model_and_trainer_steps
    num = sum(float(((s1[k].double() - s0[k].double()) ** 2).sum()) for k in P)
    den = sum(float(((s0[k].double().cpu() - P[k].double()) ** 2).sum()) for k in P)
    assert (num / den) ** 0.5 < 5e-2, (num / den) ** 0.5
    assert G.ops.sra_wait_timeouts() == 0      # no bounded wait inside the SRA kernels ran out anywhere in this test session


# ------------------------------------------------------------------------------ input side
def test_world_augmentation_matches_reference_kat(G, golden):
    """The device DataAugmentor (+ shuffle) on a collated 3-frame batch against the reference's DataAugmentor /
    shuffle_points run frame by frame (tests/golden/augment_kat.npz): same numpy stream -> same parameters and permutations
    (exact), points within fp32 rounding of the reference's matmul (1e-6 of the coordinate range)."""
    from gd_mae_b200.pcdet.datasets.augmentor.data_augmentor import DataAugmentor
    K = golden("augment_kat")
    acfg = G.config.to_attr({"DISABLE_AUG_LIST": ["placeholder"], "AUG_CONFIG_LIST": [
        {"NAME": "random_world_flip", "PROBABILITY": 0.5, "ALONG_AXIS_LIST": ["x", "y"]},
        {"NAME": "random_world_rotation", "PROBABILITY": 1.0, "WORLD_ROT_ANGLE": [-0.78539816, 0.78539816]},
        {"NAME": "random_world_scaling", "PROBABILITY": 1.0, "WORLD_SCALE_RANGE": [0.95, 1.05]}]})
    aug = DataAugmentor(None, acfg, ["Vehicle", "Pedestrian", "Cyclist"])
    frames = [K[f"f{f}.points_in"] for f in range(3)]
    batch = np.concatenate([np.concatenate([np.full((p.shape[0], 1), f, np.float32), p], 1) for f, p in enumerate(frames)], 0)
    np.random.seed(1234)
    dd = aug.forward({"points": torch.from_numpy(batch).cuda(), "batch_size": 3}, shuffle=True)
    out = dd["points"].cpu().numpy()
    off = 0
    for f in range(3):
        prm = dd["transformation_3d_params"][f]
        assert ("x" in prm["random_world_flip"]) == bool(K[f"f{f}.flip_x"]) and ("y" in prm["random_world_flip"]) == bool(K[f"f{f}.flip_y"])
        assert prm["random_world_rotation"] == float(K[f"f{f}.rotation"]) and prm["random_world_scaling"] == float(K[f"f{f}.scaling"])
        n = frames[f].shape[0]
        mine, ref = out[off:off + n], K[f"f{f}.points_out"]
        assert np.all(mine[:, 0] == f)
        assert np.array_equal(mine[:, 4:], ref[:, 3:])                       # untouched features follow the permutation exactly
        assert np.abs(mine[:, 1:4] - ref[:, :3]).max() <= 1e-6 * np.abs(ref[:, :3]).max()
        off += n
    with pytest.raises(Exception):
        G.ops.world_augment(torch.from_numpy(batch), torch.zeros(3, 6))      # CPU tensor: no fallback


# ------------------------------------------------------------------------------ r2: parity at the benched size / precision
def _oracle_sra(qkv32, lutb, tau, bv, coords4, H, W, shift, d):
    """The reference's windowed cosine attention (flat2window -> _scaled_cosine_attention -> window2flat) on the oracle's
    window bookkeeping, written on q = qkv_q + lut[pos], k = qkv_k + lut[pos], v = qkv_v (+ bv on the output), the split the
    kernels take.  Returns (out (N,d), drop levels present)."""
    win, ciw, _ = O.get_window_coors(coords4, [W, H, 1], (8, 8, 1), shift == 1)
    keep, lvl = O.drop_single_shift(win)
    assert bool(keep.all())
    f2w = O.get_flat2win_inds(win, lvl)
    pos = ciw[:, 1] * 8 + ciw[:, 2]
    q = qkv32[:, :d] + lutb[pos, :d]
    k = qkv32[:, d:2 * d] + lutb[pos, d:]
    v = qkv32[:, 2 * d:]
    q3, k3, v3 = O.flat2window(q, f2w), O.flat2window(k, f2w), O.flat2window(v, f2w)
    ones3 = O.flat2window(torch.ones((q.shape[0], 1), dtype=torch.bool), f2w)
    out3 = {}
    for dl in q3:
        key_pad = ones3[dl].logical_not().squeeze(2)
        out3[dl] = O.cosine_attention_core(q3[dl], k3[dl], v3[dl], key_pad, tau, 0.01, 8)
    return O.window2flat(out3, f2w, q.shape[0]) + bv, sorted(f2w.keys())


@pytest.mark.parametrize("d", [128, 256])
def test_sra_tensor_core_kernels_match_oracle(G, d):
    """the bf16 tensor-core SRA forward AND backward directly against the oracle's window attention
    (autograd for the backward), not against the repo's own SIMT kernel.  All three drop levels occur in both shifts.
    Inputs are bf16-representable on both sides (q/k/v, dO and the positional LUT are what the kernel reads), so the
    differences are the kernel's own rounding: normalised q/k and the probabilities P are bf16 tensor-core operands
    (2^-9 relative each, two chained products) -> 1e-2 of the output range forward, 2e-2 backward; tau gradient 3e-2."""
    g = torch.Generator().manual_seed(70 + d)
    B, H, W = 2, 52, 45
    occ = torch.rand(B, H, W, generator=g) < 0.6                     # dense: windows of 32..64 tokens (level 2)
    occ[1, :, 24:] &= torch.rand(H, 21, generator=g) < 0.45          # level 1
    occ[0, 30:, :] &= torch.rand(22, W, generator=g) < 0.08          # level 0, down to single-token windows
    idx = torch.nonzero(occ).int().contiguous()
    N = idx.shape[0]
    coords4 = torch.stack([idx[:, 0], torch.zeros(N, dtype=torch.int32), idx[:, 1], idx[:, 2]], 1).long()
    qkv_b = torch.randn(N, 3 * d, generator=g).to(torch.bfloat16)
    lutb = (0.5 * torch.randn(64, 2 * d, generator=g)).to(torch.bfloat16).float()
    do_b = torch.randn(N, d, generator=g).to(torch.bfloat16)
    bv = torch.randn(d, generator=g)
    for shift in (0, 1):
        qkv32 = qkv_b.float().requires_grad_(True)
        tau = torch.tensor([0.7], requires_grad=True)
        o_ref, levels = _oracle_sra(qkv32, lutb, tau, bv, coords4, H, W, shift, d)
        assert levels == [0, 1, 2], levels
        (o_ref * do_b.float()).sum().backward()
        table = G.ops.window_table(idx.cuda(), B, H, W, shift)
        tau_c = torch.tensor([0.7]).cuda()
        o_tc, lse = G.ops.sra_fwd(qkv_b.cuda(), lutb.cuda(), tau_c, table, 0.01, 8, bv=bv.cuda())
        e_o = rel(o_tc, o_ref)
        dqkv, dts = G.ops.sra_bwd(qkv_b.cuda(), lutb.cuda(), tau_c, table, 0.01, 8, None, lse, do_b.cuda())
        errs = {nm: rel(dqkv[:, c0:c1].float(), qkv32.grad[:, c0:c1]) for nm, c0, c1 in (("dq", 0, d), ("dk", d, 2 * d), ("dv", 2 * d, 3 * d))}
        dtau_tc = -float(dts) / 0.7                                   # dS/dtau = -S / tau  (csrc/encoder_layer.cu dtau_kernel)
        e_t = abs(dtau_tc - float(tau.grad)) / max(abs(float(tau.grad)), 1e-6)
        print(f"sra tc vs oracle d={d} shift={shift}: out {e_o:.2e} {({k: f'{v:.2e}' for k, v in errs.items()})} dtau {e_t:.2e}")
        assert e_o < 1e-2, e_o
        for nm, e in errs.items():
            assert e < 2e-2, (nm, e)
        assert e_t < 3e-2, (dtau_tc, float(tau.grad))


@pytest.fixture(scope="module")
def waymo_frame_oracle():
    """One full-size Waymo-shape frame (BASELINE config C2 at B=1, ~159 k points) through the CPU oracle's training step:
    indices, loss and every parameter gradient (a few seconds of CPU time, shared by the fp32 and the bf16 test)."""
    ocfg = O.make_cfg("waymo_ssl")
    pts = torch.from_numpy(O.synth_batch([7], ocfg))
    P, Bf = O.init_params(ocfg, 11)
    keep, opts, ocoords, ovc, oinv = O.voxelize(pts, ocfg)
    noise = torch.rand(ovc.shape[0], generator=torch.Generator().manual_seed(666))
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    trace = {}
    loss, _ = O.mae_forward(leaves, pts, 1, ocfg, noise=noise, stats={k: v.clone() for k, v in Bf.items()}, trace=trace)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return dict(pts=pts, P=P, Bf=Bf, noise=noise, loss=float(loss), grads=grads, voxel_coords=ovc, inverse=oinv,
                n_tokens=[trace[f"x_conv{i + 1}.indices"].shape[0] for i in range(3)],
                x_idx=[trace[f"x_conv{i + 1}.indices"] for i in range(3)], voxel_features=trace["voxel_features"].detach(),
                pillar_features=trace["pillar_features"].detach())


def _waymo_model(G, W):
    cfg = G.config.builtin_cfg("waymo_ssl")
    model = G.config.build_mae_model(cfg).cuda()
    sd = dict(W["P"])
    sd.update(W["Bf"])
    missing = model.load_state_dict(sd, strict=False)
    assert missing.missing_keys == ["global_step"] and not missing.unexpected_keys
    return model.train(), cfg


def _grad_report(model, W):
    grads = {k: p.grad for k, p in model.named_parameters()}
    tot_ref = float(torch.sqrt(sum((g.double() ** 2).sum() for g in W["grads"].values())))
    tot = float(torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())))
    worst = []
    for k, gr in W["grads"].items():
        a, b = float(grads[k].norm()), float(gr.norm())
        worst.append((abs(a - b) / max(b, 1e-3 * tot_ref), k, a, b))
    worst.sort(reverse=True)
    return tot, tot_ref, worst


def test_waymo_size_full_step_fp32_matches_oracle(G, waymo_frame_oracle):
    """full MAE step at the Waymo C2 frame size (B=1, ~159 k points, 35 k pillars, 468x468 grid) in
    the fp32 parity configuration against the oracle: indices bit-exact, loss <= 1e-3 (north_star), token counts equal,
    features <= 1e-3, every parameter's gradient norm <= 5e-3 (relative to max(its norm, 1e-3 of the total norm))."""
    W = waymo_frame_oracle
    model, cfg = _waymo_model(G, W)
    G.config.set_precision(model, "fp32", dense_spatial_features=True)
    bd = dict(points=W["pts"].cuda(), batch_size=1, voxel_mae_noise=W["noise"].cuda())
    ret, _, _ = model(bd)
    ret["loss"].backward()
    assert torch.equal(bd["voxel_coords"].cpu(), W["voxel_coords"])
    assert torch.equal(bd["point_inverse_indices"].cpu(), W["inverse"])
    for i in range(3):
        sp = bd["multi_scale_3d_features"][f"x_conv{i + 1}"]
        assert torch.equal(sp.indices.cpu().long(), W["x_idx"][i].long()), i
    assert rel(bd["pillar_features"][::SUB], W["pillar_features"][::SUB]) < 1e-4
    assert rel(bd["voxel_features"][::SUB], W["voxel_features"][::SUB]) < 1e-3
    e_loss = abs(float(ret["loss"]) - W["loss"]) / W["loss"]
    tot, tot_ref, worst = _grad_report(model, W)
    print(f"waymo fp32: loss {float(ret['loss']):.6f} vs {W['loss']:.6f} ({e_loss:.1e}); |g| {tot:.5f} vs {tot_ref:.5f}; worst {worst[:3]}")
    assert e_loss < 1e-3
    assert abs(tot - tot_ref) / tot_ref < 2e-3
    for e, k, a, b in worst:
        assert e < (5e-2 if k.endswith(".tau") else 5e-3), (k, a, b)


def test_waymo_size_full_step_bf16_close_to_oracle(G, waymo_frame_oracle):
    """the same frame in the BENCHED configuration (bf16 GEMM operands / attention / decoder map,
    fp32 accumulation and statistics) against the fp32 oracle.  Indices stay bit-exact.  Stated tolerances and why
    (measured on a B200, r2: loss 1.3e-5, features 8.0e-3, total gradient norm 1.6e-4, worst parameter 1.2e-2):
    loss 1e-3 (north_star's figure; SURVEY 7 probe: the reference itself under bf16 autocast moves the loss by 1.1e-4 and
    decoder features by 1.6e-2 relative L2 - here q/k/v, the attention output and four GEMM outputs per layer are bf16 as
    well); decoder features at the pillars 2e-2 of their range (bf16 map: 2^-9 per value, a 3x3x384 reduction of them);
    total gradient norm 5e-3; per-parameter gradient norms 5 % of max(own norm, 1e-2 of the total) - bf16 dO / dqkv
    quantisation is the dominant term for the VFE weights at the end of the backward chain."""
    from gd_mae_b200 import fused
    W = waymo_frame_oracle
    model, cfg = _waymo_model(G, W)
    try:
        G.config.set_precision(model, "bf16", dense_spatial_features=False)
        bd = dict(points=W["pts"].cuda(), batch_size=1, voxel_mae_noise=W["noise"].cuda())
        ret, _, _ = model(bd)
        ret["loss"].backward()
        assert torch.equal(bd["voxel_coords"].cpu(), W["voxel_coords"])
        assert torch.equal(bd["point_inverse_indices"].cpu(), W["inverse"])
        for i in range(3):
            sp = bd["multi_scale_3d_features"][f"x_conv{i + 1}"]
            assert torch.equal(sp.indices.cpu().long(), W["x_idx"][i].long()), i
        e_loss = abs(float(ret["loss"]) - W["loss"]) / W["loss"]
        e_vf = rel(bd["voxel_features"][::SUB].float(), W["voxel_features"][::SUB])
        tot, tot_ref, _ = _grad_report(model, W)
        grads = {k: p.grad for k, p in model.named_parameters()}
        worst = sorted(((abs(float(grads[k].norm()) - float(g.norm())) / max(float(g.norm()), 1e-2 * tot_ref), k)
                        for k, g in W["grads"].items()), reverse=True)
        print(f"waymo bf16: loss {float(ret['loss']):.6f} vs {W['loss']:.6f} ({e_loss:.1e}); voxel_features {e_vf:.1e}; "
              f"|g| {tot:.5f} vs {tot_ref:.5f}; worst {worst[:4]}")
        assert e_loss < 1e-3
        assert e_vf < 2e-2
        assert abs(tot - tot_ref) / tot_ref < 5e-3
        for e, k in worst:
            assert e < 0.05, (k, e)
        assert G.ops.sra_wait_timeouts() == 0
    finally:
        fused.BF16_SHADOW.clear()
        G.config.set_precision(model, "fp32")


def test_trainer_matches_reference_optimizer_golden(G, golden):
    """The CUDA trainer's clip + adam_onecycle update (flat bucket, fused kernel) on the gradients of
    tests/golden/optimizer_kat.npz, against the REFERENCE's OptimWrapper / OneCycle run stored there (5 iterations)."""
    from gd_mae_b200.trainer import MAETrainer
    K = golden("optimizer_kat")
    model, cfg, ocfg, P, Bf = build(G, "tiny", 0.85, int(K["param_seed"]))
    trainer = MAETrainer(model, cfg.OPTIMIZATION, total_steps=int(K["total_steps"]))
    params = dict(model.named_parameters())
    P0 = {k: v.clone() for k, v in P.items()}
    watch = [str(k) for k in K["watch"]]
    for it in range(int(K["n_iters"])):
        g = torch.Generator().manual_seed(900 + it)
        scale = 30.0 if it == 1 else 1.0
        for k, v in P.items():
            params[k].grad.copy_((torch.randn(v.shape, generator=g) * (0.002 * scale)).cuda())
        lr, mom = trainer.optimizer_step()
        assert abs(lr - float(K["lr"][it])) <= 1e-9 * lr + 1e-12 and abs(mom - float(K["mom"][it])) <= 1e-12
        assert abs(float(trainer.sumsq.sqrt()) - float(K["total_norm"][it])) / float(K["total_norm"][it]) < 1e-5
        for k in watch:
            flat = params[k].detach().reshape(-1)
            mine = flat[::max(1, flat.numel() // 1500)].cpu()
            if O.in_optimizer(k):
                assert rel(mine, K[f"it{it}.{k}"]) < 2e-5, (it, k)
            else:
                assert np.array_equal(mine.numpy(), K[f"it{it}.{k}"]), (it, k)
    for k, n in zip([str(s) for s in K["final_keys"]], K["final_norm"]):
        assert abs(float(params[k].double().norm()) - n) <= 2e-5 * max(n, 1e-6), k
        if not O.in_optimizer(k):
            assert torch.equal(params[k].detach().cpu(), P0[k]), k


def test_prefetch_waits_for_late_points(G):
    """MAETrainer.step(batch, next_batch=nxt) WITHOUT a ready event while next_batch['points'] is still being
    produced by work queued on the main stream.  The index side stream must wait for it (an event recorded at step entry);
    it used to start at once and voxelise whatever the buffer held (here: NaNs -> no pillars)."""
    from gd_mae_b200.trainer import MAETrainer
    model, cfg, ocfg, P, Bf = build(G, "tiny", 0.85, 9)
    tr = MAETrainer(model, cfg.OPTIMIZATION, total_steps=20)
    r = np.random.RandomState(5)
    n = 2500
    pts = np.concatenate([r.randint(0, 2, (n, 1)), r.normal(0, 3, (n, 2)), r.uniform(-2, 4, (n, 1)), r.uniform(0, 1, (n, 2))], 1)
    pts = torch.from_numpy(pts[np.argsort(pts[:, 0], kind="stable")].astype(np.float32))
    _, _, _, ovc, _ = O.voxelize(pts, ocfg)
    cur = dict(points=pts.cuda(), batch_size=2)
    late = torch.full(pts.shape, float("nan"), device="cuda")
    big = torch.randn(6144, 6144, device="cuda")
    torch.cuda.synchronize()
    for _ in range(12):                      # ~tens of ms of queued main-stream work ahead of the producer of `late`
        big = (big @ big) * 1e-4
    late.copy_(pts.cuda())                   # the producer: enqueued on the main stream, far from done when step() is called
    nxt = dict(points=late, batch_size=2)
    tr.step(cur, next_batch=nxt)
    torch.cuda.synchronize()
    assert torch.equal(nxt["voxel_coords"].cpu(), ovc)
    assert nxt.get("mae_index") is not None
    loss = tr.step(nxt)
    assert torch.isfinite(loss)


def test_trainer_state_dict_roundtrip(G):
    """optimizer state (Adam moments + both counters) survives state_dict -> load_state_dict into a fresh trainer:
    the resumed trainer takes the same next step (train_utils.checkpoint_state / load_params_with_optimizer semantics)."""
    from gd_mae_b200.trainer import MAETrainer
    model, cfg, ocfg, P, Bf = build(G, "tiny", 0.85, 3)
    tr = MAETrainer(model, cfg.OPTIMIZATION, total_steps=20)
    g = torch.Generator().manual_seed(1)
    params = dict(model.named_parameters())

    def fake_step(trainer, prm):
        for k in P:
            prm[k].grad.copy_((torch.randn(P[k].shape, generator=g) * 1e-3).cuda())
        trainer.optimizer_step()

    fake_step(tr, params)
    fake_step(tr, params)
    sd_opt, sd_model = tr.state_dict(), {k: v.detach().clone() for k, v in model.state_dict().items()}
    g_state = g.get_state()
    fake_step(tr, params)
    want = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model2, *_ = build(G, "tiny", 0.85, 3)
    model2.load_state_dict(sd_model)
    tr2 = MAETrainer(model2, cfg.OPTIMIZATION, total_steps=20)
    tr2.load_state_dict(sd_opt)
    assert (tr2.it, tr2.t) == (2, 2)
    g.set_state(g_state)
    fake_step(tr2, dict(model2.named_parameters()))
    for k, v in model2.state_dict().items():
        assert torch.equal(v, want[k]), k


# ------------------------------------------------------------------------------ iou3d_nms on the device
def test_iou3d_kernels_match_oracle_and_reference_golden(G, golden):
    """csrc/iou3d_nms.cu through the reference-named Python API against (a) the known answers of the reference's own CPU
    implementation and (b) the C oracle on fresh boxes.  fp32 geometry with device sinf / cosf / atan2f: differences are
    last-bit except where a corner sits within rounding of the reference's 1e-2 margin test (the polygon then gains or loses
    that corner: <= 5e-3 in IoU), hence 99.9 % of the pairs within 1e-5 and all within 5e-3."""
    from gd_mae_b200.pcdet.ops.iou3d_nms import iou3d_nms_utils as U
    from oracle import iou3d_oracle as IO
    K = golden("iou3d_kat")
    for name in ("special", "rand_a", "rand_self"):
        iou = U.boxes_iou_bev(torch.from_numpy(K[name + ".a"]).cuda(), torch.from_numpy(K[name + ".b"]).cuda()).cpu().numpy()
        d = np.abs(iou - K[name + ".iou"])
        assert d.max() <= 5e-3 and (d <= 1e-5).mean() >= 0.999, (name, d.max(), (d <= 1e-5).mean())
    a, b = IO.random_boxes(700, 21, spread=12.0), IO.random_boxes(500, 22, spread=12.0)
    ov = torch.zeros(700, 500, device="cuda")
    U.iou3d_nms_cuda.boxes_overlap_bev_gpu(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), ov)
    d = np.abs(ov.cpu().numpy() - IO.boxes_overlap_bev(a, b))
    assert d.max() <= 5e-2 and (d <= 1e-4).mean() >= 0.999, (d.max(), (d <= 1e-4).mean())
    i3 = U.boxes_iou3d_gpu(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()
    d = np.abs(i3 - IO.boxes_iou3d(a, b))
    assert d.max() <= 5e-3 and (d <= 1e-5).mean() >= 0.999
    assert U.boxes_iou_bev(torch.zeros(0, 7).cuda(), torch.from_numpy(b).cuda()).shape == (0, 500)     # empty input


@pytest.mark.parametrize("n,spread", [(1, 5.0), (63, 4.0), (64, 4.0), (65, 4.0), (1000, 15.0), (4096, 30.0)])
def test_nms_kernels_match_oracle(G, n, spread):
    """rotated and axis-aligned NMS (mask kernel + on-device sweep) against the oracle's sweep: identical kept indices.
    Sizes straddle the 64-box word boundary and go up to the largest NMS_PRE_MAXSIZE of the configs."""
    from gd_mae_b200.pcdet.ops.iou3d_nms import iou3d_nms_utils as U
    from oracle import iou3d_oracle as IO
    boxes = IO.random_boxes(n, 100 + n, spread=spread)
    scores = np.random.RandomState(n).uniform(0, 1, n).astype(np.float32)
    bc, sc = torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda()

    def sweep(iou, thresh):
        """the reference's host loop (iou3d_nms.cpp:113-128) on a given IoU matrix of the score-sorted boxes"""
        removed, keep = np.zeros(iou.shape[0], dtype=bool), []
        for i in range(iou.shape[0]):
            if not removed[i]:
                keep.append(i)
                removed[i + 1:] |= iou[i, i + 1:] > thresh
        return np.array(keep, dtype=np.int64)

    # (1) the mask kernel + on-device sweep against the host sweep over the DEVICE's own IoU values: exact at every size
    order = np.argsort(-scores, kind="stable")
    srt = torch.from_numpy(boxes[order]).cuda()
    iou_dev = U.boxes_iou_bev(srt, srt).cpu().numpy()
    keep, _ = U.nms_gpu(bc, sc, 0.25)
    assert np.array_equal(keep.cpu().numpy(), order[sweep(iou_dev, 0.25)])
    # (2) against the oracle (host libm): a pair whose IoU lies within 1e-4 of the threshold may flip, so the threshold is
    # chosen so that no pair of the case does; with 16 M pairs (n = 4096) no such threshold exists and (1) stands alone
    if n <= 1000:
        iou = IO.boxes_iou_bev(boxes, boxes)
        thresh = next(t for t in np.arange(0.25, 0.45, 0.003) if not (np.abs(iou - t) < 1e-4).any())
        keep, _ = U.nms_gpu(bc, sc, thresh)
        assert np.array_equal(keep.cpu().numpy(), IO.nms(boxes, scores, thresh))
        keep, _ = U.nms_gpu(bc, sc, thresh, pre_maxsize=max(n // 3, 1))
        assert np.array_equal(keep.cpu().numpy(), IO.nms(boxes, scores, thresh, pre_maxsize=max(n // 3, 1)))
    keep, _ = U.nms_normal_gpu(bc, sc, 0.25)                      # plain arithmetic, identical on both sides
    assert np.array_equal(keep.cpu().numpy(), IO.nms_normal(boxes, scores, 0.25))


# ------------------------------------------------------------------------------ finetune path (config 4)
def _build_finetune(G, name, K):
    cfg = G.config.builtin_cfg(name)
    model = G.config.build_mae_model(cfg).cuda()
    ocfg = O.make_cfg("tiny" if name.startswith("tiny") else "waymo_ssl")
    P, Bf = O.init_params(ocfg, int(K["param_seed"]) if K is not None else 7)
    sd = {k: v for k, v in model.state_dict().items()}
    new = {}
    for k, v in list(P.items()) + list(Bf.items()):
        k2 = k.replace("backbone_3d.decoder_deblocks", "backbone_3d.deblocks").replace("backbone_3d.decoder_conv_out", "backbone_3d.conv_out")
        if k2 in sd and sd[k2].shape == v.shape:
            new[k2] = v
    new.update(O.finetune_head_state({k: v.shape for k, v in sd.items() if k.startswith(("backbone_2d.", "dense_head."))},
                                     seed=int(K["head_seed"]) if K is not None else 3))
    missing = model.load_state_dict(new, strict=False)
    assert missing.missing_keys == ["global_step"] and not missing.unexpected_keys, missing
    return model, cfg, ocfg


def test_center_assign_targets_kernel_matches_oracle_and_golden(G, golden):
    """ops.center_assign_targets (one launch, on the device) against the reference's CPU loop (golden) and the oracle on a
    larger random case: heat map, indices, masks, IoU boxes exact; regression targets 1e-6 (device logf / cosf / sinf)."""
    from oracle import center_oracle as CO
    K = golden("finetune_tiny")
    cfg = O.make_cfg("tiny")
    X, Y, _ = cfg["grid"]
    cmap = torch.tensor([0, 1, 2, 3], dtype=torch.int32).cuda()
    heat, tgt, iou_boxes, inds, mask = G.ops.center_assign_targets(torch.from_numpy(K["gt_boxes"]).cuda(), cmap, 3, (Y, X), cfg["pc_range"],
                                                                   cfg["voxel"], 1)
    ref_heat = np.zeros(int(np.prod(K["heatmap.shape"])), dtype=np.float32)
    ref_heat[K["heatmap.nz_index"]] = K["heatmap.nz_value"]
    assert np.abs(heat.cpu().numpy().reshape(-1) - ref_heat).max() <= 1e-7
    assert np.array_equal(inds.cpu().numpy(), K["inds"]) and np.array_equal(mask.cpu().numpy(), K["masks"])
    assert np.array_equal(iou_boxes.cpu().numpy(), K["iou_boxes"])
    assert np.abs(tgt.cpu().numpy() - K["target_boxes"]).max() <= 1e-6
    # Waymo-size map, 300 boxes in 3 frames incl. zero padding, a class outside the head, more boxes than slots
    wcfg = O.make_cfg("waymo_ssl")
    r = np.random.RandomState(0)
    n = 300
    cls = r.randint(0, 4, (3, n)).astype(np.float32)
    dims = np.array([[1, 1, 1], [4.7, 2.1, 1.7], [0.9, 0.9, 1.7], [1.8, 0.8, 1.7]])[cls.astype(int)] * r.uniform(0.8, 1.2, (3, n, 3))
    gt = np.concatenate([r.uniform(-76, 76, (3, n, 2)), r.uniform(-1, 2, (3, n, 1)), dims, r.uniform(-3.2, 3.2, (3, n, 1)), cls[..., None]], 2)
    gt[cls == 0] = 0
    gt = torch.from_numpy(gt.astype(np.float32))
    cm = [0, 1, 0, 2]                                   # the head covers classes 1 and 3 only
    want = CO.assign_targets(gt, cm, 2, 468, 468, wcfg["pc_range"], wcfg["voxel"], max_objs=120)
    got = G.ops.center_assign_targets(gt.cuda(), torch.tensor(cm, dtype=torch.int32).cuda(), 2, (468, 468), wcfg["pc_range"], wcfg["voxel"], 1,
                                      num_max_objs=120)
    assert np.abs(got[0].cpu().numpy() - want[0].numpy()).max() <= 1e-7
    assert torch.equal(got[3].cpu(), want[3]) and torch.equal(got[4].cpu(), want[4]) and torch.equal(got[2].cpu(), want[2])
    assert int(want[4].sum()) == 3 * 120 or int(want[4].sum()) > 200
    assert (got[1].cpu() - want[1]).abs().max() <= 1e-6


def test_center_focal_loss_kernel_matches_oracle(G):
    from oracle import center_oracle as CO
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(2, 3, 40, 48, generator=g) * 4                             # +-12: both clamp sides are hit
    edge = (logits.abs() - 9.2102).abs() < 0.02           # sigmoid within rounding of the clamp limits: either side is legitimate
    logits = torch.where(edge, logits * 1.01, logits).requires_grad_(True)
    heat = torch.rand(2, 3, 40, 48, generator=g) ** 6
    heat[0, 1, 5, 7] = heat[1, 2, 30, 40] = heat[1, 0, 0, 0] = 1.0
    for gt in (heat, heat.clamp(max=0.99)):                                        # second case: no positive at all
        want = CO.focal_loss_from_logits(logits, gt)
        gw, = torch.autograd.grad(want, logits)
        lc = logits.detach().cuda().requires_grad_(True)
        got = G.ops.CenterFocalLoss.apply(lc, gt.cuda())
        gg, = torch.autograd.grad(got, lc)
        assert abs(float(got) - float(want)) <= 1e-5 * abs(float(want))
        assert rel(gg, gw) < 1e-5


def test_finetune_step_matches_reference_golden(G, golden):
    """BASELINE config 4 on the tiny grid: DynVFE + SPTBackbone (no masking, all three drop levels) + SSTBEVBackbone +
    CenterHead, forward + loss + backward in the fp32 parity configuration against the unmodified reference's run
    (tests/golden/finetune_tiny.npz): indices exact, features / predictions 1e-3, loss terms 1e-3, gradient norms 5e-3."""
    K = golden("finetune_tiny")
    model, cfg, ocfg = _build_finetune(G, "tiny_iou", K)
    assert [str(k) for k in K["state_keys"]] == [k for k in model.state_dict().keys() if k != "global_step"] or \
        set(str(k) for k in K["state_keys"]) == set(model.state_dict().keys()) - {"global_step"}
    G.config.set_precision(model, "fp32")
    model.train()
    bd = dict(points=torch.from_numpy(K["points_in"]).cuda(), batch_size=int(K["batch_size"]), gt_boxes=torch.from_numpy(K["gt_boxes"]).cuda())
    ret, tb, _ = model(bd)
    ret["loss"].backward()
    for i in range(3):
        sp = bd["multi_scale_3d_features"][f"x_conv{i + 1}"]
        assert np.array_equal(sp.indices.cpu().numpy(), K[f"x_conv{i + 1}.indices"])
        assert rel(sp.features[::SUB], K[f"x_conv{i + 1}.features.sub"]) < 1e-3
    assert rel(bd["spatial_features"][:, ::16, ::5, ::5], K["spatial_features.sub"]) < 1e-3
    assert rel(bd["spatial_features_2d"][:, ::16, ::5, ::5], K["spatial_features_2d.sub"]) < 1e-3
    pd = model.dense_head.forward_ret_dict["pred_dicts"][0]
    for k, v in pd.items():
        if k == "hm":      # the reference's get_loss replaces pred_dict['hm'] by its clamped sigmoid in place (center_head.py:246)
            v = torch.clamp(v.sigmoid(), min=1e-4, max=1 - 1e-4)
        assert rel(v[:, :, ::5, ::5], K["pred." + k + ".sub"]) < 2e-3, k
    td = model.dense_head.forward_ret_dict["target_dicts"]
    assert np.array_equal(td["inds"][0].cpu().numpy(), K["inds"]) and np.array_equal(td["masks"][0].cpu().numpy(), K["masks"])
    for k in ("hm_loss_head_0", "loc_loss_head_0", "iou_loss_head_0"):
        assert abs(float(tb[k]) - float(K["tb." + k])) <= 1e-3 * abs(float(K["tb." + k])) + 1e-5, (k, float(tb[k]), float(K["tb." + k]))
    assert abs(float(ret["loss"]) - float(K["loss"])) / float(K["loss"]) < 1e-3
    grads = {k: p.grad for k, p in model.named_parameters()}
    tot_ref = float(np.sqrt((K["grad_norms"] ** 2).sum()))
    for k, gn in zip([str(s) for s in K["grad_keys"]], K["grad_norms"]):
        mine = float(grads[k].norm()) if grads[k] is not None else 0.0
        rtol = 5e-2 if k.endswith(".tau") else 5e-3
        assert abs(mine - gn) <= rtol * max(gn, 1e-4 * tot_ref) + 1e-7, (k, mine, gn)


def test_finetune_eval_decodes_and_suppresses(G, golden):
    """eval path of CenterHead (generate_predicted_boxes: top-K decode, range / score filter, IoU-rectified multi-class
    rotated NMS through the own NMS kernels): structural checks - boxes inside the limit range, labels 1..3, scores sorted
    per class by the NMS, no surviving pair of one class above its IoU threshold."""
    from gd_mae_b200.pcdet.ops.iou3d_nms import iou3d_nms_utils as U
    K = golden("finetune_tiny")
    model, cfg, ocfg = _build_finetune(G, "tiny_iou", K)
    G.config.set_precision(model, "fp32")
    model.eval()
    with torch.no_grad():
        preds, _ = model(dict(points=torch.from_numpy(K["points_in"]).cuda(), batch_size=int(K["batch_size"])))
    assert len(preds) == int(K["batch_size"])
    th = cfg.MODEL.DENSE_HEAD.POST_PROCESSING.NMS_CONFIG.NMS_THRESH
    for p in preds:
        assert p["pred_boxes"].shape[1] == 7 and p["pred_boxes"].shape[0] == p["pred_scores"].shape[0] == p["pred_labels"].shape[0]
        if p["pred_boxes"].shape[0] == 0:
            continue
        assert int(p["pred_labels"].min()) >= 1 and int(p["pred_labels"].max()) <= 3
        for c in (1, 2, 3):
            b = p["pred_boxes"][p["pred_labels"] == c]
            if b.shape[0] > 1:
                iou = U.boxes_iou_bev(b.contiguous(), b.contiguous())
                assert float(torch.triu(iou, 1).max()) <= th[c - 1] + 1e-4


def test_vfe_node_with_pillars_beyond_byte_argmax(G):
    """The fused VFE node keeps the scatter-max arg-max as one byte per (pillar, channel); pillars with more than 255 points
    saturate it and the backward pass recomputes their arg-max.  A frame with pillars of ~700 and ~300 points (and ordinary
    ones): pillar features and all six VFE parameter gradients against the oracle's autograd, fp32 configuration."""
    model, cfg, ocfg, P, Bf = build(G, "tiny", 0.85, 4)
    G.config.set_precision(model, "fp32")
    model.train()
    r = np.random.RandomState(9)
    n = 1500
    base = np.concatenate([np.zeros((n, 1)), r.normal(0, 3, (n, 2)), r.uniform(-2, 4, (n, 1)), r.uniform(0, 1, (n, 2))], 1)
    crowd1 = np.concatenate([np.zeros((700, 1)), r.uniform(0.01, 0.30, (700, 2)), r.uniform(-2, 4, (700, 1)), r.uniform(0, 1, (700, 2))], 1)
    crowd2 = np.concatenate([np.zeros((300, 1)), r.uniform(0.33, 0.63, (300, 2)), r.uniform(-2, 4, (300, 1)), r.uniform(0, 1, (300, 2))], 1)
    pts = np.concatenate([base, crowd1, crowd2], 0)
    pts = torch.from_numpy(pts[r.permutation(pts.shape[0])].astype(np.float32))
    keep, opts, ocoords, ovc, oinv = O.voxelize(pts, ocfg)
    assert int(torch.bincount(oinv).max()) >= 700
    W = torch.randn(ovc.shape[0], 128, generator=torch.Generator().manual_seed(1))
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in P.items() if k.startswith("vfe.")}
    pf, _, _ = O.vfe_forward(leaves, opts, ocoords, oinv, ovc.shape[0], ocfg)
    (pf * W).sum().backward()
    bd = model.vfe(dict(points=pts.cuda(), batch_size=1))
    assert rel(bd["pillar_features"], pf) < 1e-4
    (bd["pillar_features"] * W.cuda()).sum().backward()
    for k, v in leaves.items():
        g = dict(model.named_parameters())[k].grad
        assert rel(g, v.grad) < 2e-3, (k, rel(g, v.grad))


# ------------------------------------------------------------------------------ dcn
@pytest.mark.parametrize("modulated,groups,dg,stride,dil", [(False, 1, 1, 1, 1), (True, 1, 1, 1, 1), (True, 2, 2, 2, 1), (False, 1, 4, 1, 2)])
def test_deform_conv_matches_oracle(G, modulated, groups, dg, stride, dil):
    """pcdet.ops.dcn (deform_conv / modulated_deform_conv through the reference-named deform_conv_cuda entry points) against the
    torch-CPU oracle: output and the gradients w.r.t. input, offsets, masks, weight, bias (autograd on the oracle side).
    Offsets reach outside the plane, so the zero-padding rules of the sampler are exercised.  fp32, 1e-4 of the range
    (the GEMMs run with TF32 off in this test session)."""
    from gd_mae_b200.pcdet.ops.dcn import deform_conv, modulated_deform_conv
    from oracle import dcn_oracle as DO
    g = torch.Generator().manual_seed(3 + groups + dg)
    B, C, H, W, Cout, k = 2, 8, 11, 13, 12, 3
    pad = dil
    x = torch.randn(B, C, H, W, generator=g)
    Ho, Wo = (H + 2 * pad - (dil * (k - 1) + 1)) // stride + 1, (W + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    offset = torch.randn(B, dg * 2 * k * k, Ho, Wo, generator=g) * 1.7
    mask = torch.rand(B, dg * k * k, Ho, Wo, generator=g) if modulated else None
    weight = torch.randn(Cout, C // groups, k, k, generator=g) * 0.2
    bias = torch.randn(Cout, generator=g) if modulated else None
    go = torch.randn(B, Cout, Ho, Wo, generator=g)
    leaves = [t.clone().requires_grad_(True) for t in (x, offset, weight)] + ([mask.clone().requires_grad_(True), bias.clone().requires_grad_(True)] if modulated else [])
    ref = DO.deform_conv2d(leaves[0], leaves[1], leaves[3] if modulated else None, leaves[2], leaves[4] if modulated else None, stride, pad, dil, groups, dg)
    ref.backward(go)
    cl = [t.clone().cuda().requires_grad_(True) for t in (x, offset, weight)] + ([mask.clone().cuda().requires_grad_(True), bias.clone().cuda().requires_grad_(True)] if modulated else [])
    if modulated:
        out = modulated_deform_conv(cl[0], cl[1], cl[3], cl[2], cl[4], stride, pad, dil, groups, dg)
    else:
        out = deform_conv(cl[0], cl[1], cl[2], stride, pad, dil, groups, dg)
    out.backward(go.cuda())
    assert rel(out, ref) < 1e-4
    for name, a, b in zip(("input", "offset", "weight", "mask", "bias"), cl, leaves):
        assert rel(a.grad, b.grad) < 2e-4, (name, rel(a.grad, b.grad))
    with pytest.raises(NotImplementedError):
        deform_conv(x, offset, weight, stride, pad, dil, groups, dg)               # CPU tensors: like the reference, no CPU path
