"""Checkpoint interop: reference-format .pth files, the spconv weight-layout adaption of
Detector3DTemplate._load_state_dict and the reference's optimizer_state layout.  CPU only (host logic)."""
import os
import sys

import numpy as np
import pytest
import torch

import gd_mae_b200  # noqa: F401
from gd_mae_b200 import config, train_utils
from gd_mae_b200.trainer import MAETrainer, reference_param_groups
from oracle import gdmae_oracle as O

REF = "/root/reference"


def _model(seed):
    cfg = config.builtin_cfg("tiny")
    torch.manual_seed(seed)
    return config.build_mae_model(cfg), cfg


def test_checkpoint_file_roundtrip_and_partial_load(tmp_path):
    """save_checkpoint/checkpoint_state -> load_params_from_file (key + shape match only, the SSL -> finetune transfer rule)
    and load_params_with_optimizer (strict) restore every tensor; a foreign / mis-shaped key is skipped, not fatal."""
    m1, cfg = _model(1)
    tr1 = MAETrainer(m1, cfg.OPTIMIZATION, total_steps=10)
    tr1.t = tr1.it = 3
    tr1.exp_avg.normal_()
    tr1.exp_avg_sq.uniform_(0, 1)
    path = train_utils.save_checkpoint(train_utils.checkpoint_state(m1, tr1, epoch=2, it=3), str(tmp_path / "ckpt"))
    ck = torch.load(path, weights_only=False)
    assert set(ck) == {"epoch", "it", "model_state", "optimizer_state", "version"}
    assert set(ck["model_state"]) == set(m1.state_dict()) and "global_step" in ck["model_state"]
    m2, _ = _model(2)
    tr2 = MAETrainer(m2, cfg.OPTIMIZATION, total_steps=10)
    it, epoch = m2.load_params_with_optimizer(path, to_cpu=True, optimizer=tr2)
    assert (it, epoch) == (3, 2) and (tr2.t, tr2.it) == (3, 3)
    for k, v in m1.state_dict().items():
        assert torch.equal(v, m2.state_dict()[k]), k
    for n, (off, k) in tr1.slices.items():                  # (the zero padding between bucket entries is not part of the state)
        if off < tr1.n_opt:
            assert torch.equal(tr1.exp_avg[off:off + k], tr2.exp_avg[off:off + k]), n
            assert torch.equal(tr1.exp_avg_sq[off:off + k], tr2.exp_avg_sq[off:off + k]), n
    for n, p in m2.named_parameters():                      # parameters are still views of the trainer's flat bucket
        off, k = tr2.slices[n]
        assert p.data_ptr() == tr2.flat_params[off:off + k].data_ptr()
    # partial load: drop the decoder, add a head key that does not exist here, mis-shape one tensor
    ms = {k: v for k, v in ck["model_state"].items() if not k.startswith("backbone_3d.decoder")}
    ms["dense_head.heads_list.0.hm.weight"] = torch.zeros(3)
    ms["vfe.dvfe_mlps.0.0.weight"] = torch.zeros(5, 5)
    torch.save({"model_state": ms}, str(tmp_path / "partial.pth"))
    m3, _ = _model(3)
    before = {k: v.clone() for k, v in m3.state_dict().items()}
    n_loaded, n_total = m3.load_params_from_file(str(tmp_path / "partial.pth"), to_cpu=True)
    assert n_total == len(before) and n_loaded == len([k for k in ms if k in before]) - 1
    for k, v in m3.state_dict().items():
        untouched = k.startswith("backbone_3d.decoder") or k == "vfe.dvfe_mlps.0.0.weight"
        assert torch.equal(v, before[k] if untouched else ck["model_state"][k]), k


def test_spconv_weight_layout_adaption():
    """detector3d_template.py:361-380: sparse-conv weights of another spconv layout are permuted into this model's
    (C_out, kH, kW, C_in); everything else must match by shape or is skipped."""
    m, _ = _model(4)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    key_down, key_subm = "backbone_3d.sst_blocks.1.conv_down.0.weight", "backbone_3d.sst_blocks.0.conv_out.0.weight"
    from gd_mae_b200.pcdet.utils.spconv_utils import find_all_spconv_keys
    keys = find_all_spconv_keys(m)
    assert key_down in keys and key_subm in keys and len(keys) == 5 and all(k.endswith(".weight") for k in keys)
    disk = dict(sd)
    disk[key_down] = sd[key_down].permute(1, 2, 3, 0).contiguous()       # spconv 1.x: (kH, kW, C_in, C_out), C_in != C_out
    disk[key_subm] = sd[key_subm].transpose(-1, -2).contiguous()         # square case is ambiguous by shape: stays as stored
    m2, _ = _model(5)
    _, updated = m2._load_state_dict(disk, strict=True)
    assert torch.equal(m2.state_dict()[key_down], sd[key_down])
    assert len(updated) == len(sd)


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout (build container only)")
def test_optimizer_state_interop_with_reference_optimizer():
    """The reference's OptimWrapper(torch Adam) after two iterations -> MAETrainer.load_state_dict -> state_dict(reference_format)
    -> a fresh reference optimizer: moments land on the same parameters both ways, param-group split (non-BN | BN) and index
    order included."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import ref_harness as RH
    sys.path.insert(0, REF + "/tools")
    import train_utils.optimization as RO
    ocfg = O.make_cfg("tiny")
    mcfg = RH.load_model_cfg("tools/cfgs/waymo_models/gd_mae_ssl.yaml")
    grid = np.array(ocfg["grid"], dtype=np.int64)
    ref_model = RH.RefMAE(mcfg.MODEL, ocfg["n_feat"], ocfg["voxel"], np.array(ocfg["pc_range"], dtype=np.float32), grid)
    opt = RO.build_optimizer(ref_model, mcfg.OPTIMIZATION)
    sched, _ = RO.build_scheduler(opt, total_iters_each_epoch=12, total_epochs=1, last_epoch=-1, optim_cfg=mcfg.OPTIMIZATION)
    g = torch.Generator().manual_seed(0)
    for it in range(2):
        sched.step(it)
        for p in ref_model.parameters():
            p.grad = torch.randn(p.shape, generator=g) * 1e-3
        opt.step()
    ref_sd = opt.state_dict()
    mine, cfg = _model(6)
    names = reference_param_groups(mine)
    ref_names = dict((id(p), n) for n, p in ref_model.named_parameters())
    assert [[ref_names[id(p)] for p in g["params"]] for g in opt.opt.param_groups] == names
    tr = MAETrainer(mine, cfg.OPTIMIZATION, total_steps=12)
    tr.load_state_dict(ref_sd)
    assert tr.t == 2
    for n, p in ref_model.named_parameters():
        off, k = tr.slices[n]
        if off < tr.n_opt:
            assert torch.equal(tr.exp_avg[off:off + k], opt.opt.state[p]["exp_avg"].reshape(-1)), n
            assert torch.equal(tr.exp_avg_sq[off:off + k], opt.opt.state[p]["exp_avg_sq"].reshape(-1)), n
    opt2 = RO.build_optimizer(ref_model, mcfg.OPTIMIZATION)
    opt2.load_state_dict(tr.state_dict(reference_format=True))
    for p in ref_model.parameters():
        if p in opt.opt.state:
            assert torch.equal(opt2.opt.state[p]["exp_avg"], opt.opt.state[p]["exp_avg"])
            assert float(opt2.opt.state[p]["step"]) == 2.0


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout (build container only)")
@pytest.mark.parametrize("yaml_rel,builtin", [("tools/cfgs/waymo_models/gd_mae_ssl.yaml", "waymo_ssl"), ("tools/cfgs/once_models/gd_mae_ssl.yaml", "once_ssl"),
                                              ("tools/cfgs/kitti_models/gd_mae.yaml", "kitti"), ("tools/cfgs/waymo_models/gd_mae_iou.yaml", "waymo_iou")])
def test_yaml_loader_and_builtin_configs_against_reference_yamls(yaml_rel, builtin):
    """config.cfg_from_yaml_file (pcdet/config.py:71-85 incl. _BASE_CONFIG_ merging) on the reference's own yaml files, and the
    restated built-in configs against what the files say: every MODEL hyper-parameter the built-in carries, the point-cloud
    range, voxel size and the optimizer block.  (The yaml files are read where they lie; none is copied into the repo.)"""
    cfg = config.cfg_from_yaml_file(os.path.join(REF, yaml_rel), root=os.path.join(REF, "tools"))
    mine = config.builtin_cfg(builtin)
    assert "DATA_CONFIG" in cfg and "POINT_CLOUD_RANGE" in cfg.DATA_CONFIG                      # merged from the _BASE_CONFIG_ file
    assert [float(np.float32(v)) for v in cfg.DATA_CONFIG.POINT_CLOUD_RANGE] == [float(v) for v in mine.POINT_CLOUD_RANGE]   # stored as float32
    vox = [p for p in cfg.DATA_CONFIG.DATA_PROCESSOR if "VOXEL_SIZE" in p][0].VOXEL_SIZE       # calculate_grid_size / transform_points_to_voxels
    assert [float(v) for v in vox] == [float(v) for v in mine.VOXEL_SIZE]

    def covered(a, b, path):
        """every key the built-in carries must exist in the yaml with the same value"""
        if isinstance(a, dict):
            for k, v in a.items():
                assert k in b, path + "." + k
                covered(v, b[k], path + "." + k)
        elif isinstance(a, (list, tuple)):
            assert len(a) == len(b), path
            for i, (x, y) in enumerate(zip(a, b)):
                covered(x, y, f"{path}[{i}]")
        elif isinstance(a, float) or isinstance(b, float):
            assert abs(float(a) - float(b)) <= 1e-9 * max(1.0, abs(float(b))), (path, a, b)
        else:
            assert a == b, (path, a, b)

    if builtin == "kitti":          # the KITTI yaml is the detection config: only its VFE and first SST block are on the C1 plumbing path
        covered(mine.MODEL.VFE, cfg.MODEL.VFE, "MODEL.VFE")
        covered(mine.MODEL.BACKBONE_3D.SST_BLOCK_LIST[0], cfg.MODEL.BACKBONE_3D.SST_BLOCK_LIST[0], "SST_BLOCK_LIST[0]")
    else:
        covered(mine.MODEL, cfg.MODEL, "MODEL")
        covered({k: v for k, v in mine.OPTIMIZATION.items() if k in cfg.OPTIMIZATION}, cfg.OPTIMIZATION, "OPTIMIZATION")
        assert mine.OPTIMIZATION.OPTIMIZER == cfg.OPTIMIZATION.OPTIMIZER == "adam_onecycle"
