"""Config plumbing for the path: attribute dicts compatible with the reference's EasyDict usage
(``cfg.KEY``, ``cfg.get('KEY', default)``), a loader for the reference's yaml files
(pcdet/config.py:51-85, incl. ``_BASE_CONFIG_``) and built-in copies of the hyper-parameters
of the named configs (tools/cfgs/*/gd_mae*.yaml) for runs without the reference checkout."""
import copy
import os

import numpy as np


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def to_attr(d):
    if isinstance(d, dict):
        return AttrDict({k: to_attr(v) for k, v in d.items()})
    if isinstance(d, list):
        return [to_attr(v) for v in d]
    return d


def cfg_from_yaml_file(path, root=None):
    """pcdet/config.py:71-85 (merge ``_BASE_CONFIG_`` recursively, relative to ``root`` = tools/)."""
    import yaml
    root = root or os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(path))))

    def merge(new, cfg):
        if "_BASE_CONFIG_" in new:
            with open(os.path.join(root, new["_BASE_CONFIG_"])) as f:
                cfg.update(to_attr(yaml.safe_load(f)))
        for k, v in new.items():
            if isinstance(v, dict):
                cfg.setdefault(k, AttrDict())
                merge(v, cfg[k])
            else:
                cfg[k] = to_attr(v)
        return cfg

    with open(path) as f:
        return merge(yaml.safe_load(f), AttrDict())


def _drop_info():
    lv = {"0": {"max_tokens": 16, "drop_range": [0, 16]}, "1": {"max_tokens": 32, "drop_range": [16, 32]},
          "2": {"max_tokens": 64, "drop_range": [32, 100000]}}
    return {"train": copy.deepcopy(lv), "test": copy.deepcopy(lv)}


def _sst_block(name, d, dff, stride):
    return {"NAME": name,
            "PREPROCESS": {"WINDOW_SHAPE": [8, 8, 1], "DROP_INFO": _drop_info(), "SHUFFLE_VOXELS": False,
                           "POS_TEMPERATURE": 1000, "NORMALIZE_POS": False},
            "ENCODER": {"NUM_BLOCKS": 2, "STRIDE": stride, "D_MODEL": d, "NHEAD": 8, "DIM_FEEDFORWARD": dff, "DROPOUT": 0.0,
                        "ACTIVATION": "gelu", "LAYER_CFG": {"cosine": True, "tau_min": 0.01}}}


def _finetune_heads():
    """BACKBONE_2D + DENSE_HEAD of tools/cfgs/waymo_models/gd_mae_iou.yaml:199-254 (CenterPoint detector)"""
    conv = lambda dil: {"out_channels": 128, "kernel_size": 3, "dilation": dil, "padding": dil, "stride": 1}  # noqa: E731
    backbone_2d = {"NAME": "SSTBEVBackbone", "NUM_FILTER": 128, "CONV_KWARGS": [conv(1), conv(1), conv(2), conv(1)],
                   "CONV_SHORTCUT": [0, 1, 2]}
    head = {"out_channels": None, "num_conv": 2}
    dense_head = {
        "NAME": "CenterHead", "CLASS_AGNOSTIC": False, "CLASS_NAMES_EACH_HEAD": [["Vehicle", "Pedestrian", "Cyclist"]],
        "SHARED_CONV_CHANNEL": 64, "USE_BIAS_BEFORE_NORM": True, "NUM_HM_CONV": 2,
        "SEPARATE_HEAD_CFG": {"HEAD_ORDER": ["center", "center_z", "dim", "rot"],
                              "HEAD_DICT": {"center": dict(head, out_channels=2), "center_z": dict(head, out_channels=1),
                                            "dim": dict(head, out_channels=3), "rot": dict(head, out_channels=2),
                                            "iou": dict(head, out_channels=1)}},
        "TARGET_ASSIGNER_CONFIG": {"FEATURE_MAP_STRIDE": 1, "NUM_MAX_OBJS": 500, "GAUSSIAN_OVERLAP": 0.1, "MIN_RADIUS": 2},
        "LOSS_CONFIG": {"LOSS_WEIGHTS": {"cls_weight": 1.0, "iou_weight": 1.0, "loc_weight": 2.0, "code_weights": [1.0] * 8}},
        "POST_PROCESSING": {"SCORE_THRESH": 0.1, "POST_CENTER_LIMIT_RANGE": [-75.2, -75.2, -2, 75.2, 75.2, 4], "MAX_OBJ_PER_SAMPLE": 500,
                            "NMS_CONFIG": {"NMS_TYPE": "multi_class_nms", "NMS_THRESH": [0.8, 0.55, 0.55],
                                           "NMS_PRE_MAXSIZE": [2048, 1024, 1024], "NMS_POST_MAXSIZE": [200, 150, 150],
                                           "IOU_RECTIFIER": [0.5, 0.71, 0.65]}}}
    return backbone_2d, dense_head


def builtin_cfg(name="waymo_ssl"):
    """Model + data hyper-parameters of the named configs, restated (not copied) from
    tools/cfgs/waymo_models/gd_mae_ssl.yaml, tools/cfgs/once_models/gd_mae_ssl.yaml and
    tools/cfgs/kitti_models/gd_mae.yaml. ``tiny`` is a 40x48-pillar grid for tests.  ``waymo_iou`` / ``tiny_iou``: the
    finetune detector of tools/cfgs/waymo_models/gd_mae_iou.yaml (CenterPoint: SPTBackbone + SSTBEVBackbone + CenterHead)."""
    finetune = name.endswith("_iou")
    if finetune:
        name = {"waymo_iou": "waymo_ssl", "tiny_iou": "tiny"}[name]
    data = {"waymo_ssl": ([-74.88, -74.88, -2.0, 74.88, 74.88, 4.0], [0.32, 0.32, 6.0], 5),
            "once_ssl": ([-74.88, -74.88, -5.0, 74.88, 74.88, 3.0], [0.32, 0.32, 8.0], 4),
            "kitti": ([0.0, -39.68, -3.0, 69.12, 39.68, 1.0], [0.32, 0.32, 4.0], 4),
            "tiny": ([-6.4, -7.68, -2.0, 6.4, 7.68, 4.0], [0.32, 0.32, 6.0], 5)}[name]
    pc_range = np.array(data[0], dtype=np.float32)
    voxel = data[1]
    grid = np.round((pc_range[3:6] - pc_range[0:3]) / np.array(voxel)).astype(np.int64)
    model = {"NAME": "GDMAE",
             "VFE": {"NAME": "DynVFE", "TYPE": "mean", "WITH_DISTANCE": False, "USE_ABSLOTE_XYZ": True,
                     "USE_CLUSTER_XYZ": True, "MLPS": [[64, 128]]},
             "BACKBONE_3D": {"NAME": "SPTBackboneMAE",
                             "SST_BLOCK_LIST": [_sst_block("sst_block_x1", 128, 256, 1), _sst_block("sst_block_x2", 256, 512, 2),
                                                _sst_block("sst_block_x4", 256, 512, 2)],
                             "MASK_CONFIG": {"RATIO": 0.85, "NUM_PRD_POINTS": 16, "NUM_GT_POINTS": 64},
                             "FEATURES_SOURCE": ["x_conv1", "x_conv2", "x_conv3"],
                             "FUSE_LAYER": {"x_conv1": {"UPSAMPLE_STRIDE": 1, "NUM_FILTER": 128, "NUM_UPSAMPLE_FILTER": 128},
                                            "x_conv2": {"UPSAMPLE_STRIDE": 2, "NUM_FILTER": 256, "NUM_UPSAMPLE_FILTER": 128},
                                            "x_conv3": {"UPSAMPLE_STRIDE": 4, "NUM_FILTER": 256, "NUM_UPSAMPLE_FILTER": 128}}}}
    if finetune:
        model["NAME"] = "CenterPoint"
        model["BACKBONE_3D"]["NAME"] = "SPTBackbone"
        del model["BACKBONE_3D"]["MASK_CONFIG"]
        model["BACKBONE_2D"], model["DENSE_HEAD"] = _finetune_heads()
        model["POST_PROCESSING"] = {"RECALL_THRESH_LIST": [0.3, 0.5, 0.7], "EVAL_METRIC": "waymo_custom"}
    optim = {"BATCH_SIZE_PER_GPU": 8, "NUM_EPOCHS": 30, "OPTIMIZER": "adam_onecycle", "LR": 0.003, "WEIGHT_DECAY": 0.01,
             "MOMENTUM": 0.9, "MOMS": [0.95, 0.85], "PCT_START": 0.4, "DIV_FACTOR": 10, "GRAD_NORM_CLIP": 10}
    return to_attr({"MODEL": model, "OPTIMIZATION": optim, "POINT_CLOUD_RANGE": pc_range, "VOXEL_SIZE": voxel,
                    "GRID_SIZE": grid, "NUM_POINT_FEATURES": data[2]})


class SyntheticDataset:
    """The attributes Detector3DTemplate.build_networks reads from the dataset object
    (detector3d_template.py:45-58, dataset.py:27-41)."""

    def __init__(self, cfg, class_names=('Vehicle', 'Pedestrian', 'Cyclist')):
        self.class_names = list(class_names)
        self.point_feature_encoder = AttrDict(num_point_features=int(cfg.NUM_POINT_FEATURES))
        self.grid_size = np.asarray(cfg.GRID_SIZE)
        self.point_cloud_range = np.asarray(cfg.POINT_CLOUD_RANGE, dtype=np.float32)
        self.voxel_size = list(cfg.VOXEL_SIZE)


def build_mae_model(cfg):
    """The detector named by cfg.MODEL.NAME (GDMAE, CenterPoint) built the way tools/train.py does
    (build_network(model_cfg, num_class, dataset))."""
    from .pcdet.models import build_network
    ds = SyntheticDataset(cfg)
    return build_network(cfg.MODEL, len(ds.class_names), ds)


def set_precision(model, dtype, matmul="high", gemm_bf16=True, dense_spatial_features=None):
    """One switch for the numeric configuration of the step.
    'fp32' : parity configuration - fp32 everywhere, TF32 off, fp32 SIMT attention.
    'tf32' : TF32 GEMMs/conv, fp32 SIMT attention, fp32 decoder map.
    'bf16' : the BASELINE.json configuration - bf16 GEMM operands written by the producing kernels, bf16 q/k/v and
             tensor-core SRA kernels, bf16 row-kernel inputs (a / h / f / dgl), dense decoder map / 3x3 conv in bf16
             (the residual stream, BatchNorm / LayerNorm / softmax statistics, losses, gradients of the parameters
             and the optimizer stay fp32)."""
    import torch
    from . import ops
    assert dtype in ("fp32", "tf32", "bf16")
    fast = dtype != "fp32"
    torch.backends.cudnn.allow_tf32 = fast
    torch.backends.cuda.matmul.allow_tf32 = fast
    torch.set_float32_matmul_precision(matmul if fast else "highest")
    from . import fused
    # bf16 configuration: q/k/v leave the in-projection GEMM as bf16 and the SRA forward/backward run on the tensor
    # cores (csrc/sra_attention_tc.cu); fp32/tf32 keep the fp32 SIMT kernels
    ops.SRA_TENSOR_CORES = dtype == "bf16" and gemm_bf16
    # bf16 GEMM operands are written by the hand-written kernels themselves (no cast passes) and the GEMMs go
    # through gdmae_gemm (cublasGemmEx bf16 x bf16 -> fp32); torch.mm(out_dtype=fp32) was measured to triple the
    # host time per call (cublasLt path) and is not used.
    fused.GEMM_DTYPE = torch.bfloat16 if (dtype == "bf16" and gemm_bf16) else torch.float32
    model.backbone_3d.decoder_dtype = torch.bfloat16 if dtype == "bf16" else torch.float32
    if not hasattr(model.backbone_3d, "dense_spatial_features"):
        return model                       # finetune backbone (SPTBackbone): the dense map is always materialised
    # batch_dict['spatial_features'] (the dense BN+ReLU map the reference always writes) is a contract of its own, not a
    # precision matter: it stays on unless the caller opts out explicitly (the MAE pre-train step reads the map only at
    # the pillar cells, so bench.py / MAE training pass dense_spatial_features=False)
    if dense_spatial_features is not None:
        model.backbone_3d.dense_spatial_features = bool(dense_spatial_features)
    return model
