"""gd-mae_b200: B200-native (sm_100a) implementation of GD-MAE's MAE pre-train hot path
(pillar->BEV + Sparse Pyramid Transformer + generative decoder + chamfer head) behind the
reference's pcdet plugin/operator API.  Import as ``gd_mae_b200`` (see gd_mae_b200.py).

Layout: csrc/ CUDA kernels + C ABI (include/gdmae_b200.h), _lib.py ctypes loader, ops.py
operators + autograd, pcdet/ the host-side mirror of the reference's module tree for this
path (same class names, constructor signatures, batch_dict and state_dict keys).
"""
from . import _lib  # noqa: F401
from .build import build as build_library  # noqa: F401

__all__ = ["_lib", "build_library"]
