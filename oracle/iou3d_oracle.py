"""TEST INFRASTRUCTURE - ctypes front end of oracle/iou3d_oracle.c (the plain-C restatement of the reference's rotated-box
overlap / IoU / NMS) plus the numpy part of the Python API it mirrors:

  boxes_iou3d            pcdet/ops/iou3d_nms/iou3d_nms_utils.py:48-79 (height overlap x BEV overlap / union volume)
  nms / nms_normal       iou3d_nms_utils.py:82-116 (sort by score, optional pre_maxsize, sweep)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module."""
import ctypes
import os

import numpy as np

from . import build_oracle as _bo

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(_bo.build_oracle())
        L.oracle_nms.restype = ctypes.c_int
        _lib = L
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 7
    return a, a.ctypes.data_as(ctypes.c_void_p)


def boxes_overlap_bev(boxes_a, boxes_b):
    a, pa = _f(boxes_a)
    b, pb = _f(boxes_b)
    out = np.zeros((a.shape[0], b.shape[0]), dtype=np.float32)
    lib().oracle_boxes_overlap_bev(a.shape[0], pa, b.shape[0], pb, out.ctypes.data_as(ctypes.c_void_p))
    return out


def boxes_iou_bev(boxes_a, boxes_b):
    a, pa = _f(boxes_a)
    b, pb = _f(boxes_b)
    out = np.zeros((a.shape[0], b.shape[0]), dtype=np.float32)
    lib().oracle_boxes_iou_bev(a.shape[0], pa, b.shape[0], pb, out.ctypes.data_as(ctypes.c_void_p))
    return out


def boxes_iou3d(boxes_a, boxes_b):
    """iou3d_nms_utils.py:48-79 in float32 numpy"""
    a = np.asarray(boxes_a, dtype=np.float32)
    b = np.asarray(boxes_b, dtype=np.float32)
    a_max, a_min = (a[:, 2] + a[:, 5] / 2).reshape(-1, 1), (a[:, 2] - a[:, 5] / 2).reshape(-1, 1)
    b_max, b_min = (b[:, 2] + b[:, 5] / 2).reshape(1, -1), (b[:, 2] - b[:, 5] / 2).reshape(1, -1)
    overlaps_bev = boxes_overlap_bev(a, b)
    overlaps_h = np.clip(np.minimum(a_max, b_max) - np.maximum(a_min, b_min), 0, None)
    overlaps_3d = overlaps_bev * overlaps_h
    vol_a = (a[:, 3] * a[:, 4] * a[:, 5]).reshape(-1, 1)
    vol_b = (b[:, 3] * b[:, 4] * b[:, 5]).reshape(1, -1)
    return (overlaps_3d / np.clip(vol_a + vol_b - overlaps_3d, 1e-6, None)).astype(np.float32)


def _nms(boxes, scores, thresh, rotated, pre_maxsize=None):
    boxes = np.asarray(boxes, dtype=np.float32)
    order = np.argsort(-np.asarray(scores), kind="stable")
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    b, pb = _f(boxes[order])
    keep = np.zeros(max(b.shape[0], 1), dtype=np.int64)
    n = lib().oracle_nms(pb, b.shape[0], ctypes.c_float(thresh), int(rotated), keep.ctypes.data_as(ctypes.c_void_p))
    return order[keep[:n]]


def nms(boxes, scores, thresh, pre_maxsize=None):
    return _nms(boxes, scores, thresh, True, pre_maxsize)


def nms_normal(boxes, scores, thresh):
    return _nms(boxes, scores, thresh, False)


def random_boxes(n, seed, spread=20.0):
    """seeded detection-like boxes: clustered centres so that many pairs overlap, the three Waymo class shapes, any heading"""
    r = np.random.RandomState(seed)
    centres = r.uniform(-spread, spread, (max(n // 6, 1), 2))
    c = centres[r.randint(0, centres.shape[0], n)] + r.normal(0, 1.2, (n, 2))
    dims = np.array([[4.7, 2.1, 1.7], [0.9, 0.9, 1.7], [1.8, 0.8, 1.7]])[r.randint(0, 3, n)] * r.uniform(0.8, 1.2, (n, 3))
    z = r.uniform(-1, 1, (n, 1))
    heading = r.uniform(-np.pi, np.pi, (n, 1))
    return np.concatenate([c, z, dims, heading], 1).astype(np.float32)
