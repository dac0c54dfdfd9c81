"""pcdet.datasets.augmentor.data_augmentor on the device (the input side of the step).

Same class name, constructor signature, yaml keys and method names as the reference's ``DataAugmentor``
(pcdet/datasets/augmentor/data_augmentor.py:11-143), but ``forward`` takes the COLLATED batch - ``points (sum N, 1 + C)`` on
the device with column 0 = frame index (dataset.py:169-217), ``batch_size`` - and applies the world augmentation of all frames
with one kernel launch (csrc/augment.cu) instead of frame by frame in numpy dataset workers.  The random numbers are drawn on
the host from numpy's global generator in exactly the reference's order (per frame: flip choice per axis, rotation enable +
angle, scaling enable + factor), so a seeded run reproduces the reference's parameters.  Only the augmentations the GD-MAE
SSL configs use are implemented; the box-dependent ones (gt_sampling, local_*) belong to the finetune path and raise.
"""
from functools import partial

import numpy as np
import torch

from .... import ops as _ops


class DataAugmentor(object):
    def __init__(self, root_path, augmentor_configs, class_names, logger=None):
        self.root_path = root_path
        self.class_names = class_names
        self.logger = logger
        self.data_augmentor_queue = []
        aug_config_list = augmentor_configs.AUG_CONFIG_LIST
        for cur_cfg in aug_config_list:
            if cur_cfg.NAME in augmentor_configs.DISABLE_AUG_LIST:
                continue
            if not hasattr(self, cur_cfg.NAME):
                raise NotImplementedError(f"{cur_cfg.NAME}: not on the GD-MAE pre-train path (needs gt boxes)")
            self.data_augmentor_queue.append(getattr(self, cur_cfg.NAME)(config=cur_cfg))

    # each method only DRAWS its parameters (host, numpy stream of the reference); forward() applies them in one launch
    def random_world_flip(self, frame_params=None, config=None):
        if frame_params is None:
            return partial(self.random_world_flip, config=config)
        params = []
        for cur_axis in config['ALONG_AXIS_LIST']:
            if cur_axis not in ('x', 'y'):
                raise NotImplementedError
            enable = np.random.choice([False, True], replace=False, p=[1 - config['PROBABILITY'], config['PROBABILITY']])
            if enable:
                params.append(cur_axis)
        frame_params['random_world_flip'] = params
        return frame_params

    def random_world_rotation(self, frame_params=None, config=None):
        if frame_params is None:
            return partial(self.random_world_rotation, config=config)
        enable = np.random.choice([False, True], replace=False, p=[1 - config['PROBABILITY'], config['PROBABILITY']])
        rot_range = config['WORLD_ROT_ANGLE'] if enable else [0.0, 0.0]
        frame_params['random_world_rotation'] = np.random.uniform(rot_range[0], rot_range[1])
        return frame_params

    def random_world_scaling(self, frame_params=None, config=None):
        if frame_params is None:
            return partial(self.random_world_scaling, config=config)
        enable = np.random.choice([False, True], replace=False, p=[1 - config['PROBABILITY'], config['PROBABILITY']])
        scale_range = config['WORLD_SCALE_RANGE'] if enable else [1.0, 1.0]
        frame_params['random_world_scaling'] = np.random.uniform(scale_range[0], scale_range[1])
        return frame_params

    def forward(self, data_dict, shuffle=False):
        """data_dict: points (sum N, 1 + C) CUDA fp32 sorted by frame, batch_size.  Writes the augmented points back and
        records the drawn parameters per frame under 'transformation_3d_params' (a list of the reference's per-frame dicts).
        shuffle=True also applies DataProcessor.shuffle_points (data_processor.py:92-102) inside each frame, its permutation
        drawn right after the frame's augmentation parameters, as a dataset worker of the reference would."""
        if 'gt_boxes' in data_dict:
            raise NotImplementedError("gt_boxes: the box side of the world augmentation belongs to the finetune path")
        points, B = data_dict['points'], int(data_dict['batch_size'])
        counts = torch.bincount(points[:, 0].long(), minlength=B).tolist() if shuffle else None
        per_frame, rows, src, off = [], [], [], 0
        for b in range(B):
            fp = {}
            for cur_augmentor in self.data_augmentor_queue:
                fp = cur_augmentor(frame_params=fp)
            flips = fp.get('random_world_flip', [])
            ang, sc = float(fp.get('random_world_rotation', 0.0)), float(fp.get('random_world_scaling', 1.0))
            rows.append([float('x' in flips), float('y' in flips), np.float32(np.cos(ang)), np.float32(np.sin(ang)), np.float32(sc), 0.0])
            if shuffle:
                src.append(np.random.permutation(counts[b]) + off)
                off += counts[b]
            per_frame.append(fp)
        # parameters through a small ring of pinned staging rows: a pageable copy would block the host until everything queued
        # before it on this stream (the 30 MB H2D copy of the batch) has drained
        ring = getattr(self, '_param_ring', None)
        if ring is None or ring[0].shape[0] != B:
            ring = self._param_ring = [torch.empty((B, 6), dtype=torch.float32).pin_memory() for _ in range(8)]
            self._param_slot = 0
        stage = ring[self._param_slot % len(ring)]
        self._param_slot += 1
        stage.copy_(torch.from_numpy(np.asarray(rows, dtype=np.float32)))
        params = stage.to(points.device, non_blocking=True)
        src_index = torch.from_numpy(np.concatenate(src).astype(np.int32)).to(points.device) if shuffle and src else None
        data_dict['points'] = _ops.world_augment(points, params, src_index)
        data_dict['transformation_3d_list'] = [c.func.__name__ for c in self.data_augmentor_queue]
        data_dict['transformation_3d_params'] = per_frame
        return data_dict
