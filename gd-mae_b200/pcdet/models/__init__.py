"""build_network / model_fn_decorator of pcdet/models/__init__.py:9-39."""
from collections import namedtuple

import numpy as np
import torch

from .detectors import build_detector


def build_network(model_cfg, num_class, dataset, logger=None):
    return build_detector(model_cfg=model_cfg, num_class=num_class, dataset=dataset, logger=logger)


_HOST_ONLY_KEYS = ('frame_id', 'metadata', 'calib', 'image_shape')


def load_data_to_gpu(batch_dict):
    """every numpy array of the collated batch except the host-only bookkeeping keys becomes a float32 CUDA tensor, in
    place (pcdet/models/__init__.py:16-23); pinned inputs make the copies asynchronous"""
    for key in list(batch_dict.keys()):
        val = batch_dict[key]
        if key in _HOST_ONLY_KEYS or not isinstance(val, np.ndarray):
            continue
        batch_dict[key] = torch.from_numpy(val).float().cuda(non_blocking=True)


def model_fn_decorator():
    """-> model_func(model, batch_dict) = (loss, tb_dict, disp_dict), the callable tools/train_utils hands every batch to
    (pcdet/models/__init__.py:26-39)"""
    ModelReturn = namedtuple('ModelReturn', ['loss', 'tb_dict', 'disp_dict'])

    def model_func(model, batch_dict, **kwargs):
        load_data_to_gpu(batch_dict)
        ret_dict, tb_dict, disp_dict = model(batch_dict)
        owner = model if hasattr(model, 'update_global_step') else model.module       # DDP wraps the detector
        owner.update_global_step()
        return ModelReturn(ret_dict['loss'].mean(), tb_dict, disp_dict)

    return model_func
