"""NAME registry as in pcdet/models/detectors/__init__.py:10-29 (GD-MAE entries only)."""
from .centerpoint import CenterPoint
from .gd_mae import GDMAE, Detector3DTemplate

__all__ = {'Detector3DTemplate': Detector3DTemplate, 'GDMAE': GDMAE, 'CenterPoint': CenterPoint}


def build_detector(model_cfg, num_class, dataset, logger=None):
    return __all__[model_cfg.NAME](model_cfg=model_cfg, num_class=num_class, dataset=dataset, logger=logger)
