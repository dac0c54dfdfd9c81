"""NAME registry as in pcdet/models/backbones_3d/__init__.py:8-16 (GD-MAE entries only)."""
from .spt_backbone import SPTBackbone
from .spt_backbone_mae import SPTBackboneMAE

__all__ = {'SPTBackboneMAE': SPTBackboneMAE, 'SPTBackbone': SPTBackbone}
