"""NAME registry as in pcdet/models/backbones_2d/__init__.py (finetune-path entries only)."""
from .sst_bev_backbone import SSTBEVBackbone

__all__ = {'SSTBEVBackbone': SSTBEVBackbone}
