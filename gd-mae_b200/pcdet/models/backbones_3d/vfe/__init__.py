"""NAME registry as in pcdet/models/backbones_3d/vfe/__init__.py:7-12 (hot-path entries only)."""
from .dyn_vfe import DynVFE, VFETemplate

__all__ = {'VFETemplate': VFETemplate, 'DynVFE': DynVFE}
