"""NAME registry as in pcdet/models/dense_heads/__init__.py (finetune-path entries only)."""
from .center_head import CenterHead

__all__ = {'CenterHead': CenterHead}
