"""Alias of ffn_b200.training.models.convstack_3d."""
import sys as _sys
from ffn_b200.training.models import convstack_3d as _impl
_sys.modules[__name__] = _impl
