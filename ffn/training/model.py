"""Alias of ffn_b200.training.model."""
import sys as _sys
from ffn_b200.training import model as _impl
_sys.modules[__name__] = _impl
