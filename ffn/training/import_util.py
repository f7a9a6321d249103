"""Alias of ffn_b200.training.import_util."""
import sys as _sys
from ffn_b200.training import import_util as _impl
_sys.modules[__name__] = _impl
