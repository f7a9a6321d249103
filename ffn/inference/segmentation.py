"""Alias of ffn_b200.inference.segmentation."""
import sys as _sys
from ffn_b200.inference import segmentation as _impl
_sys.modules[__name__] = _impl
