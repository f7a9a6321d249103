"""Alias of ffn_b200.inference.inference_utils."""
import sys as _sys
from ffn_b200.inference import inference_utils as _impl
_sys.modules[__name__] = _impl
