"""`import_symbol` as used by the runner (ffn/training/import_util.py:20-23)."""

import importlib


def import_symbol(name: str, default_packages=('ffn_b200.training.models',)):
  """'convstack_3d.ConvStack3DFFNModel' -> class, searched in the default packages."""
  module_name, _, symbol = name.rpartition('.')
  errors = []
  for prefix in list(default_packages) + ['']:
    full = (prefix + '.' if prefix else '') + module_name
    try:
      return getattr(importlib.import_module(full), symbol)
    except (ImportError, AttributeError) as e:
      errors.append(repr(e))
  raise ImportError('cannot import %s (%s)' % (name, '; '.join(errors)))
