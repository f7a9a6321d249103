"""Model geometry (ffn/training/model.py:25-46).  Arrays are (x, y, z) like the reference."""

import dataclasses

import numpy as np


@dataclasses.dataclass
class ModelInfo:
  """Basic geometric information about the network; all triples in (x, y, z) order."""
  deltas: np.ndarray
  pred_mask_size: np.ndarray
  input_seed_size: np.ndarray
  input_image_size: np.ndarray
  additive: bool = False

  def __post_init__(self):
    self.deltas = np.asarray(self.deltas)
    self.pred_mask_size = np.asarray(self.pred_mask_size)
    self.input_seed_size = np.asarray(self.input_seed_size)
    self.input_image_size = np.asarray(self.input_image_size)


class FFNModel:
  """Geometry + hyper-parameter holder; the network itself lives in the CUDA engine."""
  dim = None

  def __init__(self, info: ModelInfo, batch_size=None, **kwargs):
    del kwargs
    self.info = info
    self.batch_size = batch_size
