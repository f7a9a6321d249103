"""ConvStack3DFFNModel configuration (ffn/training/models/convstack_3d.py:59-81).

The TensorFlow graph of the reference (`define_tf_graph`, :83-102) is replaced by the sm_100a
kernels in ffn_b200/csrc; this class only carries what `model_args` in an InferenceRequest sets:
fov_size / deltas (x, y, z), depth, features.
"""

from .. import model


class ConvStack3DFFNModel(model.FFNModel):
  dim = 3

  def __init__(self, fov_size=None, deltas=None, batch_size=None, depth: int = 9, features: int = 32,
               **kwargs):
    info = model.ModelInfo(deltas, fov_size, fov_size, fov_size)
    super().__init__(info, batch_size, **kwargs)
    self.depth = depth
    self.features = features
