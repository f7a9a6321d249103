"""Oracle: ConvStack3DFFNModel forward pass on the CPU (test infrastructure).

Restates ffn/training/models/convstack_3d.py:26-56 (`_predict_object_mask`), :83-95
(`define_tf_graph`: concat([patches, seed], 4) -> conv stack -> `update_seed`) and
ffn/training/model.py:168-183 (`update_seed`: logits = seed + update when pred size == seed size).

TF semantics reproduced: NDHWC tensors, DHWIO kernels, stride 1, SAME zero padding,
cross-correlation (no kernel flip), BiasAdd, ReLU.  tf_slim.convolution3d defaults to
activation_fn=relu, so the `_a` convolutions are followed by ReLU and the `_b` ones
(activation_fn=None) are linear.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


class ConvStackOracle:
  """fp32 (default) or fp64 CPU evaluation of the FFN conv stack."""

  def __init__(self, weights_dhwio, biases, dtype=torch.float32, operand_round=None):
    """weights_dhwio/biases: lists in layer order conv0_a, conv0_b, ..., conv_lom (2*depth+1)."""
    assert len(weights_dhwio) == len(biases) and len(weights_dhwio) % 2 == 1
    self.depth = (len(weights_dhwio) - 1) // 2
    self.dtype = dtype
    # Optional emulation of reduced-precision conv operands ('fp16' / 'bf16'): inputs and weights
    # of every 3x3x3 convolution are rounded, accumulation stays in `dtype`.  Used to separate
    # "kernel bug" from "expected fp16 operand rounding" in the parity report.
    self.operand_round = operand_round
    self.w = []
    self.b = []
    for w, b in zip(weights_dhwio, biases):
      wt = torch.from_numpy(np.ascontiguousarray(w)).permute(4, 3, 0, 1, 2).contiguous()  # OIDHW
      wt = self._round(wt.to(torch.float32)) if w.shape[0] == 3 else wt.to(torch.float32)
      self.w.append(wt.to(dtype))
      self.b.append(torch.from_numpy(np.ascontiguousarray(b)).to(dtype))

  def _round(self, t):
    if self.operand_round == 'fp16':
      return t.to(torch.float16).to(t.dtype)
    if self.operand_round == 'bf16':
      return t.to(torch.bfloat16).to(t.dtype)
    return t

  def _conv(self, x, i, relu):
    y = F.conv3d(self._round(x), self.w[i], self.b[i], stride=1, padding=1)
    return F.relu(y) if relu else y

  @torch.no_grad()
  def update(self, seed: np.ndarray, image: np.ndarray) -> np.ndarray:
    """Returns the logit *update* (conv_lom output) for one or a batch of (Z,Y,X) patches."""
    s = torch.from_numpy(np.ascontiguousarray(seed, dtype=np.float32))
    im = torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32))
    batched = s.dim() == 4
    if not batched:
      s, im = s[None], im[None]
    # channel 0 = image patch, channel 1 = seed (convstack_3d.py:86)
    net = torch.stack([im, s], dim=1).to(self.dtype)
    net = self._conv(net, 0, relu=True)            # conv0_a
    net = self._conv(net, 1, relu=False)           # conv0_b
    for m in range(1, self.depth):                 # residual modules (convstack_3d.py:41-49)
      skip = net
      net = F.relu(net)
      net = self._conv(net, 2 * m, relu=True)
      net = self._conv(net, 2 * m + 1, relu=False)
      net = net + skip
    net = F.relu(net)
    upd = F.conv3d(net, self.w[-1], self.b[-1])    # conv_lom, 1x1x1, linear
    upd = upd[:, 0]
    out = upd if batched else upd[0]
    return out.numpy()

  @torch.no_grad()
  def logits(self, seed: np.ndarray, image: np.ndarray) -> np.ndarray:
    """logits = seed + update (model.py:176-177), returned as float32 like the TF fetch."""
    upd = self.update(seed, image)
    s = np.asarray(seed, dtype=np.float32)
    if self.dtype == torch.float32:
      return (s + upd.astype(np.float32)).astype(np.float32)
    return (s.astype(np.float64) + upd).astype(np.float32)

  def __call__(self, seed, image):
    return self.logits(seed, image)


def flops_per_step(fov_zyx, depth, features=32):
  """2*MAC per FoV step, SAME-padded taps counted."""
  v = int(np.prod(fov_zyx))
  return 2 * v * (27 * 2 * features + (2 * depth - 1) * 27 * features * features + features)
