"""`Runner`: model + volume + executor lifecycle for dense inference over subvolumes.

Keeps the API of ffn.inference.runner.Runner (ffn/inference/runner.py:57-544): `start(request,
batch_size=1, session=None)` (:165-216), `run(corner, subvol_size, reset_counters=True)` (:484-544),
`make_canvas` (:307-414), `make_restrictor` (:218-305), `get_seed_policy` (:416-431),
`save_segmentation` (:433-482), `stop_executor` (:86-96) and the attributes `.counters`,
`.request`, `.executor`, `.canvases`.

Differences, all on the "documented API actually runs" side (SURVEY.md 8b "bitrot"):
`partial_segment_iters` defaults to 0 when there is no checkpoint; canvases keep probability
maps so the `.prob` file the README promises can be written; there is no TPUExecutor branch;
`--require_gpu` is implied (the engine has no CPU path); the TensorFlow session argument is
accepted and ignored.
"""

import copy
import functools
import json
import logging
import os

import numpy as np

from .. import _lib
from ..training import import_util
from ..utils import bounding_box
from . import align
from . import executor
from . import inference
from . import inference_pb2
from . import inference_utils
from . import movement
from . import seed
from . import storage
from .inference_utils import timer_counter


class Runner:
  """Helper for managing FFN inference runs."""

  ALL_MASKED = 1

  def __init__(self, device=0, compute_mode=None):
    self.counters = inference_utils.Counters()
    self.executor = None
    self._exec_interface = executor.ExecutorInterface()
    self.canvases = {}
    self._device = device
    if compute_mode is None:
      compute_mode = {'fp32': _lib.COMPUTE_FP32, 'fp16': _lib.COMPUTE_FP16_TC, 'x2': _lib.COMPUTE_FP16X2_TC}[
          os.environ.get('FFN_B200_COMPUTE', 'fp16')]
    self._compute_mode = compute_mode
    self.init_seg_volume = None
    self._image_volume = None
    self._mask_volumes = {}
    self._shift_mask_volume = None
    self._aligner = align.Aligner()
    self._model_info = None
    self.request = None

  def __del__(self):
    try:
      self.stop_executor()
    except Exception:  # pylint: disable=broad-except
      pass

  def stop_executor(self):
    if self.executor is not None:
      try:
        self.executor.close()
      except executor.TerminationException:
        pass
      self.executor = None

  def _init_model(self, request, batch_size):
    model_class = import_util.import_symbol(request.model_name)
    args = json.loads(request.model_args) if request.model_args else {}
    args['batch_size'] = batch_size
    model = model_class(**args)
    self._model_info = model.info
    self.executor = executor.B200Executor(
        self._exec_interface, model, self.counters, batch_size,
        checkpoint_path=request.model_checkpoint_path, device=self._device,
        compute_mode=self._compute_mode)

  def start(self, request, batch_size=1, session=None):
    """Opens input volumes and initialises the engine."""
    del session
    request = copy.deepcopy(request)
    self.request = request
    assert self.request.segmentation_output_dir
    os.makedirs(request.segmentation_output_dir, exist_ok=True)
    self.stop_executor()
    self._init_model(request, batch_size)
    with timer_counter(self.counters, 'volstore-open'):
      self._image_volume = storage.decorated_volume(request.image)
      if request.HasField('init_segmentation'):
        self.init_seg_volume = storage.decorated_volume(request.init_segmentation)
      else:
        self.init_seg_volume = None
      if request.shift_mask.WhichOneof('volume_path') is not None:
        self._shift_mask_volume = storage.decorated_volume(request.shift_mask)
      else:
        self._shift_mask_volume = None
      self._mask_volumes = {}
      opts = request.alignment_options
      if opts.type != inference_pb2.AlignmentOptions.NO_ALIGNMENT:
        raise NotImplementedError('Alignment for type %s is not implemented' %
                                  inference_pb2.AlignmentOptions.AlignType.Name(opts.type))
      self._aligner = align.Aligner()
    self.executor.start_server()

  def make_restrictor(self, corner, subvol_size, image, alignment):
    """Builds a MovementRestrictor (None when the request has no masks)."""
    kwargs = {}
    if self.request.masks:
      with timer_counter(self.counters, 'load-mask'):
        final_mask = storage.build_mask(self.request.masks, corner, subvol_size, self._mask_volumes,
                                        image, alignment)
        if np.all(final_mask):
          logging.info('Everything masked.')
          return self.ALL_MASKED
        kwargs['mask'] = final_mask
    if self.request.seed_masks:
      with timer_counter(self.counters, 'load-seed-mask'):
        seed_mask = storage.build_mask(self.request.seed_masks, corner, subvol_size, self._mask_volumes,
                                       image, alignment)
        if np.all(seed_mask):
          logging.info('All seeds masked.')
          return self.ALL_MASKED
        kwargs['seed_mask'] = seed_mask
    if self._shift_mask_volume is not None:
      # runner.py:256-300: the 2-d shift field of the subvolume, y / x at `shift_mask_scale` resolution
      with timer_counter(self.counters, 'load-shift-mask'):
        s = self.request.shift_mask_scale or 1   # unset in the request: full resolution
        scale = np.array((1, s, s))
        shift_corner = np.array(corner) // scale
        shift_size = -(-np.array(subvol_size) // scale)
        shift_alignment = alignment.rescaled(np.array((1.0, 1.0, 1.0)) / scale)
        src_corner, src_size = shift_alignment.expand_bounds(shift_corner, shift_size, forward=False)
        src_corner, src_size = storage.clip_subvolume_to_bounds(src_corner, src_size, self._shift_mask_volume)
        src_end = src_corner + src_size
        expanded = np.asarray(self._shift_mask_volume[0:2, src_corner[0]:src_end[0], src_corner[1]:src_end[1],
                                                      src_corner[2]:src_end[2]])
        shift_mask = np.array([shift_alignment.align_and_crop(src_corner, expanded[i], shift_corner, shift_size)
                               for i in range(2)])
        shift_mask = alignment.transform_shift_mask(corner, s, shift_mask)
        if self.request.HasField('shift_mask_fov'):
          fov = bounding_box.BoundingBox(start=self.request.shift_mask_fov.start, size=self.request.shift_mask_fov.size)
        else:
          diameter = np.array(self._model_info.input_image_size)
          fov = bounding_box.BoundingBox(start=-(diameter // 2), size=diameter)
        kwargs.update(shift_mask=shift_mask, shift_mask_fov=fov, shift_mask_scale=s,
                      shift_mask_threshold=self.request.shift_mask_threshold)
    return movement.MovementRestrictor(**kwargs) if kwargs else None

  def make_canvas(self, corner, subvol_size, **canvas_kwargs):
    """Builds the Canvas for a subvolume: returns (Canvas, Alignment)."""
    subvol_counters = self.counters.get_sub_counters()
    with timer_counter(subvol_counters, 'load-image'):
      logging.info('Process subvolume: %r', corner)
      alignment = self._aligner.generate_alignment(corner, subvol_size)
      dst_corner, dst_size = alignment.expand_bounds(corner, subvol_size, forward=True)
      src_corner, src_size = alignment.expand_bounds(dst_corner, dst_size, forward=False)
      src_corner, src_size = storage.clip_subvolume_to_bounds(src_corner, src_size, self._image_volume)
      src_end = np.asarray(src_corner) + np.asarray(src_size)
      sel = tuple(slice(int(a), int(b)) for a, b in zip(src_corner, src_end))
      volume = self._image_volume
      data = volume[(slice(0, 1),) + sel] if volume.ndim == 4 else volume[sel]
      data = np.asarray(data)
      if data.ndim == 4:
        data = data.squeeze(axis=0)
      image = alignment.align_and_crop(src_corner, data, dst_corner, dst_size, forward=True)
      logging.info('Image data loaded, shape: %r.', image.shape)

    restrictor = self.make_restrictor(dst_corner, dst_size, image, alignment)
    if restrictor == self.ALL_MASKED:
      return None, None

    exc = self.executor
    if exc is None:
      raise executor.TerminationException
    canvas_kwargs.setdefault('keep_probability_maps', True)
    if image.dtype == np.uint8:
      # normalisation (runner.py:383-385) happens on the device from the uint8 volume
      canvas_kwargs.update(image_mean=self.request.image_mean, image_stddev=self.request.image_stddev)
      canvas_image = image
    else:
      canvas_image = (image.astype(np.float32) - self.request.image_mean) / self.request.image_stddev
    canvas = inference.Canvas(
        self._model_info, exc.get_client(subvol_counters), canvas_image, self.request.inference_options,
        counters=subvol_counters, restrictor=restrictor,
        movement_policy_fn=movement.get_policy_fn(self.request, self._model_info),
        checkpoint_path=storage.checkpoint_path(self.request.segmentation_output_dir, corner),
        checkpoint_interval_sec=self.request.checkpoint_interval, corner_zyx=dst_corner, **canvas_kwargs)

    if self.request.HasField('init_segmentation'):
      canvas.init_segmentation_from_volume(
          self.init_seg_volume, src_corner, src_end,
          lambda im: alignment.align_and_crop(src_corner, im, dst_corner, dst_size, forward=True))
    return canvas, alignment

  def get_seed_policy(self, corner, subvol_size):
    policy_cls = getattr(seed, self.request.seed_policy)
    kwargs = {'corner': corner, 'subvol_size': subvol_size}
    if self.request.seed_policy_args:
      kwargs.update(json.loads(self.request.seed_policy_args))
    return functools.partial(policy_cls, **kwargs)

  def save_segmentation(self, canvas, alignment, target_path, prob_path):
    """Writes seg-*.npz (segmentation, origins, request, counters, overlaps) and seg-*.prob."""
    del alignment   # identity
    seg = np.asarray(canvas.segmentation)
    seg[seg < 0] = 0                                     # remove markers (runner.py:462)
    storage.save_subvolume(seg, dict(canvas.origins), target_path,
                           request=self.request.SerializeToString(), counters=canvas.counters.dumps(),
                           overlaps=canvas.overlaps)
    if canvas.seg_prob is not None:
      with storage.atomic_file(prob_path) as fd:
        storage.savez_deflate(fd, qprob=np.asarray(canvas.seg_prob))

  def run(self, corner, subvol_size, reset_counters=True):
    """Runs FFN inference over a subvolume; returns the Canvas (None if already done / masked)."""
    if reset_counters:
      self.counters.reset()
    out_dir = self.request.segmentation_output_dir
    seg_path = storage.segmentation_path(out_dir, corner)
    prob_path = storage.object_prob_path(out_dir, corner)
    cpoint_path = storage.checkpoint_path(out_dir, corner)
    if os.path.exists(seg_path):
      return None
    canvas, alignment = self.make_canvas(corner, subvol_size)
    if canvas is None:
      return None
    partial_segment_iters = 0
    if os.path.exists(cpoint_path):
      partial_segment_iters = canvas.restore_checkpoint(cpoint_path)
    if self.request.alignment_options.save_raw:
      image_path = storage.subvolume_path(out_dir, corner, 'align')
      with storage.atomic_file(image_path) as fd:
        np.savez_compressed(fd, im=canvas.image)
    self.canvases[corner] = canvas
    canvas.segment_all(seed_policy=self.get_seed_policy(corner, subvol_size),
                       partial_segment_iters=partial_segment_iters)
    self.save_segmentation(canvas, alignment, seg_path, prob_path)
    del self.canvases[corner]
    try:
      os.remove(cpoint_path)
    except OSError:
      pass
    return canvas
