"""`Canvas`: the flood-fill state of one subvolume, resident in HBM.

Keeps the public surface of ffn.inference.inference.Canvas (ffn/inference/inference.py:129-843):
constructor signature (:137-153), `segment_at` (:460-533), `segment_all` (:538-683), `update_at`
(:386-441), `predict` (:356-384), `is_valid_pos` (:312-346), `init_seed` (:443-450),
`reset_state` (:291-310), `get_next_segment_id`, `local_id`, `init_segmentation_from_volume`
(:685-726), `restore_checkpoint` / `save_checkpoint` / `_maybe_save_checkpoint` (:728-843) and the
attributes `.seed`, `.segmentation`, `.seg_prob`, `.origins`, `.overlaps`, `.counters`, `.shape`,
`.margin`, `.image`, `.options` (logit space), `.movement_policy`, `.restrictor`.

What changed underneath: `.seed` / `.segmentation` / `.seg_prob` are lazy views of device arrays,
and `segment_at` / `segment_all` / `update_at` are single calls into libffn_b200 that run the
reference's Python loops inside one persistent CUDA kernel.  The host never falls back to numpy
for the hot path: a canvas can only be built on a `B200ExecutorClient`.
"""

import logging
import os
import threading
import time

import numpy as np
from scipy.special import expit  # noqa: F401  (kept: part of the module surface callers import)
from scipy.special import logit

from .. import _lib
from .. import engine as engine_lib
from ..training import model as ffn_model
from . import executor
from . import inference_pb2
from . import movement
from . import seed
from . import storage
from .inference_utils import Counters
from .inference_utils import TimedIter  # noqa: F401
from .inference_utils import timer_counter

MSEC_IN_SEC = 1000
MAX_SELF_CONSISTENT_ITERS = 32
SEED_CHUNK = 4096   # seeds handed to the device per call (checkpoint granularity)
# keep_history: capacity of the device event log between two collections (16 bytes per event, about
# 15 events per FoV step: pushes, pops and their verdicts, the step, its history_deleted count)
HISTORY_TRACE_EVENTS = 1 << 21
_EV_STEP, _EV_SEED_START, _EV_DELETED = 6, 8, 9


class DeviceArray:
  """numpy-like lazy view of one device-resident canvas array.

  Reads pull the requested box (basic slicing) or the whole array (fancy indexing) from the GPU;
  writes go the other way.  `np.asarray(view)` materialises it.  `clear()` restores the default
  value like `storage.NumpyArray.clear`.
  """

  def __init__(self, dev: engine_lib.DeviceCanvas, which: int, dtype, default_value=0):
    self._dev = dev
    self._which = which
    self.dtype = np.dtype(dtype)
    self.shape = dev.shape
    self.ndim = 3
    self.default_value = default_value

  @property
  def size(self):
    return int(np.prod(self.shape))

  def _box(self, key):
    """key -> (lo, size, squeeze axes) for basic integer/slice indexing, else None."""
    if not isinstance(key, tuple):
      key = (key,)
    if any(k is Ellipsis for k in key):
      i = [j for j, k in enumerate(key) if k is Ellipsis][0]
      key = key[:i] + (slice(None),) * (3 - (len(key) - 1)) + key[i + 1:]
    key = key + (slice(None),) * (3 - len(key))
    if len(key) != 3:
      return None
    lo, size, squeeze = [], [], []
    for axis, (k, dim) in enumerate(zip(key, self.shape)):
      if isinstance(k, (int, np.integer)):
        k = int(k)
        if k < 0:
          k += dim
        if not 0 <= k < dim:
          raise IndexError('index out of bounds')
        lo.append(k)
        size.append(1)
        squeeze.append(axis)
      elif isinstance(k, slice):
        start, stop, step = k.indices(dim)
        if step != 1:
          return None
        lo.append(start)
        size.append(max(stop - start, 0))
      else:
        return None
    return lo, size, tuple(squeeze)

  def __array__(self, dtype=None, copy=None):
    del copy
    arr = self._dev.read(self._which)
    return arr.astype(dtype) if dtype is not None else arr

  def __getitem__(self, key):
    box = self._box(key)
    if box is None:
      return np.asarray(self)[key]
    lo, size, squeeze = box
    if min(size) == 0:
      return np.zeros([s for a, s in enumerate(size) if a not in squeeze], dtype=self.dtype)
    out = self._dev.read(self._which, lo, size)
    out = out.squeeze(axis=squeeze) if squeeze else out
    return out[()] if out.ndim == 0 else out

  def __setitem__(self, key, value):
    box = self._box(key)
    if box is None:
      full = np.asarray(self)
      full[key] = value
      self._dev.write(self._which, full)
      return
    lo, size, squeeze = box
    if min(size) == 0:
      return
    data = np.empty(size, dtype=self.dtype)
    view = data.squeeze(axis=squeeze) if squeeze else data
    view[...] = value
    self._dev.write(self._which, data, lo)

  def clear(self):
    self[...] = self.default_value

  def astype(self, dtype):
    return np.asarray(self).astype(dtype)

  def max(self):
    return np.asarray(self).max()

  def __len__(self):
    return self.shape[0]

  def __lt__(self, o): return np.asarray(self) < o
  def __le__(self, o): return np.asarray(self) <= o
  def __gt__(self, o): return np.asarray(self) > o
  def __ge__(self, o): return np.asarray(self) >= o
  def __eq__(self, o): return np.asarray(self) == o   # pylint: disable=g-bad-name
  def __ne__(self, o): return np.asarray(self) != o
  __hash__ = None


class Canvas:
  """Tracks state of the inference progress and results within a subvolume."""

  io_lock = threading.Lock()

  def __init__(self, model_info: ffn_model.ModelInfo, exec_client, image, options,
               voxel_size_zyx=(1, 1, 1), counters=None, restrictor=None, movement_policy_fn=None,
               keep_history=False, checkpoint_path=None, checkpoint_interval_sec=0, corner_zyx=None,
               storage_cls=None, keep_probability_maps=False, image_mean=None, image_stddev=None):
    """See ffn/inference/inference.py:137-181 for the shared arguments.

    Extra (optional) arguments: `image_mean` / `image_stddev` — when given, `image` is the RAW
    uint8 volume and is normalised on the device as (x - mean) / stddev (runner.py:383-385),
    which keeps a 4x smaller image resident in HBM; otherwise `image` must already be float.
    """
    if not isinstance(exec_client, executor.B200ExecutorClient):
      raise TypeError('Canvas runs its flood-fill loop on the GPU and needs a B200ExecutorClient '
                      '(from B200Executor.get_client); got %r' % type(exec_client))
    del storage_cls
    self._exec_client = exec_client
    self._exec_client_id = None
    self.voxel_size_zyx = voxel_size_zyx

    self.options = inference_pb2.InferenceOptions()
    self.options.CopyFrom(options)
    for attr in ('init_activation', 'pad_value', 'move_threshold', 'segment_threshold'):
      setattr(self.options, attr, logit(getattr(self.options, attr)))   # inference.py:186-195

    self.counters = counters if counters is not None else Counters()
    self.checkpoint_interval_sec = checkpoint_interval_sec
    self.checkpoint_path = checkpoint_path
    self.checkpoint_last = time.time()
    self._keep_history = bool(keep_history)
    self.corner_zyx = corner_zyx

    raw_u8 = image_mean is not None and image_stddev is not None
    if raw_u8:
      self._raw_image = np.ascontiguousarray(image, dtype=np.uint8)
      self._image_mean, self._image_stddev = float(image_mean), float(image_stddev)
      self._image_f32 = None
    else:
      self._raw_image = None
      self._image_f32 = np.ascontiguousarray(image, dtype=np.float32)
    self.shape = tuple(int(s) for s in image.shape)

    self.restrictor = movement.MovementRestrictor() if restrictor is None else restrictor

    self._pred_size = np.array(model_info.pred_mask_size[::-1])
    self._input_seed_size = np.array(model_info.input_seed_size[::-1])
    self._input_image_size = np.array(model_info.input_image_size[::-1])
    self.margin = self._input_image_size // 2
    self._pred_delta = (self._input_seed_size - self._pred_size) // 2
    assert np.all(self._pred_delta >= 0)
    eng = exec_client.engine
    if tuple(int(v) for v in self._input_image_size) != eng.fov_zyx:
      raise ValueError('model_info FoV %r differs from the engine FoV %r' % (self._input_image_size, eng.fov_zyx))

    if movement_policy_fn is None:
      self.movement_policy = movement.FaceMaxMovementPolicy(
          self, deltas=model_info.deltas[::-1], score_threshold=self.options.move_threshold)
    else:
      self.movement_policy = movement_policy_fn(self)
    if not isinstance(self.movement_policy, movement.FaceMaxMovementPolicy):
      raise NotImplementedError('only FaceMaxMovementPolicy runs on the device')
    if tuple(int(d) for d in self.movement_policy.deltas) != eng.deltas_zyx:
      raise ValueError('movement policy deltas differ from the engine deltas')

    mbd = self.options.min_boundary_dist
    dev_opts = _lib.Options()
    dev_opts.init_activation = self.options.init_activation
    dev_opts.pad_value = self.options.pad_value
    dev_opts.move_threshold = self.options.move_threshold
    dev_opts.segment_threshold = self.options.segment_threshold
    dev_opts.disco_seed_threshold = self.options.disco_seed_threshold
    dev_opts.policy_score_threshold = float(self.movement_policy.score_threshold)
    dev_opts.min_boundary_dist_zyx = _lib.i3((mbd.z, mbd.y, mbd.x))
    dev_opts.min_segment_size = self.options.min_segment_size
    with exec_client.engine_lock:
      if raw_u8:
        self._dev = engine_lib.DeviceCanvas(eng, self._raw_image, dev_opts, self._image_mean, self._image_stddev,
                                            keep_probability_maps=keep_probability_maps)
      else:
        self._dev = engine_lib.DeviceCanvas(eng, self._image_f32, dev_opts,
                                            keep_probability_maps=keep_probability_maps)
      # positions the FoV may not enter: the position mask, plus — evaluated once for every voxel —
      # the shift-mask rule of MovementRestrictor.is_valid_pos (movement.py:312-334)
      movement_mask = self.restrictor.movement_mask(self.shape)
      if movement_mask is not None:
        self._dev.set_mask(_lib.MASK_MOVEMENT, movement_mask)
      if self.restrictor.seed_mask is not None:
        self._dev.set_mask(_lib.MASK_SEED, self.restrictor.seed_mask)

    self.seed = DeviceArray(self._dev, _lib.ARRAY_SEED, np.float32, np.nan)
    self.segmentation = DeviceArray(self._dev, _lib.ARRAY_SEGMENTATION, np.int32, 0)
    self.keep_probability_maps = keep_probability_maps
    self.seg_prob = DeviceArray(self._dev, _lib.ARRAY_QPROB, np.uint8, 0) if keep_probability_maps else None
    if self._keep_history:
      with exec_client.engine_lock:
        self._dev.start_trace(HISTORY_TRACE_EVENTS)

    self.global_to_local_ids = {}
    self.local_to_global_ids = {}
    self.seed_policy = None
    self._seed_policy_state = None
    self._max_id = 0
    self.origins = {}
    self.overlaps = {}
    self.reset_seed_per_segment = True
    self._hosts = []
    self._last_counters = None
    self.history = []
    self.history_deleted = []
    self._min_pos = np.array((0, 0, 0))
    self._max_pos = np.array((0, 0, 0))
    self.reset_state((0, 0, 0))
    self.t_last_predict = None
    self.log_info('Constructed canvas with corner %s (zyx) and shape %s', self.corner_zyx, self.shape)

  # -- image access ----------------------------------------------------------------------------
  @property
  def image(self):
    """Normalised float32 image (host copy, created lazily when only the raw volume was given)."""
    if self._image_f32 is None:
      self._image_f32 = (self._raw_image.astype(np.float32) - np.float32(self._image_mean)) / np.float32(
          self._image_stddev)
    return self._image_f32

  # -- client registration ---------------------------------------------------------------------
  def _register_client(self):
    if self._exec_client_id is None:
      self._exec_client_id = self._exec_client.start()
      logging.info('Registered as client %d.', self._exec_client_id)

  def _deregister_client(self):
    if self._exec_client_id is not None:
      logging.info('Deregistering client %d', self._exec_client_id)
      self._exec_client.finish()
      self._exec_client_id = None

  def __del__(self):
    try:
      self._deregister_client()
      if getattr(self, '_dev', None) is not None:
        self._dev.close()
    except Exception:  # pylint: disable=broad-except
      pass

  def local_id(self, segment_id):
    return self.global_to_local_ids.get(segment_id, segment_id)

  def reset_state(self, start_pos, reset_extents=True):
    self.movement_policy.reset_state(start_pos)
    self.history = []
    self.history_deleted = []
    if reset_extents:
      self._min_pos = np.array(start_pos)
      self._max_pos = np.array(start_pos)
    self._register_client()

  # -- counters ----------------------------------------------------------------------------------
  _COUNTER_MAP = (
      ('skip_threshold', 'skip_threshold'), ('skip_invalid_pos', 'skip_invalid_pos'),
      ('skip_restricted_pos', 'skip_restriced_pos'), ('seed_got_too_weak', 'seed_got_too_weak'),
      ('voxels_segmented', 'voxels-segmented'), ('voxels_overlapping', 'voxels-overlapping'))

  def _sync_counters(self):
    """Adds what the device counted since the last sync to the reference-named counters."""
    ctr = self._dev.counters()
    prev = self._last_counters
    def delta(name):
      return getattr(ctr, name) - (getattr(prev, name) if prev is not None else 0)
    for src, dst in self._COUNTER_MAP:
      d = delta(src)
      if d:
        self.counters[dst].IncrementBy(d)
    steps = delta('inference_calls')
    if steps:
      ms = (ctr.device_seconds - (prev.device_seconds if prev is not None else 0.0)) * MSEC_IN_SEC
      for name in ('inference', 'predict', 'update_at', 'movement_policy'):
        self.counters[name + '-calls'].IncrementBy(steps)
      self.counters['inference-time-ms'].IncrementBy(ms)
      self.counters['predict-time-ms'].IncrementBy(ms)
      self.counters['update_at-time-ms'].IncrementBy(ms)
    self._last_counters = ctr
    self._max_id = max(self._max_id, int(ctr.max_id))
    return ctr

  # -- validity / single steps -----------------------------------------------------------------
  def is_valid_pos(self, pos, ignore_move_threshold=False):
    pos = tuple(int(p) for p in pos)
    if not ignore_move_threshold:
      if self.seed[pos] < self.options.move_threshold:
        self.counters['skip_threshold'].Increment()
        return False
    np_pos = np.array(pos)
    if np.any(np_pos - self.margin < 0) or np.any(np_pos + self.margin >= self.shape):
      self.counters['skip_invalid_pos'].Increment()
      return False
    if self.segmentation[pos] > 0:
      self.counters['skip_invalid_pos'].Increment()
      return False
    return True

  def _get_image(self, pos):
    start = np.array(pos) - self.margin
    end = start + self._input_image_size
    return self.image[tuple(slice(int(s), int(e)) for s, e in zip(start, end))]

  def predict(self, pos, logit_seed):
    """One network evaluation through the executor client (inference.py:356-384)."""
    with timer_counter(self.counters, 'predict'):
      with timer_counter(self.counters, 'get-image'):
        img = self._get_image(pos)
      with timer_counter(self.counters, 'inference'):
        fetches = self._exec_client.predict(logit_seed, img, ['logits'])
      self.t_last_predict = time.time()
    return fetches.pop('logits')[..., 0]

  def _collect_history(self):
    """keep_history (inference.py:168-170): turns the device event log recorded since the last call into
    `history` (positions visited, :520-521) and `history_deleted` (:420-422), then restarts the log.
    A new object (`segment_all` only) restarts both lists, like `reset_state` (:303-304)."""
    if not self._keep_history:
      return
    with self._exec_client.engine_lock:
      events, produced = self._dev.get_trace(with_total=True)
      self._dev.start_trace(HISTORY_TRACE_EVENTS)
    if produced > events.shape[0]:
      logging.warning('history log overflow: %d of %d events kept', events.shape[0], produced)
    for ev, z, y, x in events.tolist():
      if ev == _EV_SEED_START:
        self.history, self.history_deleted = [], []
      elif ev == _EV_STEP:
        self.history.append((z, y, x))
      elif ev == _EV_DELETED:
        self.history_deleted.append(z)

  def update_at(self, pos):
    """One FoV step on the device: gather, network, disco merge, paste (inference.py:386-441)."""
    with self._exec_client.engine_lock:
      pred = self._dev.update_at(tuple(int(p) for p in pos))
    self._sync_counters()
    self._collect_history()
    return pred

  def init_seed(self, pos):
    with self._exec_client.engine_lock:
      self._dev.init_seed(tuple(int(p) for p in pos))

  def get_next_segment_id(self):
    self._max_id += 1
    while self._max_id in self.origins:
      self._max_id += 1
    return self._max_id

  # -- one object --------------------------------------------------------------------------------
  def segment_at(self, start_pos, dynamic_image=None, vis_update_every=10, vis_fixed_z=False,
                 partial_segment_iters=0, max_steps=0):
    """Runs FFN segmentation from `start_pos`; returns the number of FoV steps performed."""
    del dynamic_image, vis_update_every, vis_fixed_z
    start_pos = tuple(int(p) for p in start_pos)
    with timer_counter(self.counters, 'segment_at-loop'):
      with self._exec_client.engine_lock:
        if not partial_segment_iters:
          self.reset_state(start_pos, reset_extents=self.reset_seed_per_segment)
          st = self._dev.segment_at(start_pos, reset=True, max_steps=max_steps, keep_seed=not self.reset_seed_per_segment)
        else:
          st = self._dev.segment_at(start_pos, reset=False, max_steps=max_steps)
      self._min_pos = np.array(list(st.min_pos))
      self._max_pos = np.array(list(st.max_pos))
      self._sync_counters()
      self._collect_history()
    self._last_segment_finished = bool(st.finished)
    return int(st.iters)

  def log_info(self, string, *args, **kwargs):
    logging.info('[cl %s] ' + string, self._exec_client_id, *args, **kwargs)

  # -- whole canvas --------------------------------------------------------------------------------
  def segment_all(self, seed_policy=seed.PolicyPeaks, partial_segment_iters=0):
    """Segments the input image from every seed the policy proposes (inference.py:538-683)."""
    self.seed_policy = seed_policy(self)
    if self._seed_policy_state is not None:
      self.seed_policy.set_state(self._seed_policy_state)
      self._seed_policy_state = None

    with timer_counter(self.counters, 'segment_all'):
      with timer_counter(self.counters, 'seed-policy'):
        coords = self.seed_policy.remaining()
      if (partial_segment_iters or getattr(self, '_resume_pending', False)) and coords.shape[0]:
        # A checkpoint taken inside an object stores the seed policy one step back (inference.py:745-747), so the
        # first remaining seed IS the restored in-flight object: the device finishes that object first
        # (ffn_canvas_set_resume), the list continues behind it.
        coords = coords[1:]
        self.seed_policy.idx += 1
      self._resume_pending = False
      self.counters['seed-policy-calls'].IncrementBy(max(coords.shape[0] - 1, 0))
      first = True
      pos = 0
      chunk_size = SEED_CHUNK
      while first or pos < coords.shape[0]:
        first = False
        chunk = coords[pos:pos + chunk_size]
        t_chunk = time.time()
        with self._exec_client.engine_lock:
          origins, overlaps, _ = self._dev.segment_all(chunk)
        pos += chunk.shape[0]
        if self.checkpoint_path is not None and self.checkpoint_interval_sec > 0 and chunk.shape[0]:
          # the reference looks at the checkpoint clock after every FoV step (inference.py:531); here the host is back
          # between device calls, so the seeds handed over per call are sized to last at most half an interval
          rate = chunk.shape[0] / max(time.time() - t_chunk, 1e-3)
          chunk_size = int(min(SEED_CHUNK, max(16, rate * self.checkpoint_interval_sec / 2)))
        self.seed_policy.idx += chunk.shape[0]
        ctr = self._sync_counters()
        self._collect_history()
        self.counters['segment_at-loop-calls'].Set(int(ctr.segment_at_calls))
        per_id = {}
        for ov in overlaps:
          per_id.setdefault(ov.id, []).append((ov.other_id, ov.count))
        for o in origins:
          sid = int(o.id)
          self.origins[sid] = storage.OriginInfo(tuple(int(v) for v in o.start_zyx), int(o.iters),
                                                 float(o.walltime_sec))
          pairs = sorted(per_id.get(sid, []))
          self.overlaps[sid] = np.array([[p[0] for p in pairs], [p[1] for p in pairs]], dtype=np.int64)
          self._max_id = max(self._max_id, sid)
        self._maybe_save_checkpoint(partial_segment_iters=0)
    self.log_info('Segmentation done.')
    self._deregister_client()

  def init_segmentation_from_volume(self, volume, corner, end, align_and_crop=None):
    """Starts from an existing segmentation (inference.py:685-726)."""
    init_seg = volume[:, corner[0]:end[0], corner[1]:end[1], corner[2]:end[2]]
    init_seg = np.asarray(init_seg[0, ...])
    ids = np.unique(init_seg)
    ids = ids[ids != 0]
    global_to_local = {int(g): i + 1 for i, g in enumerate(ids)}   # make_contiguous
    lut_keys = np.array(list(global_to_local.keys()), dtype=init_seg.dtype)
    lut_vals = np.array(list(global_to_local.values()), dtype=np.int32)
    local = np.zeros(init_seg.shape, dtype=np.int32)
    if lut_keys.size:
      idx = np.searchsorted(lut_keys, init_seg)
      idx = np.clip(idx, 0, lut_keys.size - 1)
      hit = lut_keys[idx] == init_seg
      local[hit] = lut_vals[idx[hit]]
    self.global_to_local_ids = global_to_local
    self.local_to_global_ids = {v: k for k, v in global_to_local.items()}
    if align_and_crop is not None:
      local = align_and_crop(local)
    self.segmentation[:] = local
    if self.keep_probability_maps:
      qp = np.zeros(self.shape, dtype=np.uint8)
      qp[local > 0] = storage.quantize_probability(np.array([1.0]))
      self.seg_prob[:] = qp
    self._max_id = int(local.max()) if local.size else 0
    with self._exec_client.engine_lock:
      self._dev.set_max_id(self._max_id)

  # -- checkpoints ---------------------------------------------------------------------------------
  def restore_checkpoint(self, path):
    """Restores state saved by `save_checkpoint` (same npz keys as inference.py:728-778)."""
    self.log_info('Restoring inference checkpoint: %s', path)
    with open(path, 'rb') as f:
      data = np.load(f, allow_pickle=True)
      self.segmentation[:] = data['segmentation']
      self.seed[:] = data['seed']
      if self.keep_probability_maps and 'seg_qprob' in data:
        self.seg_prob[:] = data['seg_qprob']
      self.history_deleted = list(data['history_deleted'])
      self.history = list(data['history'])
      self.origins = data['origins'].item()
      if 'overlaps' in data:
        self.overlaps = data['overlaps'].item()
      seg = data['segmentation']
      self.counters['voxels-segmented'].Set(int(np.sum(seg != 0)))
      self._max_id = int(np.max(seg))
      self._min_pos = data['min_pos']
      self._max_pos = data['max_pos']
      self.movement_policy.restore_state(data['movement_policy'])
      self._seed_policy_state = data['seed_policy_state']
      self.counters.loads(data['counters'].item())
      partial = int(data['partial_segment_iters']) if 'partial_segment_iters' in data else 0
      if 'hosts' in data:
        self._hosts = list(data['hosts'])
    with self._exec_client.engine_lock:
      self._dev.set_max_id(self._max_id)
    self._last_counters = self._dev.counters()
    self._resume_pending = bool(partial)
    if partial:
      # the next segment_all (or segment_at(..., partial_segment_iters=...)) first finishes this object
      with self._exec_client.engine_lock:
        self._dev.set_resume(partial, tuple(int(v) for v in self._min_pos), tuple(int(v) for v in self._max_pos))
    self.log_info('Inference checkpoint restored.')
    return partial

  def save_checkpoint(self, path, partial_segment_iters=0):
    self.log_info('Saving inference checkpoint to %s.', path)
    with timer_counter(self.counters, 'save_checkpoint'):
      os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
      with storage.atomic_file(path) as fd:
        seed_policy_state = None
        if self.seed_policy is not None:
          seed_policy_state = self.seed_policy.get_state(partial_segment_iters > 0)
        aux = {}
        if self.keep_probability_maps:
          aux['seg_qprob'] = np.asarray(self.seg_prob)
        np.savez_compressed(
            fd, movement_policy=np.asarray(self.movement_policy.get_state(), dtype=object),
            segmentation=np.asarray(self.segmentation), seed=np.asarray(self.seed),
            origins=self.origins, overlaps=self.overlaps, min_pos=self._min_pos, max_pos=self._max_pos,
            history=np.array(self.history), history_deleted=np.array(self.history_deleted),
            seed_policy_state=np.asarray(seed_policy_state, dtype=object), counters=self.counters.dumps(),
            partial_segment_iters=partial_segment_iters, hosts=self._hosts, **aux)
    self.log_info('Inference checkpoint saved.')

  def _maybe_save_checkpoint(self, partial_segment_iters=0):
    if self.checkpoint_path is None or self.checkpoint_interval_sec <= 0:
      return
    if time.time() - self.checkpoint_last < self.checkpoint_interval_sec:
      return
    with Canvas.io_lock:
      self.save_checkpoint(self.checkpoint_path, partial_segment_iters=partial_segment_iters)
    self.checkpoint_last = time.time()
