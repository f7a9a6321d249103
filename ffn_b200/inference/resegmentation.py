"""Resegmentation: local re-runs of the flood fill at decision points of an existing segmentation.

Public surface of ffn/inference/resegmentation.py — `get_starting_location` (:37-45), `get_target_path`
(:48-80), `get_canvas` (:83-111), `process_point` (:114-293), `process` (:296-301) — on the device-resident
`Canvas`: each attempt is one persistent-kernel launch (`canvas.segment_at`), and `Canvas(keep_history=True)`
reads `history` / `history_deleted` back from the device event log.  Choosing where to start (distance
transform of the original object inside the (2r+1)^3 box, retries with an exclusion radius) stays on the
host: it runs once per attempt, not per FoV step.

Result file, same keys as upstream (:278-288): probs, raw_probs, deletes, histories, start_points, request,
counters, corner_zyx, is_shift.  Ragged entries (`start_points`; `deletes` / `histories` when attempts differ
in length) are object arrays: load with `allow_pickle=True`.

Bitrot papered over: the shard digest hashes bytes (upstream feeds `str` to md5: TypeError on Python 3);
`process` forwards a voxel size (upstream calls `process_point` without one).
"""

import hashlib
import logging
import os

import numpy as np
from scipy import ndimage
from scipy.special import expit

from . import storage
from .inference_utils import timer_counter


# ------------------------------------------------------------------------------------------------ helpers
def get_starting_location(dists, exclusion_radius):
  """Deepest point of the distance map; the box around it is zeroed so that a retry picks another one."""
  peak = np.unravel_index(np.argmax(dists), dists.shape)
  reach = (exclusion_radius.z, exclusion_radius.y, exclusion_radius.x)
  dists[tuple(slice(max(p - r, 0), p + r + 1) for p, r in zip(peak, reach))] = 0
  return tuple(int(p) for p in peak)


def get_target_path(request, point_num):
  """Where the result of decision point `point_num` goes; None when that file is already there."""
  entry = request.points[point_num]
  directory = request.output_directory
  if request.subdir_digits > 1:
    digest = hashlib.md5(('%d%d' % (entry.id_a, entry.id_b)).encode('utf-8')).hexdigest()
    directory = os.path.join(directory, digest[:request.subdir_digits])
  os.makedirs(directory, exist_ok=True)
  name = '%d-%d_at_%d_%d_%d.npz' % (entry.id_a, entry.id_b, entry.point.x, entry.point.y, entry.point.z)
  path = os.path.join(directory, name)
  if os.path.exists(path):
    logging.info('Output already exists: %s', path)
    return None
  return path


def get_canvas(point, radius, runner):
  """(Canvas, Alignment) for the box of half-size `radius` around `point` (both z, y, x), with history."""
  centre, reach = np.asarray(point), np.asarray(radius)
  lo = centre - reach
  extent = 2 * reach + 1
  hi = lo + extent
  available = np.asarray(runner.init_seg_volume.shape[1:])
  if np.any(lo < 0) or np.any(available <= hi):
    logging.error('Not enough context for: %d, %d, %d; corner: %r; end: %r', point[2], point[1], point[0], lo, hi)
    return None, None
  return runner.make_canvas(lo, extent, keep_history=True)


def _as_array(items):
  """np.array(items), or an object array when the items differ in shape."""
  if len({np.shape(item) for item in items}) <= 1:
    return np.array(items)
  ragged = np.empty(len(items), dtype=object)
  for slot, item in enumerate(items):
    ragged[slot] = item
  return ragged


def _seed_distance_map(canvas, original, voxel_size, veto_box):
  """Distance to the object's border, zeroed where a FoV would leave the box or seeding is vetoed."""
  depth = ndimage.distance_transform_edt(original, sampling=voxel_size)
  for axis, margin in enumerate(int(m) for m in canvas.margin):
    edge = [slice(None)] * 3
    edge[axis] = slice(0, margin)
    depth[tuple(edge)] = 0
    edge[axis] = slice(depth.shape[axis] - margin, None)
    depth[tuple(edge)] = 0
  if veto_box is not None:
    depth[veto_box] = 0
  return depth


def _recovered(request, options, original, prob, window):
  """Did the attempt grow back enough of the object it was seeded in (inside the analysis window)?"""
  inside = original[window]
  regrown = np.sum((prob[window] >= options.segment_threshold) & inside)
  if request.segment_recovery_fraction > 0:
    return regrown / np.sum(inside) >= request.segment_recovery_fraction
  return regrown >= options.min_segment_size


# ------------------------------------------------------------------------------------------------ driver
def process_point(request, runner, point_num, voxel_size=(1, 1, 1)):
  """Re-grows the one (endpoint) or two (pair) objects at decision point `point_num`; writes the result."""
  with timer_counter(runner.counters, 'resegmentation'):
    target_path = get_target_path(request, point_num)
    if target_path is None:
      return
    entry = request.points[point_num]
    point = (entry.point.z, entry.point.y, entry.point.x)
    radius = (request.radius.z, request.radius.y, request.radius.x)
    canvas, alignment = get_canvas(point, radius, runner)
    if canvas is None:
      logging.warning('Could not get a canvas object.')
      return

    endpoint = not entry.HasField('id_b')
    wanted = [entry.id_a] if endpoint else [entry.id_a, entry.id_b]
    labels = np.asarray(canvas.segmentation)
    originals = [labels == canvas.local_id(gid) for gid in wanted]
    if not all(obj.any() for obj in originals):
      logging.warning('Segments %r (local ids %r) not found in input at %r.  Current values are: %r.', wanted,
                      [canvas.local_id(gid) for gid in wanted], point, np.unique(labels))
      canvas._deregister_client()  # pylint:disable=protected-access
      return

    # An endpoint is re-grown on an empty canvas; a pair keeps every other object as context.
    cleared = np.ones(labels.shape, dtype=bool) if endpoint else (originals[0] | originals[1])
    labels[cleared] = 0
    canvas.segmentation[...] = labels
    quantised = np.asarray(canvas.seg_prob)
    quantised[cleared] = 0
    canvas.seg_prob[...] = quantised

    local = alignment.transform(np.array([point]).T)[:, 0] - np.asarray(canvas.corner_zyx)
    veto_box = None
    if request.HasField('init_exclusion_radius'):
      ier = request.init_exclusion_radius
      veto_box = tuple(slice(max(int(c) - r, 0), int(c) + r + 1) for c, r in zip(local, (ier.z, ier.y, ier.x)))
    if request.HasField('analysis_radius'):
      ar = request.analysis_radius
      window = tuple(slice(r - a, r + a + 1) for r, a in zip(radius, (ar.z, ar.y, ar.x)))
    else:
      window = (slice(None),) * 3
    options = request.inference.inference_options

    raw_probs, probs, deletes, histories = [], [], [], []
    start_points = [[], []]
    last_prob = None
    for index, original in enumerate(originals):
      logging.info('processing object %d', index)
      with timer_counter(canvas.counters, 'edt'):
        depth = _seed_distance_map(canvas, original, voxel_size, veto_box)
        canvas.log_info('EDT computation done')

      prob, success = None, False
      for _ in range(request.max_retry_iters):
        start = get_starting_location(depth, request.exclusion_radius)
        if not original[start]:
          continue
        canvas.log_info('.. starting segmentation at (xyz): %d %d %d', start[2], start[1], start[0])
        canvas.segment_at(start)
        prob = expit(np.asarray(canvas.seed))
        start_points[index].append(start[::-1])
        success = bool(_recovered(request, options, original, prob, window))
        if success:
          break

      if prob is not None:
        last_prob = prob
        packed = storage.quantize_probability(prob)
        raw_probs.append(packed)
        probs.append(alignment.align_and_crop(canvas.corner_zyx, packed, alignment.corner, alignment.size,
                                              forward=False))
        deletes.append(np.array(canvas.history_deleted))
        histories.append(np.array(canvas.history))

      if request.terminate_early:
        if not success:
          break
        # the first object already swallowed too little of the second one: they are not the same neurite
        if request.segment_recovery_fraction > 0 and index == 0 and len(originals) > 1 and last_prob is not None:
          other = originals[1][window]
          shared = np.sum((last_prob[window] >= options.segment_threshold) & other)
          if shared / np.sum(other) < request.segment_recovery_fraction:
            break

    canvas.log_info('saving results to %s', target_path)
    ragged_starts = np.empty(2, dtype=object)
    for slot, pts in enumerate(start_points):
      ragged_starts[slot] = np.array(pts).reshape(-1, 3)
    shifted = bool(canvas.restrictor is not None and np.any(canvas.restrictor.shift_mask))
    with storage.atomic_file(target_path) as fd:
      np.savez_compressed(fd, probs=np.array(probs), raw_probs=np.array(raw_probs), deletes=_as_array(deletes),
                          histories=_as_array(histories), start_points=ragged_starts,
                          request=request.SerializeToString(), counters=canvas.counters.dumps(),
                          corner_zyx=canvas.corner_zyx, is_shift=shifted)
    canvas.log_info('.. save complete')
    canvas._deregister_client()  # pylint:disable=protected-access


def process(request, runner, voxel_size=(1, 1, 1)):
  """Every decision point of the request, in order."""
  total = len(request.points)
  for index in range(total):
    logging.info('processing %d/%d', index, total)
    process_point(request, runner, index, voxel_size)
