"""Runs BASELINE.json's configurations on this box and prints one JSON object per config.

  python tools/run_configs.py [0 1 2 4] [--size-scale S]        # single GPU
  torchrun --nproc-per-node N tools/run_configs.py 3            # config 3: slabs across N GPUs + NCCL merge

Config 0 goes through the full reference entry-point surface (InferenceRequest -> Runner.run with
PolicyPeaks seeds -> seg-*.npz); the others drive the engine API directly.  Reported: FoV steps/s
(device time), segmented voxels/s end-to-end (wall clock of segment_all incl. host seed policy),
segments, and invariants (every origin carries its own id).
"""
import json, os, sys, tempfile, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np

from ffn_b200 import _lib, engine as eng, tf_checkpoint
from ffn_b200.synthetic import interior_seed, voronoi_phantom

G = os.path.join(REPO, 'tests', 'golden')
W, B = tf_checkpoint.load_convstack_npz(os.path.join(G, 'fib25_convstack.npz'))


def device_peaks(cv, voxel=(1, 1, 1), margin=(16, 16, 16)):
  """PolicyPeaks on the device + the border filter of BaseSeedPolicy (seed.py:81-88)."""
  t0 = time.time()
  noise = np.random.RandomState(seed=42).rand(*cv.shape)
  t_noise = time.time() - t0
  coords = cv.seed_peaks(voxel, noise)
  m = np.asarray(margin)[None]
  keep = np.all((coords - m >= 0) & (coords + m < np.asarray(cv.shape)[None]), axis=1)
  return np.ascontiguousarray(coords[keep], dtype=np.int32), time.time() - t0, t_noise


def full_canvas(engine, vol, label, extra=None, voxel=(1, 1, 1), margin=(16, 16, 16)):
  t0 = time.time()
  cv = eng.DeviceCanvas(engine, vol, eng.make_options(), 128.0, 33.0)
  t_up = time.time() - t0
  seeds, tseed, tnoise = device_peaks(cv, voxel, margin)
  t0 = time.time()
  origins, overlaps, ctr = cv.segment_all(seeds, overlaps_cap=max(64 * len(seeds), 1 << 16))
  wall = time.time() - t0
  seg = cv.read(_lib.ARRAY_SEGMENTATION)
  ok = all(seg[tuple(o.start_zyx)] == o.id for o in origins)
  out = {'config': label, 'shape': list(vol.shape), 'seeds': int(len(seeds)), 'steps': int(ctr.inference_calls),
         'segments': int(ctr.segments), 'segment_at_calls': int(ctr.segment_at_calls),
         'voxels_segmented': int(ctr.voxels_segmented), 'filled_fraction': float((seg > 0).mean()),
         'device_seconds': ctr.device_seconds, 'segment_all_wall_s': wall, 'upload_s': t_up,
         'seed_policy_s': tseed, 'seed_policy_noise_s': tnoise,
         'steps_per_s_device': ctr.inference_calls / max(ctr.device_seconds, 1e-9),
         'steps_per_s_wall': ctr.inference_calls / wall, 'voxels_per_s_wall': ctr.voxels_segmented / wall,
         'voxels_per_s_incl_seed_policy': ctr.voxels_segmented / (wall + tseed + t_up),
         'kernel_launches': int(ctr.kernel_launches), 'origins_carry_own_id': bool(ok)}
  if extra:
    out.update(extra)
  cv.close()
  return out, seg, int(ctr.max_id)


def main():
  args = [a for a in sys.argv[1:] if not a.startswith('--')]
  scale = 1.0
  for a in sys.argv[1:]:
    if a.startswith('--size-scale='):
      scale = float(a.split('=')[1])
  which = [int(a) for a in args] or [1]
  rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1)); local = int(os.environ.get('LOCAL_RANK', 0))

  if 1 in which:
    e = eng.Engine(W, B, (33, 33, 33), (8, 8, 8), device=local)
    vol = voronoi_phantom((256, 256, 256), 1)
    cv = eng.DeviceCanvas(e, vol, eng.make_options(), 128.0, 33.0)
    start = interior_seed(vol, (128, 128, 128))
    cv.segment_at(start)                                           # warm-up
    c0 = cv.counters(); t0 = time.time(); st = cv.segment_at(start); wall = time.time() - t0; c1 = cv.counters()
    dev = c1.device_seconds - c0.device_seconds
    seedc = cv.read(_lib.ARRAY_SEED)
    print(json.dumps({'config': '1: single-seed flood-fill, synthetic 256^3', 'start': list(start), 'steps': int(st.iters),
                      'device_seconds': dev, 'steps_per_s_device': st.iters / dev, 'steps_per_s_wall': st.iters / wall,
                      'voxels_above_segment_threshold': int((seedc >= eng.f32_logit(0.6)).sum())}), flush=True)
    cv.close(); e.close()

  if 0 in which:
    # configs[0]: the README run, on a synthetic stand-in for training_sample2 (250^3, seed 0)
    from google.protobuf import text_format
    from ffn.inference import inference_pb2, runner as runner_mod, storage
    n = int(250 * scale)
    vol = voronoi_phantom((n, n, n), 0)
    tmp = tempfile.mkdtemp()
    np.save(os.path.join(tmp, 'vol.npy'), vol)
    req = inference_pb2.InferenceRequest()
    text_format.Parse('''image { hdf5: "%s:raw" } image_mean: 128 image_stddev: 33 checkpoint_interval: 1800
      seed_policy: "PolicyPeaks" model_checkpoint_path: "%s" model_name: "convstack_3d.ConvStack3DFFNModel"
      model_args: "{\\"depth\\": 12, \\"fov_size\\": [33, 33, 33], \\"deltas\\": [8, 8, 8]}"
      segmentation_output_dir: "%s"
      inference_options { init_activation: 0.95 pad_value: 0.05 move_threshold: 0.9 min_boundary_dist { x: 1 y: 1 z: 1}
                          segment_threshold: 0.6 min_segment_size: 1000 }''' %
                      (os.path.join(tmp, 'vol.npy'), os.path.join(G, 'fib25_convstack.npz'), os.path.join(tmp, 'out')), req)
    t0 = time.time()
    runner = runner_mod.Runner(device=local)
    runner.start(req)
    t_start = time.time() - t0
    t0 = time.time()
    canvas = runner.run((0, 0, 0), (n, n, n))
    wall = time.time() - t0
    cnt = {k: c.value for k, c in canvas.counters}
    seg, origins = storage.load_segmentation(os.path.join(tmp, 'out'), (0, 0, 0))
    print(json.dumps({'config': '0: run_inference path (Runner.run, PolicyPeaks), synthetic %d^3' % n,
                      'runner_start_s': t_start, 'runner_run_wall_s': wall, 'steps': cnt.get('inference-calls'),
                      'segments': len(origins), 'voxels_segmented': cnt.get('voxels-segmented'),
                      'filled_fraction': float((seg > 0).mean()),
                      'seed_policy_ms': cnt.get('seed-policy-time-ms'), 'segment_all_ms': cnt.get('segment_all-time-ms'),
                      'inference_ms_device': cnt.get('inference-time-ms'),
                      'steps_per_s_end_to_end': cnt.get('inference-calls') / wall,
                      'voxels_per_s_end_to_end': cnt.get('voxels-segmented') / wall,
                      'seeds_examined': cnt.get('seed-policy-calls'), 'segment_at_calls': cnt.get('segment_at-loop-calls')}), flush=True)
    runner.stop_executor()

  if 2 in which:
    e = eng.Engine(W, B, (33, 33, 33), (8, 8, 8), device=local)
    n = int(512 * scale)
    t0 = time.time(); vol = voronoi_phantom((n, n, n), 2); tg = time.time() - t0
    out, _, _ = full_canvas(e, vol, '2: full-canvas multi-seed, synthetic %d^3, device PolicyPeaks seeds' % n,
                            {'volume_gen_s': tg})
    print(json.dumps(out), flush=True)
    e.close()

  if 4 in which:
    w9, b9 = W[:18] + [W[-1]], B[:18] + [B[-1]]
    e = eng.Engine(w9, b9, (17, 33, 33), (4, 8, 8), device=local)
    shape = (int(256 * scale), int(512 * scale), int(512 * scale))
    vol = voronoi_phantom(shape, 4, sigma=(0.5, 1.0, 1.0), voxel_size_zyx=(2.0, 1.0, 1.0))
    out, _, _ = full_canvas(e, vol, '4: anisotropic fov (17,33,33) deltas (4,8,8) depth 9 (first 9 FIB-25 modules), synthetic %s' % (shape,),
                            voxel=(2, 1, 1), margin=(8, 16, 16))
    print(json.dumps(out), flush=True)
    e.close()

  if 3 in which:
    import torch, torch.distributed as dist
    from ffn_b200 import distributed as D
    torch.cuda.set_device(local)
    if world > 1:
      dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    e = eng.Engine(W, B, (33, 33, 33), (8, 8, 8), device=local)
    n = int(1024 * scale)
    boxes = D.slab_boxes((n, n, n), 8)
    mine = D.slabs_of_rank(8, rank, world)
    results = []
    t0 = time.time()
    total_max = 0
    segs = []
    for k in mine:
      lo, sz = boxes[k]
      vol = voronoi_phantom(sz, 3 * 100 + k)      # each slab is an independent synthetic volume of the slab's shape
      out, seg, max_id = full_canvas(e, vol, '3: slab %d of 8 (%s)' % (k, 'x'.join(map(str, sz))))
      results.append(out)
      seg_t = torch.from_numpy(seg).cuda()
      seg_t[seg_t > 0] += total_max
      total_max += max_id
      segs.append(seg_t)
    local_labels = torch.stack(segs) if segs else torch.zeros((0,), dtype=torch.int32, device='cuda')
    t_work = time.time() - t0
    torch.cuda.synchronize()
    t0 = time.time()
    if world > 1:
      gathered, off, total = D.merge_labels(local_labels, total_max, dst=0)
      torch.cuda.synchronize()
    else:
      gathered, off, total = [local_labels], 0, total_max
    t_merge = time.time() - t0
    steps = sum(r['steps'] for r in results); vox = sum(r['voxels_segmented'] for r in results)
    dev = sum(r['device_seconds'] for r in results)
    stats = torch.tensor([steps, vox, dev, t_work, t_merge], dtype=torch.float64, device='cuda')
    if world > 1:
      smax = stats.clone(); dist.all_reduce(smax, op=dist.ReduceOp.MAX)
      ssum = stats.clone(); dist.all_reduce(ssum, op=dist.ReduceOp.SUM)
    else:
      smax = ssum = stats
    if rank == 0:
      ids_unique = True
      if gathered is not None and world > 1:
        seen = set()
        for gt in gathered:
          u = set(torch.unique(gt[gt > 0]).tolist())
          ids_unique &= not (seen & u)
          seen |= u
      print(json.dumps({'config': '3: %d^3 as 8 slabs of %d^3 over %d GPU(s), NCCL merge' % (n, n // 2, world),
                        'total_steps': float(ssum[0]), 'total_voxels_segmented': float(ssum[1]),
                        'max_rank_work_wall_s': float(smax[3]), 'max_rank_device_s': float(smax[2]),
                        'merge_s': float(smax[4]), 'steps_per_s_wall': float(ssum[0]) / float(smax[3]),
                        'voxels_per_s_wall': float(ssum[1]) / float(smax[3]), 'total_ids': int(total),
                        'ids_disjoint_after_merge': bool(ids_unique), 'per_slab_rank0': results}), flush=True)
    if world > 1:
      dist.destroy_process_group()
    e.close()


if __name__ == '__main__':
  main()
