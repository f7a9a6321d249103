"""Perf probe of the time-multiplexed chains: 250^3 multi-seed run with 1 / 2 / 3 chains, single-seed latency,
batched predict (pure conv-stack rounds).  One JSON object per line."""
import json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from ffn_b200 import _lib, engine as eng, tf_checkpoint
from ffn_b200.synthetic import interior_seed, voronoi_phantom

G = os.path.join(REPO, 'tests', 'golden')
W, B = tf_checkpoint.load_convstack_npz(os.path.join(G, 'fib25_convstack.npz'))
FLOPS = 45831462336.0
PEAK = 1708.2e12


def main():
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 250
  prof = '--profile' in sys.argv
  e = eng.Engine(W, B, (33, 33, 33), (8, 8, 8))
  # ---- batched predict
  rng = np.random.RandomState(0)
  for batch in (1, 3, 48):
    seed = np.where(rng.rand(batch, 33, 33, 33) < 0.3, rng.randn(batch, 33, 33, 33) * 2, -2.9444).astype(np.float32)
    img = rng.randn(batch, 33, 33, 33).astype(np.float32)
    e.predict(seed, img)
    e.predict(seed, img)
    ns = e.info()['last_kernel_ns']
    print(json.dumps({'probe': 'predict', 'batch': batch, 'kernel_us': ns / 1e3, 'patches_per_s': batch / (ns * 1e-9),
                      'roofline_frac': batch / (ns * 1e-9) * FLOPS / PEAK}), flush=True)
  # ---- single seed (configs[1])
  vol1 = voronoi_phantom((256, 256, 256), 1)
  cv = eng.DeviceCanvas(e, vol1, eng.make_options(), 128.0, 33.0)
  start = interior_seed(vol1, (128, 128, 128))
  cv.segment_at(start)
  c0 = cv.counters(); st = cv.segment_at(start); c1 = cv.counters()
  dev = c1.device_seconds - c0.device_seconds
  print(json.dumps({'probe': 'single_seed_256', 'steps': int(st.iters), 'steps_per_s': st.iters / dev,
                    'us_per_step': 1e6 * dev / st.iters}), flush=True)
  cv.close()
  # ---- multi-seed canvas
  vol = voronoi_phantom((n, n, n), 0)
  for chains in (1, 2, 3):
    e.set_chains(chains)
    cv = eng.DeviceCanvas(e, vol, eng.make_options(), 128.0, 33.0)
    noise = np.random.RandomState(seed=42).rand(*cv.shape)
    coords = cv.seed_peaks((1, 1, 1), noise)
    m = np.asarray((16, 16, 16))[None]
    keep = np.all((coords - m >= 0) & (coords + m < np.asarray(cv.shape)[None]), axis=1)
    seeds = np.ascontiguousarray(coords[keep], dtype=np.int32)
    if prof:
      e.enable_profiling(True)
      e.profile(reset=True)
    t0 = time.time()
    origins, overlaps, ctr = cv.segment_all(seeds, overlaps_cap=max(64 * len(seeds), 1 << 16))
    wall = time.time() - t0
    sp = cv.spec_stats()
    out = {'probe': 'segment_all_%d' % n, 'chains': chains, 'seeds': int(len(seeds)), 'steps_counted': int(ctr.inference_calls),
           'steps_executed': sp['steps_executed'], 'spec': sp, 'segments': int(ctr.segments),
           'segment_at_calls': int(ctr.segment_at_calls), 'device_s': ctr.device_seconds, 'wall_s': wall,
           'counted_steps_per_s': ctr.inference_calls / ctr.device_seconds,
           'executed_steps_per_s': sp['steps_executed'] / ctr.device_seconds,
           'voxels_per_s': ctr.voxels_segmented / wall,
           'roofline_frac_executed': sp['steps_executed'] / ctr.device_seconds * FLOPS / PEAK,
           'seg_hash': int(np.bitwise_xor.reduce(cv.read(_lib.ARRAY_SEGMENTATION).astype(np.int64).ravel() *
                                                 np.arange(n ** 3, dtype=np.int64) % 1000003))}
    if prof:
      out['profile'] = e.profile()
      e.enable_profiling(False)
    print(json.dumps(out), flush=True)
    cv.close()
  e.close()


if __name__ == '__main__':
  main()
