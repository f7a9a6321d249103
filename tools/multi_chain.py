"""Throughput mode: N engines with disjoint SM budgets run their persistent kernels concurrently.

  python tools/multi_chain.py [chains ...]      e.g. 1 2 3
Each chain flood-fills its own 256^3 volume (single-seed objects, like bench.py) from its own host
thread.  Prints aggregate FoV steps/s and checks chain 0 against the whole-GPU result bit for bit.
"""
import json, os, sys, threading, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from ffn_b200 import _lib, engine as eng, tf_checkpoint
import bench

(W, B), _ = bench.load_weights()
STEPS = int(os.environ.get('STEPS', 512))


def run_chain(e, vol, pts, steps, out, idx, barrier):
  cv = eng.DeviceCanvas(e, vol, eng.make_options(), 128.0, 33.0)
  done, k = 0, 0
  while done < 16:                      # warm-up
    done += int(cv.segment_at(pts[k % len(pts)], max_steps=16 - done).iters); k += 1
  c0 = cv.counters()
  barrier.wait()
  t0 = time.perf_counter()
  done = 0
  while done < steps:
    st = cv.segment_at(pts[k % len(pts)], max_steps=steps - done)
    done += int(st.iters); k += 1
  wall = time.perf_counter() - t0
  c1 = cv.counters()
  out[idx] = {'steps': done, 'wall': wall, 'device': c1.device_seconds - c0.device_seconds,
              'seed_sum': float(np.nansum(cv.read(_lib.ARRAY_SEED).astype(np.float64)))}
  cv.close()


def main():
  chains_list = [int(a) for a in sys.argv[1:]] or [1, 2, 3]
  vols = {}
  ref_sum = None
  for n in chains_list:
    ctas = 148 // n
    engines = [eng.Engine(W, B, bench.FOV, bench.DELTAS, num_ctas=ctas) for _ in range(n)]
    out = [None] * n
    barrier = threading.Barrier(n)
    threads = []
    for i in range(n):
      if i not in vols:
        vols[i] = bench.make_volume(1 + i)
      pts = bench.seed_points(vols[i], 64)
      threads.append(threading.Thread(target=run_chain, args=(engines[i], vols[i], pts, STEPS, out, i, barrier)))
    t0 = time.perf_counter()
    for t in threads: t.start()
    for t in threads: t.join()
    total = sum(o['steps'] for o in out)
    wall = max(o['wall'] for o in out)
    if ref_sum is None:
      ref_sum = out[0]['seed_sum']
    print(json.dumps({'chains': n, 'ctas_per_chain': ctas, 'tiles_per_cta': -(-engines[0].info()['tiles'] // ctas),
                      'aggregate_steps_per_s_wall': total / wall,
                      'per_chain_steps_per_s_device': [o['steps'] / o['device'] for o in out],
                      'chain0_matches_whole_gpu_run': out[0]['seed_sum'] == ref_sum}), flush=True)
    for e in engines: e.close()


if __name__ == '__main__':
  main()
