"""Chains sweep of ONE engine library (FFN_B200_LIB selects a variant build, see build_variants.py):

  * ffn_predict(batch=60) patches/s for 1, 3, 4, 5 chains (as many as the build has), each bit-compared with 1 chain;
  * Canvas.segment_all on the n^3 bench canvas (device PolicyPeaks seeds): the sequential run (1 chain) as the
    reference, then every chain count under the scheduler knobs given on the command line (FFN_B200_DEBUG values,
    default 0 and 32 = without the parked-object test of the look-ahead), each compared bit for bit (labels, Canvas.seed,
    origins incl. iters, counters) and timed (flood-kernel CUDA-event seconds).

  FFN_B200_LIB=variants/libffn_b200_c5.so python tools/chain_sweep.py 250 0 32
"""
import json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from ffn_b200 import _lib, engine as eng, tf_checkpoint
import bench

G = os.path.join(REPO, 'tests', 'golden')
W, B = tf_checkpoint.load_convstack_npz(os.path.join(G, 'fib25_convstack.npz'))
FLOP = 2 * 33 ** 3 * (27 * 2 * 32 + 23 * 27 * 32 * 32 + 32)


def chain_counts(e):
  out = []
  for k in (1, 3, 4, 5):
    try:
      e.set_chains(k)
      out.append(k)
    except RuntimeError:
      break
  e.set_chains(0)
  return out


def run(e, vol, seeds, chains, debug):
  os.environ['FFN_B200_DEBUG'] = str(debug)
  e.set_chains(chains)
  cv = eng.DeviceCanvas(e, vol, eng.make_options(), 128.0, 33.0, keep_probability_maps=True)
  t0 = time.perf_counter()
  origins, _, ctr = cv.segment_all(seeds, overlaps_cap=1 << 18)
  wall = time.perf_counter() - t0
  out = dict(seg=cv.read(_lib.ARRAY_SEGMENTATION), seed=cv.read(_lib.ARRAY_SEED), qprob=cv.read(_lib.ARRAY_QPROB),
             origins=[(o.id, tuple(o.start_zyx), int(o.iters)) for o in origins],
             ctr={n: getattr(ctr, n) for n, _ in ctr._fields_ if n not in ('device_seconds', 'kernel_launches')},
             dev=float(ctr.device_seconds), wall=wall, spec=cv.spec_stats())
  cv.close()
  return out


def main():
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 250
  knobs = [int(a) for a in sys.argv[2:]] or [0, 32]
  lib = os.environ.get('FFN_B200_LIB', 'default')
  e = eng.Engine(W, B, (33, 33, 33), (8, 8, 8))
  ks = chain_counts(e)
  print(json.dumps({'lib': lib, 'chain_counts': ks, 'info': e.info()}), flush=True)
  # ---- conv stack alone
  rng = np.random.RandomState(0)
  batch = 60
  seed = np.where(rng.rand(batch, 33, 33, 33) < 0.3, rng.randn(batch, 33, 33, 33) * 2, -2.9444).astype(np.float32)
  img = rng.randn(batch, 33, 33, 33).astype(np.float32)
  ref = None
  for k in ks:
    e.set_chains(k)
    e.predict(seed, img)
    got = e.predict(seed, img)
    ns = e.info()['last_kernel_ns']
    if ref is None:
      ref = got
    rate = batch / (ns * 1e-9)
    print(json.dumps({'lib': lib, 'predict_chains': k, 'patches_per_s': round(rate, 1), 'tflops': round(rate * FLOP / 1e12, 1),
                      'equal_to_1_chain': bool(np.array_equal(got, ref))}), flush=True)
  # ---- the bench canvas
  vol = bench.make_volume((n, n, n), 0)
  e.set_chains(0)
  cv = eng.DeviceCanvas(e, vol, eng.make_options(), 128.0, 33.0)
  coords = cv.seed_peaks((1, 1, 1), np.random.RandomState(seed=42).rand(*cv.shape))
  cv.close()
  m = np.asarray((16, 16, 16))[None]
  seeds = np.ascontiguousarray(coords[np.all((coords - m >= 0) & (coords + m < n), axis=1)], dtype=np.int32)
  one = run(e, vol, seeds, 1, 0)
  steps = one['ctr']['inference_calls']
  print(json.dumps({'lib': lib, 'segment_all_chains': 1, 'steps': steps, 'dev_s': round(one['dev'], 4),
                    'steps_per_s': round(steps / one['dev'], 1), 'segments': one['ctr']['segments']}), flush=True)
  only = [int(v) for v in os.environ.get('SWEEP_CHAINS', '').split(',') if v]
  for k in ks[1:]:
    if only and k not in only:
      continue
    for dbg in knobs:
      got = run(e, vol, seeds, k, dbg)
      eq = dict(seg=bool(np.array_equal(got['seg'], one['seg'])), qprob=bool(np.array_equal(got['qprob'], one['qprob'])),
                seed=bool(np.array_equal(got['seed'], one['seed'], equal_nan=True)),
                origins=got['origins'] == one['origins'], ctr=got['ctr'] == one['ctr'])
      sp = got['spec']
      print(json.dumps({'lib': lib, 'segment_all_chains': k, 'debug': dbg, 'dev_s': round(got['dev'], 4), 'wall_s': round(got['wall'], 4),
                        'steps_per_s': round(steps / got['dev'], 1), 'frac_of_1708': round(steps / got['dev'] * FLOP / 1708.2e12, 4),
                        'all_equal': all(eq.values()), 'equal': eq, 'spec': sp,
                        'ctr_diff': {a: (one['ctr'][a], got['ctr'][a]) for a in one['ctr'] if one['ctr'][a] != got['ctr'][a]},
                        'chains_active_per_round': round(sp['steps_executed'] / max(sp['rounds'], 1), 3)}), flush=True)
  os.environ['FFN_B200_DEBUG'] = '0'
  e.close()


if __name__ == '__main__':
  main()
