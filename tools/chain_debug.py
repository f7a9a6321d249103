"""Chains vs sequential on one volume: first differing origin / counters, under the scheduler debug knobs."""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from ffn_b200 import _lib, engine as eng, tf_checkpoint
from ffn_b200.synthetic import voronoi_phantom

G = os.path.join(REPO, 'tests', 'golden')
W, B = tf_checkpoint.load_convstack_npz(os.path.join(G, 'fib25_convstack.npz'))


def run(e, vol, seeds, chains, debug):
  os.environ['FFN_B200_DEBUG'] = str(debug)
  e.set_chains(chains)
  cv = eng.DeviceCanvas(e, vol, eng.make_options(), 128.0, 33.0)
  origins, overlaps, ctr = cv.segment_all(seeds, overlaps_cap=1 << 18)
  seg = cv.read(_lib.ARRAY_SEGMENTATION)
  seedc = cv.read(_lib.ARRAY_SEED)
  out = dict(seg=seg, seed=seedc, origins=[(o.id, tuple(o.start_zyx), int(o.iters)) for o in origins], wall=[o.walltime_sec for o in origins],
             ctr={n: getattr(ctr, n) for n, _ in ctr._fields_ if n not in ('device_seconds', 'kernel_launches')},
             spec=cv.spec_stats())
  cv.close()
  return out


def main():
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 250
  vol = voronoi_phantom((n, n, n), 0)
  e = eng.Engine(W, B, (33, 33, 33), (8, 8, 8))
  cv = eng.DeviceCanvas(e, vol, eng.make_options(), 128.0, 33.0)
  coords = cv.seed_peaks((1, 1, 1), np.random.RandomState(seed=42).rand(*cv.shape))
  cv.close()
  m = np.asarray((16, 16, 16))[None]
  seeds = np.ascontiguousarray(coords[np.all((coords - m >= 0) & (coords + m < n), axis=1)], dtype=np.int32)
  ref = run(e, vol, seeds, 1, 0)
  again = run(e, vol, seeds, 1, 0)
  print(json.dumps({'case': 'sequential twice', 'equal': bool(np.array_equal(ref['seg'], again['seg']))}), flush=True)
  for chains, debug in [(3, int(a)) for a in sys.argv[2:]] or ((3, 2), (3, 1), (2, 0), (3, 0)):
    got = run(e, vol, seeds, chains, debug)
    k = 0
    while k < min(len(ref['origins']), len(got['origins'])) and ref['origins'][k] == got['origins'][k]:
      k += 1
    diff = {a: (ref['ctr'][a], got['ctr'][a]) for a in ref['ctr'] if ref['ctr'][a] != got['ctr'][a]}
    print(json.dumps({'case': 'chains=%d debug=%d' % (chains, debug), 'seg_equal': bool(np.array_equal(ref['seg'], got['seg'])), 'seed_equal': bool(np.array_equal(ref['seed'], got['seed'], equal_nan=True)),
                      'first_diff_origin': k, 'ref_origin': ref['origins'][k:k + 2], 'got_origin': got['origins'][k:k + 2],
                      'n_origins': (len(ref['origins']), len(got['origins'])), 'codes': got['wall'][max(k - 3, 0):k + 3], 'ctr_diff': diff, 'spec': got['spec']}), flush=True)
  e.close()


if __name__ == '__main__':
  main()
