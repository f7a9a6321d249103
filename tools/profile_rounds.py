"""Per-round cycle budget of Canvas.segment_all on the bench canvas (profiled kernel build): where a round's time goes
on CTA 0 (the leader) and on the last CTA.   python tools/profile_rounds.py [n=250] [chains...]"""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import bench
from ffn_b200 import _lib, engine as eng, tf_checkpoint

W, B = tf_checkpoint.load_convstack_npz(os.path.join(REPO, 'tests', 'golden', 'fib25_convstack.npz'))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 250
chain_list = [int(a) for a in sys.argv[2:]] or [1, 4]
e = eng.Engine(W, B, (33, 33, 33), (8, 8, 8))
vol = bench.make_volume((n, n, n), 0)
cv = eng.DeviceCanvas(e, vol, eng.make_options(), 128.0, 33.0)
coords = cv.seed_peaks((1, 1, 1), np.random.RandomState(seed=42).rand(*cv.shape))
cv.close()
m = np.asarray((16, 16, 16))[None]
seeds = np.ascontiguousarray(coords[np.all((coords - m >= 0) & (coords + m < n), axis=1)], dtype=np.int32)
for chains in chain_list:
  e.set_chains(chains)
  for prof in (False, True):
    e.enable_profiling(prof)
    if prof:
      e.profile(reset=True)
    cv = eng.DeviceCanvas(e, vol, eng.make_options(), 128.0, 33.0)
    _, _, ctr = cv.segment_all(seeds, overlaps_cap=1 << 18)
    sp = cv.spec_stats()
    cv.close()
    if not prof:
      plain = float(ctr.device_seconds)
      continue
    p = e.profile()
    rounds = max(sp['rounds'], 1)
    out = {'chains': chains, 'plain_dev_s': round(plain, 4), 'profiled_dev_s': round(float(ctr.device_seconds), 4),
           'rounds': rounds, 'steps_executed': sp['steps_executed'], 'us_per_round': round(1e6 * plain / rounds, 2),
           'cta0_cycles_per_round': {k: round(v / rounds) for k, v in p['cta0'].items() if v},
           'cta_last_cycles_per_round': {k: round(v / rounds) for k, v in p['cta_last'].items() if v}}
    print(json.dumps(out), flush=True)
e.close()
