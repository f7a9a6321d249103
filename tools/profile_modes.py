"""Single-chain step time and device cycle counters as a function of the SM partition size."""
import json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from ffn_b200 import _lib, engine as eng, tf_checkpoint
from ffn_b200.synthetic import interior_seed, voronoi_phantom

G = os.path.join(REPO, 'tests', 'golden')
W, B = tf_checkpoint.load_convstack_npz(os.path.join(G, 'fib25_convstack.npz'))
vol = voronoi_phantom((256, 256, 256), 1)
start = interior_seed(vol, (128, 128, 128))
PROF = '--prof' in sys.argv
for n in [int(a) for a in sys.argv[1:] if not a.startswith('--')] or [148, 74, 49, 37]:
  e = eng.Engine(W, B, (33, 33, 33), (8, 8, 8), num_ctas=n)
  cv = eng.DeviceCanvas(e, vol, eng.make_options(), 128.0, 33.0)
  cv.segment_at(start)
  c0 = cv.counters(); st = cv.segment_at(start); c1 = cv.counters()
  dev = c1.device_seconds - c0.device_seconds
  c0 = cv.counters(); st = cv.segment_at(start); c1 = cv.counters()
  dev = min(dev, c1.device_seconds - c0.device_seconds)
  per = None
  if PROF:
    e.enable_profiling(True); e.profile(reset=True)
    st2 = cv.segment_at(start)
    prof = e.profile()
    e.enable_profiling(False)
    per = {k: {s: round(v / st2.iters) for s, v in d.items()} for k, d in prof.items()}
  print(json.dumps({'lib': os.path.basename(_lib.LIB_PATH), 'ctas': n, 'steps': int(st.iters),
                    'us_per_step': dev / st.iters * 1e6, 'steps_per_s': st.iters / dev, 'cycles_per_step': per}), flush=True)
  cv.close(); e.close()
