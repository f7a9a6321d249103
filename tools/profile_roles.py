"""Per-role cycle counters (profiled kernel build) of a batched predict: where the pipeline waits."""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from ffn_b200 import engine as eng, tf_checkpoint

W, B = tf_checkpoint.load_convstack_npz(os.path.join(REPO, 'tests', 'golden', 'fib25_convstack.npz'))
e = eng.Engine(W, B, (33, 33, 33), (8, 8, 8))
rng = np.random.RandomState(0)
batch = 48
seed = np.where(rng.rand(batch, 33, 33, 33) < 0.3, rng.randn(batch, 33, 33, 33) * 2, -2.9444).astype(np.float32)
img = rng.randn(batch, 33, 33, 33).astype(np.float32)
for chains in (1, 2, 3):
  e.set_chains(chains)
  e.enable_profiling(False)
  e.predict(seed, img)
  e.predict(seed, img)
  plain_us = e.info()['last_kernel_ns'] / 1e3
  e.enable_profiling(True)
  e.profile(reset=True)
  e.predict(seed, img)
  prof = e.profile()
  us = e.info()['last_kernel_ns'] / 1e3
  layers = batch / chains * 24
  out = {'chains': chains, 'use_tmap_env': os.environ.get('FFN_B200_TMAP', '0'), 'plain_kernel_us': plain_us,
         'profiled_kernel_us': us, 'patches_per_s': batch / (plain_us * 1e-6),
         'cycles_per_round_layer': {k: round(v / layers) for k, v in prof['cta_last'].items() if v},
         'cta0_cycles_per_round_layer': {k: round(v / layers) for k, v in prof['cta0'].items() if v}}
  print(json.dumps(out), flush=True)
e.close()
