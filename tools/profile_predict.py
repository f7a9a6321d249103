"""ffn_predict(batch=48) three times: the conv stack alone, four patches per round by default (for ncu captures)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from ffn_b200 import engine as eng, tf_checkpoint

W, B = tf_checkpoint.load_convstack_npz(os.path.join(REPO, 'tests', 'golden', 'fib25_convstack.npz'))
e = eng.Engine(W, B, (33, 33, 33), (8, 8, 8))
if len(sys.argv) > 1:
  e.set_chains(int(sys.argv[1]))
rng = np.random.RandomState(0)
batch = 48
seed = np.where(rng.rand(batch, 33, 33, 33) < 0.3, rng.randn(batch, 33, 33, 33) * 2, -2.9444).astype(np.float32)
img = rng.randn(batch, 33, 33, 33).astype(np.float32)
for _ in range(3):
  e.predict(seed, img)
  print('kernel_us', e.info()['last_kernel_ns'] / 1e3, 'patches/s', batch / (e.info()['last_kernel_ns'] * 1e-9))
e.close()
