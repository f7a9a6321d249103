"""Tile timeline of one CTA during a batched predict (profiled kernel build): when each role saw / finished each tile.
Prints, for the middle layers of one round with three chains, the events relative to the round start (cycles), and
saves the raw trace.   python tools/trace_tiles.py [chains] [out.npy]
"""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from ffn_b200 import engine as eng, tf_checkpoint

W, B = tf_checkpoint.load_convstack_npz(os.path.join(REPO, 'tests', 'golden', 'fib25_convstack.npz'))
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 3
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(REPO, 'gpurun_out', 'trace_tiles_%d.npy' % chains)
e = eng.Engine(W, B, (33, 33, 33), (8, 8, 8))
e.set_chains(chains)
rng = np.random.RandomState(0)
batch = 4 * chains
seed = np.where(rng.rand(batch, 33, 33, 33) < 0.3, rng.randn(batch, 33, 33, 33) * 2, -2.9444).astype(np.float32)
img = rng.randn(batch, 33, 33, 33).astype(np.float32)
e.predict(seed, img)
e.enable_profiling(True)
e.profile(reset=True)
e.trace(reset=True)
e.predict(seed, img)
tr = e.trace()
np.save(out, tr)
nconv = 24
ntiles = 2
per_layer = chains * ntiles
names = ['bar_seen', 'load_iss', 'ops_seen', 'slot_got', 'mma_iss', 'acc_seen', 'epi_done']
# second round, layers 10..13
r0 = 1 * nconv * per_layer
t0 = tr[1, r0]
print('chains', chains, 'kernel_us', e.info()['last_kernel_ns'] / 1e3)
print('tile  layer ch j | ' + ' '.join('%9s' % n for n in names))
for i in range(r0 + 10 * per_layer, r0 + 14 * per_layer):
  l = (i - r0) // per_layer
  ch = ((i - r0) % per_layer) // ntiles
  j = (i - r0) % ntiles
  print('%5d %5d %2d %d | ' % (i, l, ch, j) + ' '.join('%9d' % (tr[ev, i] - t0 if tr[ev, i] else -1) for ev in range(7)))
sig0 = 1 * (nconv - 1) * chains
print('signals (layer, chain, time):', [(s // chains % (nconv - 1), s % chains, int(tr[7, s] - t0)) for s in range(sig0 + 10 * chains, sig0 + 14 * chains)])
# per-role summary over the round: average cost per tile
rr = slice(r0 + 2 * per_layer, r0 + 22 * per_layer)
d = lambda a, b: float(np.mean(tr[a, rr] - tr[b, rr]))
print(json.dumps({'load_wait_slot_to_issue': None, 'ops_seen_minus_load_iss': d(2, 1), 'slot_wait': d(3, 2), 'mma_issue': d(4, 3),
                  'acc_seen_minus_mma_iss': d(5, 4), 'epi_body': d(6, 5),
                  'cycles_per_tile': float((tr[6, r0 + 22 * per_layer] - tr[6, r0 + 2 * per_layer]) / (20 * per_layer))}))
e.close()
