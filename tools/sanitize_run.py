"""Small segment_all for compute-sanitizer (memcheck / racecheck): the golden 64x72x80 volume, first seeds only.
  compute-sanitizer --tool memcheck python tools/sanitize_run.py fp16 12
"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from ffn_b200 import _lib, engine as eng, tf_checkpoint

G = os.path.join(REPO, 'tests', 'golden')
W, B = tf_checkpoint.load_convstack_npz(os.path.join(G, 'fib25_convstack.npz'))
mode = {'fp16': _lib.COMPUTE_FP16_TC, 'fp32': _lib.COMPUTE_FP32, 'x2': _lib.COMPUTE_FP16X2_TC}[sys.argv[1] if len(sys.argv) > 1 else 'fp16']
nseeds = int(sys.argv[2]) if len(sys.argv) > 2 else 12
g = np.load(os.path.join(G, 'flood_fill_64.npz'))
e = eng.Engine(W, B, (33, 33, 33), (8, 8, 8), compute_mode=mode)
cv = eng.DeviceCanvas(e, g['volume'], eng.make_options(), 128.0, 33.0)
origins, overlaps, ctr = cv.segment_all(g['seeds'][:nseeds])
print('mode', sys.argv[1:] , 'steps', ctr.inference_calls, 'segments', ctr.segments, 'spec', cv.spec_stats())
seg = cv.read(_lib.ARRAY_SEGMENTATION)
print('labelled', int((seg > 0).sum()))
cv.close()
e.close()
