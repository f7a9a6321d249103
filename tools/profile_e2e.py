"""Where the end-to-end time of Runner.run goes on the bench's 250^3 canvas: cProfile of one run (after a warm-up run),
top entries by cumulative time.   FFN_B200_DEBUG=128 python tools/profile_e2e.py
"""
import cProfile, io, os, pstats, sys, tempfile, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
import bench
from ffn.inference import runner as runner_mod

shape = (250, 250, 250)
vol = bench.make_volume(shape, 0)
for it in range(2):
  tmp = tempfile.mkdtemp(prefix='ffn_prof_')
  vol_path = os.path.join(tmp, 'vol.npy')
  np.save(vol_path, vol)
  runner = runner_mod.Runner(device=0)
  runner.start(bench.request_for(vol_path, os.path.join(tmp, 'out')))
  torch.cuda.synchronize()
  pr = cProfile.Profile()
  t0 = time.perf_counter()
  pr.enable()
  rc = runner.run((0, 0, 0), shape)
  pr.disable()
  dt = time.perf_counter() - t0
  cnt = {k: c.value for k, c in rc.counters}
  print('run %d: %.3f s, %d steps, %.0f steps/s' % (it, dt, cnt['inference-calls'], cnt['inference-calls'] / dt))
  runner.stop_executor()
  if it == 1:
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(32)
    print(s.getvalue()[:6000])
