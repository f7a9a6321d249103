"""Host-path diagnosis: wall vs device time of segment_at calls, alone and from three threads."""
import json, os, sys, threading, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import bench
from ffn_b200 import _lib, engine as eng

(w, b), _ = bench.load_weights()
vol = bench.make_volume(1)
pts = bench.seed_points(vol, 64)

def calls(e, n, tag):
  cv = eng.DeviceCanvas(e, vol, eng.make_options(), 128.0, 33.0)
  cv.segment_at(pts[0], max_steps=8)
  rows = []
  for k in range(n):
    c0 = cv.counters().device_seconds
    t0 = time.perf_counter()
    st = cv.segment_at(pts[k % len(pts)], reset=True, max_steps=256)
    wall = time.perf_counter() - t0
    rows.append((int(st.iters), round(wall * 1e3, 2), round((cv.counters().device_seconds - c0) * 1e3, 2)))
  t0 = time.perf_counter(); cv.close(); tclose = time.perf_counter() - t0
  print(tag, 'iters/wall_ms/device_ms', rows, 'close_ms', round(tclose * 1e3, 2), flush=True)

e = eng.Engine(w, b, bench.FOV, bench.DELTAS)
t0 = time.perf_counter(); cv = eng.DeviceCanvas(e, vol, eng.make_options(), 128.0, 33.0); print('create_ms', round((time.perf_counter() - t0) * 1e3, 2)); cv.close()
t0 = time.perf_counter(); cv = eng.DeviceCanvas(e, vol, eng.make_options(), 128.0, 33.0); print('create_ms', round((time.perf_counter() - t0) * 1e3, 2)); cv.close()
calls(e, 4, 'grid148')
e.set_grid(49)
calls(e, 4, 'grid49')
es = [e] + [eng.Engine(w, b, bench.FOV, bench.DELTAS) for _ in range(2)]
for x in es:
  x.set_grid(49)
ths = [threading.Thread(target=calls, args=(es[i], 4, 'thread%d' % i)) for i in range(3)]
t0 = time.perf_counter()
[t.start() for t in ths]; [t.join() for t in ths]
print('three threads wall_ms', round((time.perf_counter() - t0) * 1e3, 1))
