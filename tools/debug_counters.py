import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ffn_b200 import _lib, engine as eng, tf_checkpoint
G='tests/golden'
w,b=tf_checkpoint.load_convstack_npz(G+'/fib25_convstack.npz')
g=np.load(G+'/flood_fill_64.npz')
objs=json.load(open(G+'/_debug_oracle_pops.json'))
e=eng.Engine(w,b,compute_mode=_lib.COMPUTE_FP32)
seeds=g['seeds']
for k,o in enumerate(objs):
    start=tuple(o['start'])
    idx=[i for i,s in enumerate(seeds) if tuple(s)==start][0]
    cv=eng.DeviceCanvas(e,g['volume'],eng.make_options(),128.0,33.0)
    cv.segment_all(seeds[:idx])
    c0=cv.counters()
    cv.start_trace(4096)
    st=cv.segment_at(start)
    c1=cv.counters()
    q,done,s0=cv.policy_state()
    print(k,start,'iters',st.iters,'dev invalid',c1.skip_invalid_pos-c0.skip_invalid_pos,'thr',c1.skip_threshold-c0.skip_threshold,
          'oracle invalid',sum(p[2] for p in o['pops']),'thr',sum(p[3] for p in o['pops']),'pops',len(o['pops']),'done cells',done.shape[0])
    if k==7:
        for ev in cv.get_trace(): print('   ev', ev.tolist())
    cv.close()
