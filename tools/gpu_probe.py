"""First-contact GPU probe: self tests, network parity, flood-fill parity. Writes gpurun_out/probe.json."""
import json, os, sys, time, traceback
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np

OUT = os.path.join(REPO, 'gpurun_out'); os.makedirs(OUT, exist_ok=True)
G = os.path.join(REPO, 'tests', 'golden')
res = {}

def stage(name):
  def deco(fn):
    t0 = time.time()
    try:
      res[name] = fn()
    except Exception as e:
      res[name] = {'error': repr(e), 'tb': traceback.format_exc()[-1500:]}
    res[name + '_sec'] = round(time.time() - t0, 2)
    print(name, json.dumps(res[name])[:600], flush=True)
    with open(os.path.join(OUT, 'probe.json'), 'w') as f:
      json.dump(res, f, indent=1)
  return deco

from ffn_b200 import _lib, engine as eng, tf_checkpoint
w, b = tf_checkpoint.load_convstack_npz(os.path.join(G, 'fib25_convstack.npz'))
which = sys.argv[1:] or ['selftest', 'fp32', 'tc']

for v, nm in ((0, 'kat'), (1, 'rate_1cta'), (4, 'rate_allsm'), (2, 'barrier'), (3, 'bulk'), (5, 'tmem_shift')):
  if 'selftest%d' % v in which:
    @stage('selftest_' + nm)
    def _():
      return eng.selftest(v, n_out=16)

from oracle import flood_fill as ff
pat = np.load(os.path.join(G, 'net_patches.npz'))
g64 = np.load(os.path.join(G, 'flood_fill_64.npz'))
gat = np.load(os.path.join(G, 'segment_at_64.npz'))

def run_mode(tag, mode):
  E = {}
  @stage(tag + '_engine')
  def _():
    E['e'] = eng.Engine(w, b, compute_mode=mode, num_ctas=int(os.environ.get('FFN_CTAS', '0')))
    return E['e'].info()
  if 'e' not in E:
    return
  e = E['e']

  @stage(tag + '_predict')
  def _():
    out = {}
    t0 = time.time()
    lg = e.predict(pat['seed'], pat['image'])
    out['sec_batch5'] = time.time() - t0
    out['max_abs_vs_fp32'] = float(np.abs(lg - pat['logits_fp32']).max())
    out['max_abs_vs_fp64'] = float(np.abs(lg - pat['logits_fp64']).max())
    out['mean_abs_vs_fp64'] = float(np.abs(lg - pat['logits_fp64']).mean())
    out['finite'] = bool(np.isfinite(lg).all())
    l1 = e.predict(pat['seed'][2], pat['image'][2])
    out['single_eq_batch'] = bool(np.array_equal(l1, lg[2]))
    t0 = time.time()
    for _ in range(20):
      e.predict(pat['seed'][0], pat['image'][0])
    out['ms_per_predict_call'] = (time.time() - t0) / 20 * 1e3
    out['kernel_ns'] = e.info()['last_kernel_ns']
    return out

  image = (g64['volume'].astype(np.float32) - 128.0) / 33.0
  opts = eng.make_options()

  @stage(tag + '_segment_at')
  def _():
    out = {}
    cv = eng.DeviceCanvas(e, g64['volume'], opts, 128.0, 33.0)
    start = tuple(int(v) for v in gat['start'])
    t0 = time.time()
    st = cv.segment_at(start)
    out['sec'] = time.time() - t0
    out['iters'] = int(st.iters); out['golden_iters'] = int(gat['iters'])
    out['device_seconds'] = cv.counters().device_seconds
    seed = cv.read(_lib.ARRAY_SEED)
    gs = gat['seed_canvas']
    out['nan_pattern_equal'] = bool(np.array_equal(np.isnan(seed), np.isnan(gs)))
    both = ~np.isnan(seed) & ~np.isnan(gs)
    out['seed_max_abs_diff'] = float(np.abs(seed[both] - gs[both]).max()) if both.any() else None
    q, d, s0 = cv.policy_state()
    out['queue_left'] = int(q.shape[0]); out['golden_queue_left'] = int(gat['queue'].shape[0])
    # hybrid oracle: reference loop restatement driven by the GPU network
    hyb = ff.Canvas(lambda s, im: e.predict(s, im), image, (33, 33, 33), (8, 8, 8), ff.Options())
    n = hyb.segment_at(start)
    out['hybrid_iters'] = n
    out['hybrid_seed_equal'] = bool(np.array_equal(hyb.seed, seed, equal_nan=True))
    if not out['hybrid_seed_equal']:
      bb = ~np.isnan(seed) & ~np.isnan(hyb.seed)
      out['hybrid_nan_equal'] = bool(np.array_equal(np.isnan(seed), np.isnan(hyb.seed)))
      out['hybrid_max_abs'] = float(np.abs(seed[bb] - hyb.seed[bb]).max()) if bb.any() else None
    out['hybrid_min_margin'] = hyb.min_margin
    out['golden_trace_equal_hybrid'] = bool(np.array_equal(np.asarray(hyb.trace, np.int32).reshape(-1, 3), gat['trace']))
    cv.close()
    e.enable_profiling(True)
    e.profile(reset=True)
    cv2 = eng.DeviceCanvas(e, g64['volume'], opts, 128.0, 33.0)
    st2 = cv2.segment_at(start)
    out['profile'] = e.profile()
    out['profile_device_seconds'] = cv2.counters().device_seconds
    e.enable_profiling(False)
    cv2.close()
    return out

  @stage(tag + '_segment_all')
  def _():
    out = {}
    cv = eng.DeviceCanvas(e, g64['volume'], opts, 128.0, 33.0)
    t0 = time.time()
    origins, overlaps, ctr = cv.segment_all(g64['seeds'])
    out['sec'] = time.time() - t0
    out['device_seconds'] = ctr.device_seconds
    out['steps'] = int(ctr.inference_calls); out['golden_steps'] = int(g64['trace'].shape[0])
    out['segments'] = int(ctr.segments); out['golden_segments'] = int(g64['origins'].shape[0])
    seg = cv.read(_lib.ARRAY_SEGMENTATION)
    out['seg_equal_golden'] = bool(np.array_equal(seg, g64['segmentation']))
    out['seg_mismatch_voxels'] = int((seg != g64['segmentation']).sum())
    qp = cv.read(_lib.ARRAY_QPROB)
    out['qprob_max_abs'] = int(np.abs(qp.astype(int) - g64['seg_prob'].astype(int)).max())
    out['qprob_mismatch'] = int((qp != g64['seg_prob']).sum())
    out['origins'] = [[o.id] + list(o.start_zyx) + [int(o.iters)] for o in origins]
    out['golden_origins'] = g64['origins'].tolist()
    out['counters'] = {k: getattr(ctr, k) for k, _ in ctr._fields_}
    out['golden_counters'] = json.loads(str(g64['counters']))
    hyb = ff.Canvas(lambda s, im: e.predict(s, im), image, (33, 33, 33), (8, 8, 8), ff.Options())
    hyb.segment_all(g64['seeds'])
    out['hybrid_seg_equal'] = bool(np.array_equal(hyb.segmentation, seg))
    out['hybrid_qprob_equal'] = bool(np.array_equal(hyb.seg_prob, qp))
    out['hybrid_qprob_max_abs'] = int(np.abs(hyb.seg_prob.astype(int) - qp.astype(int)).max())
    out['hybrid_steps'] = len(hyb.trace)
    out['hybrid_min_margin'] = hyb.min_margin
    cv.close()
    return out
  e.close()

if 'fp32' in which:
  run_mode('fp32', _lib.COMPUTE_FP32)
if 'tc' in which:
  run_mode('tc', _lib.COMPUTE_FP16_TC)
print('DONE')
