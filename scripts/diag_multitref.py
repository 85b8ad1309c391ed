import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from test_gpu_parity import _synthetic_case, _run_loss
from helpers import rel_err
from motionpriorcmax_b200 import synthetic
from oracle import focus_oracle as fo
for norm in ('l1', 'l2'):
  for R in (1, 2, 5):
    base = dict(synthetic.DSEC_LOSS_CONFIG, image_shape=(96, 128), num_knn=16, focus_loss_norm=norm)
    if R > 1:
        base = synthetic.multi_tref_variant(base, R)
    else:
        base.update(scale_iwe_by_dt=False, polarity_aware_batching=False)
    traj, times, ev, npos, _ = _synthetic_case(base, 3, [20000, 35000, 9000], 2, seed=11)
    r = _run_loss(base, traj, times, ev, npos)
    o = fo.FocusOracle(**base, dtype=np.float64)
    f = o.forward(traj, times, ev, npos); g = o.backward()
    o32 = fo.FocusOracle(**base, dtype=np.float32)
    f32 = o32.forward(traj, times, ev, npos); g32 = o32.backward()
    d, dr = r['dtraj'], g['dtraj']
    print(norm, R, 'loss', abs(r['loss']-f['loss'])/f['loss'], 'iwe', rel_err(r['iwes'], f['iwes']),
          'dtraj', rel_err(d, dr), 'tref part', rel_err(d[:, :R], dr[:, :R]), 'tmid part', rel_err(d[:, R:], dr[:, R:]),
          'oracle32 vs 64', rel_err(g32['dtraj'], dr))
    for rr in range(R):
        print('   r', rr, rel_err(d[:, rr], dr[:, rr]), 'max', np.abs(d[:, rr]-dr[:, rr]).max(), np.abs(dr[:, rr]).max())
    if R == 5 and norm == 'l1':
        D = g['d_iwe_raw']; 
        e = np.abs(d - dr); idx = np.unravel_index(np.argmax(e), e.shape); print('   worst', idx, d[idx], dr[idx])
