"""Profile driver: DSEC-shaped loss step with few events so the LUT-stage kernels dominate."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from motionpriorcmax_b200 import synthetic, trajectories as tj, cabi
from motionpriorcmax_b200.losses import LossFactory
B = int(sys.argv[1]) if len(sys.argv) > 1 else 14
M = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cfg = dict(synthetic.DSEC_LOSS_CONFIG)
dev = torch.device('cuda:0')
L = LossFactory.get_loss_calculator('FOCUS', dict(cfg))
cg = synthetic.make_coeff_grid(B, 1, 480, 640, sigma_px=8.0, seed=1234).to(dev).requires_grad_()
ev, npos = synthetic.make_event_batch(B, M, 480, 640, 15, True, seed=1)
ev = ev.to(dev)
times = L.get_reconstruction_times(dev); times[0] = 0.5
lib = cabi.load()
for i in range(iters + 1):
    if i == 1:
        torch.cuda.synchronize(); lib.cmax_stage_timing_enable(1)
    cg.grad = None
    traj = tj.calculate_trajectories_at_t(cg, times, 4, 1, 'polynomial')
    loss, _, _ = L.calc(traj, times, {'events': ev, 'num_pos_events': npos})
    loss.backward()
torch.cuda.synchronize()
st = cabi.stage_timing_read()
print({k: round(v[0] / max(v[1], 1), 4) for k, v in st.items()})
print('worklist', lib.cmax_last_worklist_count(None), 'of', B * 15 * 19200)
