"""Does running sub-batches on separate streams overlap the SM-bound K-NN kernels with the
L2-bound event kernels?  Timing probe only (the split losses are not the batch loss)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from motionpriorcmax_b200 import synthetic, trajectories as tj
from motionpriorcmax_b200.losses import LossFactory
import bench

dev = torch.device('cuda:0')
cfg, w = bench.workload('dsec', None, None)
cg_h, ev_h, npos, n_valid = bench.make_inputs(cfg, w, 0)
cg = cg_h.to(dev); ev = ev_h.to(dev)
L = LossFactory.get_loss_calculator('FOCUS', dict(cfg))
times = L.get_reconstruction_times(dev); times[0] = 0.5

def run_parts(nparts, streams, stagger=False):
    B = cg.shape[0]
    bounds = [round(i * B / nparts) for i in range(nparts + 1)]
    parts = []
    for i in range(nparts):
        c = cg[bounds[i]:bounds[i + 1]].clone().requires_grad_()
        parts.append((c, ev[bounds[i]:bounds[i + 1]].contiguous()))
    def once():
        losses = []
        for i, (c, e) in enumerate(parts):
            st = streams[i % len(streams)]
            with torch.cuda.stream(st):
                c.grad = None
                traj = tj.calculate_trajectories_at_t(c, times, 4, 1, 'polynomial')
                loss, _, _ = L.calc(traj, times, {'events': e, 'num_pos_events': npos})
                losses.append((st, loss))
        for st, loss in losses:
            with torch.cuda.stream(st):
                loss.backward()
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 10
    for _ in range(n):
        for st in streams:
            st.wait_stream(torch.cuda.current_stream())
        once()
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

main = torch.cuda.current_stream()
s_hi = torch.cuda.Stream(dev, priority=-1)
s_lo = torch.cuda.Stream(dev, priority=0)
s2 = torch.cuda.Stream(dev); s3 = torch.cuda.Stream(dev)
print('1 part , 1 stream :', round(run_parts(1, [s2]), 3), 'ms')
print('2 parts, 1 stream :', round(run_parts(2, [s2]), 3), 'ms')
print('2 parts, 2 streams:', round(run_parts(2, [s2, s3]), 3), 'ms')
print('2 parts, hi/lo    :', round(run_parts(2, [s_hi, s_lo]), 3), 'ms')
print('4 parts, 2 streams:', round(run_parts(4, [s2, s3]), 3), 'ms')
print('4 parts, 4 streams:', round(run_parts(4, [s2, s3, s_hi, s_lo]), 3), 'ms')
print('7 parts, 2 streams:', round(run_parts(7, [s2, s3]), 3), 'ms')
