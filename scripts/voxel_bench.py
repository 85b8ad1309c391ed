"""Time the GPU voxel grid (DSEC shape, 15 x 480 x 640, mean_std) against the torch CPU path the
reference uses inside its DataLoader workers."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from motionpriorcmax_b200.voxel_grid import VoxelGrid
from oracle import focus_oracle as fo
dev = torch.device('cuda:0')
n = 1_000_000
rng = np.random.default_rng(0)
x = (rng.random(n) * 640).astype(np.float32); y = (rng.random(n) * 480).astype(np.float32)
t = np.sort(rng.random(n)).astype(np.float32); p = (rng.random(n) < 0.5).astype(np.float32)
ev = {k: torch.as_tensor(v, device=dev) for k, v in (('x', x), ('y', y), ('t', t), ('p', p))}
vg = VoxelGrid((15, 480, 640), 'mean_std', 0)
for _ in range(3): vg.convert(ev)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): vg.convert(ev)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
t0 = time.perf_counter(); fo.voxel_grid(x, y, t, p, (15, 480, 640), 'mean_std'); cpu = time.perf_counter() - t0
print(json.dumps({'events': n, 'gpu_ms': ms, 'gpu_events_per_s': n / ms * 1e3, 'cpu_port_s': cpu, 'cpu_events_per_s': n / cpu}))
