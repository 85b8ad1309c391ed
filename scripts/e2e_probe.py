"""Where the end-to-end step goes: per-iteration copy / compute durations inside the double-buffered
loop (CUDA events on both streams), for 2 and 3 buffers.   python scripts/e2e_probe.py"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from motionpriorcmax_b200 import io as cio, trajectories as tj
from motionpriorcmax_b200.losses import LossFactory

dev = torch.device("cuda:0")
cfg, w = bench.workload("dsec", None, None)
cg_h, ev_h, npos, n_valid = bench.make_inputs(cfg, w, 0)
L = LossFactory.get_loss_calculator("FOCUS", dict(cfg))
times = L.get_reconstruction_times(dev); times[0] = 0.5
comp = cio.pack_events_compact(ev_h, npos, L).pin_memory()
cg_d = cg_h.to(dev)

def step(cg, ev):
    cg = cg.detach().requires_grad_()
    traj = tj.calculate_trajectories_at_t(cg, times, 4, 1, "polynomial")
    loss, _, _ = L.calc(traj, times, {"events": ev})
    loss.backward()
    return loss

out = {}
for nbuf in (2, 3):
    up = cio.CompactUploader(dev, L, n_buffers=nbuf)
    cur = torch.cuda.current_stream(dev)
    K = 24
    cs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K + nbuf)]
    ks = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    pend = []
    def prefetch(j):
        cs[j][0].record(up.stream)
        _, slot = up.upload(comp)
        cs[j][1].record(up.stream)
        pend.append(slot)
    for j in range(nbuf - 1):
        prefetch(j)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(K):
        if it + nbuf - 1 < K + nbuf - 1:
            prefetch(it + nbuf - 1)
        slot = pend.pop(0)
        ks[it][0].record(cur)
        buf = up.wait(slot, cur)
        loss = step(cg_d, buf)
        up.release(slot, cur)
        ks[it][1].record(cur)
        loss.item()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / K * 1e3
    copy = [a.elapsed_time(b) for a, b in cs[4:K]]
    comp_ms = [a.elapsed_time(b) for a, b in ks[4:K]]
    out[f"buffers_{nbuf}"] = {"period_ms": wall, "copy_ms_in_loop": sum(copy) / len(copy),
                              "compute_ms_in_loop_incl_wait": sum(comp_ms) / len(comp_ms)}
# compute alone with the packed events resident
pk = cio.expand_compact(comp.to(dev), L)
for _ in range(3): step(cg_d, pk)
torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): step(cg_d, pk).item()
b.record(); torch.cuda.synchronize()
out["compute_alone_with_item_ms"] = a.elapsed_time(b) / 10
print(json.dumps(out))
