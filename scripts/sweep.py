"""Event-count sweep (BASELINE.json configs[4]): loss fwd+bwd events/s vs events per window.
Windows are generated on the device (uniform events, time sorted, loader layout).
    python scripts/sweep.py [--batch 1,14] [--events 1e5,...] > profiles/rXX_sweep.json
Under torchrun (N ranks on one node) every rank runs the same sweep on its own windows; rank 0
prints whole-job rows: time = max over ranks (CUDA events), events = sum over ranks (weak scaling,
no collective on the data path).
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from motionpriorcmax_b200 import synthetic, trajectories as tj, cabi, io as cio
from motionpriorcmax_b200.losses import LossFactory


def device_batch(B, M, H, W, nb, dev, seed):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    npos = M // 2
    ev = torch.empty(B, M, 6, device=dev)
    for b in range(B):
        for s, e, pol in ((0, npos, 1.0), (npos, M, 0.0)):
            n = e - s
            t = torch.rand(n, device=dev, generator=g).sort().values
            t = (t - t[0]) / (t[-1] - t[0])
            ev[b, s:e, 0] = torch.rand(n, device=dev, generator=g) * (H - 1e-3)
            ev[b, s:e, 1] = torch.rand(n, device=dev, generator=g) * (W - 1e-3)
            ev[b, s:e, 2] = t
            ev[b, s:e, 3] = pol
            ev[b, s:e, 4] = torch.clamp((t * nb).floor(), max=nb - 1)
            ev[b, s:e, 5] = 1.0
    return ev, npos


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", default="1,14")
    ap.add_argument("--events", default="1e5,3e5,1e6,3e6,1e7,3e7,5e7")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--layouts", default="plain,packed", help="plain = upstream [B,M,6]; packed = io.PackedEvents")
    a = ap.parse_args()
    real_stdout = os.fdopen(os.dup(1), "w")      # libraries print banners to stdout (NCCL): keep fd 1 for the JSON
    os.dup2(2, 1)
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    cfg = dict(synthetic.DSEC_LOSS_CONFIG)
    H, W = cfg["image_shape"]
    L = LossFactory.get_loss_calculator("FOCUS", dict(cfg))
    times = L.get_reconstruction_times(dev); times[0] = 0.5
    peak = 6553.6
    try:
        peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
    except Exception:
        pass
    rows = []
    for B in [int(x) for x in a.batch.split(",")]:
        cg = synthetic.make_coeff_grid(B, 1, H, W, sigma_px=8.0, seed=1234).to(dev).requires_grad_()
        for M in [int(float(x)) for x in a.events.split(",")]:
            if B * M * 24 > 40e9:
                continue
            ev, npos = device_batch(B, M, H, W, cfg["num_bins"], dev, seed=B * 1000 + 7 + 100000 * rank)

            for layout in a.layouts.split(","):
                batch = {"events": ev, "num_pos_events": npos}
                pack_ms = None
                if layout == "packed":
                    pk = cio.pack_events(ev, npos, L)
                    torch.cuda.synchronize()
                    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    p0.record()
                    pk = cio.pack_events(ev, npos, L)
                    p1.record(); torch.cuda.synchronize()
                    pack_ms = p0.elapsed_time(p1)
                    batch = {"events": pk}

                def step():
                    cg.grad = None
                    traj = tj.calculate_trajectories_at_t(cg, times, 4, 1, "polynomial")
                    loss, _, _ = L.calc(traj, times, batch)
                    loss.backward()
                for _ in range(3):
                    step()
                if dist is not None:
                    dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                cabi.load().cmax_stage_timing_enable(1)
                e0.record()
                for _ in range(a.steps):
                    step()
                e1.record(); torch.cuda.synchronize()
                st = cabi.stage_timing_read()
                cabi.load().cmax_stage_timing_enable(0)
                ms = e0.elapsed_time(e1) / a.steps
                if dist is not None:
                    tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
                    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                    ms = tmax.item()
                n_t, n, Q = 16, 19200, 15 * 19200
                ev_bytes = (32 if layout == "packed" else 48) * B * M      # records are 16 B, read twice
                bytes_alg = ev_bytes + 20 * B * 2 * H * W + 24 * B * Q + 16 * B * n_t * n
                rows.append({"n_gpus": world, "B": B, "events_per_window": M, "layout": layout, "ms_per_step": ms,
                             "events_per_s": world * B * M / ms * 1e3, "algorithmic_GB": bytes_alg / 1e9,
                             "achieved_GBps": bytes_alg / ms / 1e6, "hbm_frac": bytes_alg / ms / 1e6 / peak,
                             "event_forward_ms": st["event_forward"][0] / max(st["event_forward"][1], 1),
                             "event_backward_ms": st["event_backward"][0] / max(st["event_backward"][1], 1),
                             "device_pack_ms": pack_ms})
                if rank == 0:
                    print(json.dumps(rows[-1]), file=sys.stderr)
                batch = None
                pk = None
            del ev
            torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps({"peak_GBps": peak, "n_gpus": world, "rows": rows}, indent=1), file=real_stdout, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
