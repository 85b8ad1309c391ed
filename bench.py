#!/usr/bin/env python
"""bench.py - CMax loss forward+backward throughput (events/s) on B200, per BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--events M] [--batch B] [--variant dsec|dsec_tref5|evimo2|evimo2_tref10|k3_det]

One "step" = one pass of the loss hot path over one batch of synthetic event windows:
coeff_grid -> trajectories (fused front end) -> FocusLoss.calc -> backward to d coeff_grid.
Default workload = BASELINE.json configs[1]: DSEC training shape, per-rank batch 14, 480x640,
15 bins, K_nn 32, poly k=1, polarity-aware, as-shipped dsec.yaml loss block, event counts
~LogNormal(1e6, 0.5) per window.

Printed JSON line (rank 0): see the keys below; `value` = valid events of all ranks per second
with inputs resident in HBM (CUDA events, max over ranks); `e2e` = same metric through the
plugin API with pinned HOST inputs copied in (three buffers deep) and the loss read back (4-byte
asynchronous D2H copy, consumed by the host one step later; `e2e.blocking_item` = with a
synchronous `loss.item()` instead), every step inside the timed region - the host buffer is the loader-side `io.BitpackedEvents` (a lossless
~8-byte-per-event bit stream, one cudaMemcpyAsync per step; built by the loader workers outside the
step like the reference's own collate; `e2e.compact_12B_layout` = the uncompressed 12-byte form), `e2e_reference_layout` is the same leg on the reference's padded
`[B, M, 6]` tensor; `train_step` = UNet(15, 2K) forward -> front end -> loss -> backward ->
AdamW under DDP (NCCL all-reduce of the 31 M network gradients inside the timed region);
`roofline` = dominant kernel (by measured stage time) against the HBM peak of
MEASURED_PEAKS.json; `cpu_baseline` = the CPU oracle port timed on the host cores on a bounded
sample (one window).

`--impl reference` times the reference's algorithm on the CPU (the oracle port: the reference
is Python + pykeops and cannot run on the box) on the same config and metric.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


# ------------------------------------------------------------------------------------------
# workload definitions
# ------------------------------------------------------------------------------------------
def workload(variant: str, batch: int | None, events: int | None):
    from motionpriorcmax_b200 import synthetic
    if variant == "dsec":
        cfg = dict(synthetic.DSEC_LOSS_CONFIG)
        w = dict(name="dsec_train_b14_480x640_polyk1_pab", B=batch or 14, K=1, basis="polynomial",
                 median=events or 1_000_000, lognormal=events is None, deterministic=False)
    elif variant == "dsec_tref5":
        cfg = synthetic.multi_tref_variant(synthetic.DSEC_LOSS_CONFIG, 5)
        w = dict(name="dsec_train_b14_480x640_polyk1_tref5", B=batch or 14, K=1, basis="polynomial",
                 median=events or 1_000_000, lognormal=events is None, deterministic=False)
    elif variant == "evimo2":
        cfg = dict(synthetic.EVIMO2_LOSS_CONFIG)
        w = dict(name="evimo2_300ms_b6_384x512_bezier10", B=batch or 6, K=10, basis="bezier",
                 median=events or 1_000_000, lognormal=False, deterministic=False, integer=True)
    elif variant == "evimo2_tref10":
        # the "10 reference times" variant of the EVIMO2 configuration (BASELINE.json configs[2];
        # experiment yaml :21-35 with the only multi-reference combination focus.py:49-51 allows)
        cfg = synthetic.multi_tref_variant(synthetic.EVIMO2_LOSS_CONFIG, 10)
        w = dict(name="evimo2_300ms_b6_384x512_bezier10_tref10", B=batch or 6, K=10, basis="bezier",
                 median=events or 1_000_000, lognormal=False, deterministic=False, integer=True)
    elif variant == "k3_det":
        cfg = dict(synthetic.DSEC_LOSS_CONFIG)
        w = dict(name="dsec_k3_5Mevents_deterministic", B=batch or 1, K=3, basis="polynomial",
                 median=events or 5_000_000, lognormal=False, deterministic=True)
    else:
        raise ValueError(variant)
    return cfg, w


def make_inputs(cfg, w, rank: int):
    """CPU tensors of one batch: coeff_grid [B,1,2K,H,W], events [B,M,6], num_pos, times."""
    from motionpriorcmax_b200 import synthetic
    H, W = cfg["image_shape"]
    B = w["B"]
    if w["lognormal"]:
        counts = synthetic.lognormal_event_counts(B, median=w["median"], rank=rank)
    else:
        counts = [int(w["median"])] * B
    ev, npos = synthetic.make_event_batch(B, counts, H, W, cfg["num_bins"],
                                          cfg["polarity_aware_batching"], seed=1234, rank=rank,
                                          dist=w.get("dist", "uniform"),
                                          integer_coords=w.get("integer", False), coord_scale=0.8)
    cg = synthetic.make_coeff_grid(B, w["K"], H, W, sigma_px=8.0, seed=1234 + 1000 * rank)
    n_valid = int(ev[..., 5].sum().item())
    return cg, ev, npos, n_valid


def algorithmic_bytes(cfg, w, B, M, n):
    """bytes_alg = 48 E + 20 B R P H W + 24 B Q R + 16 B n_t n."""
    H, W = cfg["image_shape"]
    R, nb, s = cfg["num_tref"], cfg["num_bins"], cfg["lut_superpixel_size"]
    P = 2 if cfg["polarity_aware_batching"] else 1
    Q = nb * math.ceil(H / s) * math.ceil(W / s)
    return dict(events=48 * B * M, image=20 * B * R * P * H * W, lut=24 * B * Q * R,
                traj=16 * B * (R + nb) * n)


# ------------------------------------------------------------------------------------------
# clocks sampler (pynvml; nvidia-smi fallback)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
               0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clocks"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) \
                    if hasattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name not in ("gpu_idle",):
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.ok:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread:
            self._stop.set()
            self._thread.join()
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline (oracle port)
# ------------------------------------------------------------------------------------------
def cpu_reference_step(cfg, w, cg, ev, npos, sample_windows=1, split=False):
    """One loss forward+backward of the CPU oracle on the first `sample_windows` windows.
    Returns (seconds, valid events processed[, seconds of the LUT stage's K-NN search alone])."""
    from oracle import focus_oracle as fo
    B = min(sample_windows, ev.shape[0])
    evs = ev[:B].numpy()
    times = fo.reconstruction_times(cfg["num_tref"], cfg["num_bins"], 0.5)
    t0 = time.perf_counter()
    traj, _ = fo.trajectories_from_coeff_grid(cg[:B].numpy(), times, 4, w["K"], w["basis"])
    o = fo.FocusOracle(**cfg, dtype=np.float32)
    o.forward(traj, times, evs, -1 if npos is None else npos)
    g = o.backward()
    fo.trajectories_backward(g["dtraj"], times, 4, w["K"], w["basis"], tuple(cg[:B].shape),
                             dtype=np.float32)
    dt = time.perf_counter() - t0
    if not split:
        return dt, int(evs[..., 5].sum())
    # the fixed LUT-stage cost (exhaustive K-NN of interpolate_flow, focus.py:115-180) on its own:
    # timed on one window and scaled (the exhaustive search is exactly linear in the windows)
    grid, _, _ = fo.lut_grid(cfg["image_shape"], cfg["lut_superpixel_size"])
    t1 = time.perf_counter()
    fo.knn_bruteforce(np.asarray(traj, np.float32)[:1, cfg["num_tref"]:], grid, cfg["num_knn"], cfg["dist_norm"], True)
    return dt, int(evs[..., 5].sum()), (time.perf_counter() - t1) * B


# The reference arm uses every host core: the windows of the batch are independent up to the final
# mean, so a pool of worker processes takes one window each (numpy event / image stage, 1 thread
# per worker) and the exhaustive K-NN inside each worker gets the remaining cores.
_REF = {}


def _ref_init(variant, batch, events, knn_threads):
    os.environ["ORACLE_KNN_THREADS"] = str(knn_threads)
    cfg, w = workload(variant, batch, events)
    cg, ev, npos, _ = make_inputs(cfg, w, 0)                    # the same batch rank 0 of our arm gets
    _REF.update(cfg=cfg, w=w, cg=cg, ev=ev, npos=npos)


def _ref_ready(_):
    time.sleep(0.2)                                             # let every worker take one
    return os.getpid()


def _ref_window(i):
    """(seconds, valid events, seconds inside the exhaustive K-NN search) of window i."""
    from oracle import focus_oracle as fo
    r = _REF
    if "knn_orig" not in r:                                     # time the search where it is called
        r["knn_orig"] = fo.knn_bruteforce

        def timed(*a, **k):
            t = time.perf_counter()
            out = r["knn_orig"](*a, **k)
            r["knn_s"] += time.perf_counter() - t
            return out
        fo.knn_bruteforce = timed
    r["knn_s"] = 0.0
    dt, n_ev = cpu_reference_step(r["cfg"], r["w"], r["cg"][i:i + 1], r["ev"][i:i + 1], r["npos"],
                                  sample_windows=1)
    return dt, n_ev, r["knn_s"]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    cfg, w = workload(args.variant, args.batch, args.events)
    B = w["B"]
    cores = os.cpu_count() or 1
    workers = max(1, min(B, cores))
    knn_threads = max(1, cores // workers)
    warm = max(0, min(args.warmup, 1))
    steps = max(1, min(args.steps, 2))                          # a 14-window step takes tens of seconds
    with ProcessPoolExecutor(workers, mp_context=mp.get_context("spawn"), initializer=_ref_init,
                             initargs=(args.variant, args.batch, args.events, knn_threads)) as pool:
        list(pool.map(_ref_ready, range(workers)))              # all workers up, inputs built
        if warm:
            list(pool.map(_ref_window, range(min(B, workers))))
        secs, n_ev, knn_s = [], 0, 0.0
        for _ in range(steps):
            t0 = time.perf_counter()
            res = list(pool.map(_ref_window, range(B)))
            dt = time.perf_counter() - t0
            secs.append(dt)
            n_ev = sum(r[1] for r in res)
            # share of the wall time the K-NN search takes (busy seconds of the workers)
            knn_s = dt * sum(r[2] for r in res) / max(sum(r[0] for r in res), 1e-9)
    t = statistics.median(secs)
    val = n_ev / t
    line = {
        "impl": "reference", "metric": "cmax_loss_fwd_bwd_events_per_sec", "value": val,
        "unit": "events/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["name"], "per_rank_batch": B, "sample": "the full batch of rank 0 per step",
                   "events_per_step": n_ev},
        "cpu_baseline": {"value": val, "unit": "events/s", "cores": cores, "kind": "port",
                         "lut_stage_knn_s": knn_s, "event_and_image_stage_s": max(t - knn_s, 0.0),
                         "sample": f"{B} windows ({n_ev} events) fwd+bwd per step, oracle port of the reference "
                                   f"loss: {workers} worker processes take one window each (numpy event / image "
                                   f"stage) with {knn_threads} OpenMP thread(s) each for the exhaustive KNN; the "
                                   f"Python+pykeops reference cannot run on the box"},
        "e2e": {"value": val, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------
# multi-rank plumbing (no collective on the data path: windows are sharded, ranks independent)
# ------------------------------------------------------------------------------------------
def aggregate_over_ranks(ms_total, ms_e2e, n_valid, n_rows, dist, device):
    """Whole-job numbers: time = MAX over ranks, work = SUM over ranks (weak scaling).
    `dist` is torch.distributed (initialised) or None for a single process."""
    t = torch.tensor([ms_total, ms_e2e, float(n_valid), float(n_rows)], device=device,
                     dtype=torch.float64)
    if dist is None:
        return ms_total, ms_e2e, float(n_valid), float(n_rows)
    tmax, tsum = t.clone(), t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    return tmax[0].item(), tmax[1].item(), tsum[2].item(), tsum[3].item()


# ------------------------------------------------------------------------------------------
# the network around the loss (stock PyTorch, as the north star prescribes): a 5-level UNet of
# the shape of upstream src/models/unet/unet_model.py:6-36 (double 3x3 conv + BN + ReLU per level,
# max-pool down, transposed-conv up, 64..1024 channels, 31.0 M parameters for 15 -> 2 channels)
# ------------------------------------------------------------------------------------------
def make_unet(n_in: int, n_out: int):
    import torch.nn as nn

    def block(ci, co):
        return nn.Sequential(nn.Conv2d(ci, co, 3, padding=1, bias=False), nn.BatchNorm2d(co), nn.ReLU(inplace=True),
                             nn.Conv2d(co, co, 3, padding=1, bias=False), nn.BatchNorm2d(co), nn.ReLU(inplace=True))

    class UNet5(nn.Module):
        def __init__(self):
            super().__init__()
            w = [64, 128, 256, 512, 1024]
            self.enc = nn.ModuleList([block(n_in, w[0])] + [block(w[i], w[i + 1]) for i in range(4)])
            self.pool = nn.MaxPool2d(2)
            self.up = nn.ModuleList([nn.ConvTranspose2d(w[i + 1], w[i], 2, stride=2) for i in (3, 2, 1, 0)])
            self.dec = nn.ModuleList([block(2 * w[i], w[i]) for i in (3, 2, 1, 0)])
            self.head = nn.Conv2d(w[0], n_out, 1)

        def forward(self, x):
            skips = []
            for i, e in enumerate(self.enc):
                x = e(x if i == 0 else self.pool(x))
                skips.append(x)
            skips.pop()
            for up, dec in zip(self.up, self.dec):
                x = dec(torch.cat((skips.pop(), up(x)), 1))
            return self.head(x)

    return UNet5()


def train_step_leg(args, dev, dist, world, local, cfg, w, L, times, ev_d, batch_keys, lib):
    """One DSEC training step per iteration: UNet -> coeff grid -> trajectories -> CMax loss ->
    backward -> (DDP: NCCL all-reduce of the network gradients, bucketed, overlapped with the
    backward) -> AdamW.  scripts/flow_training.py:125-130, src/modules/trajectory_net.py:142-170."""
    from motionpriorcmax_b200 import cabi, trajectories as tj
    H, W = cfg["image_shape"]
    B, K = w["B"], w["K"]
    torch.manual_seed(1234)
    net = make_unet(cfg["num_bins"], 2 * K).to(dev)
    n_params = sum(p.numel() for p in net.parameters())
    model = net
    if world > 1:
        from torch.nn.parallel import DistributedDataParallel as DDP
        model = DDP(net, device_ids=[local], gradient_as_bucket_view=True)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)                # trajectory_net.py:213-219
    voxel = torch.randn((B, cfg["num_bins"], H, W), device=dev)

    def tstep():
        opt.zero_grad(set_to_none=True)
        cg = model(voxel)[:, None]                                       # [B, 1, 2K, H, W]
        traj = tj.calculate_trajectories_at_t(cg, times, 4, K, w["basis"])
        loss, _, _ = L.calc(traj, times, dict(batch_keys, events=ev_d))
        loss.backward()
        opt.step()
        return loss

    steps = max(1, min(args.steps, 10))
    for _ in range(3):
        tstep()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    lib.cmax_stage_timing_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tstep()
    e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    st = cabi.stage_timing_read()
    lib.cmax_stage_timing_enable(0)
    loss_ms = sum(v[0] for v in st.values()) / steps
    del model, net, opt, voxel
    torch.cuda.empty_cache()
    return {"ms_per_step": ms, "steps": steps, "loss_kernels_ms_per_step": loss_ms,
            "network_parameters": n_params, "allreduce_bytes_per_step": n_params * 4 if world > 1 else 0}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    from motionpriorcmax_b200 import cabi, trajectories as tj
    from motionpriorcmax_b200.losses import LossFactory

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback for the product path)"
    if world > 1:       # N ranks generate their synthetic windows at the same time: share the cores
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # host placement before any pinned allocation: this rank's staging buffers next to its GPU
    from motionpriorcmax_b200.io import bind_host_to_device_numa
    numa = {"bound": False, "skipped": "--no-numa"} if args.no_numa else bind_host_to_device_numa(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = cabi.load()

    cfg, w = workload(args.variant, args.batch, args.events)
    if args.dist != "uniform":
        w["dist"] = args.dist
        w["name"] += "_" + args.dist
    H, W = cfg["image_shape"]
    cg_h, ev_h, npos, n_valid = make_inputs(cfg, w, rank)
    B, M = ev_h.shape[0], ev_h.shape[1]
    L = LossFactory.get_loss_calculator("FOCUS", dict(cfg, deterministic=w["deterministic"]))
    times = L.get_reconstruction_times(dev)
    if cfg["num_tref"] == 1:
        times[0] = 0.5
    batch_keys = {} if npos is None else {"num_pos_events": npos}

    def step(cg_d, ev_d):
        cg_d.grad = None
        traj = tj.calculate_trajectories_at_t(cg_d, times, 4, w["K"], w["basis"])
        loss, _, _ = L.calc(traj, times, dict(batch_keys, events=ev_d))
        loss.backward()
        return loss

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm ----------------------------------------------------------------
    cg_d = cg_h.to(dev).requires_grad_()
    ev_d = ev_h.to(dev)
    n_traj = tj.tile_positions((H, W), 4).shape[0]
    for _ in range(args.prof_warmup if args.prof_warmup is not None else max(3, args.warmup)):
        step(cg_d, ev_d)
    barrier()
    launches0 = lib.cmax_launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    lib.cmax_stage_timing_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(cg_d, ev_d)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    stage = cabi.stage_timing_read()
    worklist = cabi.worklist_reasons(cabi.stream_ptr(dev))      # of the last step's K-NN stage
    lib.cmax_stage_timing_enable(0)
    launches = lib.cmax_launch_count() - launches0

    # ---- packed (tile-binned, loader-side) event layout: same step, event stage in shared memory --
    packed = None
    if not args.no_packed:
        from motionpriorcmax_b200 import io as cio
        pk_d = cio.pack_events(ev_d, npos, L)
        for _ in range(3):
            step(cg_d, pk_d)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        lib.cmax_stage_timing_enable(1)            # the three pack kernels alone (no allocator time)
        for _ in range(5):
            cio.pack_events(ev_d, npos, L)
        barrier()
        ms_pack = cabi.stage_timing_read()["pack_events"][0] / 5
        lib.cmax_stage_timing_enable(1)
        barrier()
        p0.record()
        for _ in range(args.steps):
            step(cg_d, pk_d)
        p1.record()
        barrier()
        ms_packed = p0.elapsed_time(p1)
        stage_p = cabi.stage_timing_read()
        lib.cmax_stage_timing_enable(0)
        packed = {"ms_total": ms_packed, "ms_pack": ms_pack,
                  "stage": {k: v[0] / v[1] for k, v in stage_p.items() if v[1] > 0}}

    # ---- end-to-end arms: pinned host inputs, double-buffered H2D, loss read back ---------------
    # (a) headline: the loader-side compact layout (io.CompactEvents: 12 B per valid event, built by
    #     the loader workers outside the step) - ONE copy per step + the expand kernel;
    # (b) the reference's own padded [B, M, 6] tensor through io.EventUploader (valid prefixes only).
    from motionpriorcmax_b200 import io as cio
    from motionpriorcmax_b200.io import EventUploader
    cg_p = cg_h.pin_memory()
    NBUF = 3            # two copies in flight / queued behind the step that computes (measured with
                        # scripts/e2e_probe.py: with 2 buffers the copy of step i+1 waits for the slot of
                        # step i-1 and the period is copy + part of the compute, with 3 it is the copy)
    cg_bufs = [cg_d.detach().clone() for _ in range(NBUF)]
    cg_ready = [torch.cuda.Event() for _ in range(NBUF)]

    loss_host = [torch.empty(1).pin_memory() for _ in range(NBUF)]
    loss_ev = [torch.cuda.Event() for _ in range(NBUF)]

    def run_e2e(upload, stream_of, wait, release, k, cg_from_host, blocking):
        """Pipelined loop: the copies of steps i+1 and i+2 are queued while step i computes.
        cg_from_host: also copy the coefficient grid (the NETWORK's output, 34 MB) in every step.
        Every step's loss is read on the host: blocking = `loss.item()` right after the step (the
        GPU then waits for the host to enqueue the next step); otherwise the loss goes D2H with an
        asynchronous copy into pinned memory and is consumed one step later, as a training loop that
        logs its loss does."""
        cur = torch.cuda.current_stream(dev)
        pending = {}

        def prefetch(j):
            i = j % NBUF
            buf, slot = upload()
            with torch.cuda.stream(stream_of):
                if cg_from_host:
                    cg_bufs[i].copy_(cg_p, non_blocking=True)
                cg_ready[i].record(stream_of)
            pending[j] = (buf, slot)

        for j in range(min(NBUF - 1, k)):
            prefetch(j)
        out = 0.0
        for it in range(k):
            if it + NBUF - 1 < k:
                prefetch(it + NBUF - 1)
            buf, slot = pending.pop(it)
            got = wait(slot, cur)                   # CompactUploader: the expanded PackedEvents
            buf = buf if got is None else got
            i = it % NBUF
            cur.wait_event(cg_ready[i])
            loss = step(cg_bufs[i].requires_grad_(), buf)
            release(slot, cur)
            if blocking:
                out = loss.item()                   # D2H read of the step's result, synchronous
            else:
                loss_host[i].copy_(loss.detach().reshape(1), non_blocking=True)
                loss_ev[i].record(cur)
                if it >= 1:                         # the previous step's loss has landed by now
                    loss_ev[(it - 1) % NBUF].synchronize()
                    out = float(loss_host[(it - 1) % NBUF])
            cg_bufs[i].requires_grad_(False)
        if not blocking and k >= 1:
            loss_ev[(k - 1) % NBUF].synchronize()
            out = float(loss_host[(k - 1) % NBUF])
        return out

    def time_e2e(upload, up, cg_from_host=False, blocking=False):
        run_e2e(upload, up.stream, up.wait, up.release, 4, cg_from_host, blocking)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        run_e2e(upload, up.stream, up.wait, up.release, args.steps, cg_from_host, blocking)
        f1.record()
        barrier()
        return f0.elapsed_time(f1)

    ms_e2e = ms_e2e_ref = float("nan")
    e2e_info = {}
    if not args.no_e2e:
        t0 = time.perf_counter()
        wire_h = cio.pack_events_bitpacked(ev_h, npos, L)           # loader side, outside the step
        pack_s = time.perf_counter() - t0
        try:
            wire_p = wire_h.pin_memory(write_combined=True)
            e2e_host_mem = "cudaHostAllocWriteCombined"
        except Exception:
            wire_p = wire_h.pin_memory()
            e2e_host_mem = "pinned"
        cup = cio.CompactUploader(dev, L, n_buffers=NBUF)
        ms_e2e = time_e2e(lambda: cup.upload(wire_p), cup)
        ms_e2e_blocking = time_e2e(lambda: cup.upload(wire_p), cup, blocking=True)
        ms_e2e_cg = time_e2e(lambda: cup.upload(wire_p), cup, cg_from_host=True)
        # copy alone (no compute in flight): what PCIe gives this rank
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(cup.stream)
        for _ in range(5):
            _, sl = cup.upload(wire_p)
            cup.release(sl, cup.stream)
        c1.record(cup.stream)
        barrier()
        copy_ms = c0.elapsed_time(c1) / 5
        e2e_info = {"h2d_bytes": int(cup.bytes_last), "host_issue_ms": cup.issue_ms_last,
                    "copy_alone_ms": copy_ms, "host_pack_s_per_batch": pack_s,
                    "cg_bytes": int(cg_p.numel() * 4), "ms_with_cg": ms_e2e_cg, "ms_blocking": ms_e2e_blocking}
        # the uncompressed 12-byte wire layout for comparison
        comp_p = cio.pack_events_compact(ev_h, npos, L).pin_memory()
        cup12 = cio.CompactUploader(dev, L, n_buffers=NBUF)
        e2e_info["ms_compact12"] = time_e2e(lambda: cup12.upload(comp_p), cup12)
        e2e_info["compact12_bytes"] = int(cup12.bytes_last)
        del comp_p, cup12
        ev_p = ev_h.pin_memory()
        up = EventUploader(dev, n_buffers=NBUF)
        ms_e2e_ref = time_e2e(lambda: up.upload(ev_p, npos), up)
        e2e_info["ref_h2d_bytes"] = int(up.bytes_last)
        del ev_p

    # ---- training-step leg: the network and the DDP all-reduce around the loss ---------------------
    train = None
    if not args.no_train:
        train = train_step_leg(args, dev, dist, world, local, cfg, w, L, times, ev_d, batch_keys, lib)
    clocks = sampler.stop()

    # ---- max over ranks -----------------------------------------------------------------------
    ms_total, ms_e2e, ev_all, rows_all = aggregate_over_ranks(ms_total, ms_e2e, n_valid, B * M, dist, dev)
    extra = torch.tensor([ms_e2e_ref, train["ms_per_step"] if train else float("nan"),
                          e2e_info.get("copy_alone_ms", float("nan")), e2e_info.get("host_issue_ms", float("nan"))],
                         device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(extra, op=dist.ReduceOp.MAX)
    ms_e2e_ref, ms_train, copy_alone_ms, host_issue_ms = extra.tolist()

    if rank == 0:
        ms_step = ms_total / args.steps
        value = ev_all / (ms_step * 1e-3)
        e2e_val = ev_all / (ms_e2e / args.steps * 1e-3)
        # dominant kernel by measured stage time (this rank)
        per_launch = {k: (v[0] / v[1]) for k, v in stage.items() if v[1] > 0}
        dom = max(per_launch, key=per_launch.get)
        ab = algorithmic_bytes(cfg, w, B, M, n_traj)
        stage_bytes = {
            "event_forward": ab["events"] / 2, "event_backward": ab["events"] / 2,
            "image_forward": ab["image"] * 12 / 20, "image_backward": ab["image"] * 8 / 20,
            "knn_select": ab["lut"] / 3 + ab["traj"] / 2, "lut_backward": ab["lut"] / 3 + ab["traj"] / 2,
            "smooth_forward": ab["lut"] / 3, "smooth_backward": 2 * ab["lut"] / 3,
            "bin_points": ab["traj"] / 2,
        }
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        dom_bytes = stage_bytes.get(dom, 0.0)
        achieved = dom_bytes / (per_launch[dom] * 1e-3) / 1e9
        total_bytes = sum(ab.values())
        traffic, issue = None, None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")))
            traffic = prof.get(dom)
            issue = prof.get("_issue", {}).get(dom)         # ncu: the compute-side ceiling of this stage
        except Exception:
            pass
        stage_kernels = {
            "knn_select": "knn_fast_kernel x num_bins (per-bin launches) + knn_warp_kernel + knn_heap_kernel "
                          "[+ lut_accumulate_kernel]",
            "lut_backward": "lut_backward_tile_kernel + lut_backward_assemble_kernel",
            "event_forward": "event_forward_kernel", "event_backward": "event_backward_kernel"}
        # the event kernels are the HBM/atomic-bound ones: report them against both ceilings
        # atomic side of the roofline: the L2 atomic unit retires ~192 G REQUESTS/s whatever their
        # width (scalar, v2, v4 all measured: profiles/r02_atomic_microbench.json), so the bound
        # counts requests: forward 2 per valid event when both corner pairs go out as one vector red
        # each (2.5 measured: unaligned pairs split), backward 1 (the (g_y, g_x) pair)
        r_atomic = None
        try:
            mb = json.load(open(os.path.join(ROOT, "profiles", "r02_atomic_microbench.json")))
            r_atomic = mb["red_global_v2_f32/batch14_pab_34MB"]["Grequests_per_s"]
        except Exception:
            pass
        ev_roof = {}
        for k, reqs_per_event in (("event_forward", 2.5), ("event_backward", 1.0)):
            if k in per_launch:
                t = per_launch[k] * 1e-3
                ev_roof[k] = {"ms": per_launch[k], "hbm_GBps": stage_bytes[k] / t / 1e9,
                              "hbm_frac": stage_bytes[k] / t / 1e9 / peak,
                              "red_requests_per_s_G": reqs_per_event * n_valid / t / 1e9,
                              "atomic_peak_G": r_atomic,
                              "atomic_frac": (reqs_per_event * n_valid / t / 1e9 / r_atomic) if r_atomic else None}
        line = {
            "metric": "cmax_loss_fwd_bwd_events_per_sec", "value": value, "unit": "events/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "per_rank_batch": B, "event_rows_per_window": M,
                       "events_per_step_all_ranks": ev_all, "event_rows_per_step_all_ranks": rows_all,
                       "image": [H, W], "num_bins": cfg["num_bins"], "num_tref": cfg["num_tref"],
                       "num_knn": cfg["num_knn"], "trajectories": n_traj, "basis": w["basis"],
                       "basis_order": w["K"], "deterministic": w["deterministic"],
                       "l2_policy": "inputs larger than L2 (events %.0f MB per rank)" % (B * M * 24 / 1e6),
                       "parallelism": f"dp{world} (windows sharded, no collective in the loss)"},
            "roofline": {"bound": "hbm", "kernel": dom, "kernels_in_stage": stage_kernels.get(dom, dom),
                         "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                         "algorithmic_bytes_per_launch": dom_bytes,
                         "whole_step": {"algorithmic_bytes": total_bytes,
                                        "achieved": total_bytes / (ms_step * 1e-3) / 1e9,
                                        "frac": total_bytes / (ms_step * 1e-3) / 1e9 / peak,
                                        "hbm_bound_ms": total_bytes / peak / 1e6,
                                        "atomic_requests": 3.0 * n_valid,
                                        "atomic_bound_ms": (3.0 * n_valid / r_atomic / 1e6) if r_atomic else None,
                                        "combined_bound_ms": max(total_bytes / peak / 1e6,
                                                                 (3.0 * n_valid / r_atomic / 1e6) if r_atomic else 0.0),
                                        "frac_of_combined_bound": max(total_bytes / peak / 1e6,
                                                                      (3.0 * n_valid / r_atomic / 1e6) if r_atomic else 0.0)
                                        / ms_step,
                                        "note": "combined = max(HBM time of the algorithmic bytes, L2-atomic time of "
                                                "3 vector-red requests per valid event at the measured request rate)"},
                         "event_kernels": ev_roof,
                         "issue_bound_ncu": issue,
                         "note": "the dominant stage is the exact K-NN LUT build: a fixed per-window "
                                 "cost that is instruction/latency bound, not HBM bound",
                         "stage_ms_per_launch": per_launch,
                         "knn_worklist_cells": worklist},
            "e2e": {"value": e2e_val, "unit": "events/s",
                    "h2d_bytes_per_step": e2e_info.get("h2d_bytes", 0),
                    "h2d_note": "the step's host input = the event windows the loader delivers, as io.BitpackedEvents: "
                                "a LOSSLESS bit stream of per-run bit-pattern deltas of (y, x, t), ~8.2 B per valid "
                                "event + run tables, one cudaMemcpyAsync per step from pinned memory, decoded to "
                                "16-byte records on the device (cmax_expand_bitpacked); built by the loader "
                                "workers outside the step.  The coefficient grid is the network's output and stays "
                                "on the device (see train_step); with_coeff_grid_from_host ships it over PCIe as "
                                "well, every step; compact_12B_layout is the same leg on the uncompressed 12-byte "
                                "wire layout (io.CompactEvents)",
                    "compact_12B_layout": {
                        "ms_per_step_rank0": e2e_info.get("ms_compact12", float("nan")) / args.steps,
                        "h2d_bytes_per_step": e2e_info.get("compact12_bytes", 0)},
                    "loss_readback": "every step's loss is copied device -> host (4 B, asynchronous copy into pinned "
                                     "memory) and read by the host one step later; blocking_item = the same leg with a "
                                     "synchronous loss.item() after every step (the GPU then idles while the host "
                                     "enqueues the next step's ~30 launches)",
                    "blocking_item": {"ms_per_step_rank0": e2e_info.get("ms_blocking", float("nan")) / args.steps},
                    "with_coeff_grid_from_host": {
                        "ms_per_step_rank0": e2e_info.get("ms_with_cg", float("nan")) / args.steps,
                        "h2d_bytes_per_step": e2e_info.get("h2d_bytes", 0) + e2e_info.get("cg_bytes", 0)},
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                    "h2d_GBps_per_rank_copy_alone": (e2e_info.get("h2d_bytes", 0) / (copy_alone_ms * 1e-3) / 1e9)
                    if copy_alone_ms == copy_alone_ms else None,
                    "copy_alone_ms": copy_alone_ms, "host_issue_ms_per_step": host_issue_ms,
                    "host_pack_s_per_batch_rank0": e2e_info.get("host_pack_s_per_batch"),
                    "host_numa_binding_rank0": numa, "host_buffer": e2e_host_mem if not args.no_e2e else None},
            "e2e_reference_layout": {
                "value": ev_all / (ms_e2e_ref / args.steps * 1e-3), "unit": "events/s",
                "ms_per_step": ms_e2e_ref / args.steps, "h2d_bytes_per_step": e2e_info.get("ref_h2d_bytes", 0),
                "note": "the reference's padded [B, M, 6] tensor from pinned memory through io.EventUploader "
                        "(valid prefixes only, one copy per window and polarity group)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if train is not None:
            line["train_step"] = {
                "what": "UNet(15, 2K) 31 M parameters (own definition of the shape of upstream "
                        "src/models/unet/unet_model.py) -> front end -> CMax loss -> backward -> AdamW, per-rank "
                        "batch %d, fp32 parameters, TF32 tensor-core convolutions (PyTorch default); under DDP the "
                        "NCCL all-reduce of the network gradients is inside the timed region" % B,
                "ms_per_step": ms_train, "steps": train["steps"],
                "windows_per_s_all_ranks": world * B / (ms_train * 1e-3),
                "events_per_s_all_ranks": ev_all / (ms_train * 1e-3),
                "loss_kernels_ms_per_step_rank0": train["loss_kernels_ms_per_step"],
                "loss_share_of_step": train["loss_kernels_ms_per_step"] / ms_train,
                "network_parameters": train["network_parameters"],
                "allreduce_bytes_per_step": train["allreduce_bytes_per_step"],
                "collective": "torch DDP / NCCL all-reduce (bucketed, overlapped with backward)" if world > 1 else "none (1 GPU)"}
        if packed is not None:
            # per-rank numbers of rank 0 (ranks are independent; the headline keys above are max-over-ranks)
            ms_p = packed["ms_total"] / args.steps
            line["packed_layout"] = {
                "what": "same step with batch['events'] = io.PackedEvents (16 B records of the valid events, "
                        "pre-binned by 32x32 px source tile: SURVEY 8f rank 2); event stage accumulates in "
                        "shared memory; binning done once per window outside the step (loader side)",
                "value_rank0": n_valid / (ms_p * 1e-3), "ms_per_step": ms_p, "unit": "events/s",
                "device_pack_ms": packed["ms_pack"],
                "value_rank0_with_device_pack_each_step": n_valid / ((ms_p + packed["ms_pack"]) * 1e-3),
                "stage_ms_per_launch": packed["stage"]}
        if world == 1 and not args.no_cpu:
            dt, n_ev, knn_s = cpu_reference_step(cfg, w, cg_h, ev_h, npos, split=True)
            cores = os.cpu_count() or 1
            line["cpu_baseline"] = {
                "value": n_ev / dt, "unit": "events/s", "cores": cores, "kind": "port",
                "lut_stage_knn_s": knn_s, "event_and_image_stage_s": max(dt - knn_s, 0.0),
                "sample": f"1 window ({n_ev} events) fwd+bwd in {dt:.2f} s: numpy event stage "
                          f"(1 thread) + C/OpenMP exhaustive KNN ({cores} threads, {knn_s:.2f} s of it)"}
        emit(line)
    if dist is not None:
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception as exc:            # never lose the printed line to a teardown problem
            print(f"process-group teardown: {exc}", file=sys.stderr)


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line of the contract, on the process's original stdout."""
    out = _REAL_STDOUT or sys.stdout
    print(json.dumps(line), file=out, flush=True)


def main():
    global _REAL_STDOUT
    # Libraries print banners to stdout (NCCL: "NCCL version ..."): keep fd 1 for the JSON line
    # only, send everything else to stderr.
    try:
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        sys.stdout.flush()
        os.dup2(2, 1)
    except OSError:
        _REAL_STDOUT = None
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", default="dsec",
                    choices=["dsec", "dsec_tref5", "evimo2", "evimo2_tref10", "k3_det"])
    ap.add_argument("--events", type=int, default=None, help="events per window (default: lognormal ~1e6)")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--dist", default="uniform", choices=["uniform", "edges"],
                    help="spatial distribution of the synthetic events (edges = ~200 line segments, "
                         "realistic atomic contention; SURVEY.md section 8d)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-packed", action="store_true", help="skip the packed-layout legs")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-input legs (profiling runs)")
    ap.add_argument("--no-train", action="store_true", help="skip the UNet + DDP training-step leg")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the process to the GPU's NUMA node")
    ap.add_argument("--prof-warmup", type=int, default=None, help="override the >=3 warm-up rule (ncu runs only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
