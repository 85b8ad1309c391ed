"""Event-sharded mode (second mode): event rows of the same windows split over
ranks, two all-reduces (raw IWE, dLUT) through the phased C-ABI calls.

CPU: the row-sharding helper alone and under a world_size-2 gloo group (the shards of the two
ranks partition the rows; the in-place reduction the mode relies on sums a workspace-like buffer
view).  GPU (-m gpu): two *virtual* ranks on one device drive the phased calls with a hand-made
reduction; the result must equal the unsharded loss - bit for bit in deterministic mode.
"""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-5


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("npos", [None, 0, 7, 20])
def test_shard_event_rows_partitions_the_rows(world, npos):
    from motionpriorcmax_b200.losses.sharded import shard_event_rows
    M = 20
    ev = torch.arange(2 * M * 6, dtype=torch.float32).reshape(2, M, 6)
    seen_pos, seen_neg = [], []
    for r in range(world):
        sh, np_r = shard_event_rows(ev, npos, r, world)
        ids = (sh[0, :, 0] / 6).long().tolist()
        if npos is None:
            assert np_r is None
            seen_pos += ids
        else:
            assert all(i < npos for i in ids[:np_r]) and all(i >= npos for i in ids[np_r:])
            seen_pos += ids[:np_r]
            seen_neg += ids[np_r:]
    assert sorted(seen_pos + seen_neg) == list(range(M))            # every row exactly once
    sizes = [shard_event_rows(ev, npos, r, world)[0].shape[1] for r in range(world)]
    assert max(sizes) - min(sizes) <= 2


def _gloo_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from motionpriorcmax_b200 import synthetic
    from motionpriorcmax_b200.losses.sharded import shard_event_rows, _default_reduce
    ev, npos = synthetic.make_event_batch(2, [900, 500], 48, 64, 5, True, seed=4)      # same on every rank
    sh, np_r = shard_event_rows(ev, npos, rank, world)
    # what the mode exchanges: an in-place SUM over ranks of a view into a byte workspace
    ws = torch.zeros(4096, dtype=torch.uint8)
    sec = ws[256:256 + 8 * 16].view(torch.int64)
    sec += int(sh[..., 5].sum())                                     # this rank's valid rows
    _default_reduce(None)(sec)
    digest = torch.tensor([float(sh.double().sum()), float(sh.shape[1]), float(np_r)], dtype=torch.float64)
    gathered = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(gathered, digest)
    if rank == 0:
        out.put(dict(total_valid=int(ev[..., 5].sum()), reduced=sec.tolist(), npos=npos, M=ev.shape[1],
                     full_sum=float(ev.double().sum()), parts=[g.tolist() for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_reduction_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 30100 + os.getpid() % 500
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=100)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert res["reduced"] == [res["total_valid"]] * 16                # in-place sum over both ranks
    assert abs(sum(p[0] for p in res["parts"]) - res["full_sum"]) < 1e-6 * abs(res["full_sum"])
    assert sum(p[1] for p in res["parts"]) == res["M"] and sum(p[2] for p in res["parts"]) == res["npos"]


# ------------------------------------------------------------------------------------------
# GPU: virtual ranks on one device
# ------------------------------------------------------------------------------------------
def _virtual_ranks(cfg, traj, times, ev, npos, world, deterministic):
    from motionpriorcmax_b200.losses import LossFactory
    from motionpriorcmax_b200.losses.sharded import PhasedLoss, shard_event_rows
    dev = torch.device("cuda:0")
    L = LossFactory.get_loss_calculator("FOCUS", dict(cfg, deterministic=deterministic))
    t = torch.as_tensor(traj, device=dev)
    tm = torch.as_tensor(times, device=dev)
    evd = torch.as_tensor(ev, device=dev)
    ranks = []
    for r in range(world):
        sh, np_r = shard_event_rows(evd, npos if npos >= 0 else None, r, world)
        ranks.append(PhasedLoss(L._cfg, t, tm, sh, np_r))

    def allreduce(secs):
        tot = torch.stack(secs).sum(0)
        for s in secs:
            s.copy_(tot)
    allreduce([p.forward_accumulate() for p in ranks])
    fwd = [p.forward_finish() for p in ranks]
    g = torch.ones(1, device=dev)
    allreduce([p.backward_accumulate(g, include_smooth=(i == 0)) for i, p in enumerate(ranks)])
    dtr = [p.backward_finish() for p in ranks]
    torch.cuda.synchronize()
    return [dict(loss=float(f[0][0]), focus=float(f[0][1]), smooth=float(f[0][2]),
                 iwes=f[1].cpu().numpy(), dtraj=d.cpu().numpy()) for f, d in zip(fwd, dtr)]


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["dsec_pab", "multi_tref3", "next_smooth"])
@pytest.mark.parametrize("world", [2, 3])
def test_virtual_ranks_match_unsharded(variant, world):
    from motionpriorcmax_b200 import synthetic
    from test_gpu_parity import _run_loss, _synthetic_case
    cfg = dict(synthetic.DSEC_LOSS_CONFIG, image_shape=(96, 128), num_knn=16)
    if variant == "multi_tref3":
        cfg = synthetic.multi_tref_variant(cfg, 3)
    elif variant == "next_smooth":
        cfg.update(smooth_type="on_flow_to_next", smooth_weight=0.06, num_bins=9)
    traj, times, ev, npos, _ = _synthetic_case(cfg, 2, [30000, 12000], 2, seed=13)
    # deterministic: bit-identical to the unsharded call, on every virtual rank
    ref = _run_loss(cfg, traj, times, ev, npos, deterministic=True)
    for r in _virtual_ranks(cfg, traj, times, ev, npos, world, True):
        assert r["loss"] == ref["loss"] and r["smooth"] == ref["smooth"]
        assert np.array_equal(r["iwes"].reshape(ref["iwes"].shape), ref["iwes"])
        assert np.array_equal(r["dtraj"], ref["dtraj"])
    # float atomics (l2 focus norm: smooth gradient, see test_gpu_parity._assert_grad_close)
    cfg2 = dict(cfg, focus_loss_norm="l2")
    ref2 = _run_loss(cfg2, traj, times, ev, npos, deterministic=True)
    for r in _virtual_ranks(cfg2, traj, times, ev, npos, world, False):
        assert abs(r["loss"] - ref2["loss"]) <= TOL * abs(ref2["loss"])
        assert rel_err(r["iwes"].reshape(ref2["iwes"].shape), ref2["iwes"]) < TOL
        assert rel_err(r["dtraj"], ref2["dtraj"]) < TOL


def test_sharding_is_additive_for_the_reference_algorithm():
    """CPU, oracle only: the raw IWE of the reference algorithm is additive over any split of the
    event rows (what the first all-reduce of the sharded mode relies on), and the row-sharding
    helper produces such a split."""
    from motionpriorcmax_b200 import synthetic
    from motionpriorcmax_b200.losses.sharded import shard_event_rows
    from oracle import focus_oracle as fo
    cfg = dict(synthetic.DSEC_LOSS_CONFIG, image_shape=(48, 64), num_knn=6, num_bins=5)
    H, W = cfg["image_shape"]
    times = fo.reconstruction_times(1, cfg["num_bins"], 0.6)
    cg = synthetic.make_coeff_grid(2, 1, H, W, sigma_px=5.0, seed=8, coarse=(4, 5)).numpy()
    traj, _ = fo.trajectories_from_coeff_grid(cg, times, 4, 1, "polynomial")
    ev, npos = synthetic.make_event_batch(2, [3000, 1700], H, W, cfg["num_bins"], True, seed=8)
    full = fo.FocusOracle(**cfg, dtype=np.float64).forward(traj, times, ev.numpy(), npos)
    for world in (2, 3):
        raw = 0.0
        for r in range(world):
            sh, np_r = shard_event_rows(ev, npos, r, world)
            raw = raw + fo.FocusOracle(**cfg, dtype=np.float64).forward(traj, times, sh.numpy(), np_r)["iwe_raw"]
        assert rel_err(raw, full["iwe_raw"]) < 1e-12
