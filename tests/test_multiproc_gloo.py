"""world_size-2 checks of the multi-rank host logic on CPU (gloo): windows are sharded across
ranks with rank-dependent seeds, there is no collective on the data path, and the whole-job
number is (sum of events) / (max over ranks of the time)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from motionpriorcmax_b200 import synthetic
    cfg, w = bench.workload("dsec", batch=2, events=300)
    cg, ev, npos, n_valid = bench.make_inputs(cfg, w, rank)
    # every rank holds its own windows; nothing about the loss needs the other rank's data
    digest = torch.tensor([float(ev.double().sum()), float(cg.double().sum())], dtype=torch.float64)
    gathered = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(gathered, digest)
    ms_total = 10.0 * (rank + 1)            # pretend rank 1 is the slow one
    tot, e2e, ev_all, rows_all = bench.aggregate_over_ranks(ms_total, 2 * ms_total, n_valid,
                                                            ev.shape[0] * ev.shape[1], dist, "cpu")
    counts = synthetic.lognormal_event_counts(4, rank=rank)
    if rank == 0:
        out.put(dict(tot=tot, e2e=e2e, ev_all=ev_all, rows_all=rows_all, n_valid=n_valid,
                     digests=[g.tolist() for g in gathered], counts=counts))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharding_and_aggregation_world_size_2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=100)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert res["tot"] == 20.0 and res["e2e"] == 40.0                 # MAX over ranks
    assert res["ev_all"] == 2 * res["n_valid"]                        # SUM over ranks (weak scaling)
    assert res["digests"][0] != res["digests"][1]                     # different windows per rank
    assert all(2e5 <= c <= 4e6 for c in res["counts"])
