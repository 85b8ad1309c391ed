"""Host -> device copy rate of one 180 MB buffer per rank, all ranks copying at once (torchrun):
torch pinned memory vs write-combined pinned memory (cudaHostAllocWriteCombined), 1 vs 2 streams.
    python -m torch.distributed.run --nproc-per-node N scripts/h2d_probe.py"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def wc_pinned(nbytes):
    rt = ctypes.CDLL("libcudart.so.12")
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04))   # WriteCombined
    assert rc == 0, rc
    buf = (ctypes.c_char * nbytes).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.float32)


def main():
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n = 45_000_000                                     # floats = 180 MB
    dst = torch.empty(n, device=dev)
    srcs = {"torch_pinned": torch.empty(n).pin_memory()}
    try:
        w = wc_pinned(n * 4)
        w[:1024].fill_(1.0)
        srcs["write_combined"] = w
    except Exception as exc:
        srcs_err = str(exc)
    out = {}
    for name, src in srcs.items():
        for nstream in (1, 2):
            streams = [torch.cuda.Stream(dev) for _ in range(nstream)]
            def go():
                k = n // nstream
                for i, s in enumerate(streams):
                    with torch.cuda.stream(s):
                        dst[i * k:(i + 1) * k].copy_(src[i * k:(i + 1) * k], non_blocking=True)
            for _ in range(3):
                go()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for s in streams:
                s.wait_event(e0)
            for _ in range(10):
                go()
            for s in streams:
                torch.cuda.current_stream().wait_stream(s)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out[f"{name}/{nstream}_stream"] = {"ms_max_over_ranks": t.item(), "GBps_per_rank": n * 4 / t.item() / 1e6}
    if rank == 0:
        print(json.dumps({"world": world, "bytes": n * 4, "results": out}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
