"""Measure the atomic side of the roofline on the box (R_atomic is not
given anywhere): float32 red.global.add to pseudo-random bilinear-vote-shaped addresses inside an
L2-resident region, shared-memory atomics, and int64 global atomics.  Writes JSON to stdout."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from motionpriorcmax_b200 import cabi

lib = cabi.load()
dev = torch.device("cuda:0")
out = {}
n_ops = 1 << 28
for name, floats in (("iwe_plane_1.2MB", 480 * 640), ("batch14_pab_34MB", 14 * 2 * 480 * 640),
                     ("larger_than_L2_512MB", 128 * 1024 * 1024)):
    region = torch.zeros(floats, dtype=torch.float32, device=dev)
    for mode, mname in ((0, "red_global_f32"), (1, "atom_shared_f32"), (2, "atom_global_u64"),
                        (3, "red_global_v2_f32"), (4, "red_global_v4_f32")):
        if mode == 2 and floats % 2:
            continue
        for _ in range(2):
            cabi.check(lib.cmax_atomic_microbench(cabi.ptr(region), floats, n_ops, mode, None), "bench")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            cabi.check(lib.cmax_atomic_microbench(cabi.ptr(region), floats, n_ops, mode, None), "bench")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out[f"{mname}/{name}"] = {"ops": n_ops, "ms": ms, "Gops_per_s": n_ops / ms / 1e6}
        if mode >= 3:            # n_ops scalar votes travel in n_ops / 2 vector requests
            out[f"{mname}/{name}"]["Grequests_per_s"] = n_ops / 2 / ms / 1e6
print(json.dumps(out, indent=1))
