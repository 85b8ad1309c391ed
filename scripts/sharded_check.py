"""Event-sharded mode on real ranks (second mode): run under torchrun on N GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/sharded_check.py [--events 20000000]

Every rank builds the same single DSEC-shaped window, keeps its slice of the event rows and calls
FocusLoss.calc_event_sharded (two NCCL all-reduces inside).  Checks: deterministic mode is
bit-identical to the unsharded call on every rank; prints the strong-scaling time of one huge window
(max over ranks, CUDA events) next to the single-GPU time.
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from motionpriorcmax_b200 import synthetic, trajectories as tj
from motionpriorcmax_b200.losses import LossFactory
from motionpriorcmax_b200.losses.sharded import shard_event_rows
from sweep import device_batch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--events", type=float, default=2e7)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = dict(synthetic.DSEC_LOSS_CONFIG)
    H, W = cfg["image_shape"]
    M = int(a.events)
    ev, npos = device_batch(1, M, H, W, cfg["num_bins"], dev, seed=99)         # same window on every rank
    cg = synthetic.make_coeff_grid(1, 1, H, W, sigma_px=8.0, seed=1234).to(dev)
    out = {}
    for det, norm in ((True, "l1"), (False, "l1"), (False, "l2")):
        L = LossFactory.get_loss_calculator("FOCUS", dict(cfg, deterministic=det, focus_loss_norm=norm))
        times = L.get_reconstruction_times(dev)
        times[0] = 0.5
        sh, np_r = shard_event_rows(ev, npos, rank, world)
        sh = sh.contiguous()

        def full():
            c = cg.clone().requires_grad_()
            loss, _, misc = L.calc(tj.calculate_trajectories_at_t(c, times, 4, 1, "polynomial"), times,
                                   {"events": ev, "num_pos_events": npos})
            loss.backward()
            return loss.detach(), misc["iwes"], c.grad

        def sharded():
            c = cg.clone().requires_grad_()
            loss, _, misc = L.calc_event_sharded(tj.calculate_trajectories_at_t(c, times, 4, 1, "polynomial"),
                                                 times, {"events": sh, "num_pos_events": np_r})
            loss.backward()
            return loss.detach(), misc["iwes"], c.grad

        a_, b_ = full(), sharded()
        torch.cuda.synchronize()
        rel = [((x - y).norm() / y.norm()).item() for x, y in zip(a_, b_)]
        if det:
            same = all(torch.equal(x, y) for x, y in zip(a_, b_))
        else:
            # float atomics: loss and IWE to 1e-5; the gradient of the l1 focus norm contains
            # sign(Sobel) of ~0 responses, which flips with the summation order (reported only)
            same = rel[0] <= 1e-5 and rel[1] <= 1e-5
            # Evidence for that statement: count the pixels whose Sobel sign differs between the two
            # IWEs and bound their response; with the l2 norm (no sign in the backward) the gradients
            # must agree like everything else.
            kx = torch.tensor([[-1., 0., 1.], [-2., 0., 2.], [-1., 0., 1.]], device=dev)[None, None]
            def sob(im):
                im = im.reshape(-1, 1, H, W)
                return torch.nn.functional.conv2d(im, kx, padding=1), torch.nn.functional.conv2d(im, kx.transpose(2, 3), padding=1)
            (ax, ay), (bx, by) = sob(a_[1]), sob(b_[1])
            fx, fy = torch.sign(ax) != torch.sign(bx), torch.sign(ay) != torch.sign(by)
            scale = (ax.abs().mean() + ay.abs().mean()) / 2
            worst = max(float(ax[fx].abs().max()) if fx.any() else 0.0, float(ay[fy].abs().max()) if fy.any() else 0.0)
            flips = {"sign_flipped_pixels": int((fx | fy).sum()), "pixels": int(fx.numel()),
                     "largest_flipped_response_over_mean": worst / float(scale)}
            if norm == "l2":
                same = same and rel[2] <= 1e-5
            # the yardstick for the l1 gradient: two UNSHARDED float runs differ from each other by the
            # same amount (float atomics commute differently from run to run; sign(Sobel) of responses
            # that cancel to ~0 follows the last bits).  Not a property of the sharding.
            c_ = full()
            torch.cuda.synchronize()
            flips["unsharded_run_to_run_rel_diff_dcoeff"] = ((a_[2] - c_[2]).norm() / a_[2].norm()).item()
            flips["unsharded_run_to_run_rel_diff_iwes"] = ((a_[1] - c_[1]).norm() / a_[1].norm()).item()
        flag = torch.tensor([1.0 if same else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        res = {"matches_unsharded_on_every_rank": bool(flag.item()),
               "rel_err_loss_iwes_dcoeff_rank0": rel}
        if not det:
            res["sobel_sign_flips_between_sharded_and_unsharded_iwe"] = flips
        for name, fn in (("single_gpu_ms", full), ("sharded_ms", sharded)):
            for _ in range(3):
                fn()
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                fn()
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[name] = t.item()
        res["events_per_s_sharded"] = M / res["sharded_ms"] * 1e3
        out["deterministic" if det else f"float_{norm}_focus_norm"] = res
    if rank == 0:
        print(json.dumps({"world": world, "events_in_window": M, "results": out}))
    assert all(v["matches_unsharded_on_every_rank"] for v in out.values())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
