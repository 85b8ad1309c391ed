"""Minimal data-parallel training step around the B200 CMax loss (one process per GPU).

Mirrors the shape of the reference's training loop (`scripts/flow_training.py:125-130` ->
`TrajectoryNet.training_step`, `src/modules/trajectory_net.py:142-170`): a stock-PyTorch network
maps the voxel grid to a coefficient grid, the fused front end turns it into trajectories, the
CMax loss scores them on this rank's event windows, and DDP all-reduces the *network* gradients
over NCCL (NVLink / NVSwitch).  The loss issues no collective.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        examples/ddp_flow_training_step.py --steps 3
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motionpriorcmax_b200 import synthetic, trajectories as tj      # noqa: E402
from motionpriorcmax_b200.losses import LossFactory                  # noqa: E402


class TinyFlowNet(nn.Module):
    """Stand-in for the reference UNet(15, 2K) (stock PyTorch; the real network stays stock too)."""

    def __init__(self, bins, out):
        super().__init__()
        self.net = nn.Sequential(nn.Conv2d(bins, 16, 3, padding=1), nn.ReLU(),
                                 nn.Conv2d(16, 16, 3, padding=1), nn.ReLU(),
                                 nn.Conv2d(16, out, 3, padding=1))

    def forward(self, x):
        return self.net(x)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--events", type=int, default=50_000)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = dict(synthetic.DSEC_LOSS_CONFIG, image_shape=(96, 128), num_knn=16)
    H, W = cfg["image_shape"]
    torch.manual_seed(0)                                   # same initial weights on every rank
    net = TinyFlowNet(cfg["num_bins"], 2).to(dev)
    model = nn.parallel.DistributedDataParallel(net, device_ids=[local]) if world > 1 else net
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)   # trajectory_net.py:213-219
    loss_calc = LossFactory.get_loss_calculator("FOCUS", cfg)
    for step in range(a.steps):
        ev, npos = synthetic.make_event_batch(a.batch, a.events, H, W, cfg["num_bins"], True,
                                              seed=100 + step, rank=rank)
        ev = ev.to(dev)
        voxel = torch.randn(a.batch, cfg["num_bins"], H, W, device=dev)
        coeff_grid = model(voxel)[:, None]                 # [B, 1, 2K, H, W]
        times = loss_calc.get_reconstruction_times(dev)
        traj = tj.calculate_trajectories_at_t(coeff_grid, times, 4, 1, "polynomial")
        loss, log, _ = loss_calc.calc(traj, times, {"events": ev, "num_pos_events": npos})
        opt.zero_grad(set_to_none=True)
        loss.backward()                                    # DDP all-reduces the network grads here
        opt.step()
        gsum = torch.stack([p.grad.double().abs().sum() for p in net.parameters()]).sum()
        if world > 1:
            lo, hi = gsum.clone(), gsum.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            assert torch.allclose(lo, hi), "network gradients differ across ranks after DDP"
        if rank == 0:
            print(f"step {step}: loss={loss.item():.6f} focus={log['focus_loss'].item():.6f} "
                  f"|grad|={gsum.item():.4e} ranks={world}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
