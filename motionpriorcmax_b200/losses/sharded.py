"""Event-sharded CMax loss: one batch of windows whose EVENT ROWS are split over the ranks
(second mode - a single huge window, or windows too large for one GPU).

Every rank holds the same trajectories and its own slice of the event rows.  The loss is the
reference's `FocusLoss.calc` (upstream src/losses/focus.py:66-113) over the union of the rows:

    forward   LUT (replicated) -> local splat -> all-reduce(raw IWE) -> image stage (replicated)
    backward  image stage^T (replicated) -> local gather -> all-reduce(dLUT) -> LUT^T (replicated)

through the phased C-ABI calls (`cmax_forward_accumulate` / `_finish`, `cmax_backward_accumulate`
/ `_finish`, include/cmax_b200.h).  The two exchanged sections are R*P*H*W and Q*R*2 values per
sample (2.4 MB and 2.3 MB for one DSEC window); they are reduced in place inside the workspace
with `torch.distributed.all_reduce` (NCCL over NVLink).  In deterministic mode the sections are
int64 fixed point, so the result is bit-identical to the unsharded call for ANY split of the rows.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch

from .. import cabi


def shard_event_rows(events: torch.Tensor, num_pos_events: Optional[int], rank: int, world: int
                     ) -> Tuple[torch.Tensor, Optional[int]]:
    """Rows of `events [B, M, 6]` owned by `rank`: a contiguous slice of the positive group and one
    of the negative group (so the slice keeps the positives-first layout), balanced to +-1 row.
    Returns (view, num_pos_events of the slice)."""
    assert 0 <= rank < world
    M = events.shape[1]

    def span(a, e):
        n = e - a
        return a + (n * rank) // world, a + (n * (rank + 1)) // world

    if num_pos_events is None or num_pos_events < 0:
        a, e = span(0, M)
        return events[:, a:e], None
    npos = int(num_pos_events)
    pa, pe = span(0, npos)
    na, ne = span(npos, M)
    return torch.cat((events[:, pa:pe], events[:, na:ne]), dim=1), pe - pa


class PhasedLoss:
    """Thin host wrapper of the phased C-ABI calls for ONE rank; the reduction between the phases
    is the caller's (`section` returns a tensor view of the workspace to all-reduce in place)."""

    def __init__(self, cfg, trajectories, times, events, num_pos_events):
        self.lib = cabi.load()
        self.cfg = cfg
        dev = trajectories.device
        self.traj = trajectories.detach().to(torch.float32).contiguous()
        self.times = times.detach().to(device=dev, dtype=torch.float32).contiguous()
        self.ev = events.detach().to(torch.float32).contiguous()
        B, n_t, n, two = self.traj.shape
        assert two == 2 and n_t == cfg.num_tref + cfg.num_bins
        assert self.ev.dim() == 3 and self.ev.shape[0] == B and self.ev.shape[2] == 6
        self.B, self.n, self.M = B, n, self.ev.shape[1]
        self.npos = -1 if num_pos_events is None else int(num_pos_events)
        self.need = self.lib.cmax_workspace_bytes(cfg, B, self.M, n)
        if self.need == 0:
            raise RuntimeError("cmax_workspace_bytes returned 0 (invalid configuration or shape)")
        self.ws = torch.empty(self.need, dtype=torch.uint8, device=dev)
        self.dev = dev

    def section(self, which: int) -> torch.Tensor:
        off, size, i64 = cabi.workspace_section(self.cfg, self.B, self.M, self.n, which)
        return self.ws[off:off + size].view(torch.int64 if i64 else torch.float32)

    def forward_accumulate(self) -> torch.Tensor:
        rc = self.lib.cmax_forward_accumulate(self.cfg, cabi.ptr(self.traj), cabi.ptr(self.times),
                                              cabi.ptr(self.ev), self.B, self.M, self.n, self.npos, None,
                                              cabi.ptr(self.ws), self.need, cabi.stream_ptr(self.dev))
        cabi.check(rc, "cmax_forward_accumulate")
        return self.section(cabi.SECTION_RAW_IWE)

    def forward_finish(self):
        cfg = self.cfg
        P = 2 if cfg.polarity_aware_batching else 1
        iwes = torch.empty((self.B * cfg.num_tref, P, cfg.height, cfg.width), dtype=torch.float32,
                           device=self.dev)
        losses = torch.empty(3, dtype=torch.float32, device=self.dev)
        rc = self.lib.cmax_forward_finish(cfg, self.B, self.M, self.n, cabi.ptr(iwes), cabi.ptr(losses),
                                          cabi.ptr(self.ws), self.need, cabi.stream_ptr(self.dev))
        cabi.check(rc, "cmax_forward_finish")
        return losses, iwes

    def backward_accumulate(self, grad_loss: torch.Tensor, include_smooth: bool) -> torch.Tensor:
        self.g = grad_loss.detach().to(device=self.dev, dtype=torch.float32).reshape(1).contiguous()
        rc = self.lib.cmax_backward_accumulate(self.cfg, cabi.ptr(self.traj), cabi.ptr(self.times),
                                               cabi.ptr(self.ev), self.B, self.M, self.n, self.npos,
                                               cabi.ptr(self.g), int(bool(include_smooth)),
                                               cabi.ptr(self.ws), self.need, cabi.stream_ptr(self.dev))
        cabi.check(rc, "cmax_backward_accumulate")
        return self.section(cabi.SECTION_DLUT)

    def backward_finish(self) -> torch.Tensor:
        dtraj = torch.empty_like(self.traj)
        rc = self.lib.cmax_backward_finish(self.cfg, cabi.ptr(self.traj), self.B, self.M, self.n,
                                           cabi.ptr(self.g), cabi.ptr(dtraj), cabi.ptr(self.ws),
                                           self.need, cabi.stream_ptr(self.dev))
        cabi.check(rc, "cmax_backward_finish")
        return dtraj


def _default_reduce(group):
    import torch.distributed as dist

    def reduce(t: torch.Tensor):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return reduce


class _ShardedLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, trajectories, times, events, cfg, num_pos_events, reduce: Callable, is_root: bool):
        ph = PhasedLoss(cfg, trajectories, times, events, num_pos_events)
        reduce(ph.forward_accumulate())
        losses, iwes = ph.forward_finish()
        ctx.ph, ctx.reduce, ctx.is_root = ph, reduce, is_root
        ctx.mark_non_differentiable(iwes, losses)
        return losses[0].clone(), losses, iwes

    @staticmethod
    def backward(ctx, grad_loss, *unused):
        ph = ctx.ph
        ctx.reduce(ph.backward_accumulate(grad_loss, ctx.is_root))
        dtraj = ph.backward_finish()
        ctx.ph = None
        return dtraj, None, None, None, None, None, None


def calc_event_sharded(loss, trajectories, times, batch, group=None, reduce: Optional[Callable] = None,
                       is_root: Optional[bool] = None):
    """`FocusLoss.calc` for a batch whose event rows are sharded over the ranks of `group`.

    `batch['events']` holds THIS rank's rows (see `shard_event_rows`), `trajectories` / `times` are
    the same on every rank.  Returns the same triple as `calc`, identical on every rank.
    `reduce(tensor)` (in-place sum over ranks) and `is_root` default to torch.distributed."""
    import torch.distributed as dist
    events = batch['events']
    if not (trajectories.is_cuda and events.is_cuda):
        raise RuntimeError("FocusLoss (B200) needs CUDA tensors; there is no CPU fallback")
    num_pos_events = batch['num_pos_events'] if 'num_pos_events' in batch else -1
    assert not loss.polarity_aware_batching or num_pos_events > -1
    if reduce is None:
        reduce = _default_reduce(group)
    if is_root is None:
        is_root = dist.get_rank(group) == 0
    out, losses, iwes = _ShardedLossFunction.apply(trajectories, times, events, loss._cfg,
                                                   int(num_pos_events), reduce, bool(is_root))
    h, w = loss.image_shape
    b = trajectories.shape[0]
    iwes = iwes.reshape(b, loss.num_tref, 2, h, w) if loss.polarity_aware_batching \
        else iwes.reshape(b, loss.num_tref, h, w)
    return out, {'focus_loss': losses[1].detach(), 'smoothness_loss': losses[2].detach()}, \
        {'iwes': iwes.detach()}
