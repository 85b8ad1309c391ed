"""Host -> device staging of collated event batches.

The reference collate (`src/loader/dsec/loader.py:360-415`) pads every window of a batch with
all-zero rows up to the batch maximum, separately for the positive and the negative group, so a
`[B, M, 6]` batch is typically ~50 % padding.  Padding rows carry `valid = 0` and are inert in
the loss; copying them over PCIe is pure waste.  `EventUploader` copies only the valid prefix of
each group from pinned host memory (asynchronously, on a copy stream, double buffered) and keeps
the rest of the device buffer zero, so the device tensor is bit-identical to a full copy.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch


def valid_prefix_lengths(events: torch.Tensor, num_pos_events: Optional[int]) -> np.ndarray:
    """Length of the `valid == 1` prefix of every (sample, polarity group): int64 [B, G]
    (G = 2 when polarity aware, else 1).  Vectorised binary search on the host tensor; relies on
    the collate contract that `valid` is 1...1 0...0 inside each group."""
    B, M, _ = events.shape
    col = events[:, :, 5].numpy()
    bounds = [(0, M)] if num_pos_events is None or num_pos_events < 0 else \
        [(0, int(num_pos_events)), (int(num_pos_events), M)]
    out = np.zeros((B, len(bounds)), np.int64)
    rows = np.arange(B)
    for gi, (s, e) in enumerate(bounds):
        lo = np.zeros(B, np.int64)           # invariant: col[s + lo - 1] == 1 (or lo == 0)
        hi = np.full(B, e - s, np.int64)     # invariant: col[s + hi] == 0 (or hi == e - s)
        while (lo < hi).any():
            mid = (lo + hi) // 2
            act = lo < hi
            v = np.where(act, col[rows, np.minimum(s + mid, M - 1)], 0) > 0
            lo = np.where(act & v, mid + 1, lo)
            hi = np.where(act & ~v, mid, hi)
        out[:, gi] = lo
    return out


def _subtract(old, keep):
    """Parts of the intervals `old` not covered by the intervals `keep` (half-open, sorted)."""
    out = []
    for a, b in old:
        cur = a
        for s, e in sorted(keep):
            if e <= cur or s >= b:
                continue
            if s > cur:
                out.append((cur, s))
            cur = max(cur, e)
        if cur < b:
            out.append((cur, b))
    return out


class EventUploader:
    def __init__(self, device, n_buffers: int = 2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.n = n_buffers
        self.bufs = [None] * n_buffers
        self.extent = [None] * n_buffers          # rows currently non-zero per (sample, group)
        self.ready = [torch.cuda.Event() for _ in range(n_buffers)]
        self.free = [torch.cuda.Event() for _ in range(n_buffers)]
        self.turn = 0
        self.bytes_last = 0
        for ev in self.free:
            ev.record(torch.cuda.current_stream(self.device))

    def upload(self, events_host: torch.Tensor, num_pos_events: Optional[int] = None):
        """Start the asynchronous upload of one batch. Returns (device tensor, slot); call
        `wait(slot)` on the consumer stream before use and `release(slot)` after the last use."""
        assert events_host.is_pinned() and events_host.dtype == torch.float32
        B, M, C = events_host.shape
        slot = self.turn
        self.turn = (self.turn + 1) % self.n
        lens = valid_prefix_lengths(events_host, num_pos_events)
        starts = [0] if lens.shape[1] == 1 else [0, int(num_pos_events)]
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.free[slot])
            buf = self.bufs[slot]
            if buf is None or buf.shape != events_host.shape:
                buf = torch.zeros(events_host.shape, dtype=torch.float32, device=self.device)
                self.bufs[slot] = buf
                self.extent[slot] = None
            old = self.extent[slot]                   # per sample: [(start, stop)] still non-zero
            nbytes = 0
            new_ext = []
            for b in range(B):
                keep = [(s, s + int(lens[b, gi])) for gi, s in enumerate(starts) if lens[b, gi]]
                for s0, s1 in _subtract(old[b] if old else [], keep):
                    buf[b, s0:s1].zero_()             # stale rows of the previous batch
                for s0, s1 in keep:
                    buf[b, s0:s1].copy_(events_host[b, s0:s1], non_blocking=True)
                    nbytes += (s1 - s0) * C * 4
                new_ext.append(keep)
            lens = new_ext
            self.extent[slot] = lens
            self.bytes_last = nbytes
            self.ready[slot].record(self.stream)
        return buf, slot

    def wait(self, slot: int, stream=None):
        (stream or torch.cuda.current_stream(self.device)).wait_event(self.ready[slot])

    def release(self, slot: int, stream=None):
        self.free[slot].record(stream or torch.cuda.current_stream(self.device))
