"""Host -> device staging of collated event batches.

The reference collate (`src/loader/dsec/loader.py:360-415`) pads every window of a batch with
all-zero rows up to the batch maximum, separately for the positive and the negative group, so a
`[B, M, 6]` batch is typically ~50 % padding.  Padding rows carry `valid = 0` and are inert in
the loss; copying them over PCIe is pure waste.  `EventUploader` copies only the valid prefix of
each group from pinned host memory (asynchronously, on a copy stream, double buffered) and keeps
the rest of the device buffer zero, so the device tensor is bit-identical to a full copy.

`PackedEvents` is the loader-side layout of SURVEY.md 8f rank 2 (include/cmax_b200.h, "packed,
tile-binned event layout"): 16-byte records of the valid events only, grouped by polarity group
and 32x32-pixel source tile, so the event kernels accumulate in shared memory.  Build it in the
loader workers with `pack_events_host` (torch CPU ops, no GPU) or on the device with
`pack_events` (C ABI `cmax_pack_events`); pass it as `batch['events']` to `FocusLoss.calc`.

`CompactEvents` is the WIRE form of the same layout for the host -> device copy: 12 bytes per valid
event ((y, x, t) float32; the time bin is implied by the run an event sits in, the LUT cell by its
coordinates), all windows of a batch back to back in ONE pinned buffer, so a step's events cross
PCIe in one `cudaMemcpyAsync`.  `pack_events_compact` builds it in the loader workers (C++ /
OpenMP), `CompactUploader` copies it (double buffered) and rebuilds the 16-byte records on the
device with one kernel (`cmax_expand_compact`); `FocusLoss.calc` also accepts it directly.
`BitpackedEvents` shrinks the wire form further, losslessly: every (group, tile, bin) run stores
fixed-width bit-pattern deltas of (y, x, t) - about 8 bytes per event for a DSEC window - decoded by
`cmax_expand_bitpacked` into the very same records.
"""
from __future__ import annotations

from dataclasses import dataclass
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


# ------------------------------------------------------------------------------------------------
# packed, tile-binned event layout
# ------------------------------------------------------------------------------------------------
@dataclass
class PackedEvents:
    """records [B, M, 4] float32 (y, x, t, meta bits), seg_start [B, G*NT+1] int32; see
    include/cmax_b200.h.  `skipped` (optional) int64[2]: valid rows dropped because their LUT cell
    is outside the table / rows whose `valid` is neither 0 nor 1."""
    records: torch.Tensor
    seg_start: torch.Tensor
    skipped: Optional[torch.Tensor] = None

    @property
    def is_cuda(self):
        return self.records.is_cuda

    @property
    def device(self):
        return self.records.device

    def num_events(self) -> torch.Tensor:
        return self.seg_start[:, -1]

    def to(self, device, non_blocking: bool = False):
        return PackedEvents(self.records.to(device, non_blocking=non_blocking),
                            self.seg_start.to(device, non_blocking=non_blocking),
                            None if self.skipped is None else self.skipped.to(device, non_blocking=non_blocking))

    def pin_memory(self):
        return PackedEvents(self.records.pin_memory(), self.seg_start.pin_memory(),
                            None if self.skipped is None else self.skipped.pin_memory())


def _cfg_of(loss_or_cfg):
    return getattr(loss_or_cfg, "_cfg", loss_or_cfg)


def pack_events(events: torch.Tensor, num_pos_events: Optional[int], loss_or_cfg) -> PackedEvents:
    """Device-side packing (three kernels) of an upstream-layout `[B, M, 6]` CUDA events tensor."""
    from . import cabi
    lib = cabi.load()
    cfg = _cfg_of(loss_or_cfg)
    if not events.is_cuda:
        raise RuntimeError("pack_events needs a CUDA tensor (use pack_events_host in loader workers)")
    ev = events.detach().to(torch.float32).contiguous()
    B, M, six = ev.shape
    assert six == 6, "events must be [B, M, 6]"
    _, nty, ntx, G = cabi.pack_layout(cfg)
    n_seg = G * nty * ntx
    rec = torch.empty((B, M, 4), dtype=torch.float32, device=ev.device)
    seg = torch.empty((B, n_seg + 1), dtype=torch.int32, device=ev.device)
    scratch = torch.empty((B, n_seg), dtype=torch.int32, device=ev.device)
    skipped = torch.empty(2, dtype=torch.int64, device=ev.device)
    npos = -1 if num_pos_events is None else int(num_pos_events)
    if not cfg.polarity_aware_batching:
        npos = 0
    cabi.check(lib.cmax_pack_events(cfg, cabi.ptr(ev), B, M, npos, cabi.ptr(rec), cabi.ptr(seg),
                                    cabi.ptr(scratch), cabi.ptr(skipped), cabi.stream_ptr(ev.device)),
               "cmax_pack_events")
    return PackedEvents(rec, seg, skipped)


def _check_binary_valid(odd: int, strict: bool):
    if odd and strict:
        raise ValueError(
            f"{odd} event rows have a `valid` value that is neither 0 nor 1: the packed layout keeps no "
            "per-event weight (the reference multiplies the vote by `valid`, focus.py:201). Use the "
            "[B, M, 6] layout for weighted events, or pass strict=False to treat them as valid = 1.")


def pack_events_host(events: torch.Tensor, num_pos_events: Optional[int], loss_or_cfg,
                     layout=None, strict: bool = True) -> PackedEvents:
    """The same layout built with torch CPU ops - for DataLoader workers / the collate function.
    `layout` = (ct, tiles_y, tiles_x, G) avoids loading the CUDA library in a worker process; by
    default it is asked from the library (`cabi.pack_layout`)."""
    cfg = _cfg_of(loss_or_cfg)
    if layout is None:
        from . import cabi
        layout = cabi.pack_layout(cfg)
    ct, nty, ntx, G = layout
    s, nb = cfg.lut_superpixel_size, cfg.num_bins
    Hq = (cfg.height + s - 1) // s
    Wq = (cfg.width + s - 1) // s
    nt = nty * ntx
    ev = events.detach().to(torch.float32).cpu()
    B, M, _ = ev.shape
    npos = int(num_pos_events) if (cfg.polarity_aware_batching and num_pos_events is not None) else None
    if cfg.polarity_aware_batching and (npos is None or npos < 0 or npos > M):
        raise ValueError("polarity-aware packing needs 0 <= num_pos_events <= M")
    # focus.py:185-187 - the reference's own index arithmetic (float floor division, truncation)
    it = ev[..., 4].to(torch.int64)
    fy, fx = ev[..., 0] // s, ev[..., 1] // s
    ok = (ev[..., 5] != 0) & (it >= 0) & (it < nb) & (fy >= 0) & (fy < Hq) & (fx >= 0) & (fx < Wq) \
        & (ev[..., 4] == ev[..., 4])
    dropped = int(((ev[..., 5] != 0) & ~ok).sum())
    odd = int(((ev[..., 5] != 0) & (ev[..., 5] != 1)).sum())
    _check_binary_valid(odd, strict)
    iy = torch.where(ok, fy, torch.zeros_like(fy)).to(torch.int64)
    ix = torch.where(ok, fx, torch.zeros_like(fx)).to(torch.int64)
    it = torch.where(ok, it, torch.zeros_like(it))
    grp = torch.zeros((B, M), dtype=torch.int64)
    if npos is not None:
        grp[:, npos:] = 1
    key = grp * nt + (iy // ct) * ntx + ix // ct
    key = torch.where(ok, key, torch.full_like(key, G * nt))          # dropped rows sort last
    order = torch.argsort(key, dim=1, stable=True)
    counts = torch.zeros((B, G * nt + 1), dtype=torch.int64)
    counts.scatter_add_(1, key, torch.ones_like(key))
    seg = torch.zeros((B, G * nt + 1), dtype=torch.int64)
    seg[:, 1:] = torch.cumsum(counts[:, :G * nt], dim=1)
    Mp = max(int(seg[:, -1].max()), 1)
    meta = ((it << 24) | (iy << 12) | ix).to(torch.int32)
    rec = torch.stack((ev[..., 0], ev[..., 1], ev[..., 2], meta.view(torch.float32)), dim=-1)
    rec = torch.gather(rec, 1, order[:, :Mp, None].expand(B, Mp, 4)).contiguous()
    return PackedEvents(rec, seg.to(torch.int32), torch.tensor([dropped, odd], dtype=torch.int64))


def pack_events_native(events: torch.Tensor, num_pos_events: Optional[int], loss_or_cfg,
                       strict: bool = True) -> PackedEvents:
    """`pack_events_host` through the C ABI (`cmax_pack_events_host`: C++ / OpenMP counting sort,
    one thread per window): the same records and segments (byte for byte), several times faster
    than the torch sort (8 windows of 1 M events on 8 cores: 0.18 s vs 0.62 s).  CPU tensors
    in, CPU `PackedEvents` out (pin it and hand it to `PackedUploader`, or `.to(device)`)."""
    from . import cabi
    import ctypes
    lib = cabi.load()
    cfg = _cfg_of(loss_or_cfg)
    ev = events.detach().to(torch.float32).cpu().contiguous()
    B, M, six = ev.shape
    assert six == 6, "events must be [B, M, 6]"
    _, nty, ntx, G = cabi.pack_layout(cfg)
    n_seg = G * nty * ntx
    npos = int(num_pos_events) if (cfg.polarity_aware_batching and num_pos_events is not None) else 0
    seg = torch.empty((B, n_seg + 1), dtype=torch.int32)
    skipped = torch.zeros(2, dtype=torch.int64)
    p = lambda t: ctypes.c_void_p(t.data_ptr())             # noqa: E731
    # pass 1: segment sizes only (records = NULL) -> the record capacity the batch needs
    cabi.check(lib.cmax_pack_events_host(cfg, p(ev), B, M, npos, None, 0, p(seg), p(skipped)),
               "cmax_pack_events_host")
    _check_binary_valid(int(skipped[1]), strict)
    Mp = max(int(seg[:, -1].max()) if B else 0, 1)
    rec = torch.zeros((B, Mp, 4), dtype=torch.float32)
    cabi.check(lib.cmax_pack_events_host(cfg, p(ev), B, M, npos, p(rec), Mp, p(seg), p(skipped)),
               "cmax_pack_events_host")
    return PackedEvents(rec, seg, skipped)


def unpack_events(packed: PackedEvents, loss_or_cfg, layout=None) -> "tuple[torch.Tensor, Optional[int]]":
    """Inverse for tests / debugging: an upstream-layout `[B, M', 6]` tensor (positives first,
    zero padding, p = +1 / 0 by group) and `num_pos_events`, from a packed batch (CPU)."""
    cfg = _cfg_of(loss_or_cfg)
    if layout is None:
        from . import cabi
        layout = cabi.pack_layout(cfg)
    _, nty, ntx, G = layout
    nt = nty * ntx
    rec, seg = packed.records.cpu(), packed.seg_start.cpu().to(torch.int64)
    B = rec.shape[0]
    meta = rec[..., 3].contiguous().view(torch.int32).to(torch.int64)
    cnt = [[int(seg[b, (g + 1) * nt] - seg[b, g * nt]) for g in range(G)] for b in range(B)]
    cap = [max(max(c[g] for c in cnt), 0) for g in range(G)]
    out = torch.zeros((B, max(sum(cap), 1), 6), dtype=torch.float32)
    for b in range(B):
        off = 0
        for g in range(G):
            a, e = int(seg[b, g * nt]), int(seg[b, (g + 1) * nt])
            rows = out[b, off:off + (e - a)]
            rows[:, 0:3] = rec[b, a:e, 0:3]
            rows[:, 3] = 1.0 if g == 0 else 0.0
            rows[:, 4] = (meta[b, a:e] >> 24).to(torch.float32)
            rows[:, 5] = 1.0
            off += cap[g]
    return out, (cap[0] if G == 2 else None)


class PackedUploader:
    """Double-buffered H2D staging of host-packed batches (pinned `PackedEvents`): per sample only
    the `seg_start[b, -1]` records in use cross PCIe (16 B per valid event, no padding rows)."""

    def __init__(self, device, n_buffers: int = 2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.n = n_buffers
        self.bufs = [None] * n_buffers
        self.ready = [torch.cuda.Event() for _ in range(n_buffers)]
        self.free = [torch.cuda.Event() for _ in range(n_buffers)]
        self.turn = 0
        self.bytes_last = 0
        for ev in self.free:
            ev.record(torch.cuda.current_stream(self.device))

    def upload(self, packed: PackedEvents, counts=None):
        assert packed.records.is_pinned() and packed.seg_start.is_pinned()
        slot = self.turn
        self.turn = (self.turn + 1) % self.n
        if counts is None:
            counts = packed.seg_start[:, -1].tolist()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.free[slot])
            buf = self.bufs[slot]
            if buf is None or buf.records.shape != packed.records.shape:
                buf = PackedEvents(torch.empty(packed.records.shape, dtype=torch.float32, device=self.device),
                                   torch.empty(packed.seg_start.shape, dtype=torch.int32, device=self.device))
                self.bufs[slot] = buf
            nbytes = packed.seg_start.numel() * 4
            buf.seg_start.copy_(packed.seg_start, non_blocking=True)
            for b, c in enumerate(counts):
                if c:
                    buf.records[b, :c].copy_(packed.records[b, :c], non_blocking=True)
                    nbytes += c * 16
            self.bytes_last = nbytes
            self.ready[slot].record(self.stream)
        return buf, slot

    def wait(self, slot: int, stream=None):
        (stream or torch.cuda.current_stream(self.device)).wait_event(self.ready[slot])

    def release(self, slot: int, stream=None):
        self.free[slot].record(stream or torch.cuda.current_stream(self.device))


# ------------------------------------------------------------------------------------------------
# compact wire layout: 12 B per valid event, one buffer per batch
# ------------------------------------------------------------------------------------------------
@dataclass
class CompactEvents:
    """coords [T, 3] float32 (y, x, t), fine_start [B, G*NT*nb + 1] int32, sample_off [B + 1] int64;
    see include/cmax_b200.h ("Compact WIRE layout").  `max_count` = largest window (host int)."""
    coords: torch.Tensor
    fine_start: torch.Tensor
    sample_off: torch.Tensor
    max_count: int
    skipped: Optional[torch.Tensor] = None

    @property
    def is_cuda(self):
        return self.coords.is_cuda

    @property
    def device(self):
        return self.coords.device

    def num_events(self) -> torch.Tensor:
        return self.fine_start[:, -1]

    def nbytes(self) -> int:
        return int(self.coords.numel() * 4 + self.fine_start.numel() * 4 + self.sample_off.numel() * 8)

    def to(self, device, non_blocking: bool = False):
        return CompactEvents(self.coords.to(device, non_blocking=non_blocking),
                             self.fine_start.to(device, non_blocking=non_blocking),
                             self.sample_off.to(device, non_blocking=non_blocking), self.max_count, self.skipped)

    def pin_memory(self, write_combined: bool = False):
        """Page-locked copy for `CompactUploader`.  write_combined=True puts the (write-once,
        never read by the CPU) coordinate buffer into cudaHostAllocWriteCombined memory: the DMA
        engine reads it without snooping the CPU caches (46 -> 55 GB/s for a single rank in
        scripts/h2d_probe.py; no difference once 8 ranks share the host)."""
        coords = pinned_write_combined_copy(self.coords) if write_combined else self.coords.pin_memory()
        return CompactEvents(coords, self.fine_start.pin_memory(), self.sample_off.pin_memory(),
                             self.max_count, self.skipped)


class _HostAlloc:
    """Owner of one cudaHostAlloc'ed buffer (freed with the last tensor that views it)."""

    def __init__(self, nbytes: int, flags: int):
        import ctypes
        self._rt = ctypes.CDLL("libcudart.so.12")
        p = ctypes.c_void_p()
        rc = self._rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(max(nbytes, 16)), ctypes.c_uint(flags))
        if rc != 0:
            raise RuntimeError(f"cudaHostAlloc failed with code {rc}")
        self.ptr = p.value

    def __del__(self):
        try:
            import ctypes
            self._rt.cudaFreeHost(ctypes.c_void_p(self.ptr))
        except Exception:
            pass


def pinned_write_combined_copy(t: torch.Tensor) -> torch.Tensor:
    """Copy of a CPU tensor in write-combined page-locked memory (cudaHostAllocWriteCombined).
    Only ever write it sequentially and never read it on the CPU: it is uncached."""
    import ctypes
    src = t.detach().cpu().contiguous()
    nbytes = src.numel() * src.element_size()
    owner = _HostAlloc(nbytes, 0x04)
    arr = (ctypes.c_char * max(nbytes, 16)).from_address(owner.ptr)
    arr._owner = owner                                   # torch keeps `arr` alive, `arr` keeps the allocation
    out = torch.frombuffer(arr, dtype=src.dtype, count=src.numel()).reshape(src.shape)
    out.copy_(src)
    return out


def pack_events_compact(events: torch.Tensor, num_pos_events: Optional[int], loss_or_cfg,
                        strict: bool = True) -> CompactEvents:
    """Loader-side builder of the compact wire layout from an upstream-layout `[B, M, 6]` CPU tensor
    (`cmax_pack_events_host_compact`: C++ / OpenMP stable counting sort, one thread per window)."""
    from . import cabi
    import ctypes
    lib = cabi.load()
    cfg = _cfg_of(loss_or_cfg)
    ev = events.detach().to(torch.float32).cpu().contiguous()
    B, M, six = ev.shape
    assert six == 6, "events must be [B, M, 6]"
    _, nty, ntx, G = cabi.pack_layout(cfg)
    F = G * nty * ntx * cfg.num_bins
    npos = int(num_pos_events) if (cfg.polarity_aware_batching and num_pos_events is not None) else 0
    fine = torch.empty((B, F + 1), dtype=torch.int32)
    off = torch.zeros(B + 1, dtype=torch.int64)
    skipped = torch.zeros(2, dtype=torch.int64)
    p = lambda t: ctypes.c_void_p(t.data_ptr())             # noqa: E731
    cabi.check(lib.cmax_pack_events_host_compact(cfg, p(ev), B, M, npos, None, 0, p(fine), p(off), p(skipped)),
               "cmax_pack_events_host_compact")
    _check_binary_valid(int(skipped[1]), strict)
    T = int(off[-1])
    coords = torch.empty((max(T, 1), 3), dtype=torch.float32)
    cabi.check(lib.cmax_pack_events_host_compact(cfg, p(ev), B, M, npos, p(coords), coords.shape[0], p(fine),
                                                 p(off), p(skipped)), "cmax_pack_events_host_compact")
    return CompactEvents(coords, fine, off, max(int(fine[:, -1].max()) if B else 0, 1), skipped)


def _expand_scratch(out: PackedEvents, lib, cfg, B: int) -> torch.Tensor:
    """Device scratch of the expand kernels, kept on the output object (reused with it)."""
    need = int(lib.cmax_expand_scratch_ints(cfg, B, out.records.shape[1]))
    sc = getattr(out, "_scratch", None)
    if sc is None or sc.numel() < need or sc.device != out.records.device:
        sc = torch.empty(max(need, 1), dtype=torch.int32, device=out.records.device)
        out._scratch = sc
    return sc


def expand_compact(compact: CompactEvents, loss_or_cfg, out: Optional[PackedEvents] = None) -> PackedEvents:
    """Device: compact wire layout -> `PackedEvents` (one kernel, on the current stream)."""
    from . import cabi
    lib = cabi.load()
    cfg = _cfg_of(loss_or_cfg)
    if not compact.is_cuda:
        raise RuntimeError("expand_compact needs CUDA tensors (CompactEvents.to(device) first)")
    dev = compact.device
    B = compact.fine_start.shape[0]
    _, nty, ntx, G = cabi.pack_layout(cfg)
    Mp = int(compact.max_count)
    if out is None or out.records.shape[0] != B or out.records.shape[1] < Mp:
        out = PackedEvents(torch.empty((B, Mp, 4), dtype=torch.float32, device=dev),
                           torch.empty((B, G * nty * ntx + 1), dtype=torch.int32, device=dev))
    scratch = _expand_scratch(out, lib, cfg, B)
    with torch.cuda.device(dev):
        cabi.check(lib.cmax_expand_compact(cfg, cabi.ptr(compact.coords), cabi.ptr(compact.fine_start),
                                           cabi.ptr(compact.sample_off), B, out.records.shape[1],
                                           cabi.ptr(out.records), cabi.ptr(out.seg_start), cabi.ptr(scratch),
                                           cabi.stream_ptr(dev)), "cmax_expand_compact")
    return out


@dataclass
class BitpackedEvents:
    """Bit-packed wire layout (include/cmax_b200.h): words [total] int32 (bit stream), fine_start
    [B, F + 1] int32, run_hdr [B, F, 4] int32, run_word [B, F + 1] int32, word_off [B + 1] int64."""
    words: torch.Tensor
    fine_start: torch.Tensor
    run_hdr: torch.Tensor
    run_word: torch.Tensor
    word_off: torch.Tensor
    max_count: int
    skipped: Optional[torch.Tensor] = None

    _TENSORS = ("words", "fine_start", "run_hdr", "run_word", "word_off")

    @property
    def is_cuda(self):
        return self.words.is_cuda

    @property
    def device(self):
        return self.words.device

    def num_events(self) -> torch.Tensor:
        return self.fine_start[:, -1]

    def nbytes(self) -> int:
        return int(sum(getattr(self, k).numel() * getattr(self, k).element_size() for k in self._TENSORS))

    def to(self, device, non_blocking: bool = False):
        return BitpackedEvents(*(getattr(self, k).to(device, non_blocking=non_blocking) for k in self._TENSORS),
                               self.max_count, self.skipped)

    def pin_memory(self, write_combined: bool = False):
        words = pinned_write_combined_copy(self.words) if write_combined else self.words.pin_memory()
        return BitpackedEvents(words, *(getattr(self, k).pin_memory() for k in self._TENSORS[1:]),
                               self.max_count, self.skipped)


_BITPACK_ONE_CALL_MAX_WORDS = 1 << 29      # 2 GiB of (mostly untouched) scratch at most


def pack_events_bitpacked(events: torch.Tensor, num_pos_events: Optional[int], loss_or_cfg,
                          strict: bool = True) -> BitpackedEvents:
    """Loader-side builder of the bit-packed wire layout from an upstream-layout `[B, M, 6]` CPU
    tensor (`cmax_pack_events_host_bitpacked`, C++ / OpenMP).  Lossless for any float32 values."""
    from . import cabi
    import ctypes
    lib = cabi.load()
    cfg = _cfg_of(loss_or_cfg)
    ev = events.detach().to(torch.float32).cpu().contiguous()
    B, M, six = ev.shape
    assert six == 6, "events must be [B, M, 6]"
    _, nty, ntx, G = cabi.pack_layout(cfg)
    F = G * nty * ntx * cfg.num_bins
    npos = int(num_pos_events) if (cfg.polarity_aware_batching and num_pos_events is not None) else 0
    fine = torch.empty((B, F + 1), dtype=torch.int32)
    hdr = torch.zeros((B, F, 4), dtype=torch.int32)
    rword = torch.empty((B, F + 1), dtype=torch.int32)
    woff = torch.zeros(B + 1, dtype=torch.int64)
    skipped = torch.zeros(2, dtype=torch.int64)
    p = lambda t: ctypes.c_void_p(t.data_ptr())             # noqa: E731
    # one call: tables and streams together, into a buffer that always suffices (3 words per row +
    # one alignment word per run + the slack; untouched pages of it are never committed), then the
    # used part is copied out - cheaper than a sizing call, which repeats the pass over the rows
    cap = max(B * (3 * M + F + 2), 4)
    if cap > _BITPACK_ONE_CALL_MAX_WORDS:                   # too much address space: size it exactly instead
        cabi.check(lib.cmax_pack_events_host_bitpacked(cfg, p(ev), B, M, npos, None, 0, p(fine), p(hdr), p(rword),
                                                       p(woff), p(skipped)), "cmax_pack_events_host_bitpacked")
        cap = max(int(woff[-1]), 4)
    scratch = torch.empty(cap, dtype=torch.int32)
    cabi.check(lib.cmax_pack_events_host_bitpacked(cfg, p(ev), B, M, npos, p(scratch), cap, p(fine), p(hdr),
                                                   p(rword), p(woff), p(skipped)), "cmax_pack_events_host_bitpacked")
    _check_binary_valid(int(skipped[1]), strict)
    words = scratch[:max(int(woff[-1]), 4)].clone()
    del scratch
    return BitpackedEvents(words, fine, hdr, rword, woff, max(int(fine[:, -1].max()) if B else 0, 1), skipped)


def expand_bitpacked(bp: BitpackedEvents, loss_or_cfg, out: Optional[PackedEvents] = None) -> PackedEvents:
    """Device: bit-packed wire layout -> `PackedEvents` (one kernel, on the current stream)."""
    from . import cabi
    lib = cabi.load()
    cfg = _cfg_of(loss_or_cfg)
    if not bp.is_cuda:
        raise RuntimeError("expand_bitpacked needs CUDA tensors (BitpackedEvents.to(device) first)")
    dev = bp.device
    B = bp.fine_start.shape[0]
    _, nty, ntx, G = cabi.pack_layout(cfg)
    Mp = int(bp.max_count)
    if out is None or out.records.shape[0] != B or out.records.shape[1] < Mp:
        out = PackedEvents(torch.empty((B, Mp, 4), dtype=torch.float32, device=dev),
                           torch.empty((B, G * nty * ntx + 1), dtype=torch.int32, device=dev))
    scratch = _expand_scratch(out, lib, cfg, B)
    with torch.cuda.device(dev):
        cabi.check(lib.cmax_expand_bitpacked(cfg, cabi.ptr(bp.words), cabi.ptr(bp.fine_start), cabi.ptr(bp.run_hdr),
                                             cabi.ptr(bp.run_word), cabi.ptr(bp.word_off), B, out.records.shape[1],
                                             cabi.ptr(out.records), cabi.ptr(out.seg_start), cabi.ptr(scratch),
                                             cabi.stream_ptr(dev)), "cmax_expand_bitpacked")
    return out


def expand_wire(wire, loss_or_cfg, out: Optional[PackedEvents] = None) -> PackedEvents:
    """`expand_compact` or `expand_bitpacked`, by the type of the wire layout."""
    if isinstance(wire, BitpackedEvents):
        return expand_bitpacked(wire, loss_or_cfg, out)
    return expand_compact(wire, loss_or_cfg, out)


class CompactUploader:
    """Double-buffered H2D staging of `CompactEvents` / `BitpackedEvents` batches: ONE big copy of the
    event payload (12 B x events, or the ~8 B x events bit stream) plus a few small tables from pinned
    memory on a copy stream; `wait(slot)` makes the consumer's stream wait for it and returns the
    device `PackedEvents` rebuilt there by the expand kernel."""

    def __init__(self, device, loss_or_cfg, n_buffers: int = 2):
        self.device = torch.device(device)
        self.cfg = _cfg_of(loss_or_cfg)
        self.stream = torch.cuda.Stream(self.device)
        self.n = n_buffers
        self.wire = [None] * n_buffers           # device wire-layout buffers (capacity may exceed use)
        self.out = [None] * n_buffers            # device PackedEvents buffers
        self.ready = [torch.cuda.Event() for _ in range(n_buffers)]
        self.free = [torch.cuda.Event() for _ in range(n_buffers)]
        self.turn = 0
        self.bytes_last = 0
        self.issue_ms_last = 0.0                 # host time spent issuing the copies
        for ev in self.free:
            ev.record(torch.cuda.current_stream(self.device))

    @staticmethod
    def _payload(wire):
        """(name of the big tensor, rows in use, names of the small tables)"""
        if isinstance(wire, BitpackedEvents):
            return "words", int(wire.word_off[-1]), ("fine_start", "run_hdr", "run_word", "word_off")
        return "coords", int(wire.sample_off[-1]), ("fine_start", "sample_off")

    def upload(self, wire_host):
        import time
        big, used, small = self._payload(wire_host)
        assert getattr(wire_host, big).is_pinned() and all(getattr(wire_host, k).is_pinned() for k in small)
        t0 = time.perf_counter()
        slot = self.turn
        self.turn = (self.turn + 1) % self.n
        src = getattr(wire_host, big)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.free[slot])
            w = self.wire[slot]
            if (w is None or type(w) is not type(wire_host) or getattr(w, big).shape[0] < max(used, 1)
                    or w.fine_start.shape != wire_host.fine_start.shape):
                cap = (max(int(used * 1.25), 4),) + tuple(src.shape[1:])
                tensors = {big: torch.empty(cap, dtype=src.dtype, device=self.device)}
                for k in small:
                    t = getattr(wire_host, k)
                    tensors[k] = torch.empty(t.shape, dtype=t.dtype, device=self.device)
                w = type(wire_host)(**tensors, max_count=0)
                self.wire[slot] = w
            nbytes = 0
            if used:
                getattr(w, big)[:used].copy_(src[:used], non_blocking=True)       # the one big copy
                nbytes += used * src[0].numel() * src.element_size()
            for k in small:
                getattr(w, k).copy_(getattr(wire_host, k), non_blocking=True)
                nbytes += getattr(wire_host, k).numel() * getattr(wire_host, k).element_size()
            w.max_count = wire_host.max_count
            self.bytes_last = nbytes
            self.ready[slot].record(self.stream)
        self.issue_ms_last = (time.perf_counter() - t0) * 1e3
        return w, slot

    def wait(self, slot: int, stream=None) -> PackedEvents:
        """Make `stream` (default: the current one) wait for the slot's copy, then rebuild the
        16-byte records there (the expand kernel runs on the CONSUMER's stream: on the copy stream
        it would queue behind the consumer's long kernels and hold up the next copy)."""
        stream = stream or torch.cuda.current_stream(self.device)
        stream.wait_event(self.ready[slot])
        with torch.cuda.stream(stream):
            self.out[slot] = expand_wire(self.wire[slot], self.cfg, self.out[slot])
        return self.out[slot]

    def release(self, slot: int, stream=None):
        self.free[slot].record(stream or torch.cuda.current_stream(self.device))


# ------------------------------------------------------------------------------------------------
# host placement: keep a rank's pinned staging buffers on the NUMA node of its GPU
# ------------------------------------------------------------------------------------------------
def _parse_cpulist(text: str):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return cpus


def bind_host_to_device_numa(device_index: int) -> dict:
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off, so that the pinned
    buffers it allocates afterwards (first touch) and the loader-side packing sit next to that GPU's
    PCIe root.  With one process per GPU on a two-socket host this keeps N concurrent host->device
    streams off the inter-socket link.  Best effort: returns what it found / did, never raises."""
    import os
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = device_index
        if vis:
            ids = vis.split(",")
            if device_index < len(ids) and ids[device_index].strip().isdigit():
                phys = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:                     # nvml prints an 8-digit domain, sysfs uses 4
            bus = bus[4:]
        info["pci_bus_id"] = bus
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return info
        os.sched_setaffinity(0, allowed)
        info.update(bound=True, cpus=len(allowed))
    except Exception as exc:                                 # no NUMA information, containers, ...
        info["error"] = f"{type(exc).__name__}: {exc}"
    return info
