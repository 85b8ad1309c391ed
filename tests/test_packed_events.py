"""Packed, tile-binned event layout (include/cmax_b200.h).

CPU part: the host packer (`io.pack_events_host`) against a brute-force restatement of the
layout contract, and its round trip.  GPU part (-m gpu, through the C ABI): the tile kernels
(`cmax_forward_packed` / `cmax_backward_packed`) against the reference goldens, against the
float64 oracle, and - in deterministic mode, where sums are order independent - BIT-IDENTICAL to
the unpacked path; the device packer against the host packer.
"""
import numpy as np
import pytest
import torch

from helpers import LOSS_CASES, load_case, rel_err

TOL = 1e-5


def _cfg(d, **kw):
    from motionpriorcmax_b200 import cabi
    keys = ("image_shape", "num_tref", "num_bins", "num_knn", "smooth_weight", "lut_superpixel_size",
            "focus_loss_norm", "dist_norm", "scale_iwe_by_dt", "mask_image_border",
            "polarity_aware_batching", "interpolation_scheme", "smooth_type")
    return cabi.make_config(**{k: d[k] for k in keys}, **kw)


def _segments(packed, layout):
    """{(b, segment): sorted list of (y, x, t, meta) tuples}"""
    ct, nty, ntx, G = layout
    rec = packed.records.cpu().numpy()
    seg = packed.seg_start.cpu().numpy()
    out = {}
    for b in range(rec.shape[0]):
        for k in range(G * nty * ntx):
            a, e = seg[b, k], seg[b, k + 1]
            if e > a:
                rows = rec[b, a:e].view(np.uint32)
                out[(b, k)] = sorted(map(tuple, rows.tolist()))
    return out


# ------------------------------------------------------------------------------------------
# CPU: host packer
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("s,pab", [(4, True), (3, False), (8, True), (64, False)])
def test_host_packer_layout_contract(s, pab):
    from motionpriorcmax_b200 import io, synthetic
    H, W, nb = 50, 70, 5
    d = dict(synthetic.DSEC_LOSS_CONFIG, image_shape=(H, W), num_bins=nb, lut_superpixel_size=s,
             num_knn=2, polarity_aware_batching=pab)
    cfg = _cfg(d)
    ct = max(32 // s, 1)
    Hq, Wq = -(-H // s), -(-W // s)
    layout = (ct, -(-Hq // ct), -(-Wq // ct), 2 if pab else 1)
    ev, npos = synthetic.make_event_batch(3, [700, 300, 0], H, W, nb, pab, seed=3)
    ev = ev.clone()
    # adversarial rows: coordinates a few ulp around cell edges, outside the table, NaN, padding
    ev[0, 0, :2] = torch.tensor([np.float32(s) - np.float32(1e-6), 2.0 * s])
    ev[0, 1, :2] = torch.tensor([-0.5, 3.0])              # iy = -1 -> dropped
    ev[0, 2, :2] = torch.tensor([float(H + 3 * s), 1.0])  # iy >= Hq -> dropped
    ev[0, 3, 0] = float("nan")                            # dropped
    ev[0, 4, 4] = float(nb)                               # bin out of range -> dropped
    ev[0, 5, 5] = 0.0                                     # padding row
    pk = io.pack_events_host(ev, npos, cfg, layout=layout)
    G, nt = layout[3], layout[1] * layout[2]
    assert pk.seg_start.shape == (3, G * nt + 1) and pk.seg_start.dtype == torch.int32
    assert pk.records.dtype == torch.float32 and pk.records.shape[2] == 4
    # brute force, with the reference's own index arithmetic (focus.py:185-187)
    want = {}
    dropped = 0
    for b in range(3):
        for m in range(ev.shape[1]):
            y, x, t, _, bn, v = (ev[b, m, k] for k in range(6))
            if v == 0:
                continue
            it, iy, ix = bn.to(torch.int64) if bn == bn else torch.tensor(-1), y // s, x // s
            if not (bn == bn and 0 <= it < nb and 0 <= iy < Hq and 0 <= ix < Wq):
                dropped += 1
                continue
            it, iy, ix = int(it), int(iy), int(ix)
            g = 1 if (pab and m >= npos) else 0
            k = g * nt + (iy // ct) * layout[2] + ix // ct
            meta = (it << 24) | (iy << 12) | ix
            row = tuple(np.array([y, x, t], np.float32).view(np.uint32).tolist()) + (meta,)
            want.setdefault((b, k), []).append(row)
    want = {k: sorted(v) for k, v in want.items()}
    assert dropped >= 4 and int(pk.skipped[0]) == dropped
    assert _segments(pk, layout) == want
    assert pk.seg_start[:, -1].tolist() == [sum(len(v) for (b, _), v in want.items() if b == i) for i in range(3)]
    # round trip: unpack -> pack gives the same segments
    ev2, npos2 = io.unpack_events(pk, cfg, layout=layout)
    pk2 = io.pack_events_host(ev2, npos2, cfg, layout=layout)
    assert _segments(pk2, layout) == want


@pytest.mark.parametrize("s,pab", [(4, True), (3, False), (16, True)])
def test_native_host_packer_equals_torch_host_packer(s, pab):
    """cmax_pack_events_host (C++ / OpenMP, host pointers only - no GPU involved) against
    io.pack_events_host: same segments, same records in the same order, same drop counters."""
    from motionpriorcmax_b200 import io, synthetic
    H, W, nb = 120, 150, 9
    d = dict(synthetic.DSEC_LOSS_CONFIG, image_shape=(H, W), num_bins=nb, lut_superpixel_size=s,
             num_knn=2, polarity_aware_batching=pab)
    cfg = _cfg(d)
    ev, npos = synthetic.make_event_batch(4, [9000, 4000, 0, 12000], H, W, nb, pab, seed=6)
    ev = ev.clone()
    rng = np.random.default_rng(1)
    k = 400
    # coordinates a few ulp around cell edges, rows outside the table, NaN / inf, odd `valid`
    edge = (rng.integers(0, H // s + 2, k) * s).astype(np.float32)
    ev[0, :k, 0] = torch.as_tensor(edge + rng.choice(np.array([-2e-6, -1e-6, 0, 1e-6, 1e-5], np.float32), k))
    ev[0, k:k + 50, 1] = -0.25
    ev[1, :30, 4] = float(nb)
    ev[1, 40, 0] = float("nan")
    ev[1, 41, 1] = float("inf")
    ev[1, 42, 4] = float("nan")
    ev[3, 7, 5] = 0.5
    for packer in (io.pack_events_host, io.pack_events_native):      # a weighted row is refused by default
        with pytest.raises(ValueError, match="neither 0 nor 1"):
            packer(ev, npos, cfg)
    a = io.pack_events_host(ev, npos, cfg, strict=False)
    b = io.pack_events_native(ev, npos, cfg, strict=False)
    assert torch.equal(a.seg_start, b.seg_start)
    assert a.skipped.tolist() == b.skipped.tolist() and int(a.skipped[0]) >= 80 and int(a.skipped[1]) == 1
    for i, c in enumerate(a.seg_start[:, -1].tolist()):
        assert torch.equal(a.records[i, :c].view(torch.int32), b.records[i, :c].view(torch.int32)), i
    # capacity too small for a window -> error code, not an overflow
    from motionpriorcmax_b200 import cabi
    import ctypes
    lib = cabi.load()
    rec = torch.zeros((4, 10, 4))
    seg = torch.zeros_like(a.seg_start)
    e = ev.contiguous()
    rc = lib.cmax_pack_events_host(cfg, ctypes.c_void_p(e.data_ptr()), 4, e.shape[1], npos if pab else 0,
                                   ctypes.c_void_p(rec.data_ptr()), 10, ctypes.c_void_p(seg.data_ptr()), None)
    assert rc == -2


def test_packed_collate_example_pickles_and_packs():
    """examples/packed_collate.py: the drop-in collate survives pickling (worker processes) and
    yields the layout of the host packer."""
    import os
    import pickle
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
    from packed_collate import PackedCollate
    from motionpriorcmax_b200 import io, synthetic
    cfg = dict(synthetic.DSEC_LOSS_CONFIG, image_shape=(64, 96), num_bins=5, num_knn=4)
    ev, npos = synthetic.make_event_batch(3, [3000, 1000, 2000], 64, 96, 5, True, seed=12)
    collate = pickle.loads(pickle.dumps(PackedCollate(cfg)))
    batch = collate({"events": ev, "num_pos_events": npos, "other": 1})
    ref = io.pack_events_host(ev, npos, collate._config())
    assert batch["other"] == 1 and torch.equal(batch["events"].seg_start, ref.seg_start)
    for b, c in enumerate(ref.seg_start[:, -1].tolist()):
        assert torch.equal(batch["events"].records[b, :c].view(torch.int32), ref.records[b, :c].view(torch.int32))
    pickle.dumps(collate)                       # still picklable after the config was built


def test_pack_layout_query_and_limits():
    from motionpriorcmax_b200 import cabi, synthetic
    lib = cabi.load()
    cfg = _cfg(synthetic.DSEC_LOSS_CONFIG)
    assert cabi.pack_layout(cfg) == (8, 15, 20, 2)
    ev = _cfg(synthetic.EVIMO2_LOSS_CONFIG)
    assert cabi.pack_layout(ev) == (8, 12, 16, 2)
    big = _cfg(dict(synthetic.DSEC_LOSS_CONFIG, num_bins=300))
    out = (cabi.c_int32 * 4)()
    assert lib.cmax_pack_layout(big, out) == -5        # CMAX_ERR_UNSUPPORTED: bin does not fit the meta word
    assert lib.cmax_pack_layout(None, out) == -1
    # argument validation of the packed entry points without a GPU
    assert lib.cmax_forward_packed(cfg, None, None, None, None, 1, 10, 19200, None, None, None, None, 0, None) == -2
    assert lib.cmax_backward_packed(cfg, None, None, None, None, 1, 10, 19200, None, None, None, 0, None) == -2
    assert lib.cmax_pack_events(cfg, None, 1, 10, 5, None, None, None, None, None) == -2


# ------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("s,pab", [(4, True), (3, False), (16, True)])
def test_compact_host_packer_layout_contract(s, pab):
    """cmax_pack_events_host_compact (host pointers only): every (group, tile, bin) run holds exactly
    the rows of the reference layout that belong there, in their original order; 12 bytes per row;
    the windows sit back to back in one buffer; same drop counters as the 16-byte packer."""
    from motionpriorcmax_b200 import cabi, io, synthetic
    H, W, nb = 120, 150, 9
    d = dict(synthetic.DSEC_LOSS_CONFIG, image_shape=(H, W), num_bins=nb, lut_superpixel_size=s,
             num_knn=2, polarity_aware_batching=pab)
    cfg = _cfg(d)
    ct, nty, ntx, G = cabi.pack_layout(cfg)
    ev, npos = synthetic.make_event_batch(4, [9000, 4000, 0, 12000], H, W, nb, pab, seed=6)
    ev = ev.clone()
    ev[0, 400:450, 1] = -0.25                  # outside the table
    ev[1, :30, 4] = float(nb)                  # bin out of range
    ev[1, 40, 0] = float("nan")
    c = io.pack_events_compact(ev, npos, cfg)
    ref = io.pack_events_native(ev, npos, cfg)
    assert c.skipped.tolist() == ref.skipped.tolist()
    assert c.coords.shape[1] == 3 and c.coords.dtype == torch.float32
    assert c.sample_off.tolist() == [0] + torch.cumsum(ref.seg_start[:, -1].long(), 0).tolist()
    assert torch.equal(c.fine_start[:, ::nb], ref.seg_start)              # coarse table = 16-byte layout's
    assert c.max_count == int(ref.seg_start[:, -1].max())
    Hq, Wq = -(-H // s), -(-W // s)
    e = ev.numpy()
    for b in range(4):
        it = e[b, :, 4]
        ok = (e[b, :, 5] != 0) & (it == it) & (np.trunc(np.nan_to_num(it)) >= 0) & (np.trunc(np.nan_to_num(it)) < nb)
        fy, fx = (torch.as_tensor(e[b, :, 0]) // s).numpy(), (torch.as_tensor(e[b, :, 1]) // s).numpy()
        ok &= (fy >= 0) & (fy < Hq) & (fx >= 0) & (fx < Wq)
        grp = (np.arange(e.shape[1]) >= npos).astype(np.int64) if pab else np.zeros(e.shape[1], np.int64)
        key = np.where(ok, (grp * nty * ntx + (np.nan_to_num(fy).astype(np.int64) // ct) * ntx
                            + np.nan_to_num(fx).astype(np.int64) // ct) * nb
                       + np.trunc(np.nan_to_num(it)).astype(np.int64), -1)
        base = int(c.sample_off[b])
        fs = c.fine_start[b].numpy()
        for k in np.unique(key[key >= 0]):
            rows = e[b, key == k, :3]                                     # original order
            got = c.coords[base + fs[k]:base + fs[k + 1]].numpy()
            assert got.shape == rows.shape and np.array_equal(got.view(np.uint32), rows.view(np.uint32)), (b, k)
        assert fs[-1] == int((key >= 0).sum())


def _decode_bitpacked(bp):
    """numpy restatement of the bit-packed layout contract (include/cmax_b200.h): uint32 [T, 3]."""
    words = bp.words.numpy().view(np.uint32)
    hdr = bp.run_hdr.numpy().view(np.uint32)
    out = []
    for b in range(bp.fine_start.shape[0]):
        w = words[int(bp.word_off[b]):]
        fs = bp.fine_start[b].numpy()
        for f in np.nonzero(np.diff(fs))[0]:
            wd = [int(hdr[b, f, 3] >> (8 * c)) & 255 for c in range(3)]
            for k in range(int(fs[f + 1] - fs[f])):
                bit = int(bp.run_word[b, f]) * 32 + k * sum(wd)
                row = []
                for c in range(3):
                    v = 0
                    if wd[c]:
                        wi, sh = bit >> 5, bit & 31
                        v = ((int(w[wi]) | (int(w[wi + 1]) << 32)) >> sh) & ((1 << wd[c]) - 1)
                    row.append((int(hdr[b, f, c]) + v) & 0xffffffff)
                    bit += wd[c]
                out.append(row)
    return np.array(out, np.uint32).reshape(-1, 3)


@pytest.mark.parametrize("s,pab", [(4, True), (3, False)])
def test_bitpacked_host_packer_is_lossless(s, pab):
    """cmax_pack_events_host_bitpacked: decoding the bit stream with a numpy restatement of the
    contract gives back the compact layout's (y, x, t) BIT FOR BIT - also for values that make a run
    wide (-0.0, tiny and huge t, NaN t, coordinates a few ulp around cell edges) - in fewer bytes."""
    from motionpriorcmax_b200 import io, synthetic
    H, W, nb = 120, 150, 9
    d = dict(synthetic.DSEC_LOSS_CONFIG, image_shape=(H, W), num_bins=nb, lut_superpixel_size=s,
             num_knn=2, polarity_aware_batching=pab)
    cfg = _cfg(d)
    ev, npos = synthetic.make_event_batch(3, [9000, 0, 4000], H, W, nb, pab, seed=6)
    ev = ev.clone()
    ev[0, 0, 0] = -0.0
    ev[0, 1, 2] = float("nan")
    ev[0, 2, 2] = 1e-30
    ev[0, 3, 2] = -3.0
    ev[0, 4, 2] = 1e30
    rng = np.random.default_rng(4)
    k = 300
    edge = (rng.integers(0, H // s, k) * s).astype(np.float32)
    ev[2, :k, 0] = torch.as_tensor(edge + rng.choice(np.array([0, 1e-6, 1e-5], np.float32), k))
    c = io.pack_events_compact(ev, npos, cfg)
    b = io.pack_events_bitpacked(ev, npos, cfg)
    assert torch.equal(c.fine_start, b.fine_start) and c.skipped.tolist() == b.skipped.tolist()
    T = int(c.sample_off[-1])
    assert np.array_equal(_decode_bitpacked(b), c.coords.numpy().view(np.uint32)[:T])
    assert b.max_count == c.max_count
    # the sizing-call path (taken for very large batches) writes the same bytes as the one-call path
    old_limit, io._BITPACK_ONE_CALL_MAX_WORDS = io._BITPACK_ONE_CALL_MAX_WORDS, 16
    try:
        b2 = io.pack_events_bitpacked(ev, npos, cfg)
    finally:
        io._BITPACK_ONE_CALL_MAX_WORDS = old_limit
    assert all(torch.equal(getattr(b, k), getattr(b2, k)) for k in ("words", "run_hdr", "run_word", "word_off", "fine_start"))
    # a realistic window: well under 12 bytes per event
    ev2, npos2 = synthetic.make_event_batch(1, [300_000], 480, 640, 15, True, seed=2)
    big = io.pack_events_bitpacked(ev2, npos2, _cfg(dict(synthetic.DSEC_LOSS_CONFIG)))
    assert big.nbytes() < 9.5 * 300_000


def test_bitpacked_round_trip_on_arbitrary_bit_patterns():
    """Property test (hypothesis): any float32 bit pattern in t (NaN payloads, infinities, denormals,
    either sign) and any in-range coordinate mantissa survives pack -> decode bit for bit, whatever
    the run widths come out as (0 ... 32 bits per field)."""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st
    from hypothesis.extra import numpy as hnp
    from motionpriorcmax_b200 import io, synthetic
    H, W, nb, s = 16, 24, 3, 4
    cfg = _cfg(dict(synthetic.DSEC_LOSS_CONFIG, image_shape=(H, W), num_bins=nb, lut_superpixel_size=s,
                    num_knn=2, polarity_aware_batching=True))

    @settings(max_examples=40, deadline=None)
    @given(hnp.arrays(np.uint32, (37,), elements=st.integers(0, 2 ** 32 - 1)),
           hnp.arrays(np.uint32, (37, 2), elements=st.integers(0, 2 ** 23 - 1)),
           st.integers(0, 37))
    def check(tbits, mant, npos):
        n = len(tbits)
        ev = np.zeros((1, n, 6), np.float32)
        # coordinates in [0, H) x [0, W): integer part from the high mantissa bits, arbitrary low bits
        ev[0, :, 0] = (mant[:, 0].astype(np.float64) / 2 ** 23 * H).astype(np.float32).clip(0, np.nextafter(np.float32(H), np.float32(0)))
        ev[0, :, 1] = (mant[:, 1].astype(np.float64) / 2 ** 23 * W).astype(np.float32).clip(0, np.nextafter(np.float32(W), np.float32(0)))
        ev[0, :, 2] = tbits.view(np.float32)
        ev[0, :, 3] = 1.0
        ev[0, :, 4] = (np.arange(n) % nb).astype(np.float32)
        ev[0, :, 5] = 1.0
        e = torch.from_numpy(ev)
        c = io.pack_events_compact(e, npos, cfg)
        b = io.pack_events_bitpacked(e, npos, cfg)
        T = int(c.sample_off[-1])
        assert T == n and torch.equal(c.fine_start, b.fine_start)
        assert np.array_equal(_decode_bitpacked(b), c.coords.numpy().view(np.uint32)[:T])

    check()


def _cuda():
    assert torch.cuda.is_available(), "these tests need a GPU (run with -m gpu on a B200)"
    return torch.device("cuda:0")


def _run(cfg, traj, times, events, npos, mode, deterministic=False):
    """mode: 'plain' (upstream layout), 'packed_dev' (cmax_pack_events), 'packed_host'."""
    from motionpriorcmax_b200 import io
    from motionpriorcmax_b200.losses import LossFactory
    dev = _cuda()
    L = LossFactory.get_loss_calculator("FOCUS", dict(cfg, deterministic=deterministic))
    t = torch.as_tensor(traj, device=dev).clone().requires_grad_()
    ev = torch.as_tensor(events)
    if mode == "plain":
        batch = {"events": ev.to(dev)}
        if npos >= 0:
            batch["num_pos_events"] = npos
    elif mode == "packed_dev":
        batch = {"events": io.pack_events(ev.to(dev), npos if npos >= 0 else None, L)}
    elif mode == "compact":      # 12-byte wire layout, expanded on the device inside calc
        batch = {"events": io.pack_events_compact(ev, npos if npos >= 0 else None, L).to(dev)}
    elif mode == "bitpacked":    # ~8-byte lossless wire layout (bit-pattern deltas per run)
        batch = {"events": io.pack_events_bitpacked(ev, npos if npos >= 0 else None, L).to(dev)}
    else:
        batch = {"events": io.pack_events_host(ev, npos if npos >= 0 else None, L).to(dev)}
    loss, log, misc = L.calc(t, torch.as_tensor(times, device=dev), batch, return_flow_lut=True)
    loss.backward()
    torch.cuda.synchronize()
    return dict(loss=loss.item(), focus=log["focus_loss"].item(), smooth=log["smoothness_loss"].item(),
                iwes=misc["iwes"].cpu().numpy(), lut=misc["flow_lut"].cpu().numpy(),
                dtraj=t.grad.cpu().numpy(), packed=batch["events"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", LOSS_CASES)
@pytest.mark.parametrize("mode", ["packed_dev", "packed_host", "compact", "bitpacked"])
def test_packed_matches_reference_golden(name, mode):
    from test_gpu_parity import _assert_grad_close
    c = load_case(name)
    r = _run(c["cfg"], c["trajectories"], c["times"], c["events"], c["num_pos_events"], mode)
    assert abs(r["loss"] - float(c["loss"])) <= TOL * abs(float(c["loss"]))
    assert abs(r["focus"] - float(c["focus_loss"])) <= TOL * abs(float(c["focus_loss"]))
    assert rel_err(r["iwes"], c["iwes"]) < TOL
    _assert_grad_close(r["dtraj"], c["dtraj"], c["cfg"], c["trajectories"], c["times"], c["events"],
                       c["num_pos_events"], tol=TOL, gpu_iwes=r["iwes"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", LOSS_CASES)
def test_packed_deterministic_is_bit_identical_to_unpacked(name):
    """int64 fixed-point sums do not depend on the order or the grouping of the events: the tile
    kernels must reproduce the unpacked kernels bit for bit (IWE, loss, gradients)."""
    c = load_case(name)
    a = _run(c["cfg"], c["trajectories"], c["times"], c["events"], c["num_pos_events"], "plain", True)
    for mode in ("packed_dev", "packed_host", "compact", "bitpacked"):
        b = _run(c["cfg"], c["trajectories"], c["times"], c["events"], c["num_pos_events"], mode, True)
        assert a["loss"] == b["loss"] and a["focus"] == b["focus"]
        assert np.array_equal(a["iwes"], b["iwes"])
        assert np.array_equal(a["dtraj"], b["dtraj"])


@pytest.mark.gpu
def test_compact_expands_to_the_host_packed_layout():
    """cmax_expand_compact: the 16-byte records rebuilt on the device from the 12-byte wire layout
    are, segment by segment, the records of the host packer (bit for bit: the LUT cell in `meta`
    is recomputed with the reference's float floor division), through the uploader too."""
    from motionpriorcmax_b200 import cabi, io, synthetic
    dev = _cuda()
    for d, B, M in ((dict(synthetic.DSEC_LOSS_CONFIG), 3, [150_000, 40_000, 0]),
                    (dict(synthetic.DSEC_LOSS_CONFIG, image_shape=(50, 70), lut_superpixel_size=3,
                          polarity_aware_batching=False, num_bins=5, num_knn=2), 2, [5000, 7000])):
        cfg = _cfg(d)
        H, W = d["image_shape"]
        ev, npos = synthetic.make_event_batch(B, M, H, W, d["num_bins"], d["polarity_aware_batching"], seed=8)
        ev = ev.clone()
        k = 300                                # coordinates a few ulp around cell edges
        s = d["lut_superpixel_size"]
        rng = np.random.default_rng(2)
        edge = (rng.integers(0, H // s, k) * s).astype(np.float32)
        ev[0, :k, 0] = torch.as_tensor(edge + rng.choice(np.array([-2e-6, -1e-6, 0, 1e-6, 1e-5], np.float32), k)).clamp_(0)
        # arbitrary float32 bit patterns in t (NaN payloads, infinities, denormals, both signs): the
        # device decoder must hand back every one of them (runs up to 32 bits wide)
        tb = rng.integers(0, 2 ** 32, 2000, dtype=np.uint64).astype(np.uint32)
        ev[1, 100:2100, 2] = torch.from_numpy(tb.view(np.float32).copy())
        layout = cabi.pack_layout(cfg)
        host = io.pack_events_native(ev, npos, cfg)
        comp = io.pack_events_compact(ev, npos, cfg)
        a = io.expand_compact(comp.to(dev), cfg)
        up = io.CompactUploader(dev, cfg)
        _, slot = up.upload(comp.pin_memory())
        b = up.wait(slot)
        torch.cuda.synchronize()
        assert up.bytes_last == 12 * int(comp.sample_off[-1]) + comp.fine_start.numel() * 4 + comp.sample_off.numel() * 8
        bp = io.pack_events_bitpacked(ev, npos, cfg)
        c = io.expand_bitpacked(bp.to(dev), cfg)
        up2 = io.CompactUploader(dev, cfg)
        _, slot2 = up2.upload(bp.pin_memory())
        d2 = up2.wait(slot2)
        torch.cuda.synchronize()
        assert up2.bytes_last == bp.nbytes() - (bp.words.numel() - int(bp.word_off[-1])) * 4
        nrec = host.records.shape[1]                     # bit for bit the records of the 12-byte layout
        for bb, cnt in enumerate(host.seg_start[:, -1].tolist()):
            assert torch.equal(c.records[bb, :cnt].view(torch.int32), a.records[bb, :cnt].view(torch.int32))
        for got in (a, b, c, d2):
            assert torch.equal(host.seg_start, got.seg_start.cpu())
            assert _segments(host, layout) == _segments(
                io.PackedEvents(got.records[:, :host.records.shape[1]], got.seg_start), layout)


@pytest.mark.gpu
def test_device_packer_equals_host_packer():
    from motionpriorcmax_b200 import cabi, io, synthetic
    dev = _cuda()
    for d, B, M in ((dict(synthetic.DSEC_LOSS_CONFIG), 3, [150_000, 40_000, 0]),
                    (dict(synthetic.DSEC_LOSS_CONFIG, image_shape=(50, 70), lut_superpixel_size=3,
                          polarity_aware_batching=False, num_bins=5, num_knn=2), 2, [5000, 7000])):
        cfg = _cfg(d)
        H, W = d["image_shape"]
        ev, npos = synthetic.make_event_batch(B, M, H, W, d["num_bins"], d["polarity_aware_batching"], seed=8)
        ev = ev.clone()
        ev[0, :50, 0] = -1.0                  # outside the table
        ev[1, 10:20, 4] = 99.0                # bin out of range
        ev[0, 100, 1] = float("inf")
        layout = cabi.pack_layout(cfg)
        host = io.pack_events_host(ev, npos, cfg)
        devp = io.pack_events(ev.to(dev), npos, cfg)
        torch.cuda.synchronize()
        assert torch.equal(host.seg_start, devp.seg_start.cpu())
        assert _segments(host, layout) == _segments(
            io.PackedEvents(devp.records[:, :host.records.shape[1]], devp.seg_start), layout)
        assert devp.skipped.cpu().tolist() == host.skipped.tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["dsec_edges", "big_flow", "multi_tref3", "s2_det", "evimo_next"])
def test_packed_matches_oracle_midsize(variant):
    """Mid-size windows against the float64 oracle; `big_flow` pushes most votes out of the
    shared-memory window (global fallback), `s2_det` uses 16x16-cell tiles in fixed point."""
    from motionpriorcmax_b200 import synthetic
    from oracle import focus_oracle as fo
    from test_gpu_parity import _assert_grad_close, _synthetic_case
    base = dict(synthetic.DSEC_LOSS_CONFIG, image_shape=(96, 128), num_knn=16)
    det, dist = False, "uniform"
    if variant == "dsec_edges":
        dist = "edges"
    elif variant == "multi_tref3":
        base = synthetic.multi_tref_variant(base, 3)
    elif variant == "s2_det":
        base.update(lut_superpixel_size=2)
        det = True
    elif variant == "evimo_next":
        base.update(smooth_type="on_flow_to_next", smooth_weight=0.06, num_bins=9)
    traj, times, ev, npos, _ = _synthetic_case(base, 3, [20000, 35000, 9000], 2, seed=21, dist=dist)
    if variant == "big_flow":
        pos = traj[:, -1:, :, :] * 0 + fo.tile_positions((96, 128), 4).astype(np.float32)[None, None]
        traj = (pos + (traj - pos) * 6.0).astype(np.float32)        # flows of several tens of pixels
    r = _run(base, traj, times, ev, npos, "packed_dev", det)
    o = fo.FocusOracle(**base, dtype=np.float64)
    f = o.forward(traj, times, ev, npos)
    g = o.backward()
    assert abs(r["loss"] - f["loss"]) <= TOL * abs(f["loss"])
    assert abs(r["smooth"] - f["smoothness_loss"]) <= TOL * max(abs(f["smoothness_loss"]), 1e-3)
    assert rel_err(r["iwes"], f["iwes"]) < TOL
    _assert_grad_close(r["dtraj"], g["dtraj"], base, traj, times, ev, npos, gpu_iwes=r["iwes"])


@pytest.mark.gpu
def test_packed_full_size_properties():
    """DSEC-size batch: deterministic packed == deterministic unpacked bit for bit, with large
    segments (slices per tile) and one empty window; float mode within tolerance."""
    from motionpriorcmax_b200 import synthetic
    from test_gpu_parity import _synthetic_case
    cfg = dict(synthetic.DSEC_LOSS_CONFIG)
    traj, times, ev, npos, _ = _synthetic_case(cfg, 3, [1_500_000, 300_000, 0], 1, seed=31)
    a = _run(cfg, traj, times, ev, npos, "plain", True)
    b = _run(cfg, traj, times, ev, npos, "packed_dev", True)
    assert a["loss"] == b["loss"]
    assert np.array_equal(a["iwes"], b["iwes"]) and np.array_equal(a["dtraj"], b["dtraj"])
    # float atomics: the l2 focus norm keeps the gradient smooth (with l1, sign(Sobel) of a ~0
    # response flips with the summation order - see test_gpu_parity._assert_grad_close)
    cfg2 = dict(cfg, focus_loss_norm="l2")
    a2 = _run(cfg2, traj, times, ev, npos, "plain", True)
    c = _run(cfg2, traj, times, ev, npos, "packed_dev", False)
    assert abs(a2["loss"] - c["loss"]) <= TOL * abs(a2["loss"])
    assert rel_err(c["iwes"], a2["iwes"]) < TOL and rel_err(c["dtraj"], a2["dtraj"]) < TOL


@pytest.mark.gpu
def test_packed_slices_empty_and_padding():
    """Segments larger than 8 k events are cut into slices (several CTAs per tile); windows that are
    empty or all padding, and M = 0, behave like the unpacked call."""
    from motionpriorcmax_b200 import synthetic
    from test_gpu_parity import _synthetic_case
    cfg = dict(synthetic.DSEC_LOSS_CONFIG, image_shape=(96, 128), num_knn=16)
    # 12 tiles x 2 groups and 400 k events per window -> ~17 k events per segment -> 3 slices
    traj, times, ev, npos, _ = _synthetic_case(cfg, 2, [400_000, 150_000], 2, seed=41, dist="edges")
    a = _run(cfg, traj, times, ev, npos, "plain", True)
    b = _run(cfg, traj, times, ev, npos, "packed_dev", True)
    assert a["loss"] == b["loss"] and np.array_equal(a["iwes"], b["iwes"]) and np.array_equal(a["dtraj"], b["dtraj"])
    # second window entirely padding
    ev2 = ev.copy()
    ev2[1] = 0
    a = _run(cfg, traj, times, ev2, npos, "plain", True)
    for mode in ("packed_dev", "packed_host"):
        b = _run(cfg, traj, times, ev2, npos, mode, True)
        assert a["loss"] == b["loss"] and np.array_equal(a["iwes"], b["iwes"]) and np.array_equal(a["dtraj"], b["dtraj"])
        assert b["packed"].seg_start[:, -1].tolist() == [int(ev2[0, :, 5].sum()), 0]
    b = _run(cfg, traj, times, ev2, npos, "compact", True)          # the 12-byte wire layout, sliced segments
    assert a["loss"] == b["loss"] and np.array_equal(a["iwes"], b["iwes"]) and np.array_equal(a["dtraj"], b["dtraj"])
    assert b["packed"].sample_off.tolist() == [0, int(ev2[0, :, 5].sum()), int(ev2[0, :, 5].sum())]
    # M = 0: no event kernel runs; 1 / mean(0) = inf like the reference
    c = _run(cfg, traj, times, ev[:, :0], 0, "packed_dev")
    assert np.isinf(c["loss"]) and np.abs(c["iwes"]).max() == 0.0
    # every window empty / all padding through the compact layout (T = 0: nothing crosses PCIe)
    c = _run(cfg, traj, times, np.zeros_like(ev), npos, "compact")
    assert np.isinf(c["loss"]) and np.abs(c["iwes"]).max() == 0.0 and int(c["packed"].sample_off[-1]) == 0


@pytest.mark.gpu
def test_packed_votes_leaving_the_image():
    """Flows that push events across the image border (out-of-bounds corners, border mask on and
    off): deterministic packed == deterministic unpacked, bit for bit."""
    from motionpriorcmax_b200 import synthetic
    from oracle import focus_oracle as fo
    from test_gpu_parity import _synthetic_case
    for mask in (True, False):
        cfg = dict(synthetic.DSEC_LOSS_CONFIG, image_shape=(64, 96), num_knn=8, num_bins=5,
                   mask_image_border=mask)
        traj, times, ev, npos, _ = _synthetic_case(cfg, 2, [40_000, 25_000], 1, seed=43)
        pos = fo.tile_positions((64, 96), 4).astype(np.float32)[None, None]
        traj = (pos + (traj - pos) * 4.0 + np.array([9.0, -7.0], np.float32)).astype(np.float32)
        a = _run(cfg, traj, times, ev, npos, "plain", True)
        b = _run(cfg, traj, times, ev, npos, "packed_host", True)
        assert a["loss"] == b["loss"] and np.array_equal(a["iwes"], b["iwes"]) and np.array_equal(a["dtraj"], b["dtraj"])


def test_packing_is_neutral_for_the_reference_algorithm():
    """CPU, oracle only: the reference algorithm (float64 restatement) gives the same loss, IWEs
    and gradients on unpack(pack(events)) as on the collated batch - dropping the padding rows and
    regrouping the events by tile changes nothing but the summation order."""
    from motionpriorcmax_b200 import io, synthetic
    from oracle import focus_oracle as fo
    cfg = dict(synthetic.DSEC_LOSS_CONFIG, image_shape=(48, 64), num_knn=6, num_bins=5)
    H, W = cfg["image_shape"]
    times = fo.reconstruction_times(1, cfg["num_bins"], 0.37)
    cg = synthetic.make_coeff_grid(2, 2, H, W, sigma_px=5.0, seed=4, coarse=(4, 5)).numpy()
    traj, _ = fo.trajectories_from_coeff_grid(cg, times, 4, 2, "polynomial")
    ev, npos = synthetic.make_event_batch(2, [2500, 1200], H, W, cfg["num_bins"], True, seed=4)
    c = _cfg(cfg)
    ct = 8
    layout = (ct, -(-(H // 4) // ct), -(-(W // 4) // ct), 2)
    pk = io.pack_events_host(ev, npos, c, layout=layout)
    assert int(pk.skipped[0]) == 0
    ev2, npos2 = io.unpack_events(pk, c, layout=layout)
    a = fo.FocusOracle(**cfg, dtype=np.float64)
    fa = a.forward(traj, times, ev.numpy(), npos)
    ga = a.backward()
    b = fo.FocusOracle(**cfg, dtype=np.float64)
    fb = b.forward(traj, times, ev2.numpy(), npos2)
    gb = b.backward()
    assert abs(fa["loss"] - fb["loss"]) <= 1e-12 * abs(fa["loss"])
    assert rel_err(fb["iwes"], fa["iwes"]) < 1e-12
    assert rel_err(gb["dtraj"], ga["dtraj"]) < 1e-10
