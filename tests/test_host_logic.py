"""Host-side mirror of the reference plugin interface and the synthetic loader layout (CPU)."""
import numpy as np
import pytest
import torch

from helpers import LOSS_CASES, load_case


def test_synthetic_batch_follows_loader_layout():
    from motionpriorcmax_b200 import synthetic
    ev, npos = synthetic.make_event_batch(3, [500, 900, 200], 48, 64, 15, True, seed=5)
    assert ev.dtype == torch.float32 and ev.shape[0] == 3 and ev.shape[2] == 6
    for b in range(3):
        pos, neg = ev[b, :npos], ev[b, npos:]
        for part, pol in ((pos, 1.0), (neg, 0.0)):
            valid = part[:, 5] == 1
            n = int(valid.sum())
            assert valid[:n].all() and not valid[n:].any()          # padding at the end
            assert (part[n:] == 0).all()                            # loader.py:360-364
            assert (part[:n, 3] == pol).all()
            assert (part[:n, 2].diff() >= 0).all()                  # time sorted
            assert ((part[:n, 0] >= 0) & (part[:n, 0] < 48) & (part[:n, 1] >= 0) & (part[:n, 1] < 64)).all()
            t = part[:n, 2].numpy()
            edges = np.linspace(0, 1, 16, dtype=np.float32)
            want = np.clip(np.searchsorted(edges, t) - 1, 0, 14)
            assert np.array_equal(part[:n, 4].numpy(), want.astype(np.float32))
    assert npos == max(int((ev[b, :, 3] == 1).sum()) for b in range(3))
    ev2, none = synthetic.make_event_batch(2, 100, 48, 64, 15, False, seed=5)
    assert none is None and ev2.shape == (2, 100, 6)
    # different ranks draw different windows, same rank is reproducible
    a, _ = synthetic.make_event_batch(1, 50, 48, 64, 15, False, seed=1, rank=0)
    b, _ = synthetic.make_event_batch(1, 50, 48, 64, 15, False, seed=1, rank=1)
    c, _ = synthetic.make_event_batch(1, 50, 48, 64, 15, False, seed=1, rank=0)
    assert not torch.equal(a, b) and torch.equal(a, c)


def test_plugin_interface_mirrors_reference():
    from motionpriorcmax_b200.losses import LossFactory, TrajectoryLossBase, FocusLoss
    from motionpriorcmax_b200 import synthetic
    with pytest.raises(ValueError, match="Unsupported loss type"):
        LossFactory.get_loss_calculator("OTHER", {})
    L = LossFactory.get_loss_calculator("FOCUS", dict(synthetic.DSEC_LOSS_CONFIG), profiler=None)
    assert isinstance(L, FocusLoss) and isinstance(L, TrajectoryLossBase)
    assert L.is_needing_offsets is True and hasattr(L.imager, "create_iwe")
    # extra kwargs are swallowed like the reference's **kwargs (flow_training.py:34-52 injects some)
    LossFactory.get_loss_calculator("FOCUS", dict(synthetic.DSEC_LOSS_CONFIG, patch_size=4, loss_name="FOCUS"))
    for bad in (dict(num_tref=2), dict(num_tref=2, scale_iwe_by_dt=False),
                dict(num_tref=2, scale_iwe_by_dt=False, polarity_aware_batching=False,
                     smooth_type="on_flow_to_next")):
        with pytest.raises(AssertionError):
            LossFactory.get_loss_calculator("FOCUS", dict(synthetic.DSEC_LOSS_CONFIG, **bad))
    with pytest.raises(TypeError):       # a required keyword is missing, as in the reference
        FocusLoss(image_shape=(4, 4))


def test_reconstruction_times_match_reference_formula():
    from motionpriorcmax_b200.losses import LossFactory
    from motionpriorcmax_b200 import synthetic
    from oracle import focus_oracle as fo
    L = LossFactory.get_loss_calculator("FOCUS", dict(synthetic.DSEC_LOSS_CONFIG))
    torch.manual_seed(3)
    t = L.get_reconstruction_times("cpu")
    torch.manual_seed(3)
    want_ref = torch.rand(1)
    assert t.shape == (16,) and t[0] == want_ref[0]                 # same RNG draw as focus.py:57
    assert np.allclose(t[1:].numpy(), fo.reconstruction_times(1, 15, 0.0)[1:])
    L3 = LossFactory.get_loss_calculator("FOCUS", synthetic.multi_tref_variant(synthetic.DSEC_LOSS_CONFIG, 3))
    assert np.allclose(L3.get_reconstruction_times("cpu").numpy(), fo.reconstruction_times(3, 15))
    for name in LOSS_CASES:                                          # times stored with the goldens
        c = load_case(name)
        if c["cfg"]["num_tref"] > 1:
            assert np.allclose(c["times"], fo.reconstruction_times(c["cfg"]["num_tref"], c["cfg"]["num_bins"]))


def test_cpu_tensors_are_rejected_not_silently_computed():
    from motionpriorcmax_b200.losses import LossFactory
    from motionpriorcmax_b200.utils import EventImageConverter
    from motionpriorcmax_b200 import synthetic, trajectories as tj
    L = LossFactory.get_loss_calculator("FOCUS", dict(synthetic.DSEC_LOSS_CONFIG))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L.calc(torch.zeros(1, 16, 19200, 2), torch.zeros(16),
               {"events": torch.zeros(1, 8, 6), "num_pos_events": 4})
    with pytest.raises(RuntimeError, match="no CPU path"):
        EventImageConverter((8, 8)).create_iwe(torch.zeros(4, 4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        tj.calculate_trajectories_at_t(torch.zeros(1, 2, 8, 8), torch.zeros(3), 4, 1)


def test_tile_mask_and_basis_table_match_oracle():
    from motionpriorcmax_b200 import trajectories as tj
    from oracle import focus_oracle as fo
    for shape, p in (((480, 640), 4), ((30, 45), 4), ((16, 24), 2), ((9, 9), 3)):
        mask = tj.get_optical_flow_tile_mask(shape, p)
        pos = torch.nonzero(mask)
        assert torch.equal(pos, tj.tile_positions(shape, p))
        assert np.array_equal(pos.numpy(), fo.tile_positions(shape, p))
    times = torch.tensor(fo.reconstruction_times(1, 7, 0.3))
    for basis, K in (("polynomial", 3), ("dct", 4), ("bezier", 10)):
        got = tj.basis_table(times, K, basis).numpy()
        want = fo.basis_matrix(times.numpy(), K, basis) - fo.basis_matrix([0.0], K, basis)
        assert np.allclose(got, want, rtol=2e-6, atol=1e-7), basis


def test_valid_prefix_lengths_match_the_valid_column():
    from motionpriorcmax_b200 import synthetic
    from motionpriorcmax_b200.io import valid_prefix_lengths
    ev, npos = synthetic.make_event_batch(4, [5000, 9000, 100, 0], 48, 64, 15, True, seed=2)
    got = valid_prefix_lengths(ev, npos)
    want = np.stack([[int((ev[b, :npos, 5] > 0).sum()), int((ev[b, npos:, 5] > 0).sum())] for b in range(4)])
    assert np.array_equal(got, want)
    ev2, _ = synthetic.make_event_batch(3, [50, 0, 7], 48, 64, 15, False, seed=2)
    assert valid_prefix_lengths(ev2, None)[:, 0].tolist() == [50, 0, 7]


def test_oracle_integer_semantics_equal_torch_property():
    """Property test (hypothesis): the oracle's integer work is bit-identical to the torch
    expressions the reference evaluates - floor(v + 1e-6) in float32 (event_image_converter.py:357),
    float floor division `//` and truncating `.to(int)` (focus.py:185-187)."""
    from hypothesis import given, settings, strategies as st
    from oracle import focus_oracle as fo

    f32 = st.floats(min_value=-700.0, max_value=700.0, allow_nan=False, width=32)
    near_int = st.builds(lambda k, e: np.float32(np.float32(k) + np.float32(e)),
                         st.integers(-3, 650), st.sampled_from([-2e-6, -1e-6, -5e-7, -1e-7, 0.0, 1e-7, 5e-7, 1e-6, 0.5]))

    @settings(max_examples=300, deadline=None)
    @given(st.lists(st.one_of(f32, near_int), min_size=2, max_size=40), st.sampled_from([1, 2, 3, 4, 5, 7, 8]))
    def check(vals, s):
        v = np.asarray(vals, np.float32)
        n = len(v) // 2
        yx = np.stack((v[:n], v[n:2 * n]), -1)[None]
        # corners / masks
        H, W = 480, 640
        inds, mask, frac = fo.vote_corners(yx, (H, W))
        t = torch.from_numpy(yx)
        fl = torch.floor(t + 1e-6)
        y1, x1 = fl[..., 0].long(), fl[..., 1].long()
        ref_inds = torch.stack((x1 + y1 * W, x1 + (y1 + 1) * W, (x1 + 1) + y1 * W, (x1 + 1) + (y1 + 1) * W), -1)
        ref_mask = torch.stack(((0 <= x1) * (x1 < W) * (0 <= y1) * (y1 < H),
                                (0 <= x1) * (x1 < W) * (0 <= y1 + 1) * (y1 + 1 < H),
                                (0 <= x1 + 1) * (x1 + 1 < W) * (0 <= y1) * (y1 < H),
                                (0 <= x1 + 1) * (x1 + 1 < W) * (0 <= y1 + 1) * (y1 + 1 < H)), -1)
        assert np.array_equal(mask, ref_mask.numpy())
        assert np.array_equal(inds, (ref_inds * ref_mask).numpy())
        assert np.array_equal(frac, (t - fl).numpy())
        # LUT cell indices
        ev = np.zeros((1, n, 6), np.float32)
        ev[0, :, 0], ev[0, :, 1], ev[0, :, 4] = np.abs(v[:n]), np.abs(v[n:2 * n]), np.abs(v[:n]) % 15
        it, iy, ix = fo.lut_cell_indices(ev, s)
        te = torch.from_numpy(ev)
        assert np.array_equal(it, te[..., 4].to(int).numpy())
        assert np.array_equal(iy, (te[..., 0] // s).to(int).numpy())
        assert np.array_equal(ix, (te[..., 1] // s).to(int).numpy())

    check()
