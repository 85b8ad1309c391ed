"""Pin the CPU oracle against golden vectors produced by the real reference.

The goldens (tests/golden/*.npz) were written by oracle/make_golden.py, which executes the
unmodified upstream ``src/losses/focus.py`` on CPU with the two import stubs under
``oracle/ref_stubs``.  Tolerances: float32 oracle mirrors the reference op for op, so only
summation order differs (1e-6 norm-wise); the float64 oracle is the "truth" used by the GPU
parity tests and must agree with the float32 reference to 5e-6.
"""
import numpy as np
import pytest

from helpers import LOSS_CASES, GOLDEN_DIR, load_case, rel_err
from oracle import focus_oracle as fo


@pytest.mark.parametrize("name", LOSS_CASES)
@pytest.mark.parametrize("dtype,tol", [(np.float32, 2e-6), (np.float64, 5e-6)])
def test_loss_forward_backward_matches_reference(name, dtype, tol):
    c = load_case(name)
    o = fo.FocusOracle(**c["cfg"], dtype=dtype)
    f = o.forward(c["trajectories"], c["times"], c["events"], c["num_pos_events"])
    g = o.backward()
    assert abs(float(f["loss"]) - float(c["loss"])) <= tol * abs(float(c["loss"]))
    assert abs(float(f["focus_loss"]) - float(c["focus_loss"])) <= tol * abs(float(c["focus_loss"]))
    assert abs(float(f["smoothness_loss"]) - float(c["smoothness_loss"])) <= tol * max(
        1e-3, abs(float(c["smoothness_loss"])))
    assert f["iwes"].shape == c["iwes"].shape
    assert rel_err(f["iwes"], c["iwes"]) < tol
    assert f["flow_lut"].shape == c["flow_lut"].shape
    assert rel_err(f["flow_lut"], c["flow_lut"]) < tol
    if "flow_to_next" in c:
        ref = c["flow_to_next"].reshape(f["flow_to_next"].shape)
        assert rel_err(f["flow_to_next"], ref) < tol
    assert rel_err(g["dtraj"], c["dtraj"]) < 4 * tol


@pytest.mark.parametrize("name", [n for n in LOSS_CASES if n != "free_points_b1"])
def test_front_end_matches_reference(name):
    c = load_case(name)
    basis = str(c["basis"])
    tr, pos = fo.trajectories_from_coeff_grid(c["coeff_grid"], c["times"], int(c["patch"]),
                                              int(c["num_basis"]), basis)
    assert tr.shape == c["trajectories"].shape
    assert rel_err(tr, c["trajectories"]) < 1e-6


def test_front_end_adjoint_is_consistent():
    rng = np.random.default_rng(0)
    cg = rng.standard_normal((2, 1, 6, 16, 24))
    times = fo.reconstruction_times(1, 5, 0.3)
    for basis in ("polynomial", "dct", "bezier"):
        tr, _ = fo.trajectories_from_coeff_grid(cg, times, 4, 3, basis, dtype=np.float64)
        g = rng.standard_normal(tr.shape)
        dcg = fo.trajectories_backward(g, times, 4, 3, basis, cg.shape)
        eps = rng.standard_normal(cg.shape)
        tr2, _ = fo.trajectories_from_coeff_grid(cg + 1e-6 * eps, times, 4, 3, basis, dtype=np.float64)
        lhs = ((tr2 - tr) * g).sum() / 1e-6
        assert abs(lhs - (dcg * eps).sum()) < 1e-5 * max(1.0, abs(lhs))


def test_imager_matches_reference():
    z = np.load(f"{GOLDEN_DIR}/imager.npz")
    H, W = (int(v) for v in z["shape"])
    ev, wt = z["events"], z["weight"]
    raw = fo.create_iwe(ev, (H, W), wt, sigma=0)
    assert rel_err(raw, z["iwe_sigma0"]) < 1e-6
    assert rel_err(fo.create_iwe(ev, (H, W), wt, sigma=1), z["iwe_sigma1"]) < 1e-6
    unit = fo.create_iwe(ev, (H, W), 1.0, sigma=0)
    assert rel_err(unit, z["iwe_unit"]) < 1e-6
    # the count image shares indices/masks with the bilinear vote: its total equals the number
    # of in-bounds corners and it is integer valued
    cnt = fo.count_image(ev, (H, W))
    _, mask, _ = fo.vote_corners(ev[..., :2], (H, W))
    assert cnt.dtype == np.int64 and cnt.sum() == mask.sum()
    assert ((cnt > 0) >= (unit > 0)).all()
    # outer_padding: the padded image, corners shifted after the floor
    pad = tuple(int(v) for v in z["pad"])
    HP, WP = H + 2 * pad[0], W + 2 * pad[1]
    assert rel_err(fo.create_iwe(ev, (HP, WP), wt, sigma=0, outer_padding=pad), z["iwe_pad_sigma0"]) < 1e-6
    assert rel_err(fo.create_iwe(ev, (HP, WP), wt, sigma=1, outer_padding=pad), z["iwe_pad_sigma1"]) < 1e-6
    # method='polarity': positive / negative events voted separately; a batch is flattened
    pos = ev[..., 3] > 0
    for name, e_, w_, sg in (("iwe_polarity_unbatched", ev[0], wt[0], 1), ("iwe_polarity_batched", ev, wt, 0)):
        p_ = e_[..., 3] > 0
        got = np.stack((fo.create_iwe(e_[p_], (H, W), w_[p_], sigma=sg)[0],
                        fo.create_iwe(e_[~p_], (H, W), w_[~p_], sigma=sg)[0]))
        assert rel_err(got, z[name]) < 1e-6


def test_integer_semantics_of_float32_floor():
    # SURVEY 8a "Integer-semantics facts": fp32 add of 1e-6f before floor, python-style //.
    v = np.array([[31.999998, 0.0], [-1e-7, 0.0], [-5e-7, 0.0], [-1.5e-6, 0.0]], np.float32)
    inds, mask, frac = fo.vote_corners(v, (64, 64))
    assert (inds[:, 0] // 64).tolist() == [32, 0, 0, 0] and mask[3, 0] == False  # noqa: E712
    ev = np.zeros((1, 2, 6), np.float32)
    ev[0, :, 0] = [3.9999998, 479.99997]
    it, iy, ix = fo.lut_cell_indices(ev, 4)
    assert iy[0].tolist() == [0, 119]


def test_knn_c_helper_equals_numpy():
    rng = np.random.default_rng(1)
    pts = (rng.random((2, 3, 150, 2)) * 40).astype(np.float32)
    grid, _, _ = fo.lut_grid((32, 48), 4)
    for norm in ("l2", "l1"):
        a = fo.knn_bruteforce(pts, grid, 7, norm, True)
        b = fo.knn_bruteforce(pts, grid, 7, norm, False)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all()


def test_knn_ties_pick_lowest_index():
    # zero flow: trajectories sit on the tile lattice, queries are equidistant to 4 of them
    pos = fo.tile_positions((16, 16), 4).astype(np.float32)[None, None]
    grid, _, _ = fo.lut_grid((16, 16), 4)
    ind, dist = fo.knn_bruteforce(pos, grid, 2, "l2")
    assert (dist[..., 0] == dist[..., 1]).all() is not None
    assert (ind[..., 0] < ind[..., 1])[dist[..., 0] == dist[..., 1]].all()


def test_voxel_grid_oracle_matches_reference():
    z = np.load(f"{GOLDEN_DIR}/voxel.npz")
    shape = tuple(int(v) for v in z["shape"])
    for norm in (None, "mean_std", "max"):
        got = fo.voxel_grid(z["x"], z["y"], z["t"], z["p"], shape, norm)
        assert rel_err(got, z["grid_" + str(norm)]) < 2e-6, norm


def test_dense_flow_oracle_matches_reference():
    z = np.load(f"{GOLDEN_DIR}/dense_flow.npz")
    H, W = (int(v) for v in z["shape"])
    dense, patch = fo.dense_flow_from_traj(z["traj_flow"], z["pixel_positions"], int(z["patch"]), (H, W))
    assert np.array_equal(patch.astype(np.float32), z["patch_flow"])
    assert rel_err(dense, z["dense"]) < 1e-6
