"""Generate golden vectors from the *real* reference loss (TEST INFRASTRUCTURE).

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python oracle/make_golden.py            # writes tests/golden/*.npz

The reference (upstream ``src/losses/focus.py``) is imported unmodified with two import
stubs from ``oracle/ref_stubs`` (pykeops -> dense torch KNN, pytorch_lightning -> empty shell;
both packages are absent from this image and there is no network).  For every case of the
variant matrix below the script stores the inputs (trajectories, times, events,
num_pos_events, config) and the reference's outputs: loss, focus_loss, smoothness_loss,
iwes, flow_lut, flow_to_next, ind_k and d loss / d trajectories (autograd), all float32 as
the reference computes them on CPU.  Front-end cases additionally store coeff_grid ->
trajectories as computed by the reference's own ``coeffs_grid_to_list`` + ``compute_basis``
(polynomial / dct) and ``BezierCurves._compute_flow_from_timestamps`` (bezier).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("CMAX_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "ref_stubs"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from src.losses import LossFactory            # noqa: E402  (the reference)
from src.utils import trajectories as ref_traj  # noqa: E402
from src.utils import basis as ref_basis       # noqa: E402

from motionpriorcmax_b200 import synthetic    # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

BASE = dict(image_shape=(32, 48), num_tref=1, num_bins=5, num_knn=8, smooth_weight=0.003,
            lut_superpixel_size=4, focus_loss_norm="l1", dist_norm="l2", scale_iwe_by_dt=True,
            mask_image_border=True, polarity_aware_batching=True, interpolation_scheme="mean",
            smooth_type="on_flow_to_tref")


def cfgv(**kw):
    c = dict(BASE)
    c.update(kw)
    return c


# name -> (loss config, generator options)
CASES = {
    "dsec_like_pab": (cfgv(), dict(B=2, M=700, K=1, basis="polynomial", patch=4)),
    "l2norm_iwd_l1dist": (cfgv(focus_loss_norm="l2", dist_norm="l1", interpolation_scheme="iwd",
                               smooth_weight=0.01),
                          dict(B=2, M=600, K=2, basis="polynomial", patch=4)),
    "multi_tref3": (cfgv(num_tref=3, scale_iwe_by_dt=False, polarity_aware_batching=False,
                         num_knn=4), dict(B=2, M=500, K=3, basis="polynomial", patch=4)),
    "evimo_like_next_bezier": (cfgv(image_shape=(24, 32), num_bins=7, smooth_weight=0.06,
                                    smooth_type="on_flow_to_next", num_knn=6),
                               dict(B=2, M=500, K=4, basis="bezier", patch=4, integer=True)),
    "knn1_nomask_noscale": (cfgv(num_knn=1, mask_image_border=False, scale_iwe_by_dt=False,
                                 smooth_weight=0.0), dict(B=1, M=400, K=1, basis="polynomial",
                                                          patch=4, big_flow=True)),
    "nopab_scale_s8": (cfgv(polarity_aware_batching=False, lut_superpixel_size=8, num_knn=5),
                       dict(B=3, M=[300, 500, 120], K=2, basis="dct", patch=4)),
    "free_points_b1": (cfgv(image_shape=(40, 56), num_knn=12, smooth_weight=0.02),
                       dict(B=1, M=800, K=1, basis="free", n_free=333, patch=4, big_flow=True)),
    "s2_patch2_iwd": (cfgv(image_shape=(16, 24), lut_superpixel_size=2, num_knn=9,
                           interpolation_scheme="iwd"),
                      dict(B=2, M=300, K=1, basis="polynomial", patch=2)),
    # the reference's other focus functional (src/utils/loss.py:14-16); FocusLoss.calc hard-codes
    # 'gradient_magnitude' (focus.py:90), so the call is redirected for this case only
    "variance_functional": (cfgv(), dict(B=2, M=700, K=1, basis="polynomial", patch=4, variance=True)),
    # six windows: the batch size from which the CUDA path chains its per-bin K-NN launches
    "chain_b6_k12": (cfgv(image_shape=(48, 64), num_bins=7, num_knn=12),
                     dict(B=6, M=[1500, 2200, 900, 1800, 2500, 1200], K=2, basis="polynomial", patch=4)),
}


def ref_trajectories(coeff_grid, times, patch, K, basis, anchor=0.0):
    """The reference's own coeff_grid -> trajectories (trajectory_net.py:101-119)."""
    H, W = coeff_grid.shape[-2:]
    mask = ref_traj.get_optical_flow_tile_mask((H, W), patch)
    if basis == "bezier":
        from src.models.raft_spline.curves import BezierCurves
        params = coeff_grid[:, 0]
        # Bezier params are (x, y)-major (curves/base.py:88-89); our coeff_grid stores y first,
        # so hand the reference the swapped halves.
        params_xy = torch.cat((params[:, K:], params[:, :K]), 1)
        curves = BezierCurves(params_xy)
        flow = curves._compute_flow_from_timestamps(times.double().numpy())       # [T,B,2(x,y),H,W]
        flow0 = curves._compute_flow_from_timestamps(np.array([anchor], np.float64))
        flow = (flow - flow0).flip(2)                                              # -> (y, x)
        pos = torch.nonzero(mask)
        tr = flow[:, :, :, pos[:, 0], pos[:, 1]].permute(1, 0, 3, 2)               # [B,T,n,2]
        return (tr + pos[None, None].float()).contiguous()
    coeffs, pos, _ = ref_traj.coeffs_grid_to_list(coeff_grid, mask, num_coeffs=K)
    a = ref_basis.compute_basis(coeffs, torch.full((1,), anchor, dtype=coeffs.dtype), K, basis)
    tr = ref_basis.compute_basis(coeffs, times, K, basis)
    tr = tr - a + pos[None, :, None, :]
    return tr.permute(0, 2, 1, 3).contiguous()


def build_case(name, cfg, opt, seed):
    torch.manual_seed(seed)
    H, W = cfg["image_shape"]
    B, K = opt["B"], opt["K"]
    L = LossFactory.get_loss_calculator("FOCUS", dict(cfg))
    if cfg["num_tref"] == 1:
        t_ref = torch.tensor([0.37 if seed % 2 else 0.81])
        edges = torch.linspace(0, 1, cfg["num_bins"] + 1)
        times = torch.cat((t_ref, (edges[:-1] + edges[1:]) / 2))
    else:
        times = L.get_reconstruction_times("cpu")
    extra = {}
    if opt["basis"] == "free":
        n = opt["n_free"]
        p0 = torch.rand(B, 1, n, 2) * torch.tensor([H, W]).float()
        v = torch.randn(B, 1, n, 2) * 6
        traj = (p0 + v * times[None, :, None, None]).contiguous()
    else:
        sig = 10.0 if opt.get("big_flow") else 4.0
        cg = synthetic.make_coeff_grid(B, K, H, W, sigma_px=sig, seed=seed, coarse=(4, 5))
        cg = cg + 0.05 * torch.randn_like(cg)
        traj = ref_trajectories(cg, times, opt["patch"], K, opt["basis"])
        extra = dict(coeff_grid=cg.numpy(), patch=opt["patch"], num_basis=K, basis=opt["basis"])
    traj = traj.detach().float().requires_grad_()
    ev, npos = synthetic.make_event_batch(B, opt["M"], H, W, cfg["num_bins"],
                                          cfg["polarity_aware_batching"], seed=seed,
                                          integer_coords=opt.get("integer", False),
                                          coord_scale=0.8)
    batch = {"events": ev}
    if npos is not None:
        batch["num_pos_events"] = npos

    # capture intermediates by wrapping interpolate_flow
    cap = {}
    orig = L.interpolate_flow

    def wrapped(a, b):
        lut, nxt = orig(a, b)
        cap["lut"], cap["next"] = lut, nxt
        return lut, nxt
    L.interpolate_flow = wrapped
    import src.losses.focus as ref_focus
    orig_focus = ref_focus.utils.calculate_focus_loss
    if opt.get("variance"):
        ref_focus.utils.calculate_focus_loss = lambda iw, loss_type, norm: orig_focus(iw, loss_type="variance", norm=norm)
    try:
        loss, log, misc = L.calc(traj, times, batch)
    finally:
        ref_focus.utils.calculate_focus_loss = orig_focus
    loss.backward()
    out = dict(
        cfg=json.dumps(dict(cfg, focus_loss_type="variance") if opt.get("variance") else cfg),
        trajectories=traj.detach().numpy(), times=times.numpy(),
        events=ev.numpy(), num_pos_events=np.int64(-1 if npos is None else npos),
        loss=loss.detach().numpy(), focus_loss=log["focus_loss"].numpy(),
        smoothness_loss=log["smoothness_loss"].numpy(), iwes=misc["iwes"].numpy(),
        flow_lut=cap["lut"].detach().numpy(), dtraj=traj.grad.numpy(), **extra)
    if cap["next"] is not None:
        out["flow_to_next"] = cap["next"].detach().numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: loss={float(loss):.6f} focus={float(log['focus_loss']):.6f} "
          f"smooth={float(log['smoothness_loss']):.6f} |dtraj|={float(traj.grad.abs().sum()):.5f}")


def build_imager_case():
    """create_iwe / count image on events that hit every border case
    (event_image_converter.py:226-272, 333-391)."""
    from src.utils import EventImageConverter
    torch.manual_seed(7)
    H, W = 20, 28
    PAD = (2, 3)
    im, imp = EventImageConverter((H, W)), EventImageConverter((H, W), outer_padding=PAD)
    n = 900
    yx = torch.rand(2, n, 2) * torch.tensor([H + 6.0, W + 6.0]) - 3.0
    # exact integers, values a hair below integers, tiny negatives, the far corner
    special = torch.tensor([[0.0, 0.0], [H - 1.0, W - 1.0], [H * 1.0, W * 1.0], [-1e-7, 5.0],
                            [-5e-7, 5.0], [-1.5e-6, 5.0], [3.9999998, 7.9999995], [-1.0, -1.0],
                            [H - 1e-6, W - 1e-6], [11.999999, 0.5], [5.0, -0.9999999]])
    yx[0, :len(special)] = special
    ev = torch.cat((yx, torch.rand(2, n, 1), (torch.rand(2, n, 1) < 0.5).float()), -1)
    wt = torch.rand(2, n)
    np.savez_compressed(
        os.path.join(OUT, "imager.npz"), events=ev.numpy(), weight=wt.numpy(), shape=np.array([H, W]),
        iwe_sigma1=im.create_iwe(ev, method="bilinear_vote", sigma=1, weight=wt).numpy(),
        iwe_sigma0=im.create_iwe(ev, method="bilinear_vote", sigma=0, weight=wt).numpy(),
        iwe_unit=im.create_iwe(ev, method="bilinear_vote", sigma=0).numpy(),
        # outer_padding (:23-28, 339-343): votes land in an image enlarged by 2 * pad
        pad=np.array(PAD), iwe_pad_sigma0=imp.create_iwe(ev, method="bilinear_vote", sigma=0, weight=wt).numpy(),
        iwe_pad_sigma1=imp.create_iwe(ev, method="bilinear_vote", sigma=1, weight=wt).numpy(),
        # method='polarity' (:156-163): un-batched and batched (boolean indexing flattens the batch)
        iwe_polarity_unbatched=im.create_iwe(ev[0], method="polarity", sigma=1, weight=wt[0]).numpy(),
        iwe_polarity_batched=im.create_iwe(ev, method="polarity", sigma=0, weight=wt).numpy(),
        iwe_polarity_batched_unit=im.create_iwe(ev, method="polarity", sigma=1).numpy())
    # NOTE: the reference's count_event_tensor (event_image_converter.py:226-272) cannot be
    # executed: it scatter_adds int64 votes into a float image and torch raises
    # "Expected self.dtype to be equal to src.dtype".  The count image is therefore pinned
    # only through the indices/masks it shares with bilinear_vote_tensor (iwe_unit above).
    print("imager: ok")


def build_voxel_case():
    """Voxel grid of the reference loader (src/loader/dsec/utils.py:19-77), imported by file path
    because the loader package pulls in h5py / hdf5plugin (absent here)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_dsec_utils", os.path.join(REF, "src/loader/dsec/utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(11)
    C, H, W, n = 5, 24, 36, 6000
    x = torch.rand(n) * (W + 3) - 1.5            # includes (-1, 0): int() truncates toward zero
    y = torch.rand(n) * (H + 3) - 1.5
    t = torch.rand(n).sort().values
    t = (t - t[0]) / (t[-1] - t[0])
    p = (torch.rand(n) < 0.5).float()
    out = dict(x=x.numpy(), y=y.numpy(), t=t.numpy(), p=p.numpy(), shape=np.array([C, H, W]))
    for norm in (None, "mean_std", "max"):
        vg = mod.VoxelGrid((C, H, W), norm, 0)
        out["grid_" + str(norm)] = vg.convert({"p": p, "t": t, "x": x, "y": y}).numpy()
    vg = mod.VoxelGrid((C, H, W), "mean_std", 0.05)
    out["grid_mean_std_q05"] = vg.convert({"p": p, "t": t, "x": x, "y": y}).numpy()
    np.savez_compressed(os.path.join(OUT, "voxel.npz"), **out)
    print("voxel: ok")


def build_dense_flow_case():
    """Dense flow read-out of the reference (src/utils/flow.py:12-16)."""
    from src.utils import flow as ref_flow
    torch.manual_seed(5)
    H, W, patch = 24, 32, 4
    mask = ref_traj.get_optical_flow_tile_mask((H, W), patch)
    pos = torch.nonzero(mask)
    tf = torch.randn(2, len(pos), 2) * 5
    dense, patch_flow = ref_flow.dense_flow_from_traj(tf, pos, patch, (H, W))
    np.savez_compressed(os.path.join(OUT, "dense_flow.npz"), traj_flow=tf.numpy(), pixel_positions=pos.numpy(),
                        patch=np.int64(patch), shape=np.array([H, W]), dense=dense.numpy(),
                        patch_flow=patch_flow.numpy())
    print("dense_flow: ok")


def main():
    os.makedirs(OUT, exist_ok=True)
    for i, (name, (cfg, opt)) in enumerate(CASES.items()):
        build_case(name, cfg, opt, seed=100 + i)
    build_imager_case()
    build_voxel_case()
    build_dense_flow_case()


if __name__ == "__main__":
    main()
