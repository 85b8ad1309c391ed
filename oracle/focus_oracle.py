"""CPU oracle for the MotionPriorCMax contrast-maximisation loss path.

*** TEST INFRASTRUCTURE - NOT PRODUCT CODE. ***
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``motionpriorcmax_b200``) never does; it fails loudly when the CUDA library is absent.

This is a from-scratch numpy restatement of the reference algorithm with *closed-form*
gradients (no autograd), so it checks both the forward arithmetic and the hand-derived
backward that the CUDA kernels implement.  Each function cites the reference lines
(relative to the upstream repository root) it follows.

Parity pinning: the reference ships no tests / golden vectors.  The
oracle is pinned against the *reference itself executed in the build container* with two
import stubs (``oracle/ref_stubs``; pykeops is absent) - see ``oracle/make_golden.py`` and
``tests/test_oracle_vs_golden.py``.  What stays unpinned: KeOps' tie-breaking among exactly
equidistant neighbours (we use "lowest trajectory index wins") and its fp32 rounding of the
squared distance (we use ``fl(fl(dy*dy) + fl(dx*dx))``, the torch form).

dtype: ``np.float32`` mirrors the reference arithmetic op for op (elementwise ops are
bit-identical to torch CPU, reductions differ in summation order only); ``np.float64``
evaluates the same formulas in double (integer decisions - floor, LUT cell, KNN sets - are
still taken from the float32 inputs exactly as the reference takes them) and serves as the
"truth" for 1e-5 relative checks of float outputs.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

EPS_IWD = 1e-9          # focus.py:7
CHARBONNIER_EPS = 1e-3  # loss.py:46

_HERE = os.path.dirname(os.path.abspath(__file__))


# ----------------------------------------------------------------------------------------
# optional C helper (brute-force KNN, OpenMP) - makes the 480x640 case finish in seconds
# ----------------------------------------------------------------------------------------
_LIB = None


def build_c_helper(force: bool = False) -> Optional[str]:
    """Compile oracle/knn_bruteforce.c -> oracle/_build/liboracle_knn.so (gcc -O2 -fopenmp)."""
    src = os.path.join(_HERE, "knn_bruteforce.c")
    out_dir = os.path.join(_HERE, "_build")
    out = os.path.join(out_dir, "liboracle_knn.so")
    if os.path.exists(out) and not force and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    os.makedirs(out_dir, exist_ok=True)
    cmd = ["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-ffp-contract=off", "-o", out, src]
    subprocess.run(cmd, check=True)
    return out


def _c_lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle_knn.so")
        if not os.path.exists(path):
            try:
                build_c_helper()
            except Exception:  # pragma: no cover - gcc missing
                _LIB = False
                return None
        lib = ctypes.CDLL(path)
        lib.oracle_knn_bruteforce.restype = ctypes.c_int
        lib.oracle_knn_bruteforce.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
            ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _LIB = lib
    return _LIB or None


# ----------------------------------------------------------------------------------------
# times, lattice, front end
# ----------------------------------------------------------------------------------------
def reconstruction_times(num_tref: int, num_bins: int, t_ref: Optional[float] = None,
                         rng: Optional[np.random.Generator] = None) -> np.ndarray:
    """focus.py:53-64.  ``t_ref`` replaces the reference's ``torch.rand(1)`` draw."""
    if num_tref > 1:
        tr = np.linspace(0, 1, num_tref, dtype=np.float32)
    elif num_tref == 1:
        if t_ref is None:
            t_ref = float((rng or np.random.default_rng()).random())
        tr = np.array([t_ref], dtype=np.float32)
    else:
        raise ValueError("Invalid value for num_tref. Must be >= 1.")
    edges = np.linspace(0, 1, num_bins + 1, dtype=np.float32)
    mid = (edges[:-1] + edges[1:]) / np.float32(2)
    return np.concatenate((tr, mid)).astype(np.float32)


def lut_grid(image_shape, s: int):
    """focus.py:116-126: LUT query points at super-pixel centres, row-major (y, x)."""
    H, W = image_shape
    off = np.float32(float(s) / 2 - 0.5)
    y = np.arange(0, H, s, dtype=np.float32) + off
    x = np.arange(0, W, s, dtype=np.float32) + off
    gy, gx = np.meshgrid(y, x, indexing="ij")
    return np.stack((gy, gx), -1).reshape(-1, 2), len(y), len(x)


def tile_positions(image_shape, patch: int) -> np.ndarray:
    """trajectories.py:3-13,46: pixels mask[o::p, o::p], o = p // 2, row-major (y, x) int64."""
    H, W = image_shape
    o = patch // 2
    ys = np.arange(o, H, patch)
    xs = np.arange(o, W, patch)
    gy, gx = np.meshgrid(ys, xs, indexing="ij")
    return np.stack((gy, gx), -1).reshape(-1, 2).astype(np.int64)


def basis_matrix(times: np.ndarray, num_basis: int, basis_type: str, dtype=np.float32):
    """Basis values phi_k(t), k = 1..K -> [n_t, K].

    polynomial t**k (basis.py:29-31); dct sqrt(2) cos(pi/2 (2t+1) k) (basis.py:18-24);
    bezier C(d,k)(1-t)^(d-k) t^k without the P0 term, coefficients in float64 then cast
    (curves/bezier.py:68-113).
    """
    t = np.asarray(times, dtype=dtype).reshape(-1)
    k = np.arange(1, num_basis + 1)
    if basis_type == "polynomial":
        return (t[:, None] ** k[None, :].astype(dtype)).astype(dtype)
    if basis_type == "dct":
        a = (2 * t[:, None] + 1) * k[None, :].astype(dtype)
        return (np.sqrt(2.0) * np.cos(np.pi / 2.0 * a)).astype(dtype)
    if basis_type == "bezier":
        from math import comb
        d = num_basis
        t64 = t.astype(np.float64)
        out = np.stack([comb(d, i) * (1 - t64) ** (d - i) * t64 ** i for i in range(1, d + 1)], -1)
        return out.astype(dtype)
    raise ValueError(basis_type)


def trajectories_from_coeff_grid(coeff_grid, times, patch, num_basis, basis_type="polynomial",
                                 anchor=0.0, add_offsets=True, xy_order=False, dtype=np.float32):
    """trajectory_net.py:101-119 + trajectories.py:15-52 + basis.py:4-46.

    coeff_grid [B, S, 2K, H, W] (channels [0:K] = y, [K:2K] = x; ``xy_order`` swaps the two
    halves - the RAFT-spline (x, y) convention of raft_spline/utils.py:22-28).
    Returns trajectories [B, n_t, n, 2] (y, x) and the tile positions [n, 2].
    """
    cg = np.asarray(coeff_grid, dtype=dtype)
    if cg.ndim == 4:
        cg = cg[:, None]
    B, S, C2, H, W = cg.shape
    K = num_basis
    assert C2 == 2 * K
    pos = tile_positions((H, W), patch)
    c = cg[:, :, :, pos[:, 0], pos[:, 1]]                  # [B, S, 2K, n]
    c = c.reshape(B, S, 2, K, -1).sum(1)                     # sum over scales -> [B, 2, K, n]
    if xy_order:
        c = c[:, ::-1]
    phi = basis_matrix(times, K, basis_type, dtype) - basis_matrix([anchor], K, basis_type, dtype)
    traj = np.einsum("tk,bckn->btnc", phi, c).astype(dtype)
    if add_offsets:
        traj = traj + pos[None, None].astype(dtype)
    return traj.astype(dtype), pos


def trajectories_backward(dtraj, times, patch, num_basis, basis_type, grid_shape, anchor=0.0,
                          xy_order=False, dtype=np.float64):
    """Adjoint of :func:`trajectories_from_coeff_grid`: dense d coeff_grid [B, S, 2K, H, W],
    non-zero only at the tile pixels."""
    B, S, C2, H, W = grid_shape
    K = num_basis
    pos = tile_positions((H, W), patch)
    phi = basis_matrix(times, K, basis_type, dtype) - basis_matrix([anchor], K, basis_type, dtype)
    dc = np.einsum("tk,btnc->bckn", phi, np.asarray(dtraj, dtype))      # [B, 2, K, n]
    if xy_order:
        dc = dc[:, ::-1]
    out = np.zeros((B, S, 2 * K, H, W), dtype)
    out[:, :, :, pos[:, 0], pos[:, 1]] = dc.reshape(B, 1, 2 * K, -1)
    return out


# ----------------------------------------------------------------------------------------
# KNN (focus.py:128-137)
# ----------------------------------------------------------------------------------------
def knn_bruteforce(points: np.ndarray, grid: np.ndarray, K: int, norm: str = "l2",
                   use_c: bool = True):
    """K nearest trajectories of every LUT query, exhaustive search in float32.

    points [B, nb, n, 2] f32, grid [q, 2] f32 -> (ind int64 [B, nb, q, K] ascending distance,
    dist f32 [B, nb, q, K]).  d = fl(fl(dy*dy) + fl(dx*dx)) (l2) or fl(|dy| + |dx|) (l1);
    on equal distance the lowest trajectory index comes first.
    """
    points = np.ascontiguousarray(points, dtype=np.float32)
    grid = np.ascontiguousarray(grid, dtype=np.float32)
    B, nb, n, _ = points.shape
    q = grid.shape[0]
    assert K <= n, "num_knn larger than the number of trajectories"
    lib = _c_lib() if use_c else None
    if lib is not None:
        ind = np.empty((B, nb, q, K), np.int64)
        dist = np.empty((B, nb, q, K), np.float32)
        rc = lib.oracle_knn_bruteforce(points.ctypes.data, grid.ctypes.data, B * nb, n, q, K,
                                       1 if norm == "l1" else 0, ind.ctypes.data,
                                       dist.ctypes.data,
                                       int(os.environ.get("ORACLE_KNN_THREADS", os.cpu_count() or 1)))
        assert rc == 0
        return ind, dist
    ind = np.empty((B, nb, q, K), np.int64)
    dist = np.empty((B, nb, q, K), np.float32)
    chunk = max(1, (1 << 23) // max(1, n))
    for b in range(B):
        for t in range(nb):
            p = points[b, t]
            for s in range(0, q, chunk):
                g = grid[s:s + chunk]
                dy = g[:, None, 0] - p[None, :, 0]
                dx = g[:, None, 1] - p[None, :, 1]
                if norm == "l2":
                    d = dy * dy + dx * dx          # two separately rounded products, one add
                else:
                    d = np.abs(dy) + np.abs(dx)
                order = np.argsort(d, axis=1, kind="stable")[:, :K]
                ind[b, t, s:s + chunk] = order
                dist[b, t, s:s + chunk] = np.take_along_axis(d, order, 1)
    return ind, dist


# ----------------------------------------------------------------------------------------
# dense stencils and their adjoints (loss.py:58-87, event_image_converter.py:170-175)
# ----------------------------------------------------------------------------------------
SOBEL_X = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], np.float64)
SOBEL_Y = SOBEL_X.T.copy()


def gaussian_kernel1d(sigma: float = 1.0, ksize: int = 3, dtype=np.float32) -> np.ndarray:
    """torchvision ``_get_gaussian_kernel1d``: pdf on linspace(-h, h, k), normalised (f32)."""
    half = (ksize - 1) * 0.5
    x = np.linspace(-half, half, ksize, dtype=np.float32)
    pdf = np.exp(np.float32(-0.5) * (x / np.float32(sigma)) ** 2).astype(np.float32)
    return (pdf / pdf.sum(dtype=np.float32)).astype(dtype)


def _corr3(padded: np.ndarray, k: np.ndarray) -> np.ndarray:
    """3x3 cross-correlation over the last two axes of an already padded array."""
    Hh, Ww = padded.shape[-2] - 2, padded.shape[-1] - 2
    out = np.zeros(padded.shape[:-2] + (Hh, Ww), padded.dtype)
    for a in range(3):
        for b in range(3):
            if k[a, b] != 0:
                out += padded.dtype.type(k[a, b]) * padded[..., a:a + Hh, b:b + Ww]
    return out


def _corr3_T(g: np.ndarray, k: np.ndarray) -> np.ndarray:
    """Adjoint of :func:`_corr3` -> gradient w.r.t. the padded array."""
    Hh, Ww = g.shape[-2:]
    out = np.zeros(g.shape[:-2] + (Hh + 2, Ww + 2), g.dtype)
    for a in range(3):
        for b in range(3):
            if k[a, b] != 0:
                out[..., a:a + Hh, b:b + Ww] += g.dtype.type(k[a, b]) * g
    return out


def sobel(img: np.ndarray):
    """loss.py:58-87: depthwise Sobel-x / Sobel-y, zero padding 1."""
    p = np.pad(img, [(0, 0)] * (img.ndim - 2) + [(1, 1), (1, 1)])
    return _corr3(p, SOBEL_X), _corr3(p, SOBEL_Y)


def sobel_T(gx: np.ndarray, gy: np.ndarray) -> np.ndarray:
    return (_corr3_T(gx, SOBEL_X) + _corr3_T(gy, SOBEL_Y))[..., 1:-1, 1:-1]


def blur(img: np.ndarray, sigma: float = 1.0) -> np.ndarray:
    """event_image_converter.py:170-175 -> torchvision gaussian_blur(kernel_size=3, sigma):
    reflect padding 1, one 3x3 correlation with outer(g, g)."""
    g = gaussian_kernel1d(sigma, 3, img.dtype)
    k2 = np.outer(g, g).astype(img.dtype)
    p = np.pad(img, [(0, 0)] * (img.ndim - 2) + [(1, 1), (1, 1)], mode="reflect")
    return _corr3(p, k2)


def blur_T(gout: np.ndarray, sigma: float = 1.0) -> np.ndarray:
    g = gaussian_kernel1d(sigma, 3, gout.dtype)
    k2 = np.outer(g, g).astype(gout.dtype)
    gp = _corr3_T(gout, k2)
    # adjoint of reflect padding (pad row 0 mirrors image row 1, pad row H+1 mirrors row H-2)
    gi = gp[..., 1:-1, :].copy()
    gi[..., 1, :] += gp[..., 0, :]
    gi[..., -2, :] += gp[..., -1, :]
    gj = gi[..., :, 1:-1].copy()
    gj[..., :, 1] += gi[..., :, 0]
    gj[..., :, -2] += gi[..., :, -1]
    return gj


# ----------------------------------------------------------------------------------------
# event-stage integer work (shared by the float path and the bit-exact tests)
# ----------------------------------------------------------------------------------------
def lut_cell_indices(events: np.ndarray, s: int):
    """focus.py:185-187: it = trunc(bin), iy = int(y // s), ix = int(x // s) on float32
    (``//`` is Python-style floor division, numpy's floor_divide is the same algorithm)."""
    ev = np.asarray(events, np.float32)
    it = ev[..., 4].astype(np.int64)
    iy = np.floor_divide(ev[..., 0], np.float32(s)).astype(np.int64)
    ix = np.floor_divide(ev[..., 1], np.float32(s)).astype(np.int64)
    return it, iy, ix


def vote_corners(yx: np.ndarray, image_shape, outer_padding=(0, 0)):
    """event_image_converter.py:354-380: f = floor(yx + 1e-6) in *float32*, corner linear
    indices (OOB -> 0) and in-bounds masks for the 4 corners in reference order
    (y1,x1), (y1+1,x1), (y1,x1+1), (y1+1,x1+1).  Returns (inds int64 [..., 4],
    mask bool [..., 4], frac [..., 2] in yx's dtype).  `image_shape` is the (padded) image the
    votes land in; `outer_padding` (ph, pw) shifts the integer corner after the floor (:339-343)."""
    H, W = image_shape
    ph, pw = outer_padding
    yx32 = np.asarray(yx, np.float32)
    with np.errstate(invalid="ignore"):
        fl = np.floor(yx32 + np.float32(1e-6))
    frac = np.asarray(yx) - fl.astype(np.asarray(yx).dtype)
    fl = np.where(np.isfinite(fl), fl, -(2.0 ** 40)).astype(np.int64)
    y1, x1 = fl[..., 0] + ph, fl[..., 1] + pw
    inds = np.stack((x1 + y1 * W, x1 + (y1 + 1) * W, (x1 + 1) + y1 * W, (x1 + 1) + (y1 + 1) * W), -1)
    okx0, okx1 = (0 <= x1) & (x1 < W), (0 <= x1 + 1) & (x1 + 1 < W)
    oky0, oky1 = (0 <= y1) & (y1 < H), (0 <= y1 + 1) & (y1 + 1 < H)
    mask = np.stack((okx0 & oky0, okx0 & oky1, okx1 & oky0, okx1 & oky1), -1)
    return inds * mask, mask, frac


def count_image(events_yx: np.ndarray, image_shape, outer_padding=(0, 0)) -> np.ndarray:
    """event_image_converter.py:226-272 (count_event_tensor): 4 unit votes per event -> int64."""
    H, W = image_shape
    ev = np.asarray(events_yx)
    if ev.ndim == 2:
        ev = ev[None]
    inds, mask, _ = vote_corners(ev[..., :2], image_shape, outer_padding)
    out = np.zeros((ev.shape[0], H * W), np.int64)
    for b in range(ev.shape[0]):
        out[b] = np.bincount(inds[b].reshape(-1), weights=mask[b].reshape(-1).astype(np.float64),
                             minlength=H * W).astype(np.int64)
    return out.reshape(ev.shape[0], H, W)


def bilinear_vote(events_yx: np.ndarray, weight, image_shape, dtype=np.float32, outer_padding=(0, 0)) -> np.ndarray:
    """event_image_converter.py:333-391 (bilinear_vote_tensor) -> raw IWE [nb, H, W]."""
    H, W = image_shape
    ev = np.asarray(events_yx)
    if ev.ndim == 2:
        ev = ev[None]
    nb = ev.shape[0]
    yx = ev[..., :2].astype(dtype)
    inds, mask, frac = vote_corners(yx, image_shape, outer_padding)
    fy, fx = frac[..., 0], frac[..., 1]
    w = np.broadcast_to(np.asarray(weight, dtype), fy.shape)
    one = dtype(1)
    vals = np.stack(((one - fy) * (one - fx) * w, fy * (one - fx) * w,
                     (one - fy) * fx * w, fy * fx * w), -1) * mask
    out = np.zeros((nb, H * W), dtype)
    for b in range(nb):
        out[b] = np.bincount(inds[b].reshape(-1), weights=vals[b].reshape(-1).astype(np.float64),
                             minlength=H * W).astype(dtype)
    return out.reshape(nb, H, W)


def create_iwe(events, image_shape, weight=1.0, sigma=1, dtype=np.float32, outer_padding=(0, 0)) -> np.ndarray:
    """event_image_converter.py:45-74 with method='bilinear_vote' -> [nb, H, W] (H, W = padded sizes)."""
    img = bilinear_vote(events, weight, image_shape, dtype, outer_padding)
    return blur(img, sigma) if sigma > 0 else img


def _cubic_aa(x, a=-0.5):
    x = np.abs(x)
    return np.where(x < 1, ((a + 2) * x - (a + 3)) * x * x + 1, np.where(x < 2, a * (((x - 5) * x + 8) * x - 4), 0.0))


def resize_bicubic_antialias(img, out_h, out_w):
    """torchvision resize(BICUBIC, antialias=True) = ATen's separable anti-aliased bicubic
    (last dimension first); img [..., h, w] -> [..., out_h, out_w]."""
    def weights(n_in, n_out):
        scale = n_in / n_out
        support = 2.0 * scale if scale >= 1 else 2.0
        inv = 1.0 / scale if scale >= 1 else 1.0
        out = []
        for i in range(n_out):
            c = scale * (i + 0.5)
            lo = max(int(c - support + 0.5), 0)
            cnt = min(int(c + support + 0.5), n_in) - lo
            w = _cubic_aa((np.arange(cnt) + lo - c + 0.5) * inv)
            out.append((lo, w / w.sum()))
        return out
    img = np.asarray(img, np.float64)
    wx, wy = weights(img.shape[-1], out_w), weights(img.shape[-2], out_h)
    tmp = np.stack([(img[..., :, lo:lo + len(w)] * w).sum(-1) for lo, w in wx], -1)
    return np.stack([(tmp[..., lo:lo + len(w), :] * w[:, None]).sum(-2) for lo, w in wy], -2)


def dense_flow_from_traj(traj_flow, pixel_positions, patch, image_shape):
    """src/utils/flow.py:12-16 + list_to_grid (src/utils/trajectories.py:54-75)."""
    h, w = image_shape
    tf = np.asarray(traj_flow, np.float64)
    b, n, c = tf.shape
    grid = np.zeros((b, c, h // patch, w // patch))
    pos = np.asarray(pixel_positions) // patch
    grid[:, :, pos[:, 0], pos[:, 1]] = tf.transpose(0, 2, 1)
    return resize_bicubic_antialias(grid, h, w), grid


def voxel_grid(x, y, t, p, shape, norm_type=None):
    """src/loader/dsec/utils.py:29-77 (VoxelGrid.convert, quantile == 0): trilinear vote of 2p-1,
    int() truncation toward zero, float32 arithmetic, then 'mean_std' / 'max' normalisation."""
    C, H, W = shape
    x, y, t, p = (np.asarray(v, np.float32) for v in (x, y, t, p))
    tn = (np.float32(C - 1) * (t - t[0]) / (t[-1] - t[0])).astype(np.float32)
    x0, y0, t0 = np.trunc(x).astype(np.int64), np.trunc(y).astype(np.int64), np.trunc(tn).astype(np.int64)
    value = np.float32(2) * p - np.float32(1)
    grid = np.zeros(C * H * W, np.float64)
    for xl in (x0, x0 + 1):
        for yl in (y0, y0 + 1):
            for tl in (t0, t0 + 1):
                m = (xl < W) & (xl >= 0) & (yl < H) & (yl >= 0) & (tl >= 0) & (tl < C)
                w = value * (1 - np.abs(xl.astype(np.float32) - x)) * (1 - np.abs(yl.astype(np.float32) - y)) \
                    * (1 - np.abs(tl.astype(np.float32) - tn))
                np.add.at(grid, (H * W * tl + W * yl + xl)[m], w[m].astype(np.float64))
    grid = grid.astype(np.float32).reshape(C, H, W)
    nz = grid != 0
    if norm_type == "mean_std" and nz.any():
        mean = grid[nz].mean(dtype=np.float64)
        std = grid[nz].std(ddof=1, dtype=np.float64) if nz.sum() > 1 else np.nan
        grid[nz] = (grid[nz] - np.float32(mean)) / np.float32(std) if std > 0 else grid[nz] - np.float32(mean)
    elif norm_type == "max" and np.abs(grid).max() > 0:
        grid = grid / np.abs(grid).max()
    return grid


# ----------------------------------------------------------------------------------------
# the loss
# ----------------------------------------------------------------------------------------
@dataclass
class FocusOracle:
    """Restatement of ``FocusLoss`` (focus.py:9-246) with an explicit backward."""
    image_shape: tuple
    num_tref: int
    num_bins: int
    num_knn: int
    smooth_weight: float
    lut_superpixel_size: int
    focus_loss_norm: str
    dist_norm: str
    scale_iwe_by_dt: bool
    mask_image_border: bool
    polarity_aware_batching: bool
    interpolation_scheme: str
    smooth_type: str
    dtype: type = np.float32
    use_c_knn: bool = True
    focus_loss_type: str = "gradient_magnitude"      # upstream calc hard-codes it (focus.py:90)
    ctx: dict = field(default_factory=dict, repr=False)

    def __post_init__(self):
        # focus.py:49-51
        assert not self.scale_iwe_by_dt or self.num_tref == 1
        assert not self.polarity_aware_batching or self.num_tref == 1
        assert not self.smooth_type == "on_flow_to_next" or self.num_tref == 1

    # -- forward ---------------------------------------------------------------
    def forward(self, trajectories, times, events, num_pos_events: int = -1, ind_k=None):
        dt_ = self.dtype
        H, W = self.image_shape
        s, K, R = self.lut_superpixel_size, self.num_knn, self.num_tref
        traj32 = np.asarray(trajectories, np.float32)        # integer decisions (KNN) use float32
        traj = np.asarray(trajectories).astype(dt_)          # float math keeps the caller's precision
        times32 = np.asarray(times, np.float32)
        ev32 = np.asarray(events, np.float32)
        B, n_t, n, _ = traj.shape
        nb = n_t - R
        assert nb == self.num_bins
        assert not self.polarity_aware_batching or num_pos_events > -1        # focus.py:80
        t_ref = traj[:, :R]                      # [B, R, n, 2]
        t_mid = traj[:, R:]                      # [B, nb, n, 2]

        # ---- interpolate_flow (focus.py:115-180)
        grid, Hq, Wq = lut_grid((H, W), s)
        q = grid.shape[0]
        if ind_k is None:
            ind_k, dist_k = knn_bruteforce(traj32[:, R:], grid, K, self.dist_norm, self.use_c_knn)
        else:                                    # caller-supplied neighbour sets (tests)
            bi_ = np.arange(B)[:, None, None, None]
            ti_ = np.arange(nb)[None, :, None, None]
            pk = traj32[:, R:][bi_, ti_, ind_k]                       # [B, nb, q, K, 2] f32
            dyk = grid[None, None, :, None, 0] - pk[..., 0]
            dxk = grid[None, None, :, None, 1] - pk[..., 1]
            dist_k = dyk * dyk + dxk * dxk if self.dist_norm == "l2" else np.abs(dyk) + np.abs(dxk)
        if self.interpolation_scheme == "mean" or K == 1:
            wk = np.full((B, nb, q, K), 1.0 / K, dt_)
            mean_div = True
        elif self.interpolation_scheme == "iwd":
            if dist_k is None:
                raise ValueError("iwd needs distances")
            wk = 1 / (dist_k.astype(dt_) + dt_(EPS_IWD))
            wk = wk / wk.sum(-1, keepdims=True)
            mean_div = False
        else:
            raise ValueError
        bi = np.arange(B)[:, None, None, None]
        ti = np.arange(nb)[None, :, None, None]
        mid_k = t_mid[bi, ti, ind_k]                                   # [B, nb, q, K, 2]
        flow_lut = np.empty((B, nb, q, R, 2), dt_)
        for r in range(R):
            ref_k = t_ref[:, r][bi, ind_k]                             # [B, nb, q, K, 2]
            fk = ref_k - mid_k
            if mean_div:
                flow_lut[:, :, :, r] = fk.sum(3) / dt_(K)              # torch.mean
            else:
                flow_lut[:, :, :, r] = (wk[..., None] * fk).sum(3)
        flow_to_next = None
        if self.smooth_weight > 0 and self.smooth_type == "on_flow_to_next":
            nxt = t_mid[:, 1:] - t_mid[:, :-1]                          # [B, nb-1, n, 2]
            nk = nxt[bi, ti[:, :-1], ind_k[:, :-1]]                     # [B, nb-1, q, K, 2]
            flow_to_next = nk.sum(3) / dt_(K)
        lut6 = flow_lut.reshape(B, nb, Hq, Wq, R, 2)

        # ---- warp_events (focus.py:182-195)
        M = ev32.shape[1]
        it, iy, ix = lut_cell_indices(ev32, s)
        bidx = np.arange(B)[:, None]
        diff = lut6[bidx, it, iy, ix]                                   # [B, M, R, 2]
        warped_yx = diff.transpose(0, 2, 1, 3) + ev32[:, None, :, :2].astype(dt_)   # [B, R, M, 2]

        # ---- make_iwes weights (focus.py:201-214), no gradient
        tcol = ev32[:, None, :, 2].astype(dt_)
        weights = np.broadcast_to(ev32[:, None, :, 5].astype(dt_), (B, R, M)).copy()
        if self.scale_iwe_by_dt:
            dtm = np.clip(np.abs(tcol - times32[:R].astype(dt_)[None, :, None]), 0, 1)
            weights = (1 - dtm) * weights
        if self.mask_image_border:
            wy, wx = warped_yx[..., 0], warped_yx[..., 1]
            bm = np.ones_like(weights)
            bm[wy > H] = 0
            bm[wx > W] = 0
            bm[wy < 0] = 0
            bm[wx < 0] = 0
            weights = bm * weights
        wy_flat = warped_yx.reshape(B * R, M, 2)
        w_flat = weights.reshape(B * R, M)

        # ---- bilinear vote + blur (event_image_converter.py:333-391, 170-175)
        if self.polarity_aware_batching:
            P = 2
            npos = int(num_pos_events)
            raw = np.stack((bilinear_vote(wy_flat[:, :npos], w_flat[:, :npos], (H, W), dt_),
                            bilinear_vote(wy_flat[:, npos:], w_flat[:, npos:], (H, W), dt_)), 1)
        else:
            P = 1
            raw = bilinear_vote(wy_flat, w_flat, (H, W), dt_)[:, None]
        iwes = blur(raw, 1.0)                                            # [B*R, P, H, W]

        # ---- focus (loss.py:4-27)
        dx, dy = sobel(iwes)
        if self.focus_loss_type == "variance":               # loss.py:14-16: mean of unbiased var
            val = np.mean(np.var(iwes.astype(np.float64), axis=(-2, -1), ddof=1))
        elif self.focus_loss_norm == "l2":
            val = np.mean(dx * dx + dy * dy, dtype=np.float64)
        elif self.focus_loss_norm == "l1":
            val = np.mean(np.abs(dx) + np.abs(dy), dtype=np.float64)
        else:
            raise ValueError
        val = dt_(val)
        focus = dt_(1) / val

        # ---- smoothness (focus.py:232-246, loss.py:29-56)
        smooth = dt_(0)
        sm_field = None
        if self.smooth_weight != 0:
            if self.smooth_type == "on_flow_to_tref":
                sm_field = lut6.transpose(0, 1, 4, 5, 2, 3).reshape(-1, 2, Hq, Wq)
            elif self.smooth_type == "on_flow_to_next":
                sm_field = flow_to_next.reshape(B, nb - 1, Hq, Wq, 1, 2) \
                    .transpose(0, 1, 4, 5, 2, 3).reshape(-1, 2, Hq, Wq)
            else:
                raise ValueError
            sdx, sdy = sobel(sm_field)
            e2 = dt_(CHARBONNIER_EPS) ** 2
            cx = np.sqrt(sdx * sdx + e2)
            cy = np.sqrt(sdy * sdy + e2)
            smooth = dt_(self.smooth_weight) * dt_((np.mean(cx, dtype=np.float64)
                                                    + np.mean(cy, dtype=np.float64)) / 2.0)
        loss = focus + smooth

        self.ctx = dict(B=B, R=R, P=P, M=M, n=n, nb=nb, q=q, Hq=Hq, Wq=Wq, ind_k=ind_k, wk=wk,
                        it=it, iy=iy, ix=ix, warped=wy_flat, w=w_flat, dx=dx, dy=dy, val=val, iwes=iwes,
                        sm_field=sm_field, npos=int(num_pos_events), mean_div=mean_div)
        if sm_field is not None:
            self.ctx.update(sdx=sdx, sdy=sdy, cx=cx, cy=cy)
        iw_out = iwes.reshape(B, R, 2, H, W) if P == 2 else iwes.reshape(B, R, H, W)
        return dict(loss=loss, focus_loss=focus, smoothness_loss=smooth, iwes=iw_out,
                    iwe_raw=raw, flow_lut=lut6, flow_to_next=flow_to_next, ind_k=ind_k,
                    dist_k=dist_k, warped=warped_yx, weights=weights)

    # -- backward (closed forms of SURVEY.md section 8a) ----------------------
    def backward(self, grad_loss: float = 1.0):
        c = self.ctx
        dt_ = self.dtype
        H, W = self.image_shape
        B, R, P, M, n, nb, q, Hq, Wq = (c[k] for k in ("B", "R", "P", "M", "n", "nb", "q", "Hq", "Wq"))
        K = self.num_knn
        # focus -> blurred IWE
        N = c["dx"].size
        coef = dt_(grad_loss) * (-(dt_(1) / (c["val"] * c["val"]))) / dt_(N)
        if self.focus_loss_type == "variance":
            I = c["iwes"]
            npix = H * W
            planes = I.shape[0] * I.shape[1]
            mean = I.mean(axis=(-2, -1), keepdims=True, dtype=np.float64)
            d_blur = (dt_(grad_loss) * (-(dt_(1) / (c["val"] * c["val"]))) / dt_(planes)
                      * 2 * (I - mean) / (npix - 1)).astype(dt_)
        elif self.focus_loss_norm == "l1":
            gdx, gdy = coef * np.sign(c["dx"]), coef * np.sign(c["dy"])
            d_blur = sobel_T(gdx.astype(dt_), gdy.astype(dt_))
        else:
            gdx, gdy = coef * 2 * c["dx"], coef * 2 * c["dy"]
            d_blur = sobel_T(gdx.astype(dt_), gdy.astype(dt_))
        D = blur_T(d_blur, 1.0)                                          # dL/d raw IWE [B*R,P,H,W]

        # raw IWE -> warped coordinates -> LUT (event_image_converter.py:382-386)
        warped, w = c["warped"], c["w"]
        inds, mask, frac = vote_corners(warped, (H, W))
        fy, fx = frac[..., 0], frac[..., 1]
        plane = np.zeros((B * R, M), np.int64)
        if P == 2:
            plane[:, c["npos"]:] = 1
        Dflat = D.reshape(B * R, P, H * W)
        rows = np.arange(B * R)[:, None, None]
        Dc = Dflat[rows, plane[..., None], inds] * mask                  # [B*R, M, 4]
        D00, D10, D01, D11 = Dc[..., 0], Dc[..., 1], Dc[..., 2], Dc[..., 3]
        one = dt_(1)
        gy = w * (-(one - fx) * D00 + (one - fx) * D10 - fx * D01 + fx * D11)
        gx = w * (-(one - fy) * D00 - fy * D10 + (one - fy) * D01 + fy * D11)
        g = np.stack((gy, gx), -1).reshape(B, R, M, 2)
        dlut = np.zeros((B, nb, Hq, Wq, R, 2), dt_)
        bidx = np.broadcast_to(np.arange(B)[:, None], (B, M))
        for r in range(R):
            np.add.at(dlut[:, :, :, :, r], (bidx, c["it"], c["iy"], c["ix"]), g[:, r])

        # smoothness
        d_f2n = None
        if c["sm_field"] is not None:
            cnt = c["sdx"].size
            sc = dt_(grad_loss) * dt_(self.smooth_weight) / dt_(2 * cnt)
            gsm = sobel_T((sc * c["sdx"] / c["cx"]).astype(dt_), (sc * c["sdy"] / c["cy"]).astype(dt_))
            if self.smooth_type == "on_flow_to_tref":
                dlut += gsm.reshape(B, nb, R, 2, Hq, Wq).transpose(0, 1, 4, 5, 2, 3)
            else:
                d_f2n = gsm.reshape(B, nb - 1, 1, 2, Hq, Wq).transpose(0, 1, 4, 5, 2, 3) \
                    .reshape(B, nb - 1, q, 2)
        dlut = dlut.reshape(B, nb, q, R, 2)

        # LUT -> trajectories (focus.py:140-156)
        dtraj = np.zeros((B, R + nb, n, 2), dt_)
        ind_k, wk = c["ind_k"], c["wk"]
        bi = np.broadcast_to(np.arange(B)[:, None, None, None], ind_k.shape)
        ti = np.broadcast_to(np.arange(nb)[None, :, None, None], ind_k.shape)
        for r in range(R):
            contrib = wk[..., None] * dlut[:, :, :, None, r, :]           # [B, nb, q, K, 2]
            np.add.at(dtraj, (bi, np.full_like(bi, r), ind_k), contrib)
            np.add.at(dtraj, (bi, ti + R, ind_k), -contrib)
        if d_f2n is not None:
            contrib = np.broadcast_to(d_f2n[:, :, :, None, :] / dt_(K), ind_k[:, :-1].shape + (2,))
            np.add.at(dtraj, (bi[:, :-1], ti[:, :-1] + R + 1, ind_k[:, :-1]), contrib)
            np.add.at(dtraj, (bi[:, :-1], ti[:, :-1] + R, ind_k[:, :-1]), -contrib)
        return dict(dtraj=dtraj, dlut=dlut.reshape(B, nb, Hq, Wq, R, 2), d_iwe_raw=D)
