"""Synthetic event windows and trajectory-coefficient fields shaped like the reference's data.

There is no dataset access, so the benchmark and the tests use inputs that follow the
reference loaders' layout exactly (upstream ``src/loader/dsec/loader.py:141-182`` for a
sample, ``:360-415`` for padding / collation, ``src/loader/evimo2/datasubset.py:146-229``
for the EVIMO2 shape):

* events ``[B, M, 6]`` float32, columns ``(y, x, t, p, bin, valid)``; ``t`` sorted in [0, 1]
  with ``t.min() == 0`` and ``t.max() == 1``; ``bin = clip(searchsorted(linspace(0,1,nb+1), t) - 1, 0)``;
* polarity-aware batching: per sample the positive events (padded with all-zero rows up to
  the batch maximum) are followed by the negative events (padded likewise) and
  ``num_pos_events`` is the batch-wide maximum number of positives;
* seeds: ``1234 + 1000 * rank + sample_index``.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Union

import torch


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def make_coeff_grid(B: int, num_basis: int, H: int, W: int, sigma_px: float = 8.0,
                    seed: int = 1234, coarse: Sequence[int] = (15, 20)) -> torch.Tensor:
    """Smooth low-frequency coefficient field ``[B, 1, 2K, H, W]`` (what the UNet would emit):
    N(0, sigma_px) on a coarse grid, bicubic-upsampled; order-k coefficients shrink as 1/k."""
    out = []
    for b in range(B):
        g = _gen(seed + b)
        c = torch.randn(1, 2 * num_basis, coarse[0], coarse[1], generator=g) * sigma_px
        scale = torch.tensor([1.0 / (k + 1) for k in range(num_basis)] * 2).view(1, -1, 1, 1)
        c = torch.nn.functional.interpolate(c * scale, size=(H, W), mode="bicubic",
                                            align_corners=True)
        out.append(c)
    return torch.cat(out, 0)[:, None].contiguous()


def _sample_window(n_events: int, H: int, W: int, num_bins: int, g: torch.Generator,
                   dist: str, integer_coords: bool, coord_scale: float) -> torch.Tensor:
    """One un-padded sample ``[n, 5]`` = (y, x, t, p, bin), time-sorted."""
    n = int(n_events)
    if dist == "uniform":
        y = torch.rand(n, generator=g) * H
        x = torch.rand(n, generator=g) * W
    elif dist == "edges":
        # events on ~200 random line segments with N(0, 0.7 px) jitter
        nseg = 200
        p0 = torch.rand(nseg, 2, generator=g) * torch.tensor([H, W], dtype=torch.float32)
        ang = torch.rand(nseg, generator=g) * (2 * math.pi)
        length = 20 + torch.rand(nseg, generator=g) * 180
        seg = torch.randint(0, nseg, (n,), generator=g)
        u = torch.rand(n, generator=g)
        y = p0[seg, 0] + u * length[seg] * torch.sin(ang[seg]) + torch.randn(n, generator=g) * 0.7
        x = p0[seg, 1] + u * length[seg] * torch.cos(ang[seg]) + torch.randn(n, generator=g) * 0.7
        # the loader masks events to [0,H) x [0,W) (loader.py:160-161): wrap instead of drop
        y = torch.remainder(y, H)
        x = torch.remainder(x, W)
    else:
        raise ValueError(dist)
    if integer_coords:
        # EVIMO2: integer sensor coordinates scaled by x_scale = y_scale (datasubset.py:185-186)
        y = torch.floor(y / coord_scale) * coord_scale
        x = torch.floor(x / coord_scale) * coord_scale
    y = y.clamp_(0, math.nextafter(float(H), 0.0)).float()
    x = x.clamp_(0, math.nextafter(float(W), 0.0)).float()
    y = torch.where(y >= H, torch.full_like(y, H - 1.0), y)
    x = torch.where(x >= W, torch.full_like(x, W - 1.0), x)
    t = torch.rand(n, generator=g).sort().values
    if n > 1:
        t = (t - t[0]) / (t[-1] - t[0])
    p = (torch.rand(n, generator=g) < 0.5).float()
    edges = torch.linspace(0, 1, num_bins + 1)
    bins = (torch.searchsorted(edges, t.contiguous()) - 1).clamp_(min=0).float()
    bins = bins.clamp_(max=num_bins - 1)
    return torch.stack((y, x, t, p, bins), 1)


def _pad(ev: torch.Tensor, length: int) -> torch.Tensor:
    out = torch.zeros(length, 6, dtype=ev.dtype)        # loader.py:360-364
    out[: len(ev), :5] = ev
    out[: len(ev), 5] = 1
    return out


def make_event_batch(B: int, n_events: Union[int, Sequence[int]], H: int, W: int, num_bins: int,
                     polarity_aware_batching: bool = True, seed: int = 1234, rank: int = 0,
                     dist: str = "uniform", integer_coords: bool = False,
                     coord_scale: float = 1.0):
    """Collated batch following ``sequence_collate_fn`` (loader.py:366-415).

    Returns ``(events [B, M, 6] float32 CPU tensor, num_pos_events or None)``.
    """
    counts = [int(n_events)] * B if isinstance(n_events, int) else [int(v) for v in n_events]
    assert len(counts) == B
    samples = []
    for i, n in enumerate(counts):
        g = _gen(seed + 1000 * rank + i)
        samples.append(_sample_window(n, H, W, num_bins, g, dist, integer_coords, coord_scale))
    if not polarity_aware_batching:
        m = max(len(s) for s in samples)
        return torch.stack([_pad(s, m) for s in samples], 0), None
    pos = [s[s[:, 3] == 1] for s in samples]
    neg = [s[s[:, 3] == 0] for s in samples]
    mp = max(len(s) for s in pos)
    mn = max(len(s) for s in neg)
    ev = torch.stack([torch.cat((_pad(a, mp), _pad(b, mn)), 0) for a, b in zip(pos, neg)], 0)
    return ev, mp


def lognormal_event_counts(B: int, median: float = 1e6, sigma: float = 0.5, lo: float = 2e5,
                           hi: float = 4e6, seed: int = 1234, rank: int = 0):
    """Per-sample event counts ~ LogNormal(ln median, sigma) clipped to [lo, hi]."""
    g = _gen(seed + 1000 * rank + 777)
    z = torch.randn(B, generator=g)
    return [int(min(hi, max(lo, math.exp(math.log(median) + sigma * float(v))))) for v in z]


DSEC_LOSS_CONFIG = dict(           # upstream config/exe/flow_training/dsec.yaml:14-25 + propagate_config
    image_shape=(480, 640), num_tref=1, num_bins=15, num_knn=32, smooth_weight=0.003,
    lut_superpixel_size=4, focus_loss_norm="l1", dist_norm="l2", scale_iwe_by_dt=True,
    mask_image_border=True, polarity_aware_batching=True, interpolation_scheme="mean",
    smooth_type="on_flow_to_tref")

EVIMO2_LOSS_CONFIG = dict(         # upstream experiment yaml raft-spline_evimo2-300ms_..._Tab2L5.yaml:21-35
    image_shape=(384, 512), num_tref=1, num_bins=41, num_knn=32, smooth_weight=0.06,
    lut_superpixel_size=4, focus_loss_norm="l1", dist_norm="l2", scale_iwe_by_dt=True,
    mask_image_border=True, polarity_aware_batching=True, interpolation_scheme="mean",
    smooth_type="on_flow_to_next")


def multi_tref_variant(cfg: dict, num_tref: int) -> dict:
    """The only multi-reference-time combination upstream focus.py:49-51 allows."""
    out = dict(cfg)
    out.update(num_tref=num_tref, scale_iwe_by_dt=False, polarity_aware_batching=False,
               smooth_type="on_flow_to_tref")
    return out
