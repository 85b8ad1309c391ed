"""GPU event voxel grid - mirror of upstream ``VoxelGrid`` (``src/loader/dsec/utils.py:19-77``).

Same constructor (``input_size=(C, H, W), norm_type, quantile``) and ``convert(events)`` with
``events = {'p', 't', 'x', 'y'}`` 1-D float tensors; the tensors must live on a CUDA device and
the grid is built by ``cmax_voxel_grid`` (one streaming splat pass + the normalisation kernels)
instead of eight masked ``put_`` passes on the CPU.  For the rarely used quantile clipping
(``quantile > 0``; ``dsec.yaml`` ships 0) the order statistic is ``torch.quantile``; clipping and
normalisation of the clipped grid are ``cmax_voxel_normalize``.
"""
from __future__ import annotations

import torch

from . import cabi

_NORM = {None: 0, "mean_std": 1, "max": 2}


class VoxelGrid:
    def __init__(self, input_size: tuple, norm_type, quantile=0):
        assert len(input_size) == 3
        self.nb_channels = int(input_size[0])
        self.input_size = tuple(int(v) for v in input_size)
        self.norm_type = norm_type
        assert self.norm_type in ['mean_std', 'max', None]
        self.quantile = quantile
        assert 0 <= self.quantile < 0.15

    def convert(self, events):
        C, H, W = self.input_size
        x, y, t, p = (events[k] for k in ('x', 'y', 't', 'p'))
        if not x.is_cuda:
            raise RuntimeError("VoxelGrid (B200) needs CUDA tensors; there is no CPU path")
        x, y, t, p = (v.detach().to(torch.float32).contiguous() for v in (x, y, t, p))
        n = x.numel()
        assert y.numel() == n and t.numel() == n and p.numel() == n
        lib = cabi.load()
        grid = torch.empty((C, H, W), dtype=torch.float32, device=x.device)
        stats = torch.empty(4, dtype=torch.float64, device=x.device)
        norm = _NORM[self.norm_type] if self.quantile == 0 else 0
        rc = lib.cmax_voxel_grid(cabi.ptr(x), cabi.ptr(y), cabi.ptr(t), cabi.ptr(p), n, C, H, W, norm,
                                 cabi.ptr(grid), cabi.ptr(stats), cabi.stream_ptr(x.device))
        cabi.check(rc, "cmax_voxel_grid")
        if self.quantile > 0:
            # upstream :56-60: clip to the (1 - q) quantile of |grid|, then normalise the clipped grid.
            # The order statistic is torch's; clipping and normalisation run in the library on the
            # device-resident threshold (no host round trip).
            thr = torch.quantile(grid.abs().view(-1), 1 - self.quantile).reshape(1).contiguous()
            rc = lib.cmax_voxel_normalize(cabi.ptr(grid), C, H, W, _NORM[self.norm_type], cabi.ptr(thr),
                                          cabi.ptr(stats), cabi.stream_ptr(x.device))
            cabi.check(rc, "cmax_voxel_normalize")
        return grid
