"""Trajectory glue: tile mask, coefficient grid -> trajectories (fused CUDA front end).

Host-side mirror of upstream ``src/utils/trajectories.py`` (``get_optical_flow_tile_mask``
:3-13), ``src/utils/basis.py`` (``compute_basis`` :4-46) and
``TrajectoryNet.calculate_trajectories_at_t`` / ``calculate_coords``
(``src/modules/trajectory_net.py:101-119``), plus the Bezier basis of
``src/models/raft_spline/curves/bezier.py:68-113`` for which the reference has no adapter to
``trajectories``.

The reference gathers the tile-centre pixels with a boolean mask, builds
``[b, n, n_t, K]`` products and permutes; here one kernel reads the K coefficients of a tile
centre straight from the dense grid, evaluates the basis from a tiny ``[n_t, K]`` table held in
shared memory and writes ``[B, n_t, n, 2]``; the backward writes the sparse tile pixels of
``d coeff_grid`` directly.
"""
from __future__ import annotations

import math

import torch

from . import cabi


def get_optical_flow_tile_mask(image_shape, tile_size):
    """upstream trajectories.py:3-13 (kept in torch: it is a constant buffer)."""
    mask = torch.zeros(tuple(image_shape), dtype=torch.bool)
    s = tile_size // 2
    mask[s::tile_size, s::tile_size] = True
    return mask


def tile_positions(image_shape, tile_size, device=None):
    """``torch.nonzero(mask)`` of the tile mask: [n, 2] int64 (y, x), row-major."""
    h, w = image_shape
    o = tile_size // 2
    ys = torch.arange(o, h, tile_size, device=device)
    xs = torch.arange(o, w, tile_size, device=device)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack((gy, gx), -1).reshape(-1, 2)


def basis_table(times: torch.Tensor, num_basis: int, basis_type: str, anchor: float = 0.0):
    """phi[t, k] - phi[anchor, k] as float32 on ``times.device``; k = 1..num_basis.

    polynomial: t**k (basis.py:29-31); dct: sqrt(2) cos(pi/2 (2t+1) k) (basis.py:18-24);
    bezier: C(d,k) (1-t)^(d-k) t^k evaluated in float64 then cast (bezier.py:74-107).
    """
    t = times.reshape(-1)
    k = torch.arange(1, num_basis + 1, device=t.device)

    def phi(tt):
        if basis_type == "polynomial":
            return tt[:, None].to(torch.float32) ** k[None, :]
        if basis_type == "dct":
            a = (2 * tt[:, None].to(torch.float32) + 1) * k[None, :]
            return math.sqrt(2.0) * torch.cos(math.pi / 2.0 * a)
        if basis_type == "bezier":
            d = num_basis
            t64 = tt[:, None].to(torch.float64)
            binom = torch.tensor([math.comb(d, i) for i in range(1, d + 1)], dtype=torch.float64,
                                 device=t.device)
            return (binom[None] * (1 - t64) ** (d - k[None]) * t64 ** k[None]).to(torch.float32)
        raise ValueError(basis_type)

    anchor_t = torch.full((1,), float(anchor), device=t.device, dtype=t.dtype)
    return (phi(t) - phi(anchor_t)).to(torch.float32).contiguous()


class _TrajectoriesFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coeff_grid, phi, patch, xy_order, add_offsets):
        lib = cabi.load()
        cg = coeff_grid.detach().to(torch.float32).contiguous()
        B, S, C2, H, W = cg.shape
        n_t, K = phi.shape
        assert C2 == 2 * K
        o = patch // 2
        ny, nx = (H - o + patch - 1) // patch, (W - o + patch - 1) // patch
        out = torch.empty((B, n_t, ny * nx, 2), dtype=torch.float32, device=cg.device)
        rc = lib.cmax_trajectories_forward(cabi.ptr(cg), cabi.ptr(phi), B, S, K, H, W, patch, n_t,
                                           int(xy_order), int(add_offsets), cabi.ptr(out),
                                           cabi.stream_ptr(cg.device))
        cabi.check(rc, "cmax_trajectories_forward")
        # a learned basis (basis.py:26-27: phi = MLP(t)) needs d loss / d phi, which contracts the
        # trajectory gradient with the tile-centre coefficients: keep those (tiny) when asked for
        tile_coeff = None
        if phi.requires_grad:
            tile_coeff = cg[:, :, :, o::patch, o::patch].sum(1).reshape(B, 2 * K, ny * nx)
        ctx.save_for_backward(phi.detach(), tile_coeff)
        ctx.meta = (B, S, K, H, W, patch, n_t, int(xy_order))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = cabi.load()
        phi, tile_coeff = ctx.saved_tensors
        B, S, K, H, W, patch, n_t, xy = ctx.meta
        g = grad_out.detach().to(torch.float32).contiguous()
        dcg = None
        if ctx.needs_input_grad[0]:
            dcg = torch.empty((B, S, 2 * K, H, W), dtype=torch.float32, device=g.device)
            rc = lib.cmax_trajectories_backward(cabi.ptr(g), cabi.ptr(phi), B, S, K, H, W, patch, n_t,
                                                xy, cabi.ptr(dcg), cabi.stream_ptr(g.device))
            cabi.check(rc, "cmax_trajectories_backward")
        dphi = None
        if ctx.needs_input_grad[1] and tile_coeff is not None:
            # traj[b,t,j,a] = sum_k phi[t,k] c_a[b,k,j]  =>  dphi[t,k] = sum_{b,j,a} dtraj[b,t,j,a] c_a[b,k,j]
            first, second = tile_coeff[:, :K], tile_coeff[:, K:]
            cy, cx = (second, first) if xy else (first, second)
            dphi = torch.einsum("btj,bkj->tk", g[..., 0], cy) + torch.einsum("btj,bkj->tk", g[..., 1], cx)
        return dcg, dphi, None, None, None


def calculate_trajectories_at_t(coeff_grid: torch.Tensor, times: torch.Tensor, patch_size: int,
                                num_basis: int, basis_type: str = "polynomial",
                                anchor_time: float = 0.0, add_offsets: bool = True,
                                xy_order: bool = False) -> torch.Tensor:
    """trajectory_net.py:113-119: ``coeff_grid [B, (S,) 2K, H, W]`` -> ``[B, n_t, n, 2]`` (y, x).

    ``xy_order=True`` reads RAFT-spline style parameters whose first K channels are x.
    The learned basis (basis.py:26-27) is an MLP and stays in PyTorch: pass its output as a
    ready table through :func:`trajectories_from_table`.
    """
    if coeff_grid.dim() == 4:
        coeff_grid = coeff_grid[:, None]
    if not coeff_grid.is_cuda:
        raise RuntimeError("calculate_trajectories_at_t (B200) needs CUDA tensors; no CPU path")
    phi = basis_table(times.to(coeff_grid.device), num_basis, basis_type, anchor_time)
    return _TrajectoriesFunction.apply(coeff_grid, phi, int(patch_size), bool(xy_order),
                                       bool(add_offsets))


def trajectories_from_table(coeff_grid: torch.Tensor, phi: torch.Tensor, patch_size: int,
                            add_offsets: bool = True, xy_order: bool = False) -> torch.Tensor:
    """Same as above with a caller-built basis table ``phi [n_t, K]`` (already anchor-subtracted),
    e.g. the learned MLP basis of basis.py:26-27.  Differentiable in ``coeff_grid`` *and* ``phi``
    (the reference back-propagates into ``basis_network(times)``)."""
    if coeff_grid.dim() == 4:
        coeff_grid = coeff_grid[:, None]
    if not coeff_grid.is_cuda:
        raise RuntimeError("trajectories_from_table (B200) needs CUDA tensors; no CPU path")
    return _TrajectoriesFunction.apply(coeff_grid, phi.to(torch.float32).contiguous(),
                                       int(patch_size), bool(xy_order), bool(add_offsets))
