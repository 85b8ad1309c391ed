"""``FocusLoss`` - the motion-prior contrast-maximisation loss, computed by the sm_100a kernels.

Host-side mirror of upstream ``src/losses/focus.py:9-246``: same constructor keywords, same
attributes (``is_needing_offsets``, ``imager``), same ``get_reconstruction_times`` and the same
``calc(trajectories, times, batch) -> (loss, log_metadata, misc_metadata)`` contract, so
``scripts/flow_training.py:89`` only swaps the import.  All arithmetic runs in
``libcmax_b200.so`` through ``cmax_forward`` / ``cmax_backward`` (include/cmax_b200.h); torch
provides tensors, the current stream and the autograd graph node.  No CPU fallback.
"""
from __future__ import annotations

import torch

from . import base
from .. import cabi
from ..utils import EventImageConverter


class _CmaxLossFunction(torch.autograd.Function):
    """autograd node: forward = cmax_forward, backward = cmax_backward."""
    strict_events = False        # set per call by FocusLoss.calc (single-threaded caller per process)

    @staticmethod
    def forward(ctx, trajectories, times, events, cfg, num_pos_events, want_lut, seg_start=None):
        lib = cabi.load()
        dev = trajectories.device
        traj = trajectories.detach().to(torch.float32).contiguous()
        tms = times.detach().to(device=dev, dtype=torch.float32).contiguous()
        ev = events.detach().to(torch.float32).contiguous()
        B, n_t, n, two = traj.shape
        assert two == 2 and n_t == cfg.num_tref + cfg.num_bins, "trajectories must be [B, R+nb, n, 2]"
        packed = seg_start is not None
        if packed:       # io.PackedEvents: records [B, M, 4] + seg_start [B, G*NT+1]
            assert ev.dim() == 3 and ev.shape[0] == B and ev.shape[2] == 4, "records must be [B, M, 4]"
            seg = seg_start.detach().to(device=dev, dtype=torch.int32).contiguous()
            _, nty, ntx, G = cabi.pack_layout(cfg)
            assert tuple(seg.shape) == (B, G * nty * ntx + 1), "seg_start must be [B, G*NT+1]"
        else:
            assert ev.dim() == 3 and ev.shape[0] == B and ev.shape[2] == 6, "events must be [B, M, 6]"
            seg = None
        assert tms.numel() == n_t
        M = ev.shape[1]
        H, W = cfg.height, cfg.width
        P = 2 if cfg.polarity_aware_batching else 1
        need = lib.cmax_workspace_bytes(cfg, B, M, n)
        if need == 0:
            # let the library name the problem
            cabi.check(lib.cmax_forward(cfg, None, None, None, B, M, n, num_pos_events, None, None,
                                        None, None, 0, None), "cmax_forward")
            raise RuntimeError("cmax_workspace_bytes returned 0")
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        iwes = torch.empty((B * cfg.num_tref, P, H, W), dtype=torch.float32, device=dev)
        losses = torch.empty(3, dtype=torch.float32, device=dev)
        lut = None
        if want_lut:
            s = cfg.lut_superpixel_size
            lut = torch.empty((B, cfg.num_bins, (H + s - 1) // s, (W + s - 1) // s, cfg.num_tref, 2),
                              dtype=torch.float32, device=dev)
        if packed:
            rc = lib.cmax_forward_packed(cfg, cabi.ptr(traj), cabi.ptr(tms), cabi.ptr(ev), cabi.ptr(seg),
                                         B, M, n, cabi.ptr(iwes), cabi.ptr(losses), cabi.ptr(lut),
                                         cabi.ptr(ws), need, cabi.stream_ptr(dev))
            cabi.check(rc, "cmax_forward_packed")
        else:
            rc = lib.cmax_forward(cfg, cabi.ptr(traj), cabi.ptr(tms), cabi.ptr(ev), B, M, n,
                                  int(num_pos_events), cabi.ptr(iwes), cabi.ptr(losses), cabi.ptr(lut),
                                  cabi.ptr(ws), need, cabi.stream_ptr(dev))
            cabi.check(rc, "cmax_forward")
        if _CmaxLossFunction.strict_events:
            # opt-in debug check (synchronises): the kernels skip events whose LUT cell is outside the
            # table or whose coordinates / bin are not finite and only count them; the reference
            # raises IndexError / wraps negative indices (focus.py:188) or turns the loss into NaN
            import ctypes
            st = (ctypes.c_int64 * 4)()
            cabi.check(lib.cmax_read_status(cabi.ptr(ws), ctypes.byref(st), cabi.stream_ptr(dev)), "cmax_read_status")
            if st[0] > 0:
                raise RuntimeError(f"FocusLoss(strict_events=True): {int(st[0])} valid events were skipped (LUT "
                                   "cell out of range or non-finite coordinates / bin) - broken loader or a "
                                   "diverged network")
        ctx.cfg = cfg          # the same config (incl. the backward_follows hint) must reach cmax_backward
        ctx.dims = (B, M, n, int(num_pos_events), need)
        ctx.packed = packed
        if packed:
            ctx.save_for_backward(traj, tms, ev, ws, seg)
        else:
            ctx.save_for_backward(traj, tms, ev, ws)
        ctx.mark_non_differentiable(iwes, losses)
        if lut is not None:
            ctx.mark_non_differentiable(lut)
            return losses[0].clone(), losses, iwes, lut
        return losses[0].clone(), losses, iwes

    @staticmethod
    def backward(ctx, grad_loss, *unused):
        lib = cabi.load()
        traj, tms, ev, ws = ctx.saved_tensors[:4]
        B, M, n, npos, need = ctx.dims
        g = grad_loss.detach().to(device=traj.device, dtype=torch.float32).reshape(1).contiguous()
        dtraj = torch.empty_like(traj)
        if ctx.packed:
            seg = ctx.saved_tensors[4]
            rc = lib.cmax_backward_packed(ctx.cfg, cabi.ptr(traj), cabi.ptr(tms), cabi.ptr(ev),
                                          cabi.ptr(seg), B, M, n, cabi.ptr(g), cabi.ptr(dtraj),
                                          cabi.ptr(ws), need, cabi.stream_ptr(traj.device))
            cabi.check(rc, "cmax_backward_packed")
        else:
            rc = lib.cmax_backward(ctx.cfg, cabi.ptr(traj), cabi.ptr(tms), cabi.ptr(ev), B, M, n, npos,
                                   cabi.ptr(g), cabi.ptr(dtraj), cabi.ptr(ws), need,
                                   cabi.stream_ptr(traj.device))
            cabi.check(rc, "cmax_backward")
        return dtraj, None, None, None, None, None, None


class FocusLoss(base.TrajectoryLossBase):
    """
    Implements the Focus Loss of MotionPriorCM (https://arxiv.org/pdf/2407.10802) on B200.

    Args (identical to upstream focus.py:13-27):
        image_shape (tuple): Shape of the image as (height, width).
        num_tref (int): Number of reference time points. If 1, uses a random reference time.
        num_bins (int): Number of voxel grid channels.
        num_knn (int): Number of nearest neighbors used in interpolation.
        smooth_weight (float): Weight for smoothness loss.
        lut_superpixel_size (int): Determines size of the flow look-up-table.
        focus_loss_norm (str): 'l1' or 'l2'.
        dist_norm (str): 'l1' or 'l2'.
        scale_iwe_by_dt (bool), mask_image_border (bool), polarity_aware_batching (bool)
        interpolation_scheme (str): 'mean' or 'iwd'.
        smooth_type (str): 'on_flow_to_tref' or 'on_flow_to_next'.
    Extra (optional, new): deterministic (bool) - int64 fixed-point IWE / LUT-gradient
        accumulation, run-to-run bit-identical; strict_events (bool) - debug check that raises when
        events had to be skipped (LUT cell out of range, NaN), where upstream raises or returns NaN;
        focus_loss_type ('gradient_magnitude' as upstream
        calc hard-codes, or 'variance' = upstream utils.calculate_focus_loss(loss_type='variance')).
    """

    def __init__(self, image_shape, num_tref, num_bins, num_knn, smooth_weight,
                 lut_superpixel_size, focus_loss_norm, dist_norm,
                 scale_iwe_by_dt, mask_image_border, polarity_aware_batching,
                 interpolation_scheme, smooth_type, deterministic=False,
                 focus_loss_type='gradient_magnitude', strict_events=False, **kwargs):
        super().__init__()
        self.image_shape = tuple(image_shape)
        self.num_tref = num_tref
        self.num_bins = num_bins
        self.num_knn = num_knn
        self.smooth_weight = smooth_weight
        self.lut_superpixel_size = lut_superpixel_size
        self.focus_loss_norm = focus_loss_norm
        self.dist_norm = dist_norm
        self.scale_iwe_by_dt = scale_iwe_by_dt
        self.mask_image_border = mask_image_border
        self.polarity_aware_batching = polarity_aware_batching
        self.interpolation_scheme = interpolation_scheme
        self.smooth_type = smooth_type
        self.deterministic = bool(deterministic)
        self.focus_loss_type = focus_loss_type      # upstream calc hard-codes 'gradient_magnitude' (focus.py:90)
        self.strict_events = bool(strict_events)    # debug: raise when the kernels had to skip events
        self.is_needing_offsets = True
        self.imager = EventImageConverter(self.image_shape, deterministic=self.deterministic)

        assert not scale_iwe_by_dt or num_tref == 1                       # focus.py:49-51
        assert not polarity_aware_batching or num_tref == 1
        assert not smooth_type == 'on_flow_to_next' or num_tref == 1
        if focus_loss_norm not in cabi.NORM or dist_norm not in cabi.NORM:
            raise ValueError
        if smooth_type not in cabi.SMOOTH:
            raise ValueError
        if num_knn > 1 and interpolation_scheme not in cabi.INTERP:
            raise ValueError
        self._cfg = cabi.make_config(
            self.image_shape, num_tref, num_bins, num_knn, smooth_weight, lut_superpixel_size,
            focus_loss_norm, dist_norm, scale_iwe_by_dt, mask_image_border,
            polarity_aware_batching, interpolation_scheme if num_knn > 1 else 'mean', smooth_type,
            self.deterministic, focus_loss_type)
        self._cfg_train = cabi.with_backward_hint(self._cfg)
        cabi.load()      # fail at construction time when the CUDA library is missing

    def get_reconstruction_times(self, device):
        """focus.py:53-64 (same torch RNG draw for the random reference time)."""
        if self.num_tref > 1:
            t_ref = torch.linspace(0, 1, self.num_tref, device=device)
        elif self.num_tref == 1:
            t_ref = torch.rand(1, device=device)  # Random reference time
        else:
            raise ValueError("Invalid value for num_tref. Must be >= 1.")
        t_bins = torch.linspace(0, 1, self.num_bins + 1, device=device)
        t_mid = (t_bins[:-1] + t_bins[1:]) / 2
        return torch.concat((t_ref, t_mid), dim=0)

    def calc(self, trajectories, times, batch, return_flow_lut: bool = False):
        """focus.py:66-113.

        trajectories [B, num_tref + num_bins, n, 2] (y, x), times [num_tref + num_bins],
        batch {'events': [B, M, 6], 'num_pos_events': int}; 'events' may also be an
        `io.PackedEvents` (tile-binned loader-side layout) or an `io.CompactEvents` (its 12-byte
        wire form) - same results, faster event stage, fewer bytes over PCIe.
        Returns (loss, {'focus_loss', 'smoothness_loss'}, {'iwes': ...}).
        """
        events = batch['events']
        if not (trajectories.is_cuda and events.is_cuda):
            raise RuntimeError("FocusLoss (B200) needs CUDA tensors; there is no CPU fallback")
        _CmaxLossFunction.strict_events = self.strict_events
        # training (a backward will follow): the forward also emits dL/dIWE from its image pass
        cfg = self._cfg_train if (trajectories.requires_grad and torch.is_grad_enabled()) else self._cfg
        if hasattr(events, 'fine_start'):
            # io.CompactEvents / io.BitpackedEvents (wire layouts): rebuild the packed records on the
            # device first
            from .. import io as _io
            events = _io.expand_wire(events, self)
        if hasattr(events, 'seg_start'):
            # io.PackedEvents: the loader-side tile-binned layout; the polarity split is part of it
            out = _CmaxLossFunction.apply(trajectories, times, events.records, cfg, 0,
                                          bool(return_flow_lut), events.seg_start)
        else:
            num_pos_events = batch['num_pos_events'] if 'num_pos_events' in batch else -1
            assert not self.polarity_aware_batching or num_pos_events > -1
            out = _CmaxLossFunction.apply(trajectories, times, events, cfg,
                                          int(num_pos_events), bool(return_flow_lut))
        loss, losses, iwes = out[0], out[1], out[2]

        h, w = self.image_shape
        b = trajectories.shape[0]
        n_tref = self.num_tref
        if self.polarity_aware_batching:
            iwes = iwes.reshape(b, n_tref, 2, h, w)
        else:
            iwes = iwes.reshape(b, n_tref, h, w)

        log_metadata = {
            'focus_loss': losses[1].detach(),
            'smoothness_loss': losses[2].detach(),
        }
        misc_metadata = {
            'iwes': iwes.detach()
        }
        if return_flow_lut:
            misc_metadata['flow_lut'] = out[3]
        return loss, log_metadata, misc_metadata

    def calc_event_sharded(self, trajectories, times, batch, group=None):
        """`calc` for windows whose event rows are sharded over the ranks of `group` (SURVEY 8e,
        second mode): `batch['events']` = this rank's rows; two in-place all-reduces (raw IWE,
        dLUT) make loss, IWEs and gradients identical on every rank.  See losses/sharded.py."""
        from . import sharded
        return sharded.calc_event_sharded(self, trajectories, times, batch, group=group)
