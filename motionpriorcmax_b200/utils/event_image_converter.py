"""``EventImageConverter`` backed by the CUDA splat kernels.

Mirrors the part of upstream ``src/utils/event_image_converter.py`` the loss plugin and the
image-logging callback use (``create_iwe`` :45-74, ``create_image_from_events_tensor`` :134-176,
``bilinear_vote_tensor`` :333-391, ``count_event_tensor`` :226-272): same names, arguments and
returned shapes.  Tensors must live on a CUDA device - there is no CPU path.
"""
from __future__ import annotations

from typing import Tuple, Union

import torch

from .. import cabi


class EventImageConverter(object):
    def __init__(self, image_size: tuple, outer_padding: Union[int, Tuple[int, int]] = 0,
                 deterministic: bool = False):
        if isinstance(outer_padding, (int, float)):
            self.outer_padding = (int(outer_padding), int(outer_padding))
        else:
            self.outer_padding = tuple(outer_padding)
        self.outer_padding = tuple(int(p) for p in self.outer_padding)
        if min(self.outer_padding) < 0:
            raise ValueError("outer_padding must be >= 0")
        # upstream :28: the images this converter returns are the padded ones
        self.image_size = tuple(int(i + p * 2) for i, p in zip(image_size, self.outer_padding))
        self.deterministic = bool(deterministic)

    # -- upstream :30-43 (kept as it is there, including the single - not double - padding it adds) --
    def update_property(self, image_size=None, outer_padding=None):
        if image_size is not None:
            self.image_size = image_size
        if outer_padding is not None:
            if isinstance(outer_padding, int):
                self.outer_padding = (outer_padding, outer_padding)
            else:
                self.outer_padding = outer_padding
        self.image_size = tuple(i + p for i, p in zip(self.image_size, self.outer_padding))

    # -- upstream :45-74 ------------------------------------------------------
    def create_iwe(self, events: torch.Tensor, method: str = "bilinear_vote", sigma: int = 1,
                   weight=1.0) -> torch.Tensor:
        if hasattr(events, "run_hdr"):                       # io.BitpackedEvents: decode to records first
            raise NotImplementedError("pass io.expand_bitpacked(events, loss) (a PackedEvents) to the imager")
        if hasattr(events, "fine_start") or hasattr(events, "seg_start"):
            # loader-side layouts (io.CompactEvents / io.PackedEvents) as `batch['events']`: what the
            # image-logging callback hands over (upstream src/utils/logging.py:76-79).  Rendered from
            # the valid events; the reference's padding rows (all zero) each add one vote at pixel
            # (0, 0) there, which this rendering does not reproduce.
            events, weight = self._rows_of_packed(events, weight)
        if not isinstance(events, torch.Tensor):
            e = f"Non-supported type of events. {type(events)}"
            raise RuntimeError(e)
        iwes = self.create_image_from_events_tensor(events, method, sigma=sigma, weight=weight)
        if len(iwes.shape) == 2:
            iwes = iwes[None]
        return iwes

    # -- upstream :134-176 ----------------------------------------------------
    def create_image_from_events_tensor(self, events: torch.Tensor, method: str = "bilinear_vote",
                                        weight=1.0, sigma: int = 0) -> torch.Tensor:
        if method == "count":
            if sigma > 0:
                raise NotImplementedError("blurred count image")
            return self.count_event_tensor(events)
        if method == "bilinear_vote":
            return torch.squeeze(self._vote(events, weight, float(sigma)))
        if method == "polarity":
            # upstream :156-163: boolean indexing flattens a batch, so batched events give ONE
            # (positive, negative) image pair of all windows together - kept as it is there
            pos_flag = events[..., 3] > 0
            wt = isinstance(weight, torch.Tensor)
            pos = self._vote(events[pos_flag], weight[pos_flag] if wt else weight, float(sigma))
            neg = self._vote(events[~pos_flag], weight[~pos_flag] if wt else weight, float(sigma))
            return torch.squeeze(torch.stack([pos, neg], axis=-3))
        e = f"{method = } is not implemented"
        raise NotImplementedError(e)

    # -- upstream :333-391 ----------------------------------------------------
    def bilinear_vote_tensor(self, events: torch.Tensor, weight=1.0) -> torch.Tensor:
        return self._vote(events, weight, 0.0).squeeze()

    # -- upstream :226-272 (exact integer votes) ------------------------------
    def count_event_tensor(self, events: torch.Tensor) -> torch.Tensor:
        ev = self._prep(events)
        h, w = self.image_size
        ph, pw = self.outer_padding
        nb, m, c = ev.shape
        out = torch.empty((nb, h, w), dtype=torch.int64, device=ev.device)
        lib = cabi.load()
        with torch.cuda.device(ev.device):
            cabi.check(lib.cmax_count_image_padded(cabi.ptr(ev), nb, m, c, h, w, ph, pw, cabi.ptr(out),
                                                   cabi.stream_ptr(ev.device)), "cmax_count_image_padded")
        return out.squeeze()

    # ------------------------------------------------------------------------
    @staticmethod
    def _rows_of_packed(events, weight):
        """(rows [B, M, >=2], weight [B, M]) of a loader-side layout: records / coordinates of the
        valid events, a 0 weight on the unused tail of every window."""
        if not isinstance(weight, (int, float)) or float(weight) != 1.0:
            raise NotImplementedError("per-event weights with a packed / compact events layout")
        if hasattr(events, "fine_start"):                    # CompactEvents: ragged [T, 3]
            counts = events.fine_start[:, -1].to(torch.int64)
            off = events.sample_off.to(torch.int64)
            B, Mp = counts.shape[0], int(events.max_count)
            idx = off[:-1, None] + torch.arange(Mp, device=counts.device)[None]
            mask = torch.arange(Mp, device=counts.device)[None] < counts[:, None]
            rows = events.coords[idx.clamp_(max=events.coords.shape[0] - 1)]
        else:                                                # PackedEvents: [B, M, 4]
            counts = events.seg_start[:, -1].to(torch.int64)
            rows = events.records
            mask = torch.arange(rows.shape[1], device=rows.device)[None] < counts[:, None]
        rows = torch.where(mask[..., None], rows, torch.zeros((), dtype=rows.dtype, device=rows.device))
        return rows, mask.to(torch.float32)

    @staticmethod
    def _prep(events: torch.Tensor) -> torch.Tensor:
        if not events.is_cuda:
            raise RuntimeError("EventImageConverter (B200) needs CUDA tensors; there is no CPU path")
        if events.dim() == 2:
            events = events[None]
        if events.dim() != 3 or events.shape[-1] < 2:
            raise ValueError("events must be [(b,) n_events, >=2]")
        if events.requires_grad and torch.is_grad_enabled():
            # upstream create_iwe is differentiable in the coordinates; this stand-alone entry is a
            # forward-only renderer (the loss path has its own backward) - refuse instead of silently
            # cutting the graph
            raise RuntimeError("EventImageConverter (B200) is forward only: events require grad. Use "
                               "FocusLoss.calc for a differentiable IWE, or call under torch.no_grad() / "
                               "with events.detach().")
        return events.detach().to(torch.float32).contiguous()

    def _vote(self, events: torch.Tensor, weight, sigma: float) -> torch.Tensor:
        ev = self._prep(events)
        h, w = self.image_size
        nb, m, c = ev.shape
        wt = None
        if isinstance(weight, torch.Tensor):
            assert weight.shape == events.shape[:-1]
            if weight.requires_grad and torch.is_grad_enabled():
                raise RuntimeError("EventImageConverter (B200) is forward only: weight requires grad")
            wt = weight.detach().to(torch.float32).reshape(nb, m).contiguous()
        elif float(weight) != 1.0:
            wt = torch.full((nb, m), float(weight), dtype=torch.float32, device=ev.device)
        out = torch.empty((nb, h, w), dtype=torch.float32, device=ev.device)
        scratch = torch.empty_like(out) if sigma > 0 else None
        scratch64 = (torch.empty((nb, h, w), dtype=torch.int64, device=ev.device)
                     if self.deterministic else None)
        lib = cabi.load()
        ph, pw = self.outer_padding
        with torch.cuda.device(ev.device):
            rc = lib.cmax_create_iwe_padded(cabi.ptr(ev), cabi.ptr(wt), nb, m, c, h, w, ph, pw, sigma,
                                            cabi.ptr(out), cabi.ptr(scratch), cabi.ptr(scratch64),
                                            int(self.deterministic), cabi.stream_ptr(ev.device))
        cabi.check(rc, "cmax_create_iwe_padded")
        if sigma > 0:
            out = out[:, None]            # upstream blurs a [nb, 1, H, W] view, then squeezes
        return out
