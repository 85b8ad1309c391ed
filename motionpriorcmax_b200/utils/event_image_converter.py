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
        if self.outer_padding != (0, 0):
            raise NotImplementedError("outer_padding != 0 is not used on the loss path")
        self.image_size = tuple(int(i) for i in image_size)
        self.deterministic = bool(deterministic)

    # -- upstream :45-74 ------------------------------------------------------
    def create_iwe(self, events: torch.Tensor, method: str = "bilinear_vote", sigma: int = 1,
                   weight=1.0) -> torch.Tensor:
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
            pos_flag = events[..., 3] > 0
            if events.dim() != 2:
                raise NotImplementedError("method='polarity' needs un-batched events")
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
        nb, m, c = ev.shape
        out = torch.empty((nb, h, w), dtype=torch.int64, device=ev.device)
        lib = cabi.load()
        cabi.check(lib.cmax_count_image(cabi.ptr(ev), nb, m, c, h, w, cabi.ptr(out),
                                        cabi.stream_ptr(ev.device)), "cmax_count_image")
        return out.squeeze()

    # ------------------------------------------------------------------------
    @staticmethod
    def _prep(events: torch.Tensor) -> torch.Tensor:
        if not events.is_cuda:
            raise RuntimeError("EventImageConverter (B200) needs CUDA tensors; there is no CPU path")
        if events.dim() == 2:
            events = events[None]
        if events.dim() != 3 or events.shape[-1] < 2:
            raise ValueError("events must be [(b,) n_events, >=2]")
        return events.detach().to(torch.float32).contiguous()

    def _vote(self, events: torch.Tensor, weight, sigma: float) -> torch.Tensor:
        ev = self._prep(events)
        h, w = self.image_size
        nb, m, c = ev.shape
        wt = None
        if isinstance(weight, torch.Tensor):
            assert weight.shape == events.shape[:-1]
            wt = weight.detach().to(torch.float32).reshape(nb, m).contiguous()
        elif float(weight) != 1.0:
            wt = torch.full((nb, m), float(weight), dtype=torch.float32, device=ev.device)
        out = torch.empty((nb, h, w), dtype=torch.float32, device=ev.device)
        scratch = torch.empty_like(out) if sigma > 0 else None
        scratch64 = (torch.empty((nb, h, w), dtype=torch.int64, device=ev.device)
                     if self.deterministic else None)
        lib = cabi.load()
        rc = lib.cmax_create_iwe(cabi.ptr(ev), cabi.ptr(wt), nb, m, c, h, w, sigma, cabi.ptr(out),
                                 cabi.ptr(scratch), cabi.ptr(scratch64), int(self.deterministic),
                                 cabi.stream_ptr(ev.device))
        cabi.check(rc, "cmax_create_iwe")
        if sigma > 0:
            out = out[:, None]            # upstream blurs a [nb, 1, H, W] view, then squeezes
        return out
