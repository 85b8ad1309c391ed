"""Loader-side packing of the event batch - runs on the CPU, no GPU needed.

The reference's collate (`src/loader/dsec/loader.py:360-415`) stacks the windows of a batch into a
zero-padded `[B, M, 6]` float tensor.  This drop-in collate keeps everything else of the batch and
replaces `batch['events']` by `io.PackedEvents` (16-byte records of the valid events, grouped by
polarity and 32x32-pixel source tile), built in the DataLoader worker by the C++ / OpenMP packer
behind `cmax_pack_events_host`.  `FocusLoss.calc` accepts the packed batch unchanged:

    loader = DataLoader(dataset, batch_size=14, num_workers=8, collate_fn=PackedCollate(loss_config, upstream_collate))
    for batch in loader:
        batch['events'] = batch['events'].to(device, non_blocking=True)
        loss, log, misc = loss_calculator.calc(trajectories, times, batch)

`layout='bitpacked'` (or `'compact'`) hands out the wire layouts instead: a third of the bytes of the
padded tensor cross PCIe, `calc` decodes them on the device (`io.CompactUploader` pipelines the
copies of consecutive batches, see `bench.py`'s end-to-end leg).

    python examples/packed_collate.py        # self-check on synthetic windows
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motionpriorcmax_b200 import cabi, io as cio, synthetic      # noqa: E402


class PackedCollate:
    """Wraps the upstream collate function; picklable, so it works with worker processes."""

    PACKERS = {"packed": "pack_events_native",       # 16-byte records, ready for the kernels
               "compact": "pack_events_compact",     # 12-byte wire layout, expanded on the device
               "bitpacked": "pack_events_bitpacked"}  # lossless ~8-byte wire layout (least PCIe traffic)

    def __init__(self, loss_config: dict, upstream_collate=None, pin: bool = False, layout: str = "packed"):
        self.loss_config = dict(loss_config)
        self.upstream_collate = upstream_collate
        self.pin = pin
        self.layout = layout
        assert layout in self.PACKERS
        self._cfg = None

    def _config(self):
        if self._cfg is None:            # built lazily inside the worker (ctypes structs do not pickle)
            keys = ("image_shape", "num_tref", "num_bins", "num_knn", "smooth_weight", "lut_superpixel_size",
                    "focus_loss_norm", "dist_norm", "scale_iwe_by_dt", "mask_image_border",
                    "polarity_aware_batching", "interpolation_scheme", "smooth_type")
            self._cfg = cabi.make_config(**{k: self.loss_config[k] for k in keys})
        return self._cfg

    def __call__(self, samples):
        batch = self.upstream_collate(samples) if self.upstream_collate else samples
        pack = getattr(cio, self.PACKERS[self.layout])
        packed = pack(batch["events"], batch.get("num_pos_events"), self._config())
        batch["events"] = packed.pin_memory() if self.pin else packed
        return batch

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_cfg"] = None
        return d


def main():
    cfg = dict(synthetic.DSEC_LOSS_CONFIG)
    H, W = cfg["image_shape"]
    ev, npos = synthetic.make_event_batch(4, [800_000, 1_200_000, 400_000, 1_000_000], H, W,
                                          cfg["num_bins"], True, seed=7)
    collate = PackedCollate(cfg)
    t0 = time.perf_counter()
    batch = collate({"events": ev, "num_pos_events": npos})
    dt = time.perf_counter() - t0
    pk = batch["events"]
    n_valid = int(ev[..., 5].sum())
    assert int(pk.num_events().sum()) + int(pk.skipped[0]) == n_valid
    # round trip: the packed batch holds the same events (order inside a segment aside)
    back, npos2 = cio.unpack_events(pk, collate._config())
    again = cio.pack_events_host(back, npos2, collate._config())
    assert torch.equal(again.seg_start, pk.seg_start)
    for b, c in enumerate(pk.num_events().tolist()):
        assert torch.equal(again.records[b, :c].view(torch.int32), pk.records[b, :c].view(torch.int32))
    print(f"packed {n_valid} valid events of a [{ev.shape[0]}, {ev.shape[1]}, 6] batch "
          f"({ev.numel() * 4 / 1e6:.0f} MB) into {pk.records[:, :].numel() * 4 / 1e6:.0f} MB "
          f"+ {pk.seg_start.numel() * 4 / 1e3:.0f} KB of segment offsets in {dt * 1e3:.0f} ms")
    t0 = time.perf_counter()
    bp = PackedCollate(cfg, layout="bitpacked")({"events": ev, "num_pos_events": npos})["events"]
    print(f"bit-packed wire layout: {bp.nbytes() / 1e6:.0f} MB ({bp.nbytes() / n_valid:.1f} B per event) "
          f"in {(time.perf_counter() - t0) * 1e3:.0f} ms")


if __name__ == "__main__":
    main()
