"""Dense optical-flow read-out on the GPU - mirror of upstream ``src/utils/flow.py:8-16`` and
``list_to_grid`` (``src/utils/trajectories.py:54-75``).  Inference / logging path, forward only."""
from __future__ import annotations

import torch

from .. import cabi


def dense_flow_from_traj(traj_flow: torch.Tensor, pixel_positions: torch.Tensor, patch_size: int,
                         image_shape):
    """traj_flow [b, n, c], pixel_positions [n, 2] (y, x) -> (dense [b, c, h, w], patch [b, c, h/p, w/p])."""
    if not traj_flow.is_cuda:
        raise RuntimeError("dense_flow_from_traj (B200) needs CUDA tensors; there is no CPU path")
    h, w = (int(v) for v in image_shape)
    tf = traj_flow.detach().to(torch.float32).contiguous()
    b, n, c = tf.shape
    pos = pixel_positions.to(device=tf.device, dtype=torch.int64).contiguous()
    patch = torch.empty((b, c, h // patch_size, w // patch_size), dtype=torch.float32, device=tf.device)
    dense = torch.empty((b, c, h, w), dtype=torch.float32, device=tf.device)
    lib = cabi.load()
    rc = lib.cmax_dense_flow(cabi.ptr(tf), cabi.ptr(pos), b, n, c, int(patch_size), h, w, cabi.ptr(patch),
                             cabi.ptr(dense), cabi.stream_ptr(tf.device))
    cabi.check(rc, "cmax_dense_flow")
    return dense, patch


def list_to_grid(feature_list: torch.Tensor, pixel_positions: torch.Tensor, image_shape):
    """[b, n, c] features at integer (y, x) positions -> [b, c, h, w] (zeros elsewhere)."""
    dense, patch = dense_flow_from_traj(feature_list, pixel_positions, 1, image_shape)
    return patch
