"""ctypes binding of the C-ABI library (include/cmax_b200.h).

The product path has no CPU fallback: if ``libcmax_b200.so`` is missing or a call returns an
error code, a ``RuntimeError`` is raised.  PyTorch only supplies device memory and the stream.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_size_t, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "_lib", "libcmax_b200.so")

NORM = {"l1": 0, "l2": 1}
INTERP = {"mean": 0, "iwd": 1}
SMOOTH = {"on_flow_to_tref": 0, "on_flow_to_next": 1}
FOCUS = {"gradient_magnitude": 0, "variance": 1}        # upstream src/utils/loss.py:4-12

EXPORTS = (
    "cmax_abi_version", "cmax_error_string", "cmax_workspace_bytes", "cmax_forward",
    "cmax_backward", "cmax_create_iwe", "cmax_count_image", "cmax_create_iwe_padded", "cmax_count_image_padded", "cmax_knn_workspace_bytes",
    "cmax_knn_indices", "cmax_trajectories_forward", "cmax_trajectories_backward",
    "cmax_atomic_microbench", "cmax_read_status", "cmax_stage_count", "cmax_stage_name",
    "cmax_stage_timing_enable", "cmax_stage_timing_read", "cmax_launch_count",
    "cmax_last_worklist_count", "cmax_last_worklist_reasons", "cmax_voxel_grid", "cmax_voxel_normalize", "cmax_dense_flow",
    "cmax_pack_layout", "cmax_pack_events", "cmax_forward_packed", "cmax_backward_packed",
    "cmax_workspace_section", "cmax_forward_accumulate", "cmax_forward_finish",
    "cmax_backward_accumulate", "cmax_backward_finish", "cmax_pack_events_host",
    "cmax_pack_events_host_compact", "cmax_expand_compact",
    "cmax_pack_events_host_bitpacked", "cmax_expand_bitpacked", "cmax_expand_scratch_ints",
)


class CmaxConfig(Structure):
    _fields_ = [
        ("height", c_int32), ("width", c_int32), ("num_tref", c_int32), ("num_bins", c_int32),
        ("num_knn", c_int32), ("lut_superpixel_size", c_int32), ("focus_loss_norm", c_int32),
        ("dist_norm", c_int32), ("scale_iwe_by_dt", c_int32), ("mask_image_border", c_int32),
        ("polarity_aware_batching", c_int32), ("interpolation_scheme", c_int32),
        ("smooth_type", c_int32), ("smooth_weight", c_float), ("deterministic", c_int32),
        ("focus_functional", c_int32), ("backward_follows", c_int32), ("reserved", c_int32 * 1),
    ]


_lib = None


def load():
    """Load the shared library once; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m motionpriorcmax_b200.build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the CMax loss path.")
    lib = ctypes.CDLL(LIB_PATH)
    P = c_void_p
    lib.cmax_abi_version.restype = c_int32
    lib.cmax_error_string.restype = c_char_p
    lib.cmax_error_string.argtypes = [c_int32]
    lib.cmax_workspace_bytes.restype = c_size_t
    lib.cmax_workspace_bytes.argtypes = [POINTER(CmaxConfig), c_int64, c_int64, c_int64]
    lib.cmax_forward.restype = c_int32
    lib.cmax_forward.argtypes = [POINTER(CmaxConfig), P, P, P, c_int64, c_int64, c_int64, c_int64,
                                 P, P, P, P, c_size_t, P]
    lib.cmax_backward.restype = c_int32
    lib.cmax_backward.argtypes = [POINTER(CmaxConfig), P, P, P, c_int64, c_int64, c_int64, c_int64,
                                  P, P, P, c_size_t, P]
    lib.cmax_pack_layout.restype = c_int32
    lib.cmax_pack_layout.argtypes = [POINTER(CmaxConfig), POINTER(c_int32 * 4)]
    lib.cmax_pack_events.restype = c_int32
    lib.cmax_pack_events.argtypes = [POINTER(CmaxConfig), P, c_int64, c_int64, c_int64, P, P, P, P, P]
    lib.cmax_pack_events_host.restype = c_int32
    lib.cmax_pack_events_host.argtypes = [POINTER(CmaxConfig), P, c_int64, c_int64, c_int64, P, c_int64, P, P]
    lib.cmax_pack_events_host_compact.restype = c_int32
    lib.cmax_pack_events_host_compact.argtypes = [POINTER(CmaxConfig), P, c_int64, c_int64, c_int64, P, c_int64,
                                                  P, P, P]
    lib.cmax_pack_events_host_bitpacked.restype = c_int32
    lib.cmax_pack_events_host_bitpacked.argtypes = [POINTER(CmaxConfig), P, c_int64, c_int64, c_int64, P, c_int64,
                                                    P, P, P, P, P]
    lib.cmax_expand_bitpacked.restype = c_int32
    lib.cmax_expand_bitpacked.argtypes = [POINTER(CmaxConfig), P, P, P, P, P, c_int64, c_int64, P, P, P, P]
    lib.cmax_expand_scratch_ints.restype = c_int64
    lib.cmax_expand_scratch_ints.argtypes = [POINTER(CmaxConfig), c_int64, c_int64]
    lib.cmax_expand_compact.restype = c_int32
    lib.cmax_expand_compact.argtypes = [POINTER(CmaxConfig), P, P, P, c_int64, c_int64, P, P, P, P]
    lib.cmax_forward_packed.restype = c_int32
    lib.cmax_forward_packed.argtypes = [POINTER(CmaxConfig), P, P, P, P, c_int64, c_int64, c_int64,
                                        P, P, P, P, c_size_t, P]
    lib.cmax_backward_packed.restype = c_int32
    lib.cmax_backward_packed.argtypes = [POINTER(CmaxConfig), P, P, P, P, c_int64, c_int64, c_int64,
                                         P, P, P, c_size_t, P]
    lib.cmax_workspace_section.restype = c_int32
    lib.cmax_workspace_section.argtypes = [POINTER(CmaxConfig), c_int64, c_int64, c_int64, c_int32,
                                           POINTER(c_size_t), POINTER(c_size_t), POINTER(c_int32)]
    lib.cmax_forward_accumulate.restype = c_int32
    lib.cmax_forward_accumulate.argtypes = [POINTER(CmaxConfig), P, P, P, c_int64, c_int64, c_int64,
                                            c_int64, P, P, c_size_t, P]
    lib.cmax_forward_finish.restype = c_int32
    lib.cmax_forward_finish.argtypes = [POINTER(CmaxConfig), c_int64, c_int64, c_int64, P, P, P,
                                        c_size_t, P]
    lib.cmax_backward_accumulate.restype = c_int32
    lib.cmax_backward_accumulate.argtypes = [POINTER(CmaxConfig), P, P, P, c_int64, c_int64, c_int64,
                                             c_int64, P, c_int32, P, c_size_t, P]
    lib.cmax_backward_finish.restype = c_int32
    lib.cmax_backward_finish.argtypes = [POINTER(CmaxConfig), P, c_int64, c_int64, c_int64, P, P, P,
                                         c_size_t, P]
    lib.cmax_create_iwe.restype = c_int32
    lib.cmax_create_iwe.argtypes = [P, P, c_int64, c_int64, c_int64, c_int32, c_int32, c_float, P,
                                    P, P, c_int32, P]
    lib.cmax_count_image.restype = c_int32
    lib.cmax_count_image.argtypes = [P, c_int64, c_int64, c_int64, c_int32, c_int32, P, P]
    lib.cmax_create_iwe_padded.restype = c_int32
    lib.cmax_create_iwe_padded.argtypes = [P, P, c_int64, c_int64, c_int64, c_int32, c_int32, c_int32, c_int32,
                                           c_float, P, P, P, c_int32, P]
    lib.cmax_count_image_padded.restype = c_int32
    lib.cmax_count_image_padded.argtypes = [P, c_int64, c_int64, c_int64, c_int32, c_int32, c_int32, c_int32, P, P]
    lib.cmax_knn_workspace_bytes.restype = c_size_t
    lib.cmax_knn_workspace_bytes.argtypes = [c_int32, c_int32, c_int32, c_int64, c_int64, c_int32]
    lib.cmax_knn_indices.restype = c_int32
    lib.cmax_knn_indices.argtypes = [P, c_int64, c_int64, c_int32, c_int32, c_int32, c_int32,
                                     c_int32, P, P, P, c_size_t, P]
    lib.cmax_trajectories_forward.restype = c_int32
    lib.cmax_trajectories_forward.argtypes = [P, P, c_int64, c_int64, c_int32, c_int32, c_int32,
                                              c_int32, c_int32, c_int32, c_int32, P, P]
    lib.cmax_trajectories_backward.restype = c_int32
    lib.cmax_trajectories_backward.argtypes = [P, P, c_int64, c_int64, c_int32, c_int32, c_int32,
                                               c_int32, c_int32, c_int32, P, P]
    lib.cmax_voxel_grid.restype = c_int32
    lib.cmax_voxel_grid.argtypes = [P, P, P, P, c_int64, c_int32, c_int32, c_int32, c_int32, P, P, P]
    lib.cmax_voxel_normalize.restype = c_int32
    lib.cmax_voxel_normalize.argtypes = [P, c_int32, c_int32, c_int32, c_int32, P, P, P]
    lib.cmax_dense_flow.restype = c_int32
    lib.cmax_dense_flow.argtypes = [P, P, c_int64, c_int64, c_int32, c_int32, c_int32, c_int32, P, P, P]
    lib.cmax_atomic_microbench.restype = c_int32
    lib.cmax_atomic_microbench.argtypes = [P, c_int64, c_int64, c_int32, P]
    lib.cmax_read_status.restype = c_int32
    lib.cmax_read_status.argtypes = [P, POINTER(c_int64 * 4), P]
    lib.cmax_stage_count.restype = c_int32
    lib.cmax_stage_name.restype = c_char_p
    lib.cmax_stage_name.argtypes = [c_int32]
    lib.cmax_stage_timing_enable.restype = c_int32
    lib.cmax_stage_timing_enable.argtypes = [c_int32]
    lib.cmax_stage_timing_read.restype = c_int32
    lib.cmax_stage_timing_read.argtypes = [P, P]
    lib.cmax_launch_count.restype = c_int64
    lib.cmax_last_worklist_count.restype = c_int64
    lib.cmax_last_worklist_count.argtypes = [P]
    lib.cmax_last_worklist_reasons.restype = c_int32
    lib.cmax_last_worklist_reasons.argtypes = [P, P]
    if lib.cmax_abi_version() != 1:
        raise RuntimeError("libcmax_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().cmax_error_string(rc).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {rc})")


def make_config(image_shape, num_tref, num_bins, num_knn, smooth_weight, lut_superpixel_size,
                focus_loss_norm, dist_norm, scale_iwe_by_dt, mask_image_border,
                polarity_aware_batching, interpolation_scheme, smooth_type,
                deterministic=False, focus_loss_type="gradient_magnitude") -> CmaxConfig:
    for name, table, val in (("focus_loss_norm", NORM, focus_loss_norm),
                             ("dist_norm", NORM, dist_norm),
                             ("interpolation_scheme", INTERP, interpolation_scheme),
                             ("smooth_type", SMOOTH, smooth_type)):
        if val not in table:
            raise ValueError(f"{name}={val!r} not in {sorted(table)}")
    c = CmaxConfig()
    c.height, c.width = int(image_shape[0]), int(image_shape[1])
    c.num_tref, c.num_bins, c.num_knn = int(num_tref), int(num_bins), int(num_knn)
    c.lut_superpixel_size = int(lut_superpixel_size)
    c.focus_loss_norm, c.dist_norm = NORM[focus_loss_norm], NORM[dist_norm]
    c.scale_iwe_by_dt = int(bool(scale_iwe_by_dt))
    c.mask_image_border = int(bool(mask_image_border))
    c.polarity_aware_batching = int(bool(polarity_aware_batching))
    c.interpolation_scheme = INTERP[interpolation_scheme]
    c.smooth_type = SMOOTH[smooth_type]
    c.smooth_weight = float(smooth_weight)
    c.deterministic = int(bool(deterministic))
    if focus_loss_type not in FOCUS:
        raise ValueError(f"focus_loss_type={focus_loss_type!r} not in {sorted(FOCUS)}")
    c.focus_functional = FOCUS[focus_loss_type]
    return c


def with_backward_hint(cfg: CmaxConfig) -> CmaxConfig:
    """Copy of `cfg` with `backward_follows = 1` (the forward fuses the image-stage adjoint)."""
    c = CmaxConfig.from_buffer_copy(cfg)
    c.backward_follows = 1
    return c


def pack_layout(cfg: CmaxConfig):
    """(ct, tiles_y, tiles_x, G) of the packed event layout for this configuration."""
    out = (c_int32 * 4)()
    check(load().cmax_pack_layout(cfg, ctypes.byref(out)), "cmax_pack_layout")
    return int(out[0]), int(out[1]), int(out[2]), int(out[3])


SECTION_RAW_IWE, SECTION_DLUT = 0, 1


def workspace_section(cfg: CmaxConfig, B: int, M: int, n: int, which: int):
    """(byte offset, byte size, is_int64) of a reducible workspace section (event-sharded mode)."""
    off, size, i64 = c_size_t(), c_size_t(), c_int32()
    check(load().cmax_workspace_section(cfg, B, M, n, which, ctypes.byref(off), ctypes.byref(size),
                                        ctypes.byref(i64)), "cmax_workspace_section")
    return int(off.value), int(size.value), bool(i64.value)


def stage_timing_read():
    """{stage name: (total ms, launches)} since stage timing was enabled / last read."""
    lib = load()
    n = lib.cmax_stage_count()
    ms = (ctypes.c_double * n)()
    cnt = (ctypes.c_int64 * n)()
    check(lib.cmax_stage_timing_read(ctypes.cast(ms, c_void_p), ctypes.cast(cnt, c_void_p)),
          "cmax_stage_timing_read")
    return {lib.cmax_stage_name(i).decode(): (ms[i], cnt[i]) for i in range(n)}


def worklist_reasons(stream=None):
    """Inspection: how many LUT cells of the most recent forward left the staged fast path, by reason."""
    out = (ctypes.c_int64 * 16)()
    check(load().cmax_last_worklist_reasons(ctypes.cast(out, c_void_p), stream), "cmax_last_worklist_reasons")
    names = ("total", "window_not_staged", "no_bracket", "fewer_than_k_in_window", "more_than_255_in_window",
             "kth_beyond_window_bound", "boundary_list_full", "previous_bin_bracket_missed", "heap_fallback")
    return {k: int(out[i]) for i, k in enumerate(names)}


def ptr(t) -> c_void_p:
    """Device pointer of a torch tensor (or NULL)."""
    return c_void_p(0 if t is None else t.data_ptr())


def stream_ptr(device) -> c_void_p:
    import torch
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)
