"""The C-ABI library loads without a GPU and exports every symbol include/cmax_b200.h declares;
host-side validation (geometry, workspace sizing, error codes) is exercised - no kernel runs."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from motionpriorcmax_b200 import build, cabi
    build.build()
    return cabi.load()


def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(ROOT, "include", "cmax_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cmax_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    from motionpriorcmax_b200 import cabi
    lib = _lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(cabi.EXPORTS), declared ^ set(cabi.EXPORTS)
    assert lib.cmax_abi_version() == 1


def test_config_struct_layout_matches_header():
    from motionpriorcmax_b200 import cabi
    assert ctypes.sizeof(cabi.CmaxConfig) == 4 * 18          # 13 int32 + float + 3 int32 + 1 reserved
    assert cabi.CmaxConfig.smooth_weight.offset == 4 * 13
    assert cabi.CmaxConfig.deterministic.offset == 4 * 14


def test_workspace_bytes_and_validation():
    from motionpriorcmax_b200 import cabi, synthetic
    lib = _lib()
    cfg = cabi.make_config(**synthetic.DSEC_LOSS_CONFIG)
    small = lib.cmax_workspace_bytes(cfg, 1, 1000, 19200)
    big = lib.cmax_workspace_bytes(cfg, 14, 1_000_000, 19200)
    assert 0 < small < big < 2 * 1024 ** 3
    det = cabi.make_config(**synthetic.DSEC_LOSS_CONFIG, deterministic=True)
    assert lib.cmax_workspace_bytes(det, 14, 1_000_000, 19200) > big
    # invalid: num_knn > n, forbidden multi-tref combination (focus.py:49-51), zero sizes
    assert lib.cmax_workspace_bytes(cfg, 1, 1000, 16) == 0
    bad = cabi.make_config(**dict(synthetic.DSEC_LOSS_CONFIG, num_tref=3))
    assert lib.cmax_workspace_bytes(bad, 1, 1000, 19200) == 0
    ok = cabi.make_config(**synthetic.multi_tref_variant(synthetic.DSEC_LOSS_CONFIG, 3))
    assert lib.cmax_workspace_bytes(ok, 1, 1000, 19200) > 0
    assert lib.cmax_workspace_bytes(cfg, 0, 1000, 19200) == 0
    assert lib.cmax_knn_workspace_bytes(480, 640, 4, 15, 19200, 32) > 0
    assert lib.cmax_knn_workspace_bytes(480, 640, 4, 15, 10, 32) == 0


def test_error_codes_without_gpu():
    from motionpriorcmax_b200 import cabi, synthetic
    lib = _lib()
    cfg = cabi.make_config(**synthetic.DSEC_LOSS_CONFIG)
    # argument validation happens before any CUDA call
    rc = lib.cmax_forward(cfg, None, None, None, 1, 10, 19200, 5, None, None, None, None, 0, None)
    assert rc == -2 and b"inconsistent" in lib.cmax_error_string(rc)
    bad = cabi.make_config(**dict(synthetic.DSEC_LOSS_CONFIG, num_tref=2))
    rc = lib.cmax_forward(bad, None, None, None, 1, 10, 19200, 5, None, None, None, None, 0, None)
    assert rc == -1 and b"focus.py:49-51" in lib.cmax_error_string(rc)
    big_k = cabi.make_config(**dict(synthetic.DSEC_LOSS_CONFIG, num_knn=500))
    assert lib.cmax_workspace_bytes(big_k, 1, 10, 19200) == 0
    assert lib.cmax_create_iwe(None, None, 1, 10, 1, 4, 4, 0.0, None, None, None, 0, None) == -2
    assert lib.cmax_stage_count() == 13 and lib.cmax_stage_name(1) == b"knn_select"
    with pytest.raises(ValueError):
        cabi.make_config(**dict(synthetic.DSEC_LOSS_CONFIG, focus_loss_norm="l3"))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from motionpriorcmax_b200 import cabi
    monkeypatch.setattr(cabi, "_lib", None)
    monkeypatch.setattr(cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        cabi.load()
