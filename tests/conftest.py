import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR


def pytest_terminal_summary(terminalreporter):
    """How often the l1-norm gradient comparison needed the sign-pattern fallback, and how many
    pixels flipped (tests/test_gpu_parity.py::_assert_grad_close)."""
    try:
        import helpers
    except Exception:
        return
    checks = helpers.GRAD_CHECKS
    if not checks:
        return
    fb = [c for c in checks if c[0] != "strict"]
    terminalreporter.write_line(
        f"gradient comparisons: {len(checks)} total, {len(checks) - len(fb)} strict, {len(fb)} with the "
        f"l1 sign-pattern fallback" + (f" (flipped pixels: max {max(c[2] for c in fb)}, largest flipped "
                                        f"|response| / mean = {max(c[3] for c in fb):.2e})" if fb else ""))
