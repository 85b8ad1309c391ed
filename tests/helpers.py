"""Shared test helpers (loading golden cases, norm-wise comparisons)."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LOSS_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                    if os.path.basename(p) not in ("imager.npz", "voxel.npz", "dense_flow.npz"))


def load_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["cfg"] = json.loads(str(d["cfg"]))
    d["cfg"]["image_shape"] = tuple(d["cfg"]["image_shape"])
    d["num_pos_events"] = int(d["num_pos_events"])
    return d


def rel_err(a, b):
    """Norm-wise relative error ||a - b|| / ||b|| (max-norm guarded)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0))


def max_rel(a, b):
    """max |a-b| / max |b|."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


# (mode, rel. error, sign-flipped pixels, largest flipped |response| / mean |response|) of every
# gradient comparison of the GPU tests; printed by tests/conftest.py at the end of the run
GRAD_CHECKS = []
