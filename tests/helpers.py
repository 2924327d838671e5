"""Shared helpers for the test-suite (oracle access lives only here and in the tests)."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    """Norm-wise relative error ||a-b||_2 / ||b||_2 (the parity metric)."""
    a = np.asarray(a, dtype=np.float64).ravel() if not np.iscomplexobj(a) else np.asarray(a).ravel()
    b = np.asarray(b, dtype=np.float64).ravel() if not np.iscomplexobj(b) else np.asarray(b).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def golden_cases():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    iv = {k[3:]: z[k] for k in z.files if k.startswith("iv_")}
    iv["scan_affine"] = None
    for k in ("N_scan_slow", "N_scan_fast"):
        iv[k] = int(iv[k])
    mp = json.loads(str(z["model_params"]))
    lp = json.loads(str(z["loss_params"]))
    return z, iv, mp, lp
