"""Loader for tests/golden/*.npz (made by tests/golden/make_golden.py from the unmodified reference solver)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(name):
    """-> (case dict, reference dump dict)"""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    case, ref = {}, {}
    for k in z.files:
        (case if k.startswith("case/") else ref)[k.split("/", 1)[1]] = z[k]
    for k, v in list(case.items()):
        if v.ndim == 0:
            case[k] = float(v)
    if "IsFree" in case and "TIR" in ref and not np.all(case["IsFree"]):
        # fixed particles: the reference fills their TIR with a Shepard-smoothed value (PrimaryComputeAcceleration, Domain.cpp:1599-1698)
        # that no pair loop of the new solvers reads; neither the oracle port nor the CUDA path
        # restates it, so those rows are left out of the comparison
        tir = ref["TIR"].reshape(-1, 9).copy()
        fixed = np.nonzero(np.asarray(case["IsFree"]) == 0)[0]
        tir[fixed] = 0.0
        ref["TIR"] = tir.reshape(-1)
    return case, ref
