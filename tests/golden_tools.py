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
    return case, ref
