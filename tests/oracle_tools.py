"""Helpers to run the two oracle programs on a WFC1 case dict and get the state dump back as numpy arrays.

TEST INFRASTRUCTURE: imported only from tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke(). Nothing in weldform_b200/ imports this.
"""
import ctypes
import os
import subprocess
import tempfile

import numpy as np

from . import wfc  # noqa: F401  (package-relative when imported as tests.oracle_tools)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "wf_ref")
PORT_LIB = os.path.join(ROOT, "oracle", "liboracle_port.so")


def have_ref():
    return os.path.isfile(REF_BIN) and os.access(REF_BIN, os.X_OK)


def have_port():
    return os.path.isfile(PORT_LIB)


def run_ref(case, nproc=1, timeout=3600):
    """Run the unmodified reference solver (oracle/_ref/wf_ref) on `case`; returns the dump dict."""
    with tempfile.TemporaryDirectory() as d:
        c = dict(case)
        c["Nproc"] = float(nproc)
        wfc.write(os.path.join(d, "case.wfc"), c)
        env = dict(os.environ, OMP_NUM_THREADS=str(nproc))
        r = subprocess.run([REF_BIN, "case.wfc", "out.wfc"], cwd=d, capture_output=True, text=True, timeout=timeout, env=env)
        if r.returncode != 0:
            raise RuntimeError("wf_ref failed (%d): %s" % (r.returncode, r.stderr[-2000:]))
        return wfc.read(os.path.join(d, "out.wfc"))


_port = None


def run_port(case):
    """Run the CPU restatement (oracle/liboracle_port.so) on `case`; returns the dump dict."""
    global _port
    if _port is None:
        _port = ctypes.CDLL(PORT_LIB)
        _port.wfo_run_file.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        _port.wfo_run_file.restype = ctypes.c_int
    with tempfile.TemporaryDirectory() as d:
        wfc.write(os.path.join(d, "case.wfc"), case)
        rc = _port.wfo_run_file(os.path.join(d, "case.wfc").encode(), os.path.join(d, "out.wfc").encode())
        if rc != 0:
            raise RuntimeError("wfo_run_file failed (%d)" % rc)
        return wfc.read(os.path.join(d, "out.wfc"))


def max_rel(a, b):
    """max |a-b| / max|b| (field inf-norm relative error); 0 for two all-zero fields."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.max(np.abs(b)) if b.size else 0.0
    if scale == 0.0:
        return float(np.max(np.abs(a))) if a.size else 0.0
    return float(np.max(np.abs(a - b)) / scale)


ELEM_FLOOR = 1e-6
# Fields a particle ACCUMULATES (time integrals, monotone or slowly varying state): every entry is meaningful down to a millionth of the
# field's largest. Everything else (a, v, rates, stresses ...) is a sum of signed pair terms that cancel: an entry far below the field's top
# carries the absolute round-off of the terms it was summed from, so such entries are judged against 1 % of the top instead.
STATE_FIELDS = {"x", "Displacement", "pl_strain", "T", "T-T0", "Density", "Sigmay"}
DERIVED_FLOOR = 1e-2


def elem_floor(field):
    return ELEM_FLOOR if field.split("/")[-1] in STATE_FIELDS else DERIVED_FLOOR


def max_rel_elem(a, b, floor=ELEM_FLOOR):
    """ELEMENT-WISE relative error: max_k |a_k - b_k| / max(|b_k|, floor * max|b|). Every entry is judged against its own magnitude; the
    floor (a fraction of the field's largest entry) only keeps entries that are zero up to round-off -- sums that cancel -- from dividing
    by nothing. 0 for two all-zero fields."""
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    if not b.size:
        return 0.0
    top = float(np.max(np.abs(b)))
    if top == 0.0:
        return float(np.max(np.abs(a)))
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor * top)))
