"""CPU checks of the driver goldens (tests/golden/drivers/*.npz): every driver the GPU tests run has a golden made by the committed generator's plan, the
mirror binary list of weldform_b200/build.py covers it, and the goldens are well formed (tables of 17 columns in WriteCSV's order, kept indices ascending,
the alternative reference run of the ill-conditioned drivers has the golden's shape)."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "drivers")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def plan():
    return _load(os.path.join(GOLD, "make_driver_goldens.py"), "make_driver_goldens")


def test_every_tested_driver_has_a_golden_a_plan_and_a_mirror_binary_rule(plan):
    from tests import test_drivers_gpu as t
    from weldform_b200 import build
    tested = set(t.MORE) | {"Compression", "Compression_thermal", "Contact_Compression_1010", "FSW"}
    for name in sorted(tested):
        assert os.path.exists(os.path.join(GOLD, name + ".npz")), "no golden for %s" % name
        assert name in plan.PLAN, "%s is not in make_driver_goldens.PLAN" % name
        assert name in build.REF_DRIVERS, "%s is not built against the host mirror (weldform_b200/build.py REF_DRIVERS)" % name
    assert set(plan.PLAN) == tested
    assert set(plan.ALT) <= set(t.MORE)
    for name, files in t.NEEDS.items():
        assert plan.NEEDS.get(name) == files


def test_goldens_are_well_formed(plan):
    from tests import test_drivers_gpu as t
    for name, (key, wanted, stride) in plan.PLAN.items():
        g = np.load(os.path.join(GOLD, name + ".npz"))
        kept = g["kept"].tolist()
        assert kept == sorted(kept) and kept and set(kept) <= set(wanted), name
        assert int(g["stride"]) == stride, name
        assert [c.strip() for c in str(g["header"]).replace("CZFz", "CFz").split(",")] == t.COLS, name      # "CZFz": the reference's own typo (Output.cpp:527)
        assert g["time_at_table"].shape == (len(kept),) and np.all(np.diff(g["time_at_table"]) > 0), name
        for idx in kept:
            tab = g["table_%04d" % idx]
            rows = int(g["rows_%04d" % idx])
            assert tab.dtype == np.float32 and tab.shape == (len(range(0, rows, stride)), 17), name
            assert np.isfinite(tab[:, :16]).all(), name                 # Temp (column 16) may hold the reference's uninitialised values
            if name in plan.ALT:
                assert g["alt_%04d" % idx].shape == tab.shape, name
