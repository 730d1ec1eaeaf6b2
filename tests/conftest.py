"""pytest configuration: the `gpu` marker, and session fixtures that make sure the two native artefacts exist
(libweldform_b200.so -- the product; oracle/liboracle_port.so -- the checker)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# The parity suite pins BIT-IDENTITY with the reference wherever it can, so it runs the library in exact-arithmetic mode (wf_params::tuning
# bit 10) unless a test asks for the default (contracted arithmetic: tests/test_parity_gpu.py::test_contracted_*, tests/test_drivers_gpu.py).
os.environ.setdefault("WF_TUNING", "1024")          # weldform_b200.harness
os.environ.setdefault("WELDFORM_TUNING", "1024")    # the C++ host mirror (weldform_b200/host/Domain.h)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_port():
    """Build (if needed) and return the path of the CPU restatement library. gcc is present on every box."""
    lib = os.path.join(ROOT, "oracle", "liboracle_port.so")
    src = os.path.join(ROOT, "oracle", "wf_oracle.cpp")
    if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"], stdout=subprocess.DEVNULL)
    return lib


@pytest.fixture(scope="session")
def product_lib():
    """Build (if needed) libweldform_b200.so; nvcc cross-compiles for sm_100a without a GPU."""
    from weldform_b200 import build
    return build.build()
