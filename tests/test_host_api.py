"""The drop-in boundary above the C ABI: the host-side mirror of the reference's SPH::Domain source API
(weldform_b200/host/). CPU: the reference's own example drivers compile UNCHANGED against it (where /root/reference
exists), the repo's example driver links against libweldform_b200.so and refuses to run without a GPU. GPU: a driver
with a host GeneralAfter callback reproduces the device-side-BC run bit for bit and the oracle within tolerance, in both
callback-bridge modes."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from tests import cases, oracle_tools as ot

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "weldform_b200", "host")
REF_EXAMPLES = "/root/reference/examples"
DRIVERS = ["Compression.cpp", "Compression_thermal.cpp", "Contact_Compression_1010.cpp", "Plate-Yield_3d.cpp", "FSW.cpp"]


@pytest.fixture(scope="module")
def host_example(product_lib):
    from weldform_b200 import build
    exes = build.build_host_examples()
    exe = [e for e in exes if e.endswith("compression_host_api")]
    assert exe
    return exe[0]


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="/root/reference is only present in the build container")
@pytest.mark.parametrize("driver", DRIVERS)
def test_reference_drivers_compile_unchanged(driver):
    """The five configurations of BASELINE.json, straight from the reference tree, against the host mirror headers."""
    cmd = ["g++", "-std=c++14", "-fsyntax-only", "-fpermissive", "-w", "-fopenmp", "-D_USE_MATH_DEFINES", "-I" + HOST, "-I" + os.path.join(ROOT, "include"),
           os.path.join(REF_EXAMPLES, driver)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="/root/reference is only present in the build container")
def test_wider_driver_set_still_compiles_unchanged():
    """Beyond the five configurations: every driver listed in tests/host_parity/drivers_compiling.txt (97 of the reference's 144 examples; the
    rest need fluids, CompactNSearch, 2-D, the implicit solver, or do not compile against the reference either) keeps compiling unchanged."""
    from concurrent.futures import ThreadPoolExecutor
    names = [l.strip() for l in open(os.path.join(ROOT, "tests", "host_parity", "drivers_compiling.txt")) if l.strip() and not l.startswith("#")]
    assert len(names) >= 97 and all(d in names for d in DRIVERS)

    def compile_one(name):
        cmd = ["g++", "-std=c++14", "-fsyntax-only", "-fpermissive", "-w", "-fopenmp", "-D_USE_MATH_DEFINES", "-I" + HOST, "-I" + os.path.join(ROOT, "include"),
               os.path.join(REF_EXAMPLES, name)]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        return name, r.returncode, r.stderr[-600:]
    with ThreadPoolExecutor(max_workers=8) as ex:
        res = list(ex.map(compile_one, names))
    bad = [(n, e) for n, rc, e in res if rc != 0]
    assert not bad, bad[:3]


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="/root/reference is only present in the build container")
def test_host_generators_and_materials_match_the_reference_bit_for_bit(product_lib):
    """tests/host_parity/gen_dump.cpp compiled against the reference's own headers and against the host mirror: every lattice generator of the
    Domain API (AddCylinderLength with and without ghost rows, AddBoxLength, Add3DCubicBoxParticles + Calculate3DMass, AddTractionProbeLength,
    AddCylUniformLength full and slice, AddXYSymCylinderLength with and without the z mirror rows) gives the same positions, masses and GhostPairs, and the Hollomon / Johnson-Cook host evaluations
    (Material.cpp:6-73) the same values, to the last bit (hex floats)."""
    src = os.path.join(ROOT, "tests", "host_parity", "gen_dump.cpp")
    ref_root = os.path.dirname(REF_EXAMPLES)
    with tempfile.TemporaryDirectory() as d:
        ref_exe, our_exe = os.path.join(d, "gen_ref"), os.path.join(d, "gen_ours")
        subprocess.check_call(["g++", "-std=c++14", "-O1", "-fpermissive", "-w", "-fopenmp", "-ffp-contract=off", "-include", "iomanip", "-D_USE_MATH_DEFINES",
                               "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + os.path.join(ROOT, "oracle"), "-I" + os.path.join(ref_root, "Source"),
                               "-I" + os.path.join(ref_root, "External"), src, os.path.join(ref_root, "External", "LSDynaReader", "src", "lsdynaReader.cpp"), "-o", ref_exe])
        subprocess.check_call(["g++", "-std=c++14", "-O1", "-fpermissive", "-w", "-ffp-contract=off", "-D_USE_MATH_DEFINES", "-I" + HOST, "-I" + os.path.join(ROOT, "include"), src, "-o", our_exe,
                               "-L" + os.path.join(ROOT, "weldform_b200"), "-lweldform_b200", "-Wl,-rpath," + os.path.join(ROOT, "weldform_b200")])
        ref = [l for l in subprocess.run([ref_exe], capture_output=True, text=True, timeout=300).stdout.splitlines() if l.startswith("@ ")]
        our = [l for l in subprocess.run([our_exe], capture_output=True, text=True, timeout=300).stdout.splitlines() if l.startswith("@ ")]
    assert len(ref) > 12000 and sum(l.startswith("@ Add") for l in ref) == 9
    assert len(ref) == len(our)
    bad = [(i, a, b) for i, (a, b) in enumerate(zip(ref, our)) if a != b]
    assert not bad, "first differences (line, reference, mirror): %s" % bad[:5]


def test_example_links_and_refuses_without_gpu(host_example):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with tempfile.TemporaryDirectory() as d:
        r = subprocess.run([host_example, "0.05", "1", "0", os.path.join(d, "o.bin")], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "no CPU fallback" in (r.stdout + r.stderr)


def test_unsupported_solvers_fail_loudly(product_lib):
    src = r'''
#include "Domain.h"
int main() { SPH::Domain d; int hits = 0;
  try { d.Solve(1, 1, 1, "x", 1); } catch (Fatal * f) { ++hits; delete f; }
  try { d.SolveDiffUpdateVerlet(1, 1, 1, "x", 1); } catch (Fatal * f) { ++hits; delete f; }
  try { d.ThermalStructSolve(1, 1, 1, "x", 1); } catch (Fatal * f) { ++hits; delete f; }
  d.AddCylinderLength(1, Vec3_t(0, 0, 0), 0.1, 0.1, 0.01, 1000., 0.012, false);
  const size_t plain = d.Particles.Size();
  SPH::Domain g;                                    // ghost = true: three mirror rows below the bottom plane, GhostPairs (Source/Domain.cpp:795-826)
  g.AddCylinderLength(1, Vec3_t(0, 0, 0), 0.1, 0.1, 0.01, 1000., 0.012, false, true);
  const size_t ng = g.GhostPairs.Size();
  bool ok = ng > 0 && g.Particles.Size() == plain + ng && g.solid_part_count == (int)g.Particles.Size();
  for (size_t k = 0; k < ng && ok; ++k) {
    const SPH::Particle * in = g.Particles[g.GhostPairs[k].first], * gh = g.Particles[g.GhostPairs[k].second];
    ok = gh->not_write_surf_ID && gh->ghost_plane_axis == 2 && gh->x(0) == in->x(0) && gh->x(1) == in->x(1) && gh->x(2) < 0.0 && in->x(2) > 0.0 && gh->Mass == in->Mass;
  }
  return (hits == 3 && ok && plain > 100 && d.solid_part_count == (int)plain) ? 0 : 1; }
'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.cpp"), "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["g++", "-std=c++14", "-w", "-I" + HOST, "-I" + os.path.join(ROOT, "include"), os.path.join(d, "t.cpp"), "-o", exe,
                               "-L" + os.path.join(ROOT, "weldform_b200"), "-lweldform_b200", "-Wl,-rpath," + os.path.join(ROOT, "weldform_b200")])
        assert subprocess.run([exe], capture_output=True).returncode == 0


def _read_bin(path):
    raw = np.fromfile(path, dtype=np.uint8)
    n = int(raw[:8].view(np.int64)[0])
    rec = raw[8:].view(np.float64).reshape(n, 21)
    return dict(x=rec[:, 0:3].reshape(-1), v=rec[:, 3:6].reshape(-1), Displacement=rec[:, 6:9].reshape(-1), pl_strain=rec[:, 9], T=rec[:, 10],
                Density=rec[:, 11], Sigma=rec[:, 12:21].reshape(-1))


@pytest.mark.gpu
@pytest.mark.parametrize("sync,solver", [("full", 0), ("tagged", 0), ("push", 0), ("tagged", 1), ("tagged", 2)])
def test_host_domain_driver_matches_device_run_and_oracle(host_example, oracle_port, sync, solver):
    """The C++ driver (host GeneralAfter callback between the stages of SolveDiffUpdateFraser / KickDrift / LeapFrog) against the
    device-BC run of the same case and against the oracle."""
    from weldform_b200 import harness
    dx, nsteps, amp = 0.03, 40, 60.0
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "o.bin")
        r = subprocess.run([host_example, str(dx), str(nsteps), str(amp), out, str(solver)], capture_output=True, text=True, timeout=600, env=dict(os.environ, WELDFORM_SYNC=sync))
        assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
        got = _read_bin(out)
    case = cases.perturbed(cases.compression(dx=dx, nsteps=nsteps), amp=amp)
    case["solver"] = float(solver)
    dev = harness.run_case(case, with_pairs=False)
    for k in got:
        assert np.array_equal(got[k], dev[k]), "host-callback run differs from the device-BC run in %s (%.3e)" % (k, ot.max_rel(got[k], dev[k]))
    ref = ot.run_port(case)
    assert ref["pl_strain"].max() > 1e-3
    worst = {k: ot.max_rel(got[k], ref[k]) for k in got}
    assert max(worst.values()) < 1e-6, worst


@pytest.mark.gpu
def test_host_domain_thermal_solve_matches_device_run_and_oracle(host_example, oracle_port):
    """Domain::ThermalSolve of the host mirror (wf_thermal + the GeneralAfter callback + wf_step_end per step, Thermal.cpp:627-717)
    against wf_run with device-side BCs and against the oracle port of the reference's ThermalSolve."""
    from weldform_b200 import harness
    dx, nsteps, amp = 0.03, 30, 60.0
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "o.bin")
        r = subprocess.run([host_example, str(dx), str(nsteps), str(amp), out, "3"], capture_output=True, text=True, timeout=600, env=dict(os.environ, WELDFORM_SYNC="tagged"))
        assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
        got = _read_bin(out)
    case = cases.perturbed(cases.compression(dx=dx, nsteps=nsteps), amp=amp)
    T0 = np.array(case["T"], dtype=np.float64)
    case["solver"] = 3.0; case["dt"] = 0.05
    dev = harness.run_case(case, with_pairs=False)
    for k in got:
        assert np.array_equal(got[k], dev[k]), "host-callback run differs from the device-BC run in %s (%.3e)" % (k, ot.max_rel(got[k], dev[k]))
    ref = ot.run_port(case)
    assert np.abs(ref["T"] - T0).max() > 0.05 and not ref["Displacement"].any(), "conduction must act, nothing may move"
    worst = {k: ot.max_rel(got[k], ref[k]) for k in got}
    assert max(worst.values()) < 1e-9, worst


@pytest.mark.gpu
@pytest.mark.parametrize("solver", [0, 2])
def test_host_domain_ghost_rows_match_device_run_and_oracle(host_example, oracle_port, solver):
    """AddCylinderLength(..., ghost = true) through the host mirror: GhostPairs reach the device (wf_set_ghost_pairs), Domain::MoveGhost runs
    inside wf_fraser_stage_b1 / b2 (wf_solver_stage 1..3 for LeapFrog) around the host GeneralAfter callback."""
    from weldform_b200 import harness
    dx, nsteps, amp = 0.03, 40, 60.0
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "o.bin")
        r = subprocess.run([host_example, str(dx), str(nsteps), str(amp), out, str(solver), "1"], capture_output=True, text=True, timeout=600, env=dict(os.environ, WELDFORM_SYNC="tagged"))
        assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
        got = _read_bin(out)
    case = cases.perturbed(cases.compression(dx=dx, nsteps=nsteps, ghost=True), amp=amp)
    case["solver"] = float(solver)
    dev = harness.run_case(case, with_pairs=False)
    for k in got:
        assert np.array_equal(got[k], dev[k]), "host-callback run differs from the device-BC run in %s (%.3e)" % (k, ot.max_rel(got[k], dev[k]))
    ref = ot.run_port(case)
    assert ref["pl_strain"].max() > 1e-3
    worst = {k: ot.max_rel(got[k], ref[k]) for k in got}
    assert max(worst.values()) < 1e-6, worst


@pytest.mark.gpu
@pytest.mark.parametrize("sync", ["full", "tagged"])
def test_callback_that_pins_temperature_reaches_the_device(host_example, sync):
    """A GeneralAfter callback may assign more than v and a: a Dirichlet temperature (T = 100 on the top rows, every call). The bridge uploads
    every member the device accepts when the callback changed it (host/Domain.h sync_up), so the rows stay hot and heat their neighbours;
    before round 2 only v and a went back and the assignment was dropped silently."""
    dx, nsteps, amp = 0.03, 60, 0.0
    runs = {}
    with tempfile.TemporaryDirectory() as d:
        for pin in ("1", "0"):
            out = os.path.join(d, "o%s.bin" % pin)
            r = subprocess.run([host_example, str(dx), str(nsteps), str(amp), out, "0", "0", pin], capture_output=True, text=True, timeout=600, env=dict(os.environ, WELDFORM_SYNC=sync))
            assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
            runs[pin] = _read_bin(out)
    got, free = runs["1"], runs["0"]
    z = got["x"].reshape(-1, 3)[:, 2]
    T = got["T"]
    top = z > 0.56
    assert top.sum() > 50
    assert np.abs(T[top] - 100.0).max() < 1.0, "the pinned rows are at %.2f .. %.2f" % (T[top].min(), T[top].max())
    assert np.abs(free["T"][top] - 100.0).min() > 50.0, "without the assignment the same rows stay near their initial temperature"
    below = (z > 0.56 - 0.07) & ~top
    assert (T[below] - free["T"][below]).max() > 1e-4, "conduction from the pinned rows must heat the rows below them"
