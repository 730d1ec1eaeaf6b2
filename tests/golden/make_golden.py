"""Generates tests/golden/*.npz from the UNMODIFIED reference solver (oracle/_ref/wf_ref, built from /root/reference by
oracle/Makefile). Run in the build container only:   python tests/golden/make_golden.py

Each fixture holds the case records (prefix "case/") and the reference's state dump (prefix "ref/") of one small run
with Nproc = 1. The reference ships no golden vectors of its own; these are the pin for the
oracle port (tests/test_oracle.py) and, through it and directly, for the CUDA path (tests/test_parity_gpu.py).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import cases, oracle_tools as ot  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def golden_cases():
    """name -> case. Small lattices (hundreds of particles) so that the fixtures stay a few hundred kB each."""
    out = {}
    c = cases.perturbed(cases.compression(dx=0.04, nsteps=1)); c["mode"] = 1.0
    out["nb_only"] = c                                                                       # neighbour build: pairs in discovery order
    out["comp_1step"] = cases.perturbed(cases.compression(dx=0.04, nsteps=1), amp=60.0)     # one full Fraser step, every pair term active
    out["comp_yield_40"] = cases.perturbed(cases.compression(dx=0.04, nsteps=40), amp=60.0)  # yields -> rebuild cadence, plastic heat, conduction
    out["comp_noTI_g2_8"] = cases.perturbed(cases.compression(dx=0.04, nsteps=8, ti=0.0, gradient_type=2), amp=20.0)
    out["comp_g1_beta_8"] = cases.perturbed(cases.compression(dx=0.04, nsteps=8, gradient_type=1, beta=1.0), amp=20.0)
    out["comp_hollomon_30"] = cases.perturbed(cases.compression(dx=0.04, nsteps=30, model=cases.HOLLOMON), amp=60.0)
    out["comp_jc_30"] = cases.perturbed(cases.compression(dx=0.04, nsteps=30, model=cases.JOHNSON_COOK), amp=60.0)
    out["comp_gmt_30"] = cases.perturbed(cases.compression(dx=0.04, nsteps=30, model=cases.GMT), amp=60.0)       # GMT hardening law (Material.cpp:96-134)
    out["contact_1step"] = cases.contact_compression(dx=0.0032, nsteps=1, vmax=5.0, tau=1e-5, dens=8, cfl=0.3)
    out["contact_100"] = cases.contact_compression(dx=0.0032, nsteps=100, vmax=5.0, tau=1e-5, dens=8, cfl=0.3)   # friction stick+slide, yields
    # Domain::gradKernelCorr = true (examples/Contact_Compression_1010.cpp:125 as shipped): corrected gradients + locking summation
    def gc(c):
        c = dict(c); c["gradKernelCorr"] = 1.0
        return c
    out["gc_comp_1step"] = gc(cases.perturbed(cases.compression(dx=0.04, nsteps=1), amp=60.0))
    out["gc_comp_yield_40"] = gc(cases.perturbed(cases.compression(dx=0.04, nsteps=40), amp=60.0))      # first-yield search keeps the stale matrices
    out["gc_g2_jc_12"] = gc(cases.perturbed(cases.compression(dx=0.04, nsteps=12, gradient_type=2, model=cases.JOHNSON_COOK), amp=40.0))
    out["gc_contact_60"] = gc(cases.contact_compression(dx=0.0032, nsteps=60, vmax=5.0, tau=1e-5, dens=8, cfl=0.3))
    # SolveDiffUpdateKickDrift / SolveDiffUpdateLeapFrog, fixed and adaptive time step
    def sv(c, solver, **kw):
        c = dict(c); c["solver"] = float(solver); c.update(kw)
        return c
    base = cases.perturbed(cases.compression(dx=0.04, nsteps=30), amp=60.0)
    out["kd_comp_30"] = sv(base, 1)
    out["lf_comp_30"] = sv(base, 2)
    out["kd_auto_ts"] = sv(base, 1, auto_ts=1.0, CFL=0.1, tf=25 * base["dt"])        # CFL 0.1: the velocity criterion shortens the step
    out["lf_auto_ts"] = sv(base, 2, auto_ts=1.0, CFL=0.1, tf=25 * base["dt"])
    out["kd_auto_ts_acc"] = sv(base, 1, auto_ts_acc=1.0, tf=25 * base["dt"])             # CheckMinTSAccel: the acceleration criterion shortens the step
    out["lf_contact_60"] = sv(cases.contact_compression(dx=0.0032, nsteps=60, vmax=5.0, tau=1e-5, dens=8, cfl=0.3), 2)
    out["kd_gc_20"] = sv(gc(cases.perturbed(cases.compression(dx=0.04, nsteps=20), amp=60.0)), 1)
    # AddBoxLength lattices: examples/Plate-Yield_3d.cpp (cubic and quintic spline kernels) and an examples/FSW.cpp-like set-up
    # (Johnson-Cook, thermal, convection BC, fixed side walls, rotating + plunging tool plane, friction linear in T, friction heat)
    out["plate_cubic_40"] = cases.plate_yield(dx=0.0012, nsteps=40, amp=0.3)
    out["plate_quintic_10"] = cases.plate_yield(dx=0.0012, nsteps=10, kernel=2, hfac=0.8, amp=0.3)
    out["fsw_like_80"] = cases.fsw_like(dx=0.0009, nsteps=80, L=0.02, rt=0.004)
    out["gc_fixed_40"] = gc(cases.fsw_like(dx=0.0009, nsteps=40, L=0.02, rt=0.004))      # gradient correction with fixed walls: FSM pairs in CalcGradCorrMatrix
    # Domain::ThermalSolve (heat conduction only, one neighbour search): convection on every seventh particle, a thermal time step
    th = cases.compression(dx=0.04, nsteps=25)
    nth = len(th["T"])
    th.update(solver=3.0, dt=0.05, Thermal_BC=(np.arange(nth) % 7 == 0).astype(np.int32), h_conv=200.0 * np.ones(nth), T_inf=5.0 * np.ones(nth),
              th_exp=1e-5 * np.ones(nth))
    out["thermal_solve_25"] = th
    # the same with corrected gradients (CalcGradCorrMatrix once, Thermal.cpp:607-608; the conduction branch of Thermal.cpp:90-111) and two
    # conductivities / heat capacities, so that pairs of unlike particles use the 4 k_i k_j / (k_i + k_j) mean
    thg = gc(th)
    thg.update(nsteps=20.0, k_T=np.where(np.arange(nth) % 3 == 0, 90.0, 150.0), cp_T=np.where(np.arange(nth) % 5 == 0, 700.0, 960.0))
    out["thermal_solve_gc_20"] = thg
    # AddCylinderLength(..., ghost = true): three mirror rows below the bottom plane driven by Domain::MoveGhost; bottom rows left free
    out["ghost_comp_1step"] = cases.perturbed(cases.compression(dx=0.04, nsteps=1, ghost=True), amp=60.0)
    out["ghost_comp_yield_40"] = cases.perturbed(cases.compression(dx=0.04, nsteps=40, ghost=True), amp=60.0)
    out["ghost_kd_25"] = dict(cases.perturbed(cases.compression(dx=0.04, nsteps=25, ghost=True), amp=60.0), solver=1.0)
    out["ghost_lf_25"] = dict(cases.perturbed(cases.compression(dx=0.04, nsteps=25, ghost=True, model=cases.HOLLOMON), amp=60.0), solver=2.0)
    return out


def main():
    if not ot.have_ref():
        raise SystemExit("oracle/_ref/wf_ref missing: run `make -C oracle ref` in the container that has /root/reference")
    only = set(sys.argv[1:])            # optional: names to (re)generate; default all
    for name, case in golden_cases().items():
        if only and name not in only:
            continue
        ref = ot.run_ref(case, nproc=1)
        ref.pop("solve_seconds", None)
        rec = {"case/" + k: np.asarray(v) for k, v in case.items()}
        rec.update({"ref/" + k: v for k, v in ref.items()})
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **rec)
        print("%-18s N=%d pairs=%d maxpl=%.3g  %.0f kB" % (name, int(ref["n_particles"][0]), ref["pairs"].size // 2, ref["pl_strain"].max(),
                                                        os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
