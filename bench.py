#!/usr/bin/env python
"""bench.py -- particle-steps/s of the Fraser SPH step (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--particles P] [--impl ours|reference]

Workload (config.workload): the reference's cylinder-compression lattice (examples/Compression_thermal.cpp:116,
Domain::AddCylinderLength, R = 0.15, L = 0.56, h = 1.2 dx) scaled by dx to P particles (default 1.0e7, the configuration
BASELINE.json's metric is quoted on; it fits one GPU), elastoplastic (bilinear J2), thermal solver + plastic heating on,
artificial viscosity alpha = 1, tensile-instability TI = 0.3, velocity boundary conditions on the top/bottom rows.
All fields start from the reference's uniform state plus a smooth synthetic velocity field (so that every pair term
and the yield branch are populated); the run is spun up (untimed) until the body yields so that the timed steps
include the reference's neighbour-rebuild cadence (every ts_nb_inc = 5 steps).

One "step" = one full Fraser step (SolverFraser.cpp:133-505) over all particles. `value` is device-resident
(wf_fraser_run, boundary conditions applied on the device); `e2e` drives the step phase by phase through the C ABI
with a HOST GeneralAfter callback: every step the tagged particles' v and a are copied device->host into pinned
buffers, the callback runs on the host, the result is copied host->device (twice per step, SolverFraser.cpp:332,360),
and the step scalars are read back.

`--impl reference` times the reference's own CPU solver (oracle/_ref/wf_ref: /root/reference compiled unmodified, or
the oracle port if that binary is absent) on the host cores on a bounded sample of the same lattice.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "particle-steps/sec"
UNIT = "particle-steps/s"
R_CYL, L_CYL = 0.15, 0.56
PERTURB_AMP = 30.0


def dx_for(n_particles):
    vol = np.pi * R_CYL * R_CYL * (L_CYL + 2.0 * L_CYL / 10.0)
    return float((vol / n_particles) ** (1.0 / 3.0))


WORKLOAD_TEXT = {
    "plate": "Plate-Yield_3d box lattice (AddBoxLength 0.033 x 0.01 x 0.01, BASELINE configs[3]), bilinear yield, alpha=beta=1 TI=0.3, ends pulled apart, neighbour rebuild every 5 steps once yielding",
    "fsw": "FSW plate (AddBoxLength 0.055 x 0.003 x 0.055, BASELINE configs[4]): Johnson-Cook, thermal solver + plastic and friction heat, convection on the bottom row, fixed side walls, "
           "rotating + plunging rigid tool (penalty contact, friction linear in T)",
}


def workload_config(n_sph, extra=None, workload="compression"):
    cfg = {"workload": WORKLOAD_TEXT.get(workload) or (("Compression_thermal cylinder between two rigid planes (penalty contact), lattice" if extra and extra.get("contact") else "Compression_thermal cylinder lattice") + " (AddCylinderLength R=0.15 L=0.56 h=1.2dx), elastoplastic + thermal + plastic heating, "
                       "alpha=1 TI=0.3, yielding state, neighbour rebuild every 5 steps"),
           "particles": int(n_sph), "neighbours_per_particle": 56, "solver": "SolveDiffUpdateFraser", "ts_nb_inc": 5,
           "l2": "inputs larger than L2 (>= 1 kB of fp64 state per particle)"}
    if extra:
        cfg.update(extra)
    return cfg


# ---------------------------------------------------------------------------------------------------------------------
# algorithmic bytes per particle per launch of each kernel family: every own array read or written
# once + the neighbour-index list (4 B per neighbour) -- the unique-byte HBM lower bound of the gather formulation.
# ---------------------------------------------------------------------------------------------------------------------
def algorithmic_bytes(nb_avg, ti=True, thermal=True):
    lst = 4.0 * nb_avg + 8.0
    b = {
        # per-particle prologue when the previous step's k_post prepared the density rate (the usual case)
        "density": 8.0 + 32.0 + 48.0 + 8.0 + 32.0 + 32.0 + 32.0 + 32.0,
        # Q0,Q1,Q2 + list; ShearStress, Sigmay, pl_strain, Strain, Strain_pl in; everything the constitutive update writes out
        "rate_stress": 96.0 + lst + 20 * 8.0 + (32.0 + 32.0 + 8.0 + 44 * 8.0),
        "accel": 192.0 + 4.0 + lst + 24.0,
        # k_thermal_pre + k_post (conduction + kinetic/internal energy + next density rate + next TIR)
        "post": (8.0 + 48.0 + 48.0 + 8.0 + 64.0 + 8.0) + (96.0 + lst + 24.0 + 48.0 + 8.0 + 96.0 + 16.0 + 8.0 + 48.0 + 48.0),
        "energy": 96.0 + lst + 96.0 + 16.0,
        "move": 64.0 + 24.0 + 24.0 + 4.0 + 64.0 + 24.0 + 24.0,
    }
    return b




class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [t.strip() for t in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)), "samples": len(sm), "reasons": sorted(reasons)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------------------------------
# reference CPU solver (test infrastructure; executed only here as the baseline, never on the product path)
# ---------------------------------------------------------------------------------------------------------------------
def reference_throughput(n_sample, steps, warmup, max_threads=None, log=None):
    """particle-steps/s of the reference's own OpenMP solver on a sample lattice of ~n_sample particles.
    Time of `steps` steps = T(warmup + steps) - T(warmup) of two whole runs, so setup and the first neighbour build drop out."""
    from tests import cases, oracle_tools as ot
    dx = dx_for(n_sample)
    cores = host_cores()
    if ot.have_ref():
        kind = "reference"
        cand = sorted({c for c in (cores, 64, 32, 16, 8) if c <= cores and (max_threads is None or c <= max_threads)}, reverse=True) or [1]

        def run(nsteps, nproc):
            case = cases.compression_fast(dx=dx, nsteps=nsteps, perturb_amp=PERTURB_AMP)
            out = ot.run_ref(case, nproc=nproc, timeout=3000)
            return float(out["solve_seconds"][0]), int(out["solid_part_count"][0])
        # the reference scales poorly with threads: pick the best thread count on a 2-step probe
        best = None
        for c in cand[:4]:
            t, n = run(2, c)
            if log:
                log("reference probe: Nproc=%d  2 steps + setup %.2f s" % (c, t))
            if best is None or t < best[0]:
                best = (t, c)
        nproc = best[1]
    else:
        kind = "port"
        nproc = 1
        if not ot.have_port():
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"], stdout=subprocess.DEVNULL)

        def run(nsteps, nproc):
            case = cases.compression_fast(dx=dx, nsteps=nsteps, perturb_amp=PERTURB_AMP)
            t0 = time.time()
            out = ot.run_port(case)
            return time.time() - t0, int(out["solid_part_count"][0])
    t_a, n = run(max(warmup, 1), nproc)
    t_b, n = run(max(warmup, 1) + steps, nproc)
    dt = max(t_b - t_a, 1e-9)
    return {"value": n * steps / dt, "unit": UNIT, "cores": nproc, "kind": kind, "seconds": dt, "steps": steps, "particles": n,
            "sample": "%d-particle cylinder of the same lattice (dx=%.4g), %d Fraser steps after %d warm-up steps, Nproc=%d of %d host cores"
                      % (n, dx, steps, max(warmup, 1), nproc, cores)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = 4.0e6          # particle-steps over both runs: a few minutes of host time at ~1e5 p-steps/s
    n_sample = int(min(100000, max(20000, budget / (args.steps + 2 * max(args.warmup, 1)))))
    r = reference_throughput(n_sample, args.steps, args.warmup, log=lambda m: print(m, file=sys.stderr, flush=True))
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * r["seconds"] / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(r["particles"], {"sample_of": "the %d-particle workload of the GPU arm: the reference is timed on a %d-particle lattice of the same generator, "
                                                                                          "physics and spin-free state (particle-steps/s normalises the size)" % (args.particles, r["particles"])}),
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# e2e through the reference-style C++ API (weldform_b200/host/Domain.h), one process per mode
# ---------------------------------------------------------------------------------------------------------------------
def domain_api_e2e(n_particles, steps, warmup, device, log, gpus=1, modes=("push", "tagged", "full")):
    from weldform_b200 import build as wfbuild
    exes = wfbuild.build_host_examples()
    exe = [e for e in exes if e.endswith("bench_domain")][0]
    threads = max(1, min(32, host_cores()))
    out = {}
    # bytes moved per step by the bridge (two callbacks per step), from the mirror's own rules (host/Domain.h): full = every particle,
    # tagged = the rows with a BC ID; down: x v a Displacement (3) Sigma (9) T pl_strain Density Pressure Sigmay (1) ID (int); up: v a
    for mode, (k, w) in (("push", (steps, warmup)), ("push_indexed", (steps, warmup)), ("tagged", (max(3, steps // 4), 2)), ("full", (3, 1))):
        if mode not in modes and not (mode == "push_indexed" and "full" in modes):
            continue
        indexed = mode == "push_indexed"          # the same driver with a callback that visits its rows through index lists instead of testing every particle's ID
        if indexed:
            mode = "push"
        env = dict(os.environ, WELDFORM_SYNC=mode, WELDFORM_DEVICE=str(device), WELDFORM_GPUS=str(gpus), OMP_NUM_THREADS=str(threads))
        for k_ in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k_, None)
        t0 = time.time()
        try:
            # full mode moves every particle's mirrored members both ways at both callbacks (seconds per step): it is timed from step 3 on
            # instead of after the 20+ steps of spin-up to the yielding state (the bridge's traffic does not depend on the state)
            r = subprocess.run([exe, str(int(n_particles)), str(k), str(w), str(PERTURB_AMP), str(threads), "0" if mode == "full" else "1", "1" if indexed else "0"], capture_output=True, text=True, timeout=1500, env=env)
        except subprocess.TimeoutExpired:
            out[mode] = {"error": "timeout"}
            continue
        line = [l for l in r.stdout.splitlines() if l.startswith("BENCH ")]
        if indexed:
            mode = "push_indexed"
        if r.returncode != 0 or not line:
            out[mode] = {"error": (r.stdout + r.stderr)[-400:]}
            continue
        kv = dict(t.split("=") for t in line[0].split()[1:])
        n, sec = int(kv["particles"]), float(kv["seconds"])
        ntag = None
        for l in r.stdout.splitlines():
            if l.startswith("BRIDGE "):
                ntag = int(l.split("tagged=")[1].split()[0])
        rows = n if mode == "full" else (ntag or 0)
        down = 0 if mode.startswith("push") else 2 * rows * (4 * 24 + 72 + 5 * 8 + 4)
        up = (2 * 2 * 2 * 24) if mode.startswith("push") else 2 * rows * 48
        out[mode] = {"value": n * k / sec, "ms_per_step": 1e3 * sec / k, "steps": k, "threads": threads, "gpus": gpus, "h2d_bytes_per_step": int(up), "d2h_bytes_per_step": int(down + 3 * 200),
                     "first_timed_step": int(kv["first_step"]), "max_pl_strain": float(kv["max_pl"]), "wall_s_incl_setup": time.time() - t0}
        log("e2e Domain API, WELDFORM_SYNC=%s: %.2f ms/step (%d steps), %.3g particle-steps/s" % (mode, out[mode]["ms_per_step"], k, out[mode]["value"]))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; weldform_b200 has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from weldform_b200 import build as wfbuild, harness, api
    from tests import cases
    wfbuild.build()

    log = (lambda m: print("[bench r%d] %s" % (rank, m), file=sys.stderr, flush=True))
    # ---- setup (untimed): lattice, upload, solver init, spin-up to the yielding state
    n_target = args.particles
    dx = dx_for(n_target)
    t0 = time.time()
    contact = args.workload == "contact"
    if args.workload == "plate":
        # BASELINE configs[3]: examples/Plate-Yield_3d.cpp:60-119 scaled by dx (AddBoxLength plate 0.033 x 0.01 x 0.01, ends pulled apart, bilinear
        # yield; slabs along x on several GPUs). The shipped Quintic_Spline with h = 1.3 dx gives 250 neighbours per particle, beyond the
        # reference's own MAX_NB_PER_PART: --plate-kernel 0 (cubic spline, h = 1.2 dx, the default) or 2 (quintic, h = 0.8 dx)
        dx = float((0.033 * 0.01 * 0.01 / n_target) ** (1.0 / 3.0))
        case = cases.plate_yield(dx=dx, nsteps=1, kernel=args.plate_kernel, hfac=(1.2 if args.plate_kernel == 0 else 0.8), amp=1.0)
    elif args.workload == "fsw":
        # BASELINE configs[4]: examples/FSW.cpp:60-305 scaled by dx (AddBoxLength plate 0.055 x 0.003 x 0.055, Johnson-Cook, thermal solver, convection
        # on the bottom row, fixed side walls, friction linear in T with friction heat) with a rotating + plunging rigid tool; here the tool is an
        # AxisPlaneMesh disc-sized plane (tests/cases.fsw_like), the NASTRAN tool of the driver itself runs in tests/test_drivers_gpu.py
        dx = float((0.055 * 0.003 * 0.055 / n_target) ** (1.0 / 3.0))
        case = cases.fsw_like(dx=dx, nsteps=1)
    else:
        case = cases.compression_fast(dx=dx, nsteps=1, perturb_amp=PERTURB_AMP, contact_planes=contact)
    case["tuning"] = float(args.tuning)
    comm = None
    if world > 1:
        # the library owns its NCCL communicator; torch.distributed only carries the 128-byte id, the barriers and the max-over-ranks
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(api.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, src=0)
        comm = (rank, world, bytes(uid.cpu().numpy().tobytes()))
    ctx, info = harness.context_from_case(case, device=local, comm=comm)
    n_sph = info["n_sph_global"]            # the whole domain: `value` is the whole-job throughput
    n_local0 = ctx.n_sph
    ctx.solver_init(float(case["dt"]))
    log("setup: %d particles (%d on this rank), dx=%.5g, dt=%.3g, %.1f s" % (n_sph, n_local0, dx, case["dt"], time.time() - t0))
    spin = 0
    t0 = time.time()
    while spin < args.max_spinup:
        ctx.run(10); spin += 10
        s = ctx.reduce()
        if s.is_yielding and spin >= 20:
            break
    log("spin-up: %d steps, %.1f s, max pl_strain %.3g, yielding=%d, rebuilds=%d" % (spin, time.time() - t0, s.max_pl_strain, s.is_yielding, s.nb_rebuilds))
    nb_avg = float(ctx.download("Nb").mean())

    ext = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timed region. The clock sampler (nvidia-smi every 100 ms) starts with the warm-up steps -- the same load --
    # so that a short multi-GPU timed region (20 steps of a few ms) still has samples taken under load.
    clocks = ClockSampler(local); clocks.start()
    ctx.run(max(args.warmup, 3))
    ctx.sync()
    for _ in range(40):                              # the same 80 steps at every N: >= 0.3 s of load before the timed region at small per-GPU sizes, and
        ctx.run(2)                                   # the SAME state (step count) under the clock and in config.state_checksum whatever N is
    ctx.sync()
    ctx.prof_enable(True)
    l0 = ctx.launch_count()
    barrier()
    ev0.record(ext)
    ctx.run(args.steps)
    ev1.record(ext)
    ev1.synchronize()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    launches = ctx.launch_count() - l0
    prof = ctx.prof_get()
    ctx.prof_enable(False)
    s_after = ctx.reduce()
    # N-independence of the state, visible in the bench line itself: order-independent checksums (sums of the 32-bit halves of the fp64 bit
    # patterns, exact in int64 and independent of how particles are split over ranks) of x, pl_strain and T after spin-up + warm-up + timed steps
    def bit_checksum(field):
        a = np.ascontiguousarray(ctx.download(field)).view(np.uint64).reshape(-1)
        lo, hi = int((a & np.uint64(0xFFFFFFFF)).sum(dtype=np.uint64)), int((a >> np.uint64(32)).sum(dtype=np.uint64))
        t = torch.tensor([lo, hi], device="cuda", dtype=torch.int64)
        if world > 1:
            dist.all_reduce(t)
        return "%016x%016x" % (int(t[1].item()) & (2 ** 64 - 1), int(t[0].item()) & (2 ** 64 - 1))
    checksum = {"steps": int(s_after.steps), "max_pl_strain_hex": float(s_after.max_pl_strain).hex(),
                "x": bit_checksum("x"), "pl_strain": bit_checksum("pl_strain"), "T": bit_checksum("T")}
    if world == 1:
        tir = np.asarray(ctx.download("TIR")).reshape(-1, 9)[:ctx.n_sph]
        checksum["tir_nonzero_fraction"] = float(np.count_nonzero(np.abs(tir).max(axis=1) > 0.0)) / max(len(tir), 1)     # particles with a tensile principal stress
        del tir
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = n_sph * args.steps / (ms * 1e-3)

    # ---- e2e (N = 1, compression workload): the reference-style C++ API. examples/bench_domain.cpp is a driver written like the reference's
    # (SPH::Domain::AddCylinderLength, per-particle set-up loop, a GeneralAfter callback looping over Particles[i]->ID twice per step,
    # SolveDiffUpdateFraser) on the host mirror; it is run as its own process on the same GPU in the three callback-bridge modes.
    e2e_domain = None
    if world == 1 and args.workload == "compression" and not args.python_e2e:
        e2e_domain = domain_api_e2e(n_target, args.steps, max(args.warmup, 3), local, log)

    # ---- e2e, Python form (N > 1 and the contact workload): phase API + host GeneralAfter callback bridge with pinned host buffers
    bc_rows = {}
    if contact:
        vtop, tau = np.asarray(case["mesh0_vel"][:3], float).copy(), float(case["mesh0_vel"][3])
    else:
        for r in np.asarray(case["bc"]).reshape(-1, 8):          # (ID, vx, vy, vz, tau, set_a, ...) per boundary-condition row of the driver's callback
            bc_rows[int(r[0])] = (r[1:4].copy(), float(r[4]))
        vtop, tau = bc_rows[3]
    bridge = {"mig": None}

    def refresh_tagged():
        # local indices of the tagged particles (ID != bulk), registered on the device (wf_subset_register); re-derived when particles
        # migrated between ranks at a rebuild
        mig = (ctx.comm_info()["migrated_out"], ctx.comm_info()["migrated_in"]) if world > 1 else (0, 0)
        if bridge["mig"] == mig:
            return
        if bridge.get("handle") is not None:
            ctx.subset_release(bridge["handle"])
        ids = ctx.download("ID")
        # tagged rows grouped by ID so that the callback writes contiguous slices of the pinned buffers (ID 2 rows, then ID 3 rows)
        i2 = np.nonzero(ids == 2)[0]; i3 = np.nonzero(ids == 3)[0]
        tg = np.concatenate([i2, i3]).astype(np.int32)
        # two sets of pinned buffers, one per callback point of the step: the uploads are asynchronous and a set is rewritten only
        # after the blocking wf_reduce at the start of the next step
        mk = lambda: torch.zeros((max(len(tg), 1), 3), dtype=torch.float64).pin_memory().numpy()[:len(tg)]
        bridge.update(mig=mig, tagged=tg, n2=len(i2), handle=ctx.subset_register(tg), sets=[(mk(), mk()), (mk(), mk())], point=0)

    def general_after_contact(t_now):
        # examples/Contact_Compression_1010.cpp:18-47: the tool velocity is set from the host (TriMesh::SetVel) and the callback
        # integrates Sigma(2,2) and a(2)*Mass over the first particle row (read back from the device)
        if bridge["mig"] is None:
            zs = np.asarray(ctx.download("x")).reshape(-1, 3)[:ctx.n_sph, 2] if world == 1 else np.zeros(1)
            nrow = int(np.count_nonzero(zs <= zs.min() + 1e-9)) if world == 1 else 0
            row = np.arange(max(nrow, 1), dtype=np.int32)
            bridge.update(mig=(0, 0), tagged=row, n2=0,
                          v=torch.empty((len(row), 9), dtype=torch.float64).pin_memory().numpy(),
                          a=torch.empty((len(row), 3), dtype=torch.float64).pin_memory().numpy())
        f = min(t_now / tau, 1.0) if tau > 0 else 1.0
        ctx.mesh_set_motion(0, vtop * f)
        if world == 1:
            ctx.download("Sigma", bridge["tagged"], out=bridge["v"]); ctx.download("a", bridge["tagged"], out=bridge["a"])
            bridge["force"] = float(bridge["v"][:, 8].sum())

    def general_after(t_now):
        # examples/Compression_thermal.cpp:30-66: ID 3 rows pushed with a ramped velocity, ID 2 rows held, a = 0 for both. The
        # callback only WRITES v and a of the tagged rows, so the bridge pushes them (H2D from pinned buffers) without reading the
        # device values back first -- the host mirror's write-only mode for callbacks of this kind.
        if contact:
            return general_after_contact(t_now)
        refresh_tagged()
        tagged, n2 = bridge["tagged"], bridge["n2"]
        if len(tagged) == 0:
            return
        v_host, a_host = bridge["sets"][bridge["point"]]; bridge["point"] ^= 1
        a_host[:] = 0.0
        for (vv, tt), sl in ((bc_rows[2], slice(0, n2)), (bc_rows[3], slice(n2, None))):
            v_host[sl] = vv * (min(t_now / tt, 1.0) if tt > 0 else 1.0)
        ctx.upload_registered("v", bridge["handle"], v_host, wait=False); ctx.upload_registered("a", bridge["handle"], a_host, wait=False)

    def e2e_step():
        sc = ctx.reduce()                       # scalars the solver loop reads each step (max pl_strain, max_disp): D2H; blocking, frees the pinned sets
        bridge["point"] = 0
        ctx.phase("fraser_stage_a")
        general_after(sc.time)
        ctx.phase("fraser_stage_b1"); ctx.phase("fraser_stage_b2")
        general_after(sc.time)
        ctx.phase("fraser_stage_c")

    for _ in range(max(args.warmup, 1)):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(args.steps):
        e2e_step()
    sc = ctx.reduce()
    e1.record(ext)
    e1.synchronize()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    scal_bytes = 200
    nt = len(bridge["tagged"])
    if world > 1:
        t = torch.tensor([nt], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        nt = int(t.item())
    h2d = 2 * (2 * nt * 24)                         # v and a of the tagged rows, twice per step (all ranks); the index list is registered once
    d2h = scal_bytes * world                        # the step's scalars (max pl_strain, max_disp, energies, time), every step
    path = "wf_fraser_stage_a/b1/b2/c + wf_upload_registered (host GeneralAfter writes v, a of %d tagged particles, pinned buffers) + wf_reduce" % nt
    if contact:
        h2d = 2 * (48 + nt * 4 * 2) * world; d2h = 2 * (nt * 96) + scal_bytes * world
        path = "wf_fraser_stage_a/b1/b2/c + wf_mesh_set_motion + wf_download_subset (host callback of Contact_Compression_1010: tool velocity, Sigma and a of the first row of %d particles) + wf_reduce" % nt
    e2e = {"value": n_sph * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": ms_e2e / args.steps, "path": path}
    # ---- roofline of the dominant kernel, from the CUDA-event pairs recorded on the ctx stream inside the timed region
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak_hbm = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak_hbm = 6650.0; peak_src = "fallback (B200_PROFILING.md)"
    alg = algorithmic_bytes(nb_avg)
    # measured FP64 operation counts per launch (ncu smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on of the 10M workload,
    # profiles/fp64_ops.json written by scripts/ncu_ops.py): flop = dadd + dmul + 2 dfma, per particle
    fp64_ops = {}
    op = os.path.join(ROOT, "profiles", "fp64_ops.json")
    if os.path.exists(op):
        for kname, rec in json.load(open(op)).items():
            if rec.get("particles"):
                fp64_ops[kname] = {"flop_per_particle": (rec["dadd"] + rec["dmul"] + 2.0 * rec["dfma"]) / rec["particles"], "pipe_fp64_pct": rec.get("pipe_fp64_pct")}
    n_loc_now = ctx.n_sph                    # particles this rank's kernels process (== the domain on one GPU)
    total_prof = sum(v[0] for v in prof.values()) or 1.0
    kernels = {}
    for name, (kms, kn) in prof.items():
        if kn == 0:
            continue
        ent = {"ms_per_launch": kms / kn, "launches": int(kn), "share_of_step": kms / total_prof}
        if name in alg and alg[name] > 0:
            # one PASS over this rank's particles = every launch of the family in one step (on several GPUs a pass is an interior launch + a
            # boundary launch): algorithmic bytes of the rank / time of the pass
            per_pass = kms / args.steps
            ent["ms_per_pass"] = per_pass
            ent["alg_bytes_per_particle"] = alg[name]
            ent["hbm_gbs"] = alg[name] * n_loc_now / (per_pass * 1e-3) / 1e9
            ent["hbm_frac"] = ent["hbm_gbs"] / peak_hbm
            if name in fp64_ops:
                ent["fp64_flop_per_particle_measured"] = fp64_ops[name]["flop_per_particle"]
                ent["fp64_tflops"] = fp64_ops[name]["flop_per_particle"] * n_loc_now / (per_pass * 1e-3) / 1e12
        kernels[name] = ent
    pair_names = [k for k in ("accel", "rate_stress", "post", "energy") if k in kernels]
    dom = max(pair_names, key=lambda k: kernels[k]["share_of_step"]) if pair_names else None
    roofline = None
    if dom:
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp)).get(dom)
            if tj and abs(tj.get("particles", 0) - n_loc_now) <= 0.01 * n_loc_now:
                traffic = tj["dram_bytes_per_launch"]
        k = kernels[dom]
        roofline = {"kernel": dom, "bound": "hbm", "achieved": k["hbm_gbs"], "peak": peak_hbm, "unit": "GB/s", "frac": k["hbm_frac"], "traffic": traffic,
                    "peak_source": peak_src, "alg_bytes_per_launch": k["alg_bytes_per_particle"] * n_loc_now, "ms_per_launch": k["ms_per_pass"],
                    "launches_per_pass": k["launches"] / args.steps, "share_of_step": k["share_of_step"],
                    "note": "fp64 gather pair loop: the FP64 pipe, not HBM, is the binding roof (see fp64); HBM fraction is on unique algorithmic bytes"}
        try:
            fp64_peak = ctx.measure_fp64_peak()
            roofline["fp64"] = {"achieved_tflops": k.get("fp64_tflops"), "peak_tflops_measured_dfma": fp64_peak,
                                "frac": (k.get("fp64_tflops") / fp64_peak) if (fp64_peak and k.get("fp64_tflops")) else None,
                                "flop_per_particle_measured": k.get("fp64_flop_per_particle_measured"),
                                "ncu_pipe_fp64_pct": fp64_ops.get(dom, {}).get("pipe_fp64_pct"),
                                "source": "profiles/fp64_ops.json (ncu op counters of this kernel build); flop = dadd + dmul + 2*dfma executed"}
        except Exception as exc:       # noqa: BLE001
            roofline["fp64"] = {"error": str(exc)}

    # ---- CPU baseline (rank 0, N = 1 only): the reference solver on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = reference_throughput(args.cpu_sample, args.cpu_steps, 2, log=log)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
        except Exception as exc:       # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "failed: %s" % exc}

    comm_info = ctx.comm_info() if world > 1 else None
    ctx.close()
    if world > 1 and args.workload == "compression" and not args.python_e2e:
        # several GPUs: the same C++ driver, ONE process with one host thread and one wf_ctx per GPU (WELDFORM_GPUS, host/Backend.h), started by
        # rank 0 once every rank of this job has released its GPU; the other ranks wait at the barrier
        # (a CPU-side gloo barrier: an NCCL barrier would keep a spinning kernel on every waiting rank's GPU -- the GPUs the driver is being timed on)
        barrier()
        cpu_group = dist.new_group(backend="gloo")
        if rank == 0:
            e2e_domain = domain_api_e2e(n_target, args.steps, max(args.warmup, 3), 0, log, gpus=world, modes=("push", "push_indexed"))
        dist.barrier(group=cpu_group)
    if e2e_domain and e2e_domain.get("push") and "value" in e2e_domain["push"]:
        py = e2e
        hd = e2e_domain["push"]
        e2e = {"value": hd["value"], "unit": UNIT, "h2d_bytes_per_step": hd["h2d_bytes_per_step"], "d2h_bytes_per_step": hd["d2h_bytes_per_step"], "ms_per_step": hd["ms_per_step"],
               "path": "C++ SPH::Domain::SolveDiffUpdateFraser of the host mirror (examples/bench_domain.cpp: GeneralAfter loops over Particles[i]->ID on %d host threads, twice per step), "
                       "WELDFORM_SYNC=push (write-only bridge; uniform v / a of an ID group travel as one vector + a fill kernel; the callback overlaps the GPU's work), WELDFORM_GPUS=%d, host clock between two wf_sync" % (hd["threads"], hd.get("gpus", 1)),
               "modes": e2e_domain, "python_ctypes_push": {"value": py["value"], "ms_per_step": py["ms_per_step"], "h2d_bytes_per_step": py["h2d_bytes_per_step"]}}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(n_sph, workload=args.workload, extra={"decomposition": ("1 GPU" if world == 1 else "%d slabs along the longest axis of the cell grid (%s), NCCL halos (3 refreshes per step), rank 0: %s" % (world, {"plate": "x", "fsw": "x"}.get(args.workload, "z"), comm_info)),
                                                  "neighbours_avg": nb_avg, "spinup_steps": spin, "nb_rebuilds_in_timed_region": int(prof.get("nb_list", (0, 0))[1]),
                                                  "max_pl_strain": s_after.max_pl_strain, "state_checksum": checksum,
                                                  "yielding": bool(s_after.is_yielding),
                                                  **({"contact": "two rigid AxisPlaneMesh planes (%d elements each), Wang penalty contact, friction 0.2; contact force sum %.4g N" % (info["meshes"][0][1], s_after.contact_force_sum)} if contact else {})}),
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "kernels": kernels}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=float, default=None, help="default 1e7 (2e7 for --workload fsw)")
    ap.add_argument("--max-spinup", type=int, default=400)
    ap.add_argument("--cpu-sample", type=int, default=50000, help="particles of the CPU-baseline sample lattice")
    ap.add_argument("--cpu-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", choices=["compression", "contact", "plate", "fsw"], default="compression",
                    help="compression: Compression_thermal (velocity-prescribed end rows, BASELINE configs[1] scaled); contact: the same cylinder between two rigid planes (penalty contact); "
                         "plate: Plate-Yield_3d box lattice (BASELINE configs[3]); fsw: FSW plate with fixed walls, Johnson-Cook, thermal coupling and a rotating + plunging tool (configs[4], default 2e7 particles)")
    ap.add_argument("--plate-kernel", type=int, default=0, choices=[0, 2], help="plate workload: 0 cubic spline (h = 1.2 dx), 2 quintic spline (h = 0.8 dx)")
    ap.add_argument("--tuning", type=int, default=0, help="wf_params::tuning bit mask (kernel A/B switches)")
    ap.add_argument("--python-e2e", action="store_true", help="N = 1: time the e2e leg through the Python ctypes phase API only (skip the C++ Domain driver)")
    args = ap.parse_args()
    args.particles = int(args.particles if args.particles else (2.0e7 if args.workload == "fsw" else 1.0e7))
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
