"""Multi-GPU tests. CPU part (world_size 2, gloo): the host-side slab planner is consistent across ranks and partitions
every particle exactly once. GPU part (needs >= 2 GPUs; `gpurun --gpus 2`): the slab-decomposed run with NCCL halos and
migration is BIT-IDENTICAL per particle to the single-GPU run and within the north-star tolerance of the oracle."""
import glob
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from tests import cases, oracle_tools as ot, wfc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from weldform_b200 import api
    c = cases.perturbed(cases.compression(dx=0.02, nsteps=1))
    g = c["gen_args"]
    pts = api.gen_cylinder((g[1], g[2], g[3]), g[4], g[5], g[6])
    h = np.full(len(pts), g[8])
    rank_of, bounds = api.plan_slabs(pts, h, world, 0, c["DomMin"], c["DomMax"])
    mine = np.nonzero(rank_of == rank)[0]
    got = [None] * world
    dist.all_gather_object(got, dict(n=len(mine), first=int(mine[0]), last=int(mine[-1]), bounds=bounds.tolist(), zmin=float(pts[mine, 2].min()), zmax=float(pts[mine, 2].max()),
                                     csum=int(mine.sum())))
    if rank == 0:
        q.put((len(pts), got))
    dist.destroy_process_group()


def test_slab_planner_two_ranks_gloo(product_lib):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    n, got = q.get(timeout=120)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0]["bounds"] == got[1]["bounds"]                         # every rank derives the same plan
    assert got[0]["n"] + got[1]["n"] == n                               # every particle owned exactly once
    assert got[0]["csum"] + got[1]["csum"] == n * (n - 1) // 2
    assert abs(got[0]["n"] - got[1]["n"]) <= 0.25 * n                   # balanced by count
    assert got[0]["zmax"] < got[1]["zmin"]                              # slabs along the long (z) axis of the cylinder


@pytest.mark.parametrize("nranks", [1, 2, 3, 4, 8])
def test_slab_planner_properties(product_lib, nranks):
    from weldform_b200 import api
    pts = api.gen_cylinder((0.0, 0.0, -0.056), 0.15, 0.672, 0.0075)
    rank_of, bounds = api.plan_slabs(pts, np.full(len(pts), 0.018), nranks, 0, (-0.56, 1e11, 1e11), (0.56, -1e11, -1e11))
    assert bounds[0] == 0 and (np.diff(bounds) >= 1).all()
    cnt = np.bincount(rank_of, minlength=nranks)
    assert cnt.sum() == len(pts) and cnt.min() > 0
    assert cnt.max() <= 1.5 * len(pts) / nranks + 1400                  # layer granularity: one z-layer of cells holds ~690 particles here
    # ownership is monotone along z
    order = np.argsort(pts[:, 2], kind="stable")
    assert (np.diff(rank_of[order]) >= 0).all()


def test_rebalance_decision_moves_one_layer_towards_the_slow_rank(product_lib):
    """wf_plan_rebalance (host-only, the same code the ranks run at a rebuild): a rank that takes 20 % longer per particle hands one
    boundary layer to a neighbour; balanced ranks keep their bounds; nobody drops below 3 layers; at most one move per rank."""
    from weldform_b200 import api
    n, layer = 1.0e6, 5.0e4
    def stats(costs):
        return np.array([[c, n, layer, layer, 1.0e5, 0, 0, 0] for c in costs])
    b0 = np.array([0, 20, 40, 60, 80], np.int32)
    assert np.array_equal(api.plan_rebalance(stats([1e-9] * 4), b0), b0)
    b = api.plan_rebalance(stats([1e-9, 1.2e-9, 1e-9, 1e-9]), b0)
    assert (b[2] - b[1]) == 19 and b[0] == 0 and b[4] == 80 and np.abs(b - b0).sum() == 1     # the slow rank gives up exactly one layer
    b = api.plan_rebalance(stats([1.3e-9, 1e-9, 1e-9, 1.3e-9]), b0)
    assert b[1] == 19 and b[3] == 61                                                        # both ends shed a layer inwards
    small = np.array([0, 3, 6, 9, 12], np.int32)
    assert np.array_equal(api.plan_rebalance(stats([2e-9, 1e-9, 1e-9, 1e-9]), small), small)  # < 4 layers: no move
    tiny_cap = stats([2e-9, 1e-9, 1e-9, 1e-9]); tiny_cap[:, 4] = 1.0e4
    assert np.array_equal(api.plan_rebalance(tiny_cap, b0), b0)                                 # the layer would not fit the emigrant list


def test_planner_rejects_more_ranks_than_layers(product_lib):
    from weldform_b200 import api
    pts = np.array([[0.0, 0.0, 0.0], [0.01, 0.0, 0.0], [0.02, 0.0, 0.0]])
    with pytest.raises(api.WeldformError):
        api.plan_slabs(pts, np.full(3, 0.012), 8)


def _run_multi(case, nranks):
    from weldform_b200 import harness
    with tempfile.TemporaryDirectory() as d:
        wfc.write(os.path.join(d, "case.wfc"), case)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks), "--master-addr", "127.0.0.1",
               "--master-port", str(29900 + os.getpid() % 90), os.path.join(ROOT, "tests", "mgpu_worker.py"), os.path.join(d, "case.wfc"), d]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        dumps = []
        for f in sorted(glob.glob(os.path.join(d, "rank*.npz"))):
            z = np.load(f)
            dumps.append({k: z[k] for k in z.files})
        info = [dd.pop("comm_info") for dd in dumps]
        return harness.merge_rank_dumps(dumps), info


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("nranks,amp,nsteps", [(2, 60.0, 60), (2, 400.0, 120)])
def test_multi_gpu_bit_identical_to_single(product_lib, oracle_port, nranks, amp, nsteps):
    """amp = 400 m/s drives particles across the slab boundary within the run (migration at the rebuilds)."""
    from weldform_b200 import harness
    nranks = min(nranks, _ngpus())
    case = cases.perturbed(cases.compression(dx=0.02, nsteps=nsteps), amp=amp)
    multi, info = _run_multi(case, nranks)
    single = harness.run_case(case)
    assert all(i[1] > 0 for i in info), "every rank must hold halo particles"
    if amp > 100.0:
        assert sum(int(i[2]) for i in info) > 0, "the violent case is meant to exercise migration"
    assert np.array_equal(multi["pairs"], single["pairs"])
    bad = {k: ot.max_rel(multi[k], single[k]) for k in cases.FIELDS_F64 + ["ID", "Nb"] if not np.array_equal(multi[k], single[k])}
    assert not bad, "multi-GPU differs from single GPU per particle: %s" % bad
    for k in ("Time", "max_disp"):
        assert np.array_equal(multi[k], single[k]), k
    for k in ("int_energy_sum", "plastic_work"):          # domain totals: different summation tree across ranks
        assert ot.max_rel(multi[k], single[k]) < 1e-12, k
    # kin_energy_sum is a sum of +/- pair terms that cancels to ~1e-12 of their magnitude: compare on the scale of the terms
    scale = float(np.abs(single["dkin_energy_dt"]).sum()) * float(case["dt"]) * nsteps
    assert abs(float(multi["kin_energy_sum"][0]) - float(single["kin_energy_sum"][0])) <= 1e-12 * scale
    ref = ot.run_port(case)
    worst = {k: ot.max_rel(multi[k], ref[k]) for k in ("x", "v", "Displacement", "pl_strain", "T", "Density", "Sigma")}
    assert max(worst.values()) < 1e-6, worst


@pytest.mark.gpu
@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs (gpurun --gpus 2)")
def test_multi_gpu_contact_bit_identical_to_single(product_lib, oracle_port):
    """Rigid-mesh contact on the slab decomposition: every rank keeps a replica of the meshes and of their element
    particles (outside the cell table), contact candidates and forces are computed by the owner of the SPH particle."""
    from weldform_b200 import harness
    case = cases.contact_compression(dx=0.0024, nsteps=150, vmax=5.0, tau=1e-5, dens=10, cfl=0.3)
    multi, info = _run_multi(case, 2)
    single = harness.run_case(case)
    assert all(i[1] > 0 for i in info), "every rank must hold halo particles"
    assert single["cont_pairs"].size > 0 and np.array_equal(multi["cont_pairs"], single["cont_pairs"])
    assert np.array_equal(multi["pairs"], single["pairs"])
    bad = {k: ot.max_rel(multi[k], single[k]) for k in cases.FIELDS_F64 + ["ID", "Nb"] if not np.array_equal(multi[k], single[k])}
    assert not bad, "multi-GPU contact differs from single GPU per particle: %s" % bad
    assert (single["delta_cont"] > 0).sum() > 20
    for k in ("Time", "max_disp", "mesh0_nodes", "mesh1_nodes"):
        if k in single:
            assert np.array_equal(multi[k], single[k]), k
    for k in ("contact_force_sum", "ext_forces_work", "plastic_work"):
        assert ot.max_rel(multi[k], single[k]) < 1e-12, k
    ref = ot.run_port(case)
    worst = {k: ot.max_rel(multi[k], ref[k]) for k in ("x", "v", "Displacement", "pl_strain", "Density", "Sigma", "contforce", "delta_cont")}
    assert max(worst.values()) < 1e-6, worst


@pytest.mark.gpu
@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("solver,auto_ts", [(1, 1), (2, 0)])
def test_multi_gpu_kickdrift_leapfrog_bit_identical_to_single(product_lib, oracle_port, solver, auto_ts):
    """KickDrift / LeapFrog step orders on the slab decomposition (their own halo refresh points; the adaptive time step takes
    the minimum over the ranks)."""
    from weldform_b200 import harness
    case = cases.perturbed(cases.compression(dx=0.02, nsteps=60), amp=200.0)
    case["solver"] = float(solver)
    if auto_ts:
        case.update(auto_ts=1.0, CFL=0.15, tf=50 * case["dt"])
    multi, info = _run_multi(case, 2)
    single = harness.run_case(case)
    assert np.array_equal(multi["pairs"], single["pairs"])
    bad = {k: ot.max_rel(multi[k], single[k]) for k in cases.FIELDS_F64 + ["ID", "Nb"] if not np.array_equal(multi[k], single[k])}
    assert not bad, "multi-GPU differs from single GPU per particle: %s" % bad
    for k in ("Time", "deltat"):
        assert np.array_equal(multi[k], single[k]), k
    if auto_ts:
        assert float(single["deltat"][0]) < 0.9 * case["dt"]
    ref = ot.run_port(case)
    worst = {k: ot.max_rel(multi[k], ref[k]) for k in ("x", "v", "Displacement", "pl_strain", "T", "Density", "Sigma")}
    assert max(worst.values()) < 1e-6, worst


@pytest.mark.gpu
@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs (gpurun --gpus 2)")
def test_multi_gpu_rebalance_moves_a_layer_and_keeps_every_bit(product_lib, oracle_port):
    """Time-based slab rebalancing: wf_params::tuning bit 13 forces one boundary move (rank 0 hands its last cell layer to rank 1
    through the migration path) at the third rebuild; the state must stay bit-identical to the single-GPU run."""
    from weldform_b200 import harness
    case = cases.perturbed(cases.compression(dx=0.02, nsteps=120), amp=400.0)
    single = harness.run_case(case)
    forced = dict(case); forced["tuning"] = 8192.0 + 1024.0                     # + bit 10: exact arithmetic, like the single-GPU run it is compared with
    multi, info = _run_multi(forced, 2)
    plain, info0 = _run_multi(dict(case, tuning=16384.0 + 1024.0), 2)          # rebalancing off: the planner's bounds stay
    assert int(info[0][5]) == int(info0[0][5]) - 1 or int(info[0][5]) != int(info0[0][5]), "the forced move must change rank 0's upper layer bound"
    assert np.array_equal(multi["pairs"], single["pairs"])
    bad = {k: ot.max_rel(multi[k], single[k]) for k in cases.FIELDS_F64 + ["ID", "Nb"] if not np.array_equal(multi[k], single[k])}
    assert not bad, "rebalanced multi-GPU run differs from single GPU per particle: %s" % bad


@pytest.mark.gpu
@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("sync,amp,nsteps,ghost", [("full", 60.0, 40, 0), ("tagged", 400.0, 100, 0), ("push", 400.0, 100, 0), ("tagged", 60.0, 40, 1)])
def test_host_mirror_on_two_gpus_bit_identical_to_one(product_lib, sync, amp, nsteps, ghost):
    """The C++ host mirror of SPH::Domain on WELDFORM_GPUS=2 (one process, one host thread and one wf_ctx per GPU, host/Backend.h): the driver's
    GeneralAfter callback works on the GLOBAL Particles array while the particles live on two slabs and migrate between them at the rebuilds
    (amp = 400 m/s); every field of the final state is bit-identical to the same driver on one GPU."""
    from weldform_b200 import build
    exe = [e for e in build.build_host_examples() if e.endswith("compression_host_api")][0]
    outs = []
    with tempfile.TemporaryDirectory() as d:
        for gpus in (1, 2):
            out = os.path.join(d, "o%d.bin" % gpus)
            r = subprocess.run([exe, "0.02", str(nsteps), str(amp), out, "0", str(ghost)], capture_output=True, text=True, timeout=900,
                               env=dict(os.environ, WELDFORM_SYNC=sync, WELDFORM_GPUS=str(gpus)))
            assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
            raw = np.fromfile(out, dtype=np.uint8)
            outs.append(raw)
    assert outs[0].size == outs[1].size and outs[0].size > 8
    a, b = outs[0][8:].view(np.float64), outs[1][8:].view(np.float64)
    assert np.array_equal(a, b), "2-GPU host-mirror run differs from the 1-GPU run: max |diff| %.3e in %d of %d values" % (np.abs(a - b).max(), int((a != b).sum()), a.size)


@pytest.mark.gpu
@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("name", ["plate_cubic_40", "plate_quintic_10", "fsw_like_80", "thermal_solve_25", "thermal_solve_gc_20", "ghost_comp_yield_40", "ghost_lf_25"])
def test_multi_gpu_plate_and_fsw_goldens(product_lib, name):
    """BASELINE configs 4 and 5 in small: the Plate-Yield_3d lattice (AddBoxLength, slabs along x; cubic and quintic-spline kernels) and the
    FSW-like set-up (Johnson-Cook, convection BC, FIXED side walls, rotating + plunging tool, friction heat) on two GPUs: bit-identical per
    particle to one GPU and within the golden bar of the reference's own dump. Domain::ThermalSolve (conduction only, one neighbour search, with
    and without corrected gradients) rides along: its halo refresh points are those of the post pass. So do the ghost (symmetry) rows of
    AddCylinderLength(..., ghost = true) with Domain::MoveGhost in the Fraser and LeapFrog loops: every rank holds all GhostPairs and serves
    the pairs whose ghost it owns (wf_step.cu ghost_localize)."""
    from tests import golden_tools
    from weldform_b200 import harness
    case, ref = golden_tools.load(name)
    multi, info = _run_multi(case, 2)
    single = harness.run_case(case)
    assert all(i[1] > 0 for i in info), "every rank must hold halo particles"
    assert np.array_equal(multi["pairs"], single["pairs"]) and np.array_equal(multi["pairs"], ref["pairs"])
    bad = {k: ot.max_rel(multi[k], single[k]) for k in cases.FIELDS_F64 + ["ID", "Nb"] if not np.array_equal(multi[k], single[k])}
    assert not bad, "multi-GPU differs from single GPU per particle: %s" % bad
    worst = {k: ot.max_rel(multi[k], ref[k]) for k in ("x", "v", "Displacement", "pl_strain", "T", "Density", "Sigma") if k in ref}
    assert max(worst.values()) < 1e-7, worst
