"""Builds a Context from a "case" dict (the same WFC1 records the oracle programs take, tests/cases.py) and returns the
state as a dict of numpy arrays. Host-side setup only: lattice generation through the library's generators, rigid-mesh
construction (TriMesh::AxisPlaneMesh / CalcSpheres, Source/Mesh.cpp:82-220), property arrays. No oracle code is used.
"""
import numpy as np

from . import api

_PART_D = {"G": "G", "Cs": "Cs", "P0": "P0", "Alpha": "alpha", "Beta": "beta", "TI": "TI", "TIn": "TIn", "TIInitDist": "TIInitDist", "Ep": "Ep",
           "k_T": "k_T", "cp_T": "cp_T", "h_conv": "h_conv", "T_inf": "T_inf", "th_exp": "th_exp", "q_source": "q_source",
           "Sigmay": "sigmay", "T": "T", "RefDensity": "ref_density"}
_PART_I = {"PresEq": "pres_eq", "Fail": "fail", "Material": "material", "Material_model": "material_model", "mat_id": "mat_id",
           "Thermal_BC": "thermal_bc", "ID": "id", "IsFree": "is_free", "not_write_surf_ID": "not_write_surf_id"}


def _dotb(a, b):
    """Blitz dot order a0*b0 + (a1*b1 + a2*b2)."""
    return a[..., 0] * b[..., 0] + (a[..., 1] * b[..., 1] + a[..., 2] * b[..., 2])


def axis_plane_mesh(axis, posit, p1, p2, dens):
    """TriMesh::AxisPlaneMesh + CalcSpheres for dimension 3 (Source/Mesh.cpp:82-220): nodes, elements, normals, nfar, radius."""
    p1 = np.asarray(p1, float); p2 = np.asarray(p2, float)
    p = p2 - p1
    dirs = {0: (1, 2), 1: (0, 2)}.get(axis, (0, 1))
    x3 = p1[axis]; x2 = p1[dirs[1]]
    dl = p[dirs[0]] / dens
    nodes = []
    for _ in range(dens + 1):
        x1 = p1[dirs[0]]
        for _ in range(dens + 1):
            v = np.zeros(3); v[dirs[0]] = x1; v[dirs[1]] = x2; v[axis] = x3
            nodes.append(v)
            x1 += dl
        x2 += dl
    nodes = np.array(nodes)
    elems = []
    for j in range(dens):
        for i in range(dens):
            n0 = (dens + 1) * j + i; n1 = n0 + 1; n2 = (dens + 1) * (j + 1) + i; n3 = n2 + 1
            if posit:
                elems += [(n0, n1, n2), (n1, n3, n2)]
            else:
                elems += [(n0, n2, n1), (n1, n2, n3)]
    elems = np.array(elems, np.int32)
    A, B, Cn = nodes[elems[:, 0]], nodes[elems[:, 1]], nodes[elems[:, 2]]
    centroid = ((A + B) + Cn) / 3.0
    u, v = B - A, Cn - A
    w = np.stack([u[:, 1] * v[:, 2] - u[:, 2] * v[:, 1], u[:, 2] * v[:, 0] - u[:, 0] * v[:, 2], u[:, 0] * v[:, 1] - u[:, 1] * v[:, 0]], axis=1)
    normals = w / np.sqrt(_dotb(w, w))[:, None]
    if posit:
        normals = -normals
    radius = np.zeros(len(elems))
    for k in range(3):
        rv = nodes[elems[:, k]] - centroid
        radius = np.maximum(radius, np.sqrt(_dotb(rv, rv)))
    nfar = np.full(len(elems), 2, np.int32)      # Mesh.cpp:211-214: nfar ends as the last node
    return nodes, elems, normals, nfar, radius


def context_from_case(case, device=0, comm=None):
    """Create and fully set up a Context; returns (ctx, info). comm = (rank, nranks, uid) for the multi-GPU decomposition:
    every rank builds the whole case on the host, asks the library's planner for the slab split and uploads only its part."""
    ctx = api.Context(device)
    if comm is not None:
        ctx.comm_init(comm[0], comm[1], comm[2])
    p = ctx.default_params()
    p.dimension = int(case.get("Dimension", 3)); p.kernel_type = int(case.get("KernelType", 0)); p.gradient_type = int(case.get("GradientType", 0))
    p.ts_nb_inc = int(case.get("ts_nb_inc", 5))
    if "DomMin" in case:
        for k in range(3):
            p.dom_min[k] = float(case["DomMin"][k]); p.dom_max[k] = float(case["DomMax"][k])
    p.thermal_solver = int(case.get("thermal_solver", 0)); p.pl_heating = int(case.get("pl_heating", 0)); p.pl_work_heat_frac = float(case.get("pl_work_heat_frac", 0.9))
    mesh_count = int(case.get("mesh_count", 0))
    p.contact = int(case.get("contact", 1 if mesh_count else 0))
    p.friction_sta = float(case.get("friction_sta", 0.0)); p.friction_dyn = float(case.get("friction_dyn", 0.0))
    p.friction_function = int(case.get("friction_function", 0)); p.friction_m = float(case.get("friction_m", 0.0)); p.friction_b = float(case.get("friction_b", 0.0))
    p.pfac = float(case.get("PFAC", 0.4)); p.dfac = float(case.get("DFAC", 0.0))
    p.cont_heat_fric = int(case.get("cont_heat_fric", 0)); p.cont_heat_cond = int(case.get("cont_heat_cond", 0))
    p.tuning = int(case.get("tuning", int(__import__("os").environ.get("WF_TUNING", "0"))))
    p.grad_kernel_corr = int(case.get("gradKernelCorr", 0))
    p.solver = int(case.get("solver", 0)); p.auto_ts = int(case.get("auto_ts", 0)); p.cfl = float(case.get("CFL", 0.7)); p.auto_ts_acc = int(case.get("auto_ts_acc", 0))
    ctx.set_params(p)
    if "materials" in case:
        mt = np.asarray(case["materials"], float).reshape(-1, 16)           # [model, E, nu, p0..p12] per material
        if "materials_ext" in case:                                           # p13..p15 (GMT has 14 parameters)
            mt = np.concatenate([mt, np.asarray(case["materials_ext"], float).reshape(-1, 3)], axis=1)
        ctx.set_materials([tuple(r) for r in mt])
    # lattice
    kind = int(case.get("gen_kind", 0))
    arrays = {}
    if kind == 1:
        g = case["gen_args"]
        ghost_pairs = None
        if len(g) > 10 and g[10] != 0:                     # AddCylinderLength(..., ghost = true)
            pts, ghost_pairs = api.gen_cylinder_ghost((g[1], g[2], g[3]), g[4], g[5], g[6])
            nowrite = np.zeros(len(pts), np.int32); nowrite[ghost_pairs[:, 1]] = 1      # Domain.cpp:814
            arrays["not_write_surf_id"] = nowrite
        else:
            pts = api.gen_cylinder((g[1], g[2], g[3]), g[4], g[5], g[6])
        n = len(pts)
        mass = np.pi * g[4] * g[4] * g[5] * g[7] / n       # Domain.cpp:839-841
        arrays.update(x=pts, mass=np.full(n, mass), density=np.full(n, g[7]), h=np.full(n, g[8]), id=np.full(n, int(g[0]), np.int32))
    elif kind == 2:
        g = case["gen_args"]
        pts, dims = api.gen_box((g[1], g[2], g[3]), g[4], g[5], g[6], g[7])
        n = len(pts)
        mass = dims[0] * dims[1] * dims[2] * g[8] / n     # Domain.cpp:447-455
        arrays.update(x=pts, mass=np.full(n, mass), density=np.full(n, g[8]), h=np.full(n, g[9]), id=np.full(n, int(g[0]), np.int32))
    else:
        pts = np.asarray(case["x"], float).reshape(-1, 3)
        n = len(pts)
        arrays.update(x=pts, mass=case["Mass"], density=case["Density"], h=case["h"])
    if kind != 0:
        for key, name in (("Mass", "mass"), ("Density", "density"), ("h", "h")):
            if key in case:
                arrays[name] = case[key]
        if "x" in case:
            arrays["x"] = np.asarray(case["x"], float).reshape(-1, 3)
    if "v" in case:
        arrays["v"] = np.asarray(case["v"], float).reshape(-1, 3)
    for key, name in _PART_D.items():
        if key in case:
            arrays[name] = case[key]
    for key, name in _PART_I.items():
        if key in case:
            arrays[name] = case[key]
    meshes = []
    mesh_x, mesh_h = [], []                      # element particles (centroid, (float)hfac*radius; Contact.cpp:10-38): they shape the cell grid
    for m in range(mesh_count):
        pl = case["mesh%d_plane" % m]
        nodes, elems, normals, nfar, radius = axis_plane_mesh(int(pl[0]), pl[1] != 0, pl[2:5], pl[5:8], int(pl[8]))
        hfac = float(case.get("mesh%d_hfac" % m, 1.1))
        ctx.add_mesh(nodes, elems, normals, nfar, radius, hfac, int(case.get("mesh%d_id" % m, 10 + m)))
        meshes.append((len(nodes), len(elems)))
        mesh_x.append(((nodes[elems[:, 0]] + nodes[elems[:, 1]]) + nodes[elems[:, 2]]) / 3.0)
        mesh_h.append(float(np.float32(hfac)) * radius)
    n_global = n
    if comm is not None and comm[1] > 1:
        xg = np.asarray(arrays["x"], float).reshape(-1, 3)
        hg = np.broadcast_to(np.asarray(arrays["h"], float), (n,))
        if mesh_x:
            xg = np.concatenate([xg] + mesh_x); hg = np.concatenate([hg] + mesh_h)
        rank_of, bounds = api.plan_slabs(xg, hg, comm[1], int(case.get("KernelType", 0)), case.get("DomMin"), case.get("DomMax"), n_sph=n)
        sel = np.nonzero(rank_of[:n] == comm[0])[0]
        local = {}
        for key, val in arrays.items():
            v = np.asarray(val)
            if v.ndim == 0:
                v = np.full(n, v)
            if v.size == 3 * n and key in ("x", "v"):
                local[key] = v.reshape(-1, 3)[sel]
            elif v.shape[0] == n:
                local[key] = v[sel]
            else:
                raise api.WeldformError("multi-GPU split: array %s has unexpected size %d" % (key, v.size))
        arrays = local
        arrays["global_index"] = sel.astype(np.int64)
        n = len(sel)
    ctx.set_particles(n, n_sph_global=n_global, **arrays)
    if kind == 1 and ghost_pairs is not None:
        ctx.set_ghost_pairs(ghost_pairs)
    if "ghost_pairs" in case:                               # explicit Domain::GhostPairs (inner, ghost) + ghost_plane_axis per pair, e.g. mirrors behind x = 0 / y = 0
        ctx.set_ghost_pairs(np.asarray(case["ghost_pairs"], np.int32).reshape(-1, 2), np.asarray(case["ghost_axis"], np.int32) if "ghost_axis" in case else None)
    for m in range(mesh_count):
        if "mesh%d_w" % m in case:                         # TriMesh::SetRotAxisVel
            ctx.mesh_set_motion(m, (0.0, 0.0, 0.0), tuple(float(w) for w in case["mesh%d_w" % m][:3]))
        key = "mesh%d_vel" % m
        if key in case:
            ctx.mesh_set_ramp(m, case[key][:3], float(case[key][3]))
    if "bc" in case:
        for r in np.asarray(case["bc"], float).reshape(-1, 8):
            ctx.add_velocity_bc(int(r[0]), r[1:4], r[4], int(r[5]))
    return ctx, dict(n_sph=n, n_sph_global=n_global, meshes=meshes)


_DUMP = ["x", "v", "a", "Displacement", "normal", "contforce", "Mass", "h", "Density", "dDensity", "Pressure", "Sigma", "ShearStress", "StrainRate",
         "RotationRate", "Strain", "Strain_pl", "TIR", "pl_strain", "Sigmay", "eff_strain_rate", "T", "dTdt", "q_plheat", "q_conv", "delta_cont",
         "dkin_energy_dt", "dint_energy_dt", "ID", "Nb"]


def dump(ctx, info, with_pairs=True):
    out = {}
    for name in _DUMP:
        out[name] = ctx.download(name).reshape(-1)
    out["global_index"] = ctx.download("global_index").reshape(-1)
    if ctx.grad_kernel_corr:
        out["gradCorrM"] = ctx.download("gradCorrM").reshape(-1)
    s = ctx.reduce()
    out.update(n_particles=np.array([float(ctx.n)]), solid_part_count=np.array([float(ctx.n_sph)]), Time=np.array([s.time]), deltat=np.array([s.deltat]),
               kin_energy_sum=np.array([s.kin_energy_sum]), int_energy_sum=np.array([s.int_energy_sum]), plastic_work=np.array([s.plastic_work]),
               contact_force_sum=np.array([s.contact_force_sum]), contact_reaction_sum=np.array([s.contact_reaction_sum]),
               ext_forces_work=np.array([s.ext_forces_work]), contact_friction_work=np.array([s.contact_friction_work]),
               max_disp=np.array(s.max_disp[:]), nb_rebuilds=np.array([float(s.nb_rebuilds)]))
    if with_pairs:
        out["pairs"] = ctx.export_pairs().reshape(-1)
        out["cont_pairs"] = ctx.export_contact_pairs().reshape(-1)
    for m, (nn, ne) in enumerate(info["meshes"]):
        nodes, normals, pp = ctx.mesh_download(m, nn, ne)
        out["mesh%d_nodes" % m] = nodes.reshape(-1); out["mesh%d_normals" % m] = normals.reshape(-1); out["mesh%d_pplane" % m] = pp
    return out


def merge_rank_dumps(dumps):
    """Per-rank dumps of a multi-GPU run -> one dump in global creation order (pairs merged and sorted)."""
    n = int(max(int(d["global_index"].max()) for d in dumps)) + 1       # rigid-mesh element particles are replicated on every rank
    out = {}
    for name in _DUMP + ["global_index"]:
        first = dumps[0][name]
        nc = first.size // max(len(dumps[0]["global_index"]), 1) if len(dumps[0]["global_index"]) else 1
        full = np.empty((n, nc), first.dtype)
        for d in dumps:
            full[d["global_index"]] = d[name].reshape(len(d["global_index"]), nc)
        out[name] = full.reshape(-1)
    for k, v in dumps[0].items():
        if k not in out and k not in ("pairs", "cont_pairs"):
            out[k] = v                                     # scalars are identical on every rank
    out["n_particles"] = np.array([float(n)])
    out["solid_part_count"] = np.array([float(sum(float(d["solid_part_count"][0]) for d in dumps))])
    if "pairs" in dumps[0]:
        pr = np.concatenate([d["pairs"].reshape(-1, 2) for d in dumps])
        pr = pr[np.lexsort((pr[:, 1], pr[:, 0]))]
        out["pairs"] = pr.reshape(-1)
        cp = np.concatenate([d["cont_pairs"].reshape(-1, 2) for d in dumps])
        out["cont_pairs"] = cp[np.lexsort((cp[:, 1], cp[:, 0]))].reshape(-1)
    return out


def run_case(case, device=0, with_pairs=True, comm=None):
    """Same contract as the oracle programs: case in, state dump out (mode 0: nsteps Fraser steps; mode 1: neighbour build only)."""
    ctx, info = context_from_case(case, device, comm)
    try:
        ctx.solver_init(float(case.get("dt", 1.0e-7)))
        if int(case.get("mode", 0)) == 1:
            ctx.nb_build()
        elif float(case.get("tf", 0.0)) > 0.0:
            ctx.run_until(float(case["tf"]))               # while (Time <= tf): the stop rule of the reference's loops (adaptive time step)
        else:
            ctx.run(int(case.get("nsteps", 1)))
        return dump(ctx, info, with_pairs)
    finally:
        ctx.close()
