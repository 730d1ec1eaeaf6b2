"""ctypes binding of include/weldform_b200.h. Names and argument meaning follow the header one to one."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class WeldformError(RuntimeError):
    pass


def lib_path():
    return os.path.join(_HERE, "libweldform_b200.so")


class wf_params(C.Structure):
    _fields_ = [("dimension", C.c_int), ("kernel_type", C.c_int), ("gradient_type", C.c_int), ("ts_nb_inc", C.c_int),
                ("dom_min", C.c_double * 3), ("dom_max", C.c_double * 3),
                ("thermal_solver", C.c_int), ("pl_heating", C.c_int), ("pl_work_heat_frac", C.c_double),
                ("contact", C.c_int), ("friction_sta", C.c_double), ("friction_dyn", C.c_double),
                ("friction_function", C.c_int), ("friction_m", C.c_double), ("friction_b", C.c_double),
                ("pfac", C.c_double), ("dfac", C.c_double), ("contact_force_factor", C.c_double),
                ("cont_heat_fric", C.c_int), ("cont_heat_cond", C.c_int), ("contact_hc", C.c_double),
                ("nb_order_reference", C.c_int), ("grad_kernel_corr", C.c_int), ("solver", C.c_int), ("auto_ts", C.c_int), ("cfl", C.c_double), ("auto_ts_acc", C.c_int), ("tuning", C.c_int), ("sqrt_h_a", C.c_double)]


class wf_material(C.Structure):
    _fields_ = [("model", C.c_int), ("E", C.c_double), ("nu", C.c_double), ("p", C.c_double * 16)]


_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int32)


class wf_particles(C.Structure):
    _fields_ = [("n_sph", C.c_int64), ("x", _PD), ("v", _PD), ("mass", _PD), ("density", _PD), ("h", _PD), ("ref_density", _PD),
                ("id", _PI), ("is_free", _PI), ("not_write_surf_id", _PI),
                ("G", _PD), ("Cs", _PD), ("P0", _PD), ("alpha", _PD), ("beta", _PD), ("TI", _PD), ("TIn", _PD), ("TIInitDist", _PD),
                ("Ep", _PD), ("k_T", _PD), ("cp_T", _PD), ("h_conv", _PD), ("T_inf", _PD), ("th_exp", _PD), ("q_source", _PD),
                ("pres_eq", _PI), ("fail", _PI), ("material", _PI), ("material_model", _PI), ("mat_id", _PI), ("thermal_bc", _PI),
                ("sigmay", _PD), ("T", _PD), ("global_index", C.POINTER(C.c_int64)), ("n_sph_global", C.c_int64)]


class wf_mesh(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("n_elems", C.c_int32), ("nodes", _PD), ("elems", _PI), ("normals", _PD), ("nfar", _PI),
                ("radius", _PD), ("hfac", C.c_double), ("id", C.c_int32)]


class wf_scalars(C.Structure):
    _fields_ = [("time", C.c_double), ("max_pl_strain", C.c_double), ("max_disp", C.c_double * 3),
                ("kin_energy_sum", C.c_double), ("int_energy_sum", C.c_double), ("plastic_work", C.c_double),
                ("contact_force_sum", C.c_double), ("contact_reaction_sum", C.c_double), ("ext_forces_work", C.c_double),
                ("contact_friction_work", C.c_double), ("accum_cont_heat_cond", C.c_double), ("max_T", C.c_double), ("min_T", C.c_double),
                ("steps", C.c_int64), ("nb_rebuilds", C.c_int64), ("is_yielding", C.c_int32), ("deltat", C.c_double)]


# wf_field enum: name -> (id, components, dtype)
_F3 = ["x", "v", "a", "Displacement", "normal", "contforce"]
_F1 = ["Density", "dDensity", "Pressure", "Mass", "h", "pl_strain", "Sigmay", "eff_strain_rate", "T", "dTdt", "q_plheat", "q_conv",
       "delta_cont", "dkin_energy_dt", "dint_energy_dt"]
_F9 = ["Sigma", "ShearStress", "StrainRate", "RotationRate", "Strain", "Strain_pl", "TIR", "gradCorrM"]
_FI = ["ID", "Nb", "global_index"]
FIELDS = {}
for _k, _n in enumerate(_F3 + _F1 + _F9 + _FI):
    FIELDS[_n] = (_k, 3 if _n in _F3 else 9 if _n in _F9 else 1, np.int32 if _n in _FI else np.float64)

_lib = None


def load_library():
    """Load libweldform_b200.so; raises (no fallback) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise WeldformError("%s is missing: run `python -m weldform_b200.build` (or __graft_entry__.build()); there is no CPU fallback" % path)
    L = C.CDLL(path)
    vp = C.c_void_p
    sig = {
        "wf_create": ([C.c_int, C.POINTER(vp)], C.c_int), "wf_destroy": ([vp], None), "wf_last_error": ([vp], C.c_char_p),
        "wf_sync": ([vp], C.c_int), "wf_default_params": ([C.POINTER(wf_params)], None),
        "wf_set_params": ([vp, C.POINTER(wf_params)], C.c_int), "wf_set_materials": ([vp, C.POINTER(wf_material), C.c_int], C.c_int),
        "wf_set_particles": ([vp, C.POINTER(wf_particles)], C.c_int), "wf_add_mesh": ([vp, C.POINTER(wf_mesh)], C.c_int),
        "wf_mesh_set_motion": ([vp, C.c_int, _PD, _PD], C.c_int), "wf_mesh_set_ramp": ([vp, C.c_int, _PD, C.c_double], C.c_int),
        "wf_add_velocity_bc": ([vp, C.c_int, _PD, C.c_double, C.c_int], C.c_int), "wf_clear_bcs": ([vp], C.c_int),
        "wf_num_particles": ([vp], C.c_int64), "wf_num_sph": ([vp], C.c_int64),
        "wf_download": ([vp, C.c_int, vp], C.c_int), "wf_upload": ([vp, C.c_int, vp], C.c_int),
        "wf_download_subset": ([vp, C.c_int, _PI, C.c_int64, vp], C.c_int), "wf_upload_subset": ([vp, C.c_int, _PI, C.c_int64, vp], C.c_int),
        "wf_mesh_download": ([vp, C.c_int, _PD, _PD, _PD], C.c_int),
        "wf_solver_init": ([vp, C.c_double], C.c_int), "wf_nb_build": ([vp], C.c_int), "wf_grad_corr": ([vp], C.c_int),
        "wf_nb_export_pairs": ([vp, C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_int64)], C.c_int),
        "wf_cont_export_pairs": ([vp, C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_int64)], C.c_int),
        "wf_reduce": ([vp, C.POINTER(wf_scalars)], C.c_int), "wf_fraser_run": ([vp, C.c_int64], C.c_int),
        "wf_run": ([vp, C.c_int64], C.c_int), "wf_solver_stage": ([vp, C.c_int], C.c_int), "wf_run_until": ([vp, C.c_double, C.POINTER(C.c_int64)], C.c_int),
        "wf_gen_cylinder": ([_PD, C.c_double, C.c_double, C.c_double, _PD, C.c_int64], C.c_int64),
        "wf_gen_box": ([_PD, C.c_double, C.c_double, C.c_double, C.c_double, _PD, C.c_int64, _PD], C.c_int64),
        "wf_gen_cylinder_ghost": ([_PD, C.c_double, C.c_double, C.c_double, _PD, C.c_int64, _PI, C.c_int64, C.POINTER(C.c_int64)], C.c_int64),
        "wf_set_ghost_pairs": ([vp, _PI, _PI, C.c_int64], C.c_int), "wf_move_ghost": ([vp], C.c_int),
        "wf_get_grid": ([vp, _PD, _PD, _PD, _PI], C.c_int),
        "wf_prof_enable": ([vp, C.c_int], C.c_int), "wf_prof_get": ([vp, _PD, C.POINTER(C.c_int64), C.c_int, C.POINTER(C.c_int)], C.c_int),
        "wf_comm_unique_id": ([vp], C.c_int), "wf_comm_init": ([vp, C.c_int, C.c_int, vp], C.c_int), "wf_comm_info": ([vp, C.POINTER(C.c_int64)], C.c_int),
        "wf_plan_slabs": ([_PD, _PD, C.c_int64, C.c_int, _PD, _PD, C.c_int, _PI, _PI], C.c_int),
        "wf_subset_register": ([vp, _PI, C.c_int64, C.POINTER(C.c_int)], C.c_int), "wf_subset_release": ([vp, C.c_int], C.c_int),
        "wf_download_registered": ([vp, C.c_int, C.c_int, C.c_void_p], C.c_int), "wf_upload_registered": ([vp, C.c_int, C.c_int, C.c_void_p], C.c_int),
        "wf_upload_registered_async": ([vp, C.c_int, C.c_int, C.c_void_p], C.c_int),
        "wf_plan_rebalance": ([_PD, C.c_int, _PI], C.c_int),
        "wf_plan_slabs_ex": ([_PD, _PD, C.c_int64, C.c_int64, C.c_int, _PD, _PD, C.c_int, _PI, _PI], C.c_int),
        "wf_get_stream": ([vp], vp), "wf_measure_fp64_peak": ([vp, _PD], C.c_int),
        "wf_prof_name": ([C.c_int], C.c_char_p), "wf_launch_count": ([vp, C.POINTER(C.c_int64)], C.c_int),
        "wf_abi_sizes": ([C.POINTER(C.c_int64)], C.c_int),
    }
    for name in ["wf_step_begin", "wf_density", "wf_rate_stress", "wf_accel", "wf_contact", "wf_apply_bcs", "wf_integrate_pos", "wf_integrate_vel",
                 "wf_thermal", "wf_energy", "wf_step_end", "wf_fraser_stage_a", "wf_fraser_stage_b1", "wf_fraser_stage_b2", "wf_fraser_stage_c"]:
        sig[name] = ([vp], C.c_int)
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    sizes = (C.c_int64 * 5)()
    L.wf_abi_sizes(sizes)
    mine = [C.sizeof(t) for t in (wf_params, wf_material, wf_particles, wf_mesh, wf_scalars)]
    if list(sizes) != mine:
        raise WeldformError("%s was built from another include/weldform_b200.h: structure sizes %s, this binding declares %s (rebuild: python -m weldform_b200.build)"
                            % (path, list(sizes), mine))
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(_PD) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(_PI) if a is not None else None


def gen_cylinder(V, Rxy, Lz, r):
    """Domain::AddCylinderLength lattice (host-side generator of the library)."""
    L = load_library()
    Vc = (C.c_double * 3)(*V)
    n = L.wf_gen_cylinder(Vc, Rxy, Lz, r, None, 0)
    pts = np.empty((n, 3))
    L.wf_gen_cylinder(Vc, Rxy, Lz, r, _dp(pts), n)
    return pts


def gen_cylinder_ghost(V, Rxy, Lz, r):
    """Domain::AddCylinderLength(..., ghost = true): lattice + three mirror rows, and the (inner, ghost) index pairs."""
    L = load_library()
    Vc = (C.c_double * 3)(*V)
    npairs = C.c_int64(0)
    n = L.wf_gen_cylinder_ghost(Vc, Rxy, Lz, r, None, 0, None, 0, C.byref(npairs))
    if n < 0:
        raise WeldformError("wf_gen_cylinder_ghost: status %d" % n)
    pts = np.empty((n, 3))
    pairs = np.empty((npairs.value, 2), np.int32)
    L.wf_gen_cylinder_ghost(Vc, Rxy, Lz, r, _dp(pts), n, _ip(pairs), npairs.value, C.byref(npairs))
    return pts, pairs


def gen_box(V, Lx, Ly, Lz, r):
    L = load_library()
    Vc = (C.c_double * 3)(*V)
    dims = (C.c_double * 3)()
    n = L.wf_gen_box(Vc, Lx, Ly, Lz, r, None, 0, dims)
    pts = np.empty((n, 3))
    L.wf_gen_box(Vc, Lx, Ly, Lz, r, _dp(pts), n, dims)
    return pts, np.array(dims[:])


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 creates it, the other ranks receive it by any means)."""
    L = load_library()
    buf = (C.c_char * 128)()
    rc = L.wf_comm_unique_id(buf)
    if rc != 0:
        raise WeldformError("wf_comm_unique_id: %s" % L.wf_last_error(None).decode())
    return bytes(buf)


def plan_slabs(x, h, nranks, kernel_type=0, dom_min=None, dom_max=None, n_sph=None):
    """Host-only planner (wf_plan_slabs_ex): -> (rank_of[n], layer bounds[nranks+1]) of the slab decomposition. With rigid
    meshes x, h hold the n_sph SPH particles followed by the element particles (rank_of = -1: replicated)."""
    L = load_library()
    x = np.ascontiguousarray(x, np.float64).reshape(-1, 3); h = np.ascontiguousarray(h, np.float64)
    n = len(x)
    rank_of = np.empty(n, np.int32); bounds = np.empty(nranks + 1, np.int32)
    dmin = (C.c_double * 3)(*(dom_min if dom_min is not None else (1e11, 1e11, 1e11)))
    dmax = (C.c_double * 3)(*(dom_max if dom_max is not None else (-1e11, -1e11, -1e11)))
    rc = L.wf_plan_slabs_ex(_dp(x), _dp(h), n, int(n if n_sph is None else n_sph), int(kernel_type), dmin, dmax, int(nranks), _ip(rank_of), _ip(bounds))
    if rc != 0:
        raise WeldformError("wf_plan_slabs failed (%d): fewer cell layers than ranks, or bad arguments" % rc)
    return rank_of, bounds


def plan_rebalance(stats, bounds):
    """Host-only (wf_plan_rebalance): stats[nranks, 8], bounds[nranks + 1] -> new bounds."""
    L = load_library()
    st = np.ascontiguousarray(stats, np.float64).reshape(-1, 8); b = np.ascontiguousarray(bounds, np.int32).copy()
    rc = L.wf_plan_rebalance(_dp(st), len(st), _ip(b))
    if rc != 0:
        raise WeldformError("wf_plan_rebalance failed (%d)" % rc)
    return b


class Context:
    """One wf_ctx (one GPU). Thin, explicit wrapper: each method is the C entry point of the same name."""

    def __init__(self, device=0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.wf_create(device, C.byref(h))
        if rc != 0:
            raise WeldformError("wf_create: %s" % self.L.wf_last_error(None).decode())
        self.h = h
        self.grad_kernel_corr = 0
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.L.wf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise WeldformError("%s failed (%d): %s" % (what, rc, self.L.wf_last_error(self.h).decode()))

    def default_params(self):
        p = wf_params()
        self.L.wf_default_params(C.byref(p))
        return p

    def set_params(self, p):
        self._ck(self.L.wf_set_params(self.h, C.byref(p)), "wf_set_params")
        self.grad_kernel_corr = int(p.grad_kernel_corr)

    def set_materials(self, mats):
        arr = (wf_material * max(len(mats), 1))()
        for k, m in enumerate(mats):
            arr[k].model = int(m[0]); arr[k].E = m[1]; arr[k].nu = m[2]
            for q in range(16):
                arr[k].p[q] = m[3 + q] if 3 + q < len(m) else 0.0
        self._ck(self.L.wf_set_materials(self.h, arr, len(mats)), "wf_set_materials")

    def add_mesh(self, nodes, elems, normals, nfar, radius, hfac, mesh_id):
        nodes = np.ascontiguousarray(nodes, np.float64); elems = np.ascontiguousarray(elems, np.int32)
        normals = np.ascontiguousarray(normals, np.float64); nfar = np.ascontiguousarray(nfar, np.int32); radius = np.ascontiguousarray(radius, np.float64)
        m = wf_mesh(len(nodes), len(elems), _dp(nodes), _ip(elems), _dp(normals), _ip(nfar), _dp(radius), float(hfac), int(mesh_id))
        self._ck(self.L.wf_add_mesh(self.h, C.byref(m)), "wf_add_mesh")

    def comm_init(self, rank, nranks, uid):
        self._ck(self.L.wf_comm_init(self.h, int(rank), int(nranks), C.c_char_p(uid) if uid is not None else None), "wf_comm_init")

    def comm_info(self):
        out = (C.c_int64 * 8)()
        self._ck(self.L.wf_comm_info(self.h, out), "wf_comm_info")
        keys = ["n_local", "n_halo", "migrated_out", "migrated_in", "layer_lo", "layer_hi", "halo_bytes_per_step", "layers"]
        return dict(zip(keys, [int(v) for v in out]))

    def set_particles(self, n_sph, n_sph_global=0, **arrays):
        p = wf_particles()
        p.n_sph = n_sph
        p.n_sph_global = int(n_sph_global)
        keep = []
        for name, typ in wf_particles._fields_:
            if name in ("n_sph", "n_sph_global") or name not in arrays or arrays[name] is None:
                continue
            if name == "global_index":
                a = np.ascontiguousarray(arrays[name], np.int64)
                keep.append(a)
                p.global_index = a.ctypes.data_as(C.POINTER(C.c_int64))
                continue
            a = np.ascontiguousarray(arrays[name], np.float64 if typ is _PD else np.int32)
            keep.append(a)
            setattr(p, name, _dp(a) if typ is _PD else _ip(a))
        unknown = set(arrays) - {f[0] for f in wf_particles._fields_}
        if unknown:
            raise WeldformError("set_particles: unknown arrays %s" % sorted(unknown))
        self._ck(self.L.wf_set_particles(self.h, C.byref(p)), "wf_set_particles")

    def set_ghost_pairs(self, pairs, axis=None):
        """Domain::GhostPairs for MoveGhost: (inner, ghost) creation indices, Mirror_XYZ across ghost_plane_axis (default 2)"""
        pr = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        ax = None if axis is None else np.ascontiguousarray(axis, np.int32)
        self._ck(self.L.wf_set_ghost_pairs(self.h, _ip(pr), None if ax is None else _ip(ax), len(pr)), "wf_set_ghost_pairs")

    def get_grid(self):
        """Domain::BLPF, TRPR, CellSize, CellNo of the cell grid (after solver_init)"""
        b, t, cs, cn = np.zeros(3), np.zeros(3), np.zeros(3), np.zeros(3, np.int32)
        self._ck(self.L.wf_get_grid(self.h, _dp(b), _dp(t), _dp(cs), _ip(cn)), "wf_get_grid")
        return b, t, cs, cn

    def move_ghost(self):
        self._ck(self.L.wf_move_ghost(self.h), "wf_move_ghost")

    def mesh_set_ramp(self, mesh, vmax, tau):
        self._ck(self.L.wf_mesh_set_ramp(self.h, mesh, (C.c_double * 3)(*vmax), tau), "wf_mesh_set_ramp")

    def mesh_set_motion(self, mesh, v, w=(0.0, 0.0, 0.0)):
        self._ck(self.L.wf_mesh_set_motion(self.h, mesh, (C.c_double * 3)(*v), (C.c_double * 3)(*w)), "wf_mesh_set_motion")

    def add_velocity_bc(self, pid, vmax, tau, set_a=1):
        self._ck(self.L.wf_add_velocity_bc(self.h, int(pid), (C.c_double * 3)(*vmax), float(tau), int(set_a)), "wf_add_velocity_bc")

    @property
    def n(self):
        return self.L.wf_num_particles(self.h)

    @property
    def n_sph(self):
        return self.L.wf_num_sph(self.h)

    def download(self, name, idx=None, out=None):
        fid, nc, dt = FIELDS[name]
        cnt = self.n if idx is None else len(idx)
        if out is None:
            out = np.empty((cnt, nc) if nc > 1 else (cnt,), dt)
        elif out.dtype != dt or out.size != cnt * nc or not out.flags.c_contiguous:
            raise WeldformError("download(%s): out buffer has the wrong dtype/size/layout" % name)
        if idx is None:
            self._ck(self.L.wf_download(self.h, fid, out.ctypes.data_as(C.c_void_p)), "wf_download(%s)" % name)
        else:
            idx = np.ascontiguousarray(idx, np.int32)
            self._ck(self.L.wf_download_subset(self.h, fid, _ip(idx), len(idx), out.ctypes.data_as(C.c_void_p)), "wf_download_subset(%s)" % name)
        return out

    def upload(self, name, data, idx=None):
        fid, nc, dt = FIELDS[name]
        data = np.ascontiguousarray(data, dt)
        if idx is None:
            self._ck(self.L.wf_upload(self.h, fid, data.ctypes.data_as(C.c_void_p)), "wf_upload(%s)" % name)
        else:
            idx = np.ascontiguousarray(idx, np.int32)
            self._ck(self.L.wf_upload_subset(self.h, fid, _ip(idx), len(idx), data.ctypes.data_as(C.c_void_p)), "wf_upload_subset(%s)" % name)

    def subset_register(self, idx):
        """Keep an index list (local creation order) on the device; returns a handle for download_registered / upload_registered."""
        idx = np.ascontiguousarray(idx, np.int32); h = C.c_int()
        self._ck(self.L.wf_subset_register(self.h, _ip(idx), len(idx), C.byref(h)), "wf_subset_register")
        return int(h.value)

    def subset_release(self, handle):
        self._ck(self.L.wf_subset_release(self.h, int(handle)), "wf_subset_release")

    def download_registered(self, name, handle, out):
        fid = FIELDS[name][0]
        self._ck(self.L.wf_download_registered(self.h, fid, int(handle), out.ctypes.data_as(C.c_void_p)), "wf_download_registered(%s)" % name)
        return out

    def upload_registered(self, name, handle, data, wait=True):
        """wait=False: wf_upload_registered_async -- `data` (pinned) must stay untouched until the next blocking call (sync, reduce, download)"""
        fid = FIELDS[name][0]
        fn = self.L.wf_upload_registered if wait else self.L.wf_upload_registered_async
        self._ck(fn(self.h, fid, int(handle), data.ctypes.data_as(C.c_void_p)), "wf_upload_registered(%s)" % name)

    def mesh_download(self, mesh, n_nodes, n_elems):
        nodes = np.empty((n_nodes, 3)); normals = np.empty((n_elems, 3)); pp = np.empty(n_elems)
        self._ck(self.L.wf_mesh_download(self.h, mesh, _dp(nodes), _dp(normals), _dp(pp)), "wf_mesh_download")
        return nodes, normals, pp

    def solver_init(self, dt):
        self._ck(self.L.wf_solver_init(self.h, dt), "wf_solver_init")

    def nb_build(self):
        self._ck(self.L.wf_nb_build(self.h), "wf_nb_build")

    def grad_corr(self):
        self._ck(self.L.wf_grad_corr(self.h), "wf_grad_corr")

    def _pairs(self, fn, what):
        n = C.c_int64()
        self._ck(fn(self.h, None, 0, C.byref(n)), what)
        out = np.empty((n.value, 2), np.int64)
        if n.value:
            self._ck(fn(self.h, out.ctypes.data_as(C.POINTER(C.c_int64)), n.value, C.byref(n)), what)
        return out

    def export_pairs(self):
        return self._pairs(self.L.wf_nb_export_pairs, "wf_nb_export_pairs")

    def export_contact_pairs(self):
        return self._pairs(self.L.wf_cont_export_pairs, "wf_cont_export_pairs")

    def phase(self, name):
        """Call one per-phase entry point, e.g. phase('density') -> wf_density."""
        self._ck(getattr(self.L, "wf_" + name)(self.h), "wf_" + name)

    def run(self, nsteps):
        """nsteps whole steps of the configured solver (wf_params::solver), device-side BCs"""
        self._ck(self.L.wf_run(self.h, nsteps), "wf_run")

    def solver_stage(self, stage):
        self._ck(self.L.wf_solver_stage(self.h, int(stage)), "wf_solver_stage")

    def run_until(self, tf):
        """while (Time <= tf) step; returns the number of steps taken"""
        n = C.c_int64()
        self._ck(self.L.wf_run_until(self.h, float(tf), C.byref(n)), "wf_run_until")
        return int(n.value)

    def sync(self):
        self._ck(self.L.wf_sync(self.h), "wf_sync")

    def reduce(self):
        s = wf_scalars()
        self._ck(self.L.wf_reduce(self.h, C.byref(s)), "wf_reduce")
        return s

    def prof_enable(self, on=True):
        self._ck(self.L.wf_prof_enable(self.h, 1 if on else 0), "wf_prof_enable")

    def prof_get(self):
        ms = (C.c_double * 64)(); nl = (C.c_int64 * 64)(); used = C.c_int()
        self._ck(self.L.wf_prof_get(self.h, ms, nl, 64, C.byref(used)), "wf_prof_get")
        return {self.L.wf_prof_name(k).decode(): (ms[k], nl[k]) for k in range(used.value)}

    def stream(self):
        """cudaStream_t (as int) every kernel of this context is launched on."""
        return int(self.L.wf_get_stream(self.h) or 0)

    def measure_fp64_peak(self):
        t = C.c_double()
        self._ck(self.L.wf_measure_fp64_peak(self.h, C.byref(t)), "wf_measure_fp64_peak")
        return t.value

    def launch_count(self):
        n = C.c_int64()
        self._ck(self.L.wf_launch_count(self.h, C.byref(n)), "wf_launch_count")
        return n.value
