"""Case builders shared by the tests, the golden-fixture generator and bench.py.

A "case" is a flat dict of WFC1 records (see tests/wfc.py, oracle/ref_driver.cpp): lattice generator arguments taken
from the reference's example drivers plus per-particle property arrays, exactly what those drivers assign in their
`for (a < dom.Particles.Size())` loops.
"""
import numpy as np

BILINEAR, HOLLOMON, JOHNSON_COOK, GMT = 0, 1, 2, 3


def cylinder_points(V, Rxy, Lz, r, ghost=False):
    """numpy restatement of the lattice of Domain::AddCylinderLength (Source/Domain.cpp:708-826; ghost: the three mirror rows
    below the bottom plane); used only to size per-particle arrays in case dicts -- positions themselves always come from
    the implementation under test."""
    def half(r, R, xinc):
        yp, xp, c = r, r + (xinc - 1) * 2.0 * r, 1
        while np.sqrt(yp * yp + xp * xp) <= R - r:
            yp += 2.0 * r
            c += 1
        return c - 1
    npxy = half(r, Rxy, 1)
    pts = []
    zp = V[2] + r
    while zp <= V[2] + Lz - r:
        yp = V[1] - r - 2.0 * r * (npxy - 1)
        yinc, sign = npxy, -1
        for _ in range(2 * npxy):
            nx = half(r, Rxy, yinc)
            xp = V[0] - r - 2.0 * r * (nx - 1)
            for _ in range(2 * nx):
                pts.append((xp, yp, zp))
                xp += 2.0 * r
            yp += 2.0 * r
            yinc += sign
            if yinc < 1:
                yinc, sign = 1, 1
        zp += 2.0 * r
    if ghost:
        row = [p for p in pts if p[2] == V[2] + r]
        per_row = len(row)
        zp = V[2] - r
        for k in range(3):
            for q in range(per_row):
                pts.append((pts[k * per_row + q][0], pts[k * per_row + q][1], zp))
            zp -= 2.0 * r
    return np.array(pts, dtype=np.float64)


def box_points(V, Lx, Ly, Lz, r):
    """numpy restatement of the cubic lattice of Domain::AddBoxLength (Source/Domain.cpp:326-470, type 1, no random); used only
    to size per-particle arrays and to tag particles by position -- positions themselves come from the implementation under test."""
    def axis(v0, L):
        out, i, p = [], 0, v0
        while p <= (v0 + L - r):
            out.append(v0 + (2.0 * i + 1) * r)
            i += 1
            p = v0 + (2 * i + 1) * r
        return np.array(out)
    xs, ys, zs = axis(V[0], Lx), axis(V[1], Ly), axis(V[2], Lz)
    pts = np.empty((len(zs) * len(ys) * len(xs), 3))
    pts[:, 0] = np.tile(xs, len(ys) * len(zs)); pts[:, 1] = np.tile(np.repeat(ys, len(xs)), len(zs)); pts[:, 2] = np.repeat(zs, len(xs) * len(ys))
    return pts


def plate_yield(dx=0.0005, nsteps=1, kernel=0, hfac=1.2, vpull=0.01, amp=0.0):
    """examples/Plate-Yield_3d.cpp:60-119 scaled by dx: AddBoxLength plate, bilinear yield, ends ID 2 / 3 pulled apart. The shipped
    driver's Quintic_Spline with h = 1.3 dx gives 250 neighbours per particle, beyond the reference's MAX_NB_PER_PART (SURVEY.md
    section 8c); kernel / hfac select what the reference itself can run. `amp` adds a smooth velocity field so that a short run yields."""
    H, L = 0.01, 0.03
    rho, K, G, Fy = 1000.0, 3.25e6, 7.15e5, 4000.0
    h = hfac * dx
    Cs = np.sqrt(K / rho)
    V = (-L / 2.0 - L / 20.0, -H / 2.0, -H / 2.0)
    dims = (L + L / 10.0 + dx / 10.0, H + dx / 10.0, H + dx / 10.0)
    pts = box_points(V, dims[0], dims[1], dims[2], dx / 2.0)
    n = len(pts)
    x = pts[:, 0]
    ID = np.full(n, 1, np.int32)
    ID[x < -L / 2.0] = 2
    ID[x > L / 2.0] = 3
    ones = np.ones(n)
    case = dict(
        gen_kind=2.0, gen_args=np.array([1, V[0], V[1], V[2], dims[0], dims[1], dims[2], dx / 2.0, rho, h, 1, 0, 0, 0], float),
        Dimension=3.0, KernelType=float(kernel), GradientType=0.0, DomMax=np.array([L, -1e11, -1e11]), DomMin=np.array([-L, 1e11, 1e11]),
        ID=ID, G=G * ones, Cs=Cs * ones, PresEq=np.zeros(n, np.int32), Material=np.full(n, 2, np.int32), Fail=np.ones(n, np.int32),
        Sigmay=Fy * ones, Alpha=1.0 * ones, Beta=1.0 * ones, TI=0.3 * ones, TIInitDist=dx * ones,
        ts_nb_inc=5.0, dt=0.2 * h / Cs, nsteps=float(nsteps), mode=0.0,
        bc=np.array([3, vpull, 0, 0, 0.0, 1, 0, 0, 2, -vpull, 0, 0, 0.0, 1, 0, 0], float),
    )
    if amp:
        y, z = pts[:, 1], pts[:, 2]
        case["v"] = np.stack([amp * np.sin(300.0 * y) * np.cos(200.0 * z) + amp * 40.0 * x, amp * np.cos(250.0 * x) * np.sin(300.0 * z), amp * np.sin(400.0 * x * 1.0) * np.cos(150.0 * y)], axis=1).reshape(-1)
    return case


def fsw_like(dx=0.0012, nsteps=1, wrot=10000.0, vplunge=10.0, vfac=1.0, L=0.05, rt=0.0065):
    """examples/FSW.cpp:60-305 scaled by dx, with the NASTRAN tool replaced by a rotating, plunging AxisPlaneMesh disc-sized plane
    (the mesh file reader is out of scope): AddBoxLength plate, Johnson-Cook with the driver's argument order (FSW.cpp:100-107 passes
    C, m, n, eps_0 where the constructor takes n, c, eps0, m), thermal solver, convection on the bottom row, fixed side walls
    (IsFree = false), friction linear in temperature, contact heat from friction."""
    H = 0.003
    rho, E, nu, Fy = 2700.0, 70.0e9, 0.3, 350.0e6
    K = E / (3.0 * (1.0 - 2 * nu)); G = E / (2.0 * (1.0 + nu))
    h = 1.2 * dx
    Cs = np.sqrt(K / rho)
    ybottom = -H - dx / 1.45
    ytop = ybottom + H
    V = (-L / 2.0 - L / 20.0, ybottom, -L / 2.0 - L / 20.0)
    dims = (L + L / 10.0 + dx / 10.0, H, L + L / 10.0)
    pts = box_points(V, dims[0], dims[1], dims[2], dx / 2.0)
    n = len(pts)
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    r = np.sqrt(x * x + z * z)
    ID = np.zeros(n, np.int32); nowrite = np.zeros(n, np.int32); free = np.ones(n, np.int32)
    thbc = np.zeros(n, np.int32); hconv = np.zeros(n); tinf = np.zeros(n)
    bottom = y < (ybottom + dx)
    ID[bottom] = 2; nowrite[bottom] = 1; thbc[bottom] = 1; hconv[bottom] = 200.0 * vfac; tinf[bottom] = 20.0
    top = (r > 1.85 * rt) & (y > (ytop - dx))
    ID[top] = 2; nowrite[top] = 1
    side = (z < -L / 2.0 + dx) | (z > L / 2.0 - dx) | (x < -L / 2.0 + 2.0 * dx) | (x > L / 2.0 - 2.0 * dx)
    ID[side] = 3; nowrite[side] = 1; free[side] = 0
    ones = np.ones(n)
    A, B, C, m, n_, eps_0, T_m, T_t = 276.0e6, 255.0e6, 0.0015, 1.0, 0.3, 1.0, 775.0, 273.0
    yplane = y.max() + h                          # the tool starts exactly one h above the top row: contact begins as soon as it moves
    dens = max(4, int(np.ceil(2 * rt / (1.45 * dx))))
    sx, sz = 0.37 * dx, -0.21 * dx
    case = dict(
        gen_kind=2.0, gen_args=np.array([0, V[0], V[1], V[2], dims[0], dims[1], dims[2], dx / 2.0, rho, h, 1, 0, 0, 0], float),
        Dimension=3.0, KernelType=0.0, GradientType=0.0, DomMax=np.array([L, -1e11, -1e11]), DomMin=np.array([-L, 1e11, 1e11]),
        ID=ID, IsFree=free, not_write_surf_ID=nowrite, G=G * ones, Cs=Cs * ones, PresEq=np.zeros(n, np.int32), Material=np.full(n, 2, np.int32),
        Fail=np.ones(n, np.int32), Sigmay=Fy * ones, Alpha=1.0 * ones, TI=0.3 * ones, TIInitDist=dx * ones,
        Material_model=np.full(n, JOHNSON_COOK, np.int32), mat_id=np.zeros(n, np.int32),
        materials=np.array([JOHNSON_COOK, E, nu, A, B, C, m, n_, eps_0, T_m, T_t] + [0.0] * 5),     # constructor order: a, b, n, c, eps0, m, Tm, Tt
        k_T=130.0 * vfac * ones, cp_T=960.0 * ones, T=20.0 * ones, Thermal_BC=thbc, h_conv=hconv, T_inf=tinf,
        thermal_solver=1.0, pl_heating=1.0, cont_heat_fric=1.0, ts_nb_inc=5.0,
        contact=1.0, friction_function=1.0, friction_b=0.3, friction_m=(0.1 - 0.3) / 500.0, friction_sta=0.3, friction_dyn=0.3, PFAC=0.6, DFAC=0.0,
        mesh_count=1.0,
        mesh0_plane=np.array([1, 1, -rt + sx, yplane, -rt + sz, rt + sx, yplane, rt + sz, dens], float), mesh0_hfac=1.1, mesh0_id=10.0,
        mesh0_vel=np.array([0, -vplunge * vfac, 0, 0.0], float), mesh0_w=np.array([0.0, wrot * np.pi / 30.0 * vfac, 0.0]),
        bc=np.array([2, 0, 0, 0, 0.0, 1, 0, 0, 3, 0, 0, 0, 0.0, 1, 0, 0], float),
        dt=0.7 * h / Cs, nsteps=float(nsteps), mode=0.0,
    )
    return case


def compression(dx=0.03, nsteps=1, thermal=True, pl_heating=True, alpha=1.0, beta=0.0, ti=0.3, gradient_type=0,
                vmax=10.0, tau=0.005, cfl=0.2, fy=300.0e6, model=BILINEAR, t_field=True, ts_nb_inc=5, ghost=False):
    """examples/Compression_thermal.cpp:71-169 scaled by dx (shipped dx = 0.015 -> 12 144 particles).
    ghost: AddCylinderLength(..., ghost = true) -- three mirror rows below the bottom plane driven by MoveGhost; the bottom rows
    are then left free (no ID 2), the symmetry plane holds the body."""
    rho, K, G = 2700.0, 6.7549e10, 2.5902e10
    R, L = 0.15, 0.56
    h = 1.2 * dx
    Cs = np.sqrt(K / rho)
    V = (0.0, 0.0, -L / 10.0)
    pts = cylinder_points(V, R, L + 2.0 * L / 10.0, dx / 2.0, ghost)
    n = len(pts)
    z = pts[:, 2]
    ID = np.full(n, 1, np.int32)
    if not ghost:
        ID[z < 0] = 2
    ID[z > L] = 3
    ones = np.ones(n)
    case = dict(
        gen_kind=1.0, gen_args=np.array([1, V[0], V[1], V[2], R, L + 2.0 * L / 10.0, dx / 2.0, rho, h, 0, 1 if ghost else 0], float),
        Dimension=3.0, KernelType=0.0, GradientType=float(gradient_type), DomMax=np.array([L, -1e11, -1e11]), DomMin=np.array([-L, 1e11, 1e11]),
        ID=ID, G=G * ones, Cs=Cs * ones, PresEq=np.zeros(n, np.int32), Material=np.full(n, 2, np.int32), Fail=np.ones(n, np.int32),
        Sigmay=fy * ones, Alpha=alpha * ones, Beta=beta * ones, TI=ti * ones, TIInitDist=dx * ones,
        k_T=150.0 * ones, cp_T=960.0 * ones, th_exp=0.0 * ones,
        thermal_solver=float(thermal), pl_heating=float(pl_heating), pl_work_heat_frac=0.9, ts_nb_inc=float(ts_nb_inc),
        dt=cfl * h / Cs, nsteps=float(nsteps), mode=0.0,
        bc=np.array([3, 0, 0, -vmax, tau, 1, 0, 0, 2, 0, 0, 0, 0, 1, 0, 0], float),
    )
    # non-uniform initial temperature so that conduction is exercised (last bullet)
    case["T"] = 20.0 + (15.0 * np.sin(7.0 * pts[:, 0]) * np.cos(5.0 * z) if t_field else 0.0) * ones
    if model == HOLLOMON:
        case["Material_model"] = np.full(n, HOLLOMON, np.int32)
        case["mat_id"] = np.zeros(n, np.int32)
        case["materials"] = np.array([HOLLOMON, 70.0e9, 0.3, fy, 6.79e8, 0.15] + [0.0] * 10)
    elif model == JOHNSON_COOK:
        case["Material_model"] = np.full(n, JOHNSON_COOK, np.int32)
        case["mat_id"] = np.zeros(n, np.int32)
        # A, B, n, C, eps_0, m, T_m, T_t  (6061-T6-like numbers as in examples/FSW.cpp:113-115 semantics)
        case["materials"] = np.array([JOHNSON_COOK, 70.0e9, 0.3, fy, 114.0e6, 0.42, 0.002, 1.0, 1.34, 925.0, 0.0] + [0.0] * 5)
    elif model == GMT:
        case["Material_model"] = np.full(n, GMT, np.int32)
        case["mat_id"] = np.zeros(n, np.int32)
        # GMT(el, n1, n2, C1, C2, m1, m2, I1, I2, e_min, e_max, er_min, er_max, T_min, T_max) (Source/Material.h:168-175): the record carries the
        # first 13 parameters, materials_ext the 14th; sy = C1 exp(C2 T) e^(n1 T + n2) exp((I1 T + I2)/e) er^(m1 T + m2), ~170 MPa at the start
        case["materials"] = np.array([GMT, 70.0e9, 0.3, 1.0e-4, 0.15, 4.0e8, -0.002, 1.0e-5, 0.02, 1.0e-6, -0.002, 0.01, 3.0, 0.1, 1.0e5, 0.0])
        case["materials_ext"] = np.array([500.0, 0.0, 0.0])
    return case


def perturbed(case, amp=5.0, seed=1):
    """Give every particle a smooth non-zero initial velocity so that one step exercises every pair term."""
    c = dict(case)
    g = c["gen_args"]
    pts = cylinder_points((g[1], g[2], g[3]), g[4], g[5], g[6], len(g) > 10 and g[10] != 0)
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    v = np.stack([amp * np.sin(9.0 * y + seed) * np.cos(4.0 * z), amp * np.cos(8.0 * x - seed) * np.sin(3.0 * z), -amp * z * (1.0 + 0.3 * np.sin(11.0 * x * y))], axis=1)
    c["v"] = v.reshape(-1)
    return c


def contact_compression(dx=0.0024, nsteps=1, friction=0.2, model=HOLLOMON, vmax=1.0, tau=0.0005, cfl=0.7, dens=20, ts_nb_inc=5,
                        mesh_shift=(0.37, -0.21)):
    """examples/Contact_Compression_1010.cpp:91-268 scaled by dx (shipped dx = 0.0012), gradKernelCorr off.

    mesh_shift (in units of dx) moves the two tool planes sideways. With the driver's symmetric placement the lattice
    points with x = y sit EXACTLY on the shared diagonal edge of two triangles, where the reference's point-in-triangle
    sign test (Contact_Wang.cpp:153-164) is decided by round-off: one or two elements fire depending on the last bit of
    x - y. That configuration is ill-posed for any parity check (the reference disagrees with itself between thread
    counts), so test cases keep every particle off the mesh edges."""
    rho, E, nu = 7850.0, 200.0e9, 0.3
    K = E / (3.0 * (1.0 - 2 * nu))
    G = E / (2.0 * (1.0 + nu))
    R, L = 0.0127, 0.030
    h = 1.2 * dx
    Cs = np.sqrt(K / rho)
    fy = 260.0e6
    pts = cylinder_points((0.0, 0.0, 0.0), R, L, dx / 2.0)
    n = len(pts)
    ones = np.ones(n)
    zmax = pts[-1, 2] + h
    zmin = pts[0, 2] - h
    sx, sy = mesh_shift[0] * dx, mesh_shift[1] * dx
    case = dict(
        gen_kind=1.0, gen_args=np.array([0, 0, 0, 0, R, L, dx / 2.0, rho, h, 0, 0], float),
        Dimension=3.0, KernelType=0.0, GradientType=0.0, DomMax=np.array([L, -1e11, -1e11]), DomMin=np.array([-L, 1e11, 1e11]),
        G=G * ones, Cs=Cs * ones, PresEq=np.zeros(n, np.int32), Fail=np.ones(n, np.int32), Sigmay=fy * ones,
        Alpha=2.5 * ones, Beta=2.5 * ones, TI=0.3 * ones, TIInitDist=dx * ones,
        Material_model=np.full(n, model, np.int32), mat_id=np.zeros(n, np.int32),
        materials=np.array([HOLLOMON, E, nu, fy, 7.1568e8, 0.22] + [0.0] * 10),
        contact=1.0, friction_sta=friction, friction_dyn=friction, PFAC=0.4, DFAC=0.0, ts_nb_inc=float(ts_nb_inc),
        mesh_count=2.0,
        mesh0_plane=np.array([2, 0, -1.5 * R + sx, -1.5 * R + sy, zmax, 1.5 * R + sx, 1.5 * R + sy, zmax, dens], float), mesh0_hfac=1.1, mesh0_id=10.0,
        mesh0_vel=np.array([0, 0, -vmax, tau], float),
        mesh1_plane=np.array([2, 1, -1.5 * R + sx, -1.5 * R + sy, zmin, 1.5 * R + sx, 1.5 * R + sy, zmin, dens], float), mesh1_hfac=1.1, mesh1_id=11.0,
        mesh1_vel=np.array([0, 0, 0, 0], float),
        dt=cfl * h / (Cs + vmax), nsteps=float(nsteps), mode=0.0,
    )
    return case


FIELDS_F64 = ["x", "v", "a", "Displacement", "Density", "dDensity", "Pressure", "Sigma", "ShearStress", "StrainRate", "RotationRate",
              "Strain", "Strain_pl", "TIR", "pl_strain", "Sigmay", "eff_strain_rate", "T", "dTdt", "q_plheat", "q_conv",
              "contforce", "delta_cont", "normal", "dint_energy_dt", "dkin_energy_dt"]
SCALARS = ["Time", "int_energy_sum", "plastic_work", "contact_force_sum", "contact_reaction_sum", "ext_forces_work", "max_disp", "deltat"]


def compression_fast(dx, nsteps=1, **kw):
    """compression() for large lattices: sizes the per-particle arrays from a vectorised point count instead of the
    Python point loop (positions always come from the library's own generator; only the z of each row is needed here)."""
    rho, K, G = 2700.0, 6.7549e10, 2.5902e10
    R, L = 0.15, 0.56
    r = dx / 2.0
    V = (0.0, 0.0, -L / 10.0)
    Lz = L + 2.0 * L / 10.0

    def half(xinc):
        yp, xp, c = r, r + (xinc - 1) * 2.0 * r, 1
        while np.sqrt(yp * yp + xp * xp) <= R - r:
            yp += 2.0 * r
            c += 1
        return c - 1
    npxy = half(1)
    per_layer = 0
    xs, ys = [], []
    yp = V[1] - r - 2.0 * r * (npxy - 1)
    yinc, sign = npxy, -1
    for _ in range(2 * npxy):
        nx = half(yinc)
        x0 = V[0] - r - 2.0 * r * (nx - 1)
        row = np.empty(2 * nx)
        xp = x0
        for i in range(2 * nx):
            row[i] = xp
            xp += 2.0 * r
        xs.append(row); ys.append(np.full(2 * nx, yp))
        per_layer += 2 * nx
        yp += 2.0 * r
        yinc += sign
        if yinc < 1:
            yinc, sign = 1, 1
    zs = []
    zp = V[2] + r
    while zp <= V[2] + Lz - r:
        zs.append(zp)
        zp += 2.0 * r
    xl, yl = np.concatenate(xs), np.concatenate(ys)
    pts = np.empty((len(zs) * per_layer, 3))
    pts[:, 0] = np.tile(xl, len(zs)); pts[:, 1] = np.tile(yl, len(zs)); pts[:, 2] = np.repeat(np.array(zs), per_layer)
    return _compression_from_points(pts, dx, nsteps, **kw)


def _compression_from_points(pts, dx, nsteps, thermal=True, pl_heating=True, alpha=1.0, beta=0.0, ti=0.3, gradient_type=0,
                             vmax=10.0, tau=0.005, cfl=0.2, fy=300.0e6, ts_nb_inc=5, perturb_amp=0.0, seed=1, contact_planes=False, friction=0.2):
    rho, K, G = 2700.0, 6.7549e10, 2.5902e10
    R, L = 0.15, 0.56
    h = 1.2 * dx
    Cs = np.sqrt(K / rho)
    V = (0.0, 0.0, -L / 10.0)
    n = len(pts)
    z = pts[:, 2]
    ID = np.full(n, 1, np.int32)
    ID[z < 0] = 2
    ID[z > L] = 3
    ones = np.ones(n)
    case = dict(
        gen_kind=1.0, gen_args=np.array([1, V[0], V[1], V[2], R, L + 2.0 * L / 10.0, dx / 2.0, rho, h, 0, 0], float),
        Dimension=3.0, KernelType=0.0, GradientType=float(gradient_type), DomMax=np.array([L, -1e11, -1e11]), DomMin=np.array([-L, 1e11, 1e11]),
        ID=ID, G=G * ones, Cs=Cs * ones, PresEq=np.zeros(n, np.int32), Material=np.full(n, 2, np.int32), Fail=np.ones(n, np.int32),
        Sigmay=fy * ones, Alpha=alpha * ones, Beta=beta * ones, TI=ti * ones, TIInitDist=dx * ones,
        k_T=150.0 * ones, cp_T=960.0 * ones, th_exp=0.0 * ones,
        thermal_solver=float(thermal), pl_heating=float(pl_heating), pl_work_heat_frac=0.9, ts_nb_inc=float(ts_nb_inc),
        dt=cfl * h / Cs, nsteps=float(nsteps), mode=0.0,
        bc=np.array([3, 0, 0, -vmax, tau, 1, 0, 0, 2, 0, 0, 0, 0, 1, 0, 0], float),
        T=20.0 + 15.0 * np.sin(7.0 * pts[:, 0]) * np.cos(5.0 * z),
    )
    if perturb_amp:      # same smooth field as perturbed()
        x, y, amp = pts[:, 0], pts[:, 1], perturb_amp
        case["v"] = np.stack([amp * np.sin(9.0 * y + seed) * np.cos(4.0 * z), amp * np.cos(8.0 * x - seed) * np.sin(3.0 * z),
                              -amp * z * (1.0 + 0.3 * np.sin(11.0 * x * y))], axis=1).reshape(-1)
    if contact_planes:
        # the cylinder between two rigid planes as in examples/Contact_Compression_1010.cpp:148-156,249-250 (AxisPlaneMesh above
        # and below, the upper one pushed down with a ramped velocity) instead of velocity-prescribed end rows; the mesh density
        # keeps the element particles' h = 1.1*radius at the SPH h, as the shipped example does, so that the cell grid stays tight
        del case["bc"]
        case["ID"] = np.full(n, 1, np.int32)
        side = 2.4 * R
        dens = int(np.ceil(side / (1.45 * dx)))
        zmax, zmin = z.max() + h, z.min() - h
        sx, sy = 0.37 * dx, -0.21 * dx           # keeps lattice points off the mesh edges (see contact_compression)
        case.update(contact=1.0, friction_sta=friction, friction_dyn=friction, PFAC=0.4, DFAC=0.0, mesh_count=2.0,
                    mesh0_plane=np.array([2, 0, -side / 2 + sx, -side / 2 + sy, zmax, side / 2 + sx, side / 2 + sy, zmax, dens], float), mesh0_hfac=1.1, mesh0_id=10.0,
                    mesh0_vel=np.array([0, 0, -vmax, tau], float),
                    mesh1_plane=np.array([2, 1, -side / 2 + sx, -side / 2 + sy, zmin, side / 2 + sx, side / 2 + sy, zmin, dens], float), mesh1_hfac=1.1, mesh1_id=11.0,
                    mesh1_vel=np.array([0, 0, 0, 0], float))
    return case
