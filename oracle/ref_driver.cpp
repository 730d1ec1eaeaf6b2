// TEST INFRASTRUCTURE -- oracle only; never linked into, imported by, or executed from the product path.
//
// wf_ref: runs the UNMODIFIED reference solver (luchete80/WeldForm, /root/reference/Source, unity build through
// Domain.h) on a particle set described by a WFC1 case file and dumps every hot-path field as float64.
// Built by oracle/Makefile into oracle/_ref/wf_ref against the stand-in headers in oracle/shim/ (Blitz++/GSL/HDF5
// are not installed here). No reference source is copied: the reference files are #included where they lie.
//
//   usage: wf_ref <case.wfc> <out.wfc>
//
// What it drives (reference entry points): Domain::AddCylinderLength / AddBoxLength (Source/Domain.cpp:708,326),
// AddTrimeshParticles (Source/Contact.cpp:10), TriMesh::AxisPlaneMesh/CalcSpheres (Source/Mesh.cpp:82,206),
// CellInitiate/ListGenerate/MainNeighbourSearch (Source/Neighbour.cpp:10,141,346) and
// Domain::SolveDiffUpdateFraser (Source/SolverFraser.cpp:21). Boundary conditions are applied the way the example
// drivers do it, from the GeneralAfter callback (examples/Compression_thermal.cpp:30-66).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <algorithm>
#include <vector>

#include "Mesh.h"
#include "Domain.h"
#include "InteractionAlt.cpp"
#include "SolverKickDrift.cpp"
#include "SolverFraser.cpp"
#include "SolverLeapfrog.cpp"

#include "wfc_io.h"

using namespace SPH;

struct BCSpec { int id; double v[3]; double tau; int set_a; };
struct MeshVel { double v[3]; double tau; };
static std::vector<BCSpec> g_bc;
static std::vector<MeshVel> g_meshvel;

static double ramp(double t, double tau) { if (tau <= 0.0) return 1.0; return t < tau ? t / tau : 1.0; }

static void UserBC(SPH::Domain & domi)
{
    const double t = domi.getTime();
    for (size_t b = 0; b < g_bc.size(); ++b) {
        const BCSpec & bc = g_bc[b];
        const double f = ramp(t, bc.tau);
        for (size_t i = 0; i < domi.Particles.Size(); ++i) {
            if (domi.Particles[i]->ID == bc.id) {
                if (bc.set_a) domi.Particles[i]->a = Vec3_t(0.0, 0.0, 0.0);
                domi.Particles[i]->v = Vec3_t(bc.v[0] * f, bc.v[1] * f, bc.v[2] * f);
            }
        }
    }
    for (size_t m = 0; m < g_meshvel.size() && m < domi.trimesh.size(); ++m) {
        const double f = ramp(t, g_meshvel[m].tau);
        domi.trimesh[m]->SetVel(Vec3_t(g_meshvel[m].v[0] * f, g_meshvel[m].v[1] * f, g_meshvel[m].v[2] * f));
    }
}

static void put_vec3(wfc::File & out, const char * name, Domain & dom, Vec3_t SPH::Particle::*mem)
{
    size_t n = dom.Particles.Size(); std::vector<double> b(3 * n);
    for (size_t i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) b[3 * i + k] = (dom.Particles[i]->*mem)(k);
    out.put_f(name, b.data(), b.size());
}
static void put_mat3(wfc::File & out, const char * name, Domain & dom, Mat3_t SPH::Particle::*mem)
{
    size_t n = dom.Particles.Size(); std::vector<double> b(9 * n);
    for (size_t i = 0; i < n; ++i) for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) b[9 * i + 3 * r + c] = (dom.Particles[i]->*mem)(r, c);
    out.put_f(name, b.data(), b.size());
}
static void put_dbl(wfc::File & out, const char * name, Domain & dom, double SPH::Particle::*mem)
{
    size_t n = dom.Particles.Size(); std::vector<double> b(n);
    for (size_t i = 0; i < n; ++i) b[i] = dom.Particles[i]->*mem;
    out.put_f(name, b.data(), b.size());
}
static void put_int(wfc::File & out, const char * name, Domain & dom, int SPH::Particle::*mem)
{
    size_t n = dom.Particles.Size(); std::vector<int32_t> b(n);
    for (size_t i = 0; i < n; ++i) b[i] = dom.Particles[i]->*mem;
    out.put_i(name, b.data(), b.size());
}

static void dump_pairs(wfc::File & out, const char * name, Array<Array<std::pair<size_t, size_t> > > & P, bool sorted)
{
    std::vector<std::pair<int64_t, int64_t> > v;
    for (size_t k = 0; k < P.Size(); ++k) for (size_t a = 0; a < P[k].Size(); ++a) {
        int64_t i = (int64_t)P[k][a].first, j = (int64_t)P[k][a].second;
        if (sorted) v.push_back(std::make_pair(std::min(i, j), std::max(i, j))); else v.push_back(std::make_pair(i, j));
    }
    if (sorted) std::sort(v.begin(), v.end());
    std::vector<int64_t> flat(2 * v.size());
    for (size_t a = 0; a < v.size(); ++a) { flat[2 * a] = v[a].first; flat[2 * a + 1] = v[a].second; }
    out.put_l(name, flat.data(), flat.size());
}

static void zero_uninitialised(SPH::Particle * P)
{
    // Scalars the reference's Particle constructor leaves uninitialised (Source/Particle.cpp:28-158 never assigns
    // them) but reads before writing, e.g. T += dt*dTdt in the first ThermalCalcs (Source/Thermal.cpp:728). Give them
    // the value a fresh zero page would, so runs are reproducible.
    P->dTdt = 0.0; P->delta_cont = 0.0; P->q_fric_work = 0.0; P->dkin_energy_dt = 0.0; P->dint_energy_dt = 0.0;
    P->Et = 0.0; P->Et_m = 0.0; P->T = 0.0; P->k_T = 0.0; P->cp_T = 0.0; P->h_conv = 0.0; P->T_inf = 0.0;
}

int main(int argc, char ** argv) try
{
    if (argc < 3) { std::fprintf(stderr, "usage: wf_ref <case.wfc> <out.wfc>\n"); return 2; }
    wfc::File in;
    if (!in.read(argv[1])) { std::fprintf(stderr, "wf_ref: cannot read %s\n", argv[1]); return 2; }

    // the reference prints one line per step and per phase; keep the oracle quiet
    std::ofstream devnull("/dev/null");
    std::streambuf * cout_buf = std::cout.rdbuf(devnull.rdbuf());

    Domain dom;
    dom.Dimension = (int)in.scalar("Dimension", 3);
    dom.Nproc = (size_t)in.scalar("Nproc", 1);
    dom.Kernel_Set((Kernels_Type)(int)in.scalar("KernelType", 0));
    dom.Gradient_Approach_Set((Gradient_Type)(int)in.scalar("GradientType", 0));
    dom.Scheme = 1;
    dom.XSPH = in.scalar("XSPH", 0.0);
    dom.GeneralAfter = &UserBC;
    if (in.has("DomMax")) for (int k = 0; k < 3; ++k) dom.DomMax(k) = in.f("DomMax")[k];
    if (in.has("DomMin")) for (int k = 0; k < 3; ++k) dom.DomMin(k) = in.f("DomMin")[k];

    // ---------------------------------------------------------------- particles
    const int gen_kind = (int)in.scalar("gen_kind", 0);
    if (gen_kind == 1) {        // tag, V(3), Rxy, Lz, r, Density, h, Fixed, ghost
        const std::vector<double> & g = in.f("gen_args");
        dom.AddCylinderLength((int)g[0], Vec3_t(g[1], g[2], g[3]), g[4], g[5], g[6], g[7], g[8], g[9] != 0.0, g[10] != 0.0);
    } else if (gen_kind == 2) { // tag, V(3), Lx, Ly, Lz, r, Density, h, type, rotation, random, Fixed
        const std::vector<double> & g = in.f("gen_args");
        dom.AddBoxLength((int)g[0], Vec3_t(g[1], g[2], g[3]), g[4], g[5], g[6], g[7], g[8], g[9], (int)g[10], (int)g[11], g[12] != 0.0, g[13] != 0.0);
    } else {
        const std::vector<double> & x = in.f("x");
        const size_t n = x.size() / 3;
        const std::vector<double> & m = in.f("Mass"); const std::vector<double> & rho = in.f("Density"); const std::vector<double> & h = in.f("h");
        for (size_t i = 0; i < n; ++i) {
            int tag = in.has("ID") ? in.i("ID")[i] : 0;
            bool fixed = in.has("IsFree") ? (in.i("IsFree")[i] == 0) : false;
            dom.Particles.Push(new SPH::Particle(tag, Vec3_t(x[3 * i], x[3 * i + 1], x[3 * i + 2]), Vec3_t(0, 0, 0), m[i], rho[i], h[i], fixed));
        }
        dom.solid_part_count = (int)n;
    }
    const size_t nsph = dom.Particles.Size();
    for (size_t i = 0; i < nsph; ++i) zero_uninitialised(dom.Particles[i]);
    dom.contact_force_sum = 0.0; dom.contact_reaction_sum = 0.0; dom.max_disp = Vec3_t(0., 0., 0.);   // uninitialised in Domain::Domain (Source/Domain.cpp:78-175)

    // materials table: [model, E, nu, p0..p12] per row of 16
    std::vector<Material_ *> mats; std::vector<Elastic_ *> els;
    if (in.has("materials")) {
        const std::vector<double> & mt = in.f("materials");
        for (size_t r = 0; r + 16 <= mt.size(); r += 16) {
            Elastic_ * el = new Elastic_(mt[r + 1], mt[r + 2]); els.push_back(el);
            int model = (int)mt[r];
            if (model == HOLLOMON) mats.push_back(new Hollomon(*el, mt[r + 3], mt[r + 4], mt[r + 5]));
            else if (model == JOHNSON_COOK) mats.push_back(new JohnsonCook(*el, mt[r + 3], mt[r + 4], mt[r + 5], mt[r + 6], mt[r + 7], mt[r + 8], mt[r + 9], mt[r + 10]));
            else if (model == _GMT_) {       // GMT(el, n1, n2, C1, C2, m1, m2, I1, I2, e_min, e_max, er_min, er_max, T_min, T_max): p0..p12 + materials_ext (p13)
                const double tmax = in.has("materials_ext") ? in.f("materials_ext")[(r / 16) * 3] : 1.0e10;
                mats.push_back(new GMT(*el, mt[r + 3], mt[r + 4], mt[r + 5], mt[r + 6], mt[r + 7], mt[r + 8], mt[r + 9], mt[r + 10], mt[r + 11], mt[r + 12], mt[r + 13], mt[r + 14], mt[r + 15], tmax));
            }
            else mats.push_back(new Material_(*el));
        }
    }

#define OVR_D(NAME, MEM) if (in.has(NAME)) { const std::vector<double> & a_ = in.f(NAME); for (size_t i = 0; i < nsph; ++i) dom.Particles[i]->MEM = a_[i]; }
#define OVR_I(NAME, MEM) if (in.has(NAME)) { const std::vector<int32_t> & a_ = in.i(NAME); for (size_t i = 0; i < nsph; ++i) dom.Particles[i]->MEM = a_[i]; }
    if (gen_kind != 0) {
        if (in.has("x")) { const std::vector<double> & a_ = in.f("x"); for (size_t i = 0; i < nsph; ++i) for (int k = 0; k < 3; ++k) dom.Particles[i]->x(k) = a_[3 * i + k]; }
        OVR_D("Mass", Mass) OVR_D("Density", Density) OVR_D("h", h) OVR_I("ID", ID)
        if (in.has("IsFree")) { const std::vector<int32_t> & a_ = in.i("IsFree"); for (size_t i = 0; i < nsph; ++i) dom.Particles[i]->IsFree = (a_[i] != 0); }
    }
    if (in.has("v")) { const std::vector<double> & a_ = in.f("v"); for (size_t i = 0; i < nsph; ++i) for (int k = 0; k < 3; ++k) dom.Particles[i]->v(k) = a_[3 * i + k]; }
    OVR_D("RefDensity", RefDensity) OVR_D("G", G) OVR_D("Cs", Cs) OVR_D("P0", P0) OVR_I("PresEq", PresEq)
    OVR_D("Alpha", Alpha) OVR_D("Beta", Beta) OVR_D("TI", TI) OVR_D("TIn", TIn) OVR_D("TIInitDist", TIInitDist)
    OVR_D("Sigmay", Sigmay) OVR_D("Ep", Ep) OVR_I("Fail", Fail) OVR_I("Material", Material) OVR_I("Material_model", Material_model)
    OVR_D("k_T", k_T) OVR_D("cp_T", cp_T) OVR_D("T", T) OVR_D("h_conv", h_conv) OVR_D("T_inf", T_inf) OVR_I("Thermal_BC", Thermal_BC)
    OVR_D("th_exp", th_exp) OVR_D("q_source", q_source)
    if (in.has("not_write_surf_ID")) { const std::vector<int32_t> & a_ = in.i("not_write_surf_ID"); for (size_t i = 0; i < nsph; ++i) dom.Particles[i]->not_write_surf_ID = (a_[i] != 0); }
    if (in.has("mat_id")) { const std::vector<int32_t> & a_ = in.i("mat_id"); for (size_t i = 0; i < nsph; ++i) if (a_[i] >= 0 && (size_t)a_[i] < mats.size()) dom.Particles[i]->mat = mats[a_[i]]; }

    // ---------------------------------------------------------------- rigid meshes + contact
    const int mesh_count = (int)in.scalar("mesh_count", 0);
    std::vector<TriMesh *> meshes;
    for (int m = 0; m < mesh_count; ++m) {
        char key[64];
        std::snprintf(key, 64, "mesh%d_plane", m); const std::vector<double> & pl = in.f(key);   // axis, posit, p1(3), p2(3), dens
        TriMesh * tm = new TriMesh(); meshes.push_back(tm);
        tm->AxisPlaneMesh((int)pl[0], pl[1] != 0.0, Vec3_t(pl[2], pl[3], pl[4]), Vec3_t(pl[5], pl[6], pl[7]), (int)pl[8]);
        tm->CalcSpheres();
        std::snprintf(key, 64, "mesh%d_hfac", m); float hfac = (float)in.scalar(key, 1.1);
        std::snprintf(key, 64, "mesh%d_id", m); int id = (int)in.scalar(key, 10 + m);
        dom.AddTrimeshParticles(tm, hfac, id);
        dom.m_contact_force.push_back(Vec3_t(0.0, 0.0, 0.0));   // as Source/WeldForm.cpp:684 does; examples rely on UB here
        MeshVel mv; mv.v[0] = mv.v[1] = mv.v[2] = 0.0; mv.tau = 0.0;
        std::snprintf(key, 64, "mesh%d_vel", m);
        if (in.has(key)) { const std::vector<double> & a_ = in.f(key); mv.v[0] = a_[0]; mv.v[1] = a_[1]; mv.v[2] = a_[2]; mv.tau = a_[3]; }
        g_meshvel.push_back(mv);
        std::snprintf(key, 64, "mesh%d_w", m);
        if (in.has(key)) { const std::vector<double> & w_ = in.f(key); tm->SetRotAxisVel(Vec3_t(w_[0], w_[1], w_[2])); }
    }
    for (size_t i = nsph; i < dom.Particles.Size(); ++i) zero_uninitialised(dom.Particles[i]);
    dom.contact = in.scalar("contact", mesh_count > 0 ? 1 : 0) != 0.0;
    dom.friction_sta = in.scalar("friction_sta", 0.0);
    dom.friction_dyn = in.scalar("friction_dyn", 0.0);
    if ((int)in.scalar("friction_function", 0) == 1) { dom.friction_function = Linear; dom.friction_m = in.scalar("friction_m", 0.0); dom.friction_b = in.scalar("friction_b", 0.0); }
    dom.PFAC = in.scalar("PFAC", 0.4);
    dom.DFAC = in.scalar("DFAC", 0.0);
    dom.cont_heat_fric = in.scalar("cont_heat_fric", 0) != 0.0;
    dom.cont_heat_cond = in.scalar("cont_heat_cond", 0) != 0.0;
    dom.gradKernelCorr = in.scalar("gradKernelCorr", 0) != 0.0;
    dom.ts_nb_inc = (int)in.scalar("ts_nb_inc", 5);
    dom.thermal_solver = in.scalar("thermal_solver", 0) != 0.0;
    dom.pl_heating = in.scalar("pl_heating", 0) != 0.0;
    dom.pl_work_heat_frac = in.scalar("pl_work_heat_frac", 0.9);
    dom.BC.InOutFlow = 0;
    dom.auto_ts = in.scalar("auto_ts", 0) != 0.0;      // honoured by the KickDrift / LeapFrog solvers only (SolverFraser.cpp:500-502 is commented out)
    dom.CFL = in.scalar("CFL", 0.7);
    dom.auto_ts_acc = in.scalar("auto_ts_acc", 0) != 0.0;

    if (in.has("bc")) {
        const std::vector<double> & b = in.f("bc");
        for (size_t r = 0; r + 8 <= b.size(); r += 8) {
            BCSpec s; s.id = (int)b[r]; s.v[0] = b[r + 1]; s.v[1] = b[r + 2]; s.v[2] = b[r + 3]; s.tau = b[r + 4]; s.set_a = (int)b[r + 5];
            g_bc.push_back(s);
        }
    }

    const int mode = (int)in.scalar("mode", 0);
    const double dt = in.scalar("dt", 1.0e-7);
    const int nsteps = (int)in.scalar("nsteps", 1);

    wfc::File out;
    double solve_seconds = 0.0;
    if (mode == 1) {
        // neighbour build only: same call sequence as the solver prologue (Source/SolverFraser.cpp:42-43,87)
        dom.CellInitiate();
        dom.ListGenerate();
        dom.MainNeighbourSearch();
        dom.SaveNeighbourData();
        double cell[9] = { dom.BLPF(0), dom.BLPF(1), dom.BLPF(2), dom.CellSize(0), dom.CellSize(1), dom.CellSize(2),
                           (double)dom.CellNo[0], (double)dom.CellNo[1], (double)dom.CellNo[2] };
        out.put_f("cell_grid", cell, 9);
        dump_pairs(out, "pairs_discovery", dom.SMPairs, false);
    } else {
        // run exactly nsteps iterations of while (Time <= tf) (Source/SolverFraser.cpp:133)
        double tf = (nsteps - 0.5) * dt;
        if (in.scalar("tf", 0.0) > 0.0) tf = in.scalar("tf", 0.0);     // adaptive time step: the loop ends on Time, not on a step count
        const int solver = (int)in.scalar("solver", 0);
        auto t0 = std::chrono::steady_clock::now();
        if (solver == 1) dom.SolveDiffUpdateKickDrift(tf, dt, 1.0e30, "wfref", 1000000);
        else if (solver == 2) dom.SolveDiffUpdateLeapFrog(tf, dt, 1.0e30, "wfref", 1000000);
        else if (solver == 3) dom.ThermalSolve(tf, dt, 1.0e30, "wfref", 1000000);     // while (Time < tf): the same nsteps with tf = (nsteps - 0.5) dt
        else dom.SolveDiffUpdateFraser(tf, dt, 1.0e30, "wfref", 1000000);
        solve_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    std::cout.rdbuf(cout_buf);

    // ---------------------------------------------------------------- dump
    out.put_scalar("n_particles", (double)dom.Particles.Size());
    out.put_scalar("solid_part_count", (double)dom.solid_part_count);
    out.put_scalar("Time", dom.getTime());
    out.put_scalar("solve_seconds", solve_seconds);
    out.put_scalar("deltat", dom.getStepSize());
    out.put_scalar("kin_energy_sum", dom.kin_energy_sum);
    out.put_scalar("int_energy_sum", dom.int_energy_sum);
    out.put_scalar("plastic_work", dom.plastic_work);
    out.put_scalar("contact_force_sum", dom.contact_force_sum);
    out.put_scalar("contact_reaction_sum", dom.contact_reaction_sum);
    out.put_scalar("ext_forces_work", dom.ext_forces_work);
    out.put_scalar("contact_friction_work", dom.contact_friction_work);
    { double md[3] = { dom.max_disp(0), dom.max_disp(1), dom.max_disp(2) }; out.put_f("max_disp", md, 3); }
    put_vec3(out, "x", dom, &SPH::Particle::x); put_vec3(out, "v", dom, &SPH::Particle::v); put_vec3(out, "a", dom, &SPH::Particle::a);
    put_vec3(out, "Displacement", dom, &SPH::Particle::Displacement); put_vec3(out, "normal", dom, &SPH::Particle::normal);
    put_vec3(out, "contforce", dom, &SPH::Particle::contforce);
    put_dbl(out, "Mass", dom, &SPH::Particle::Mass); put_dbl(out, "h", dom, &SPH::Particle::h);
    put_dbl(out, "Density", dom, &SPH::Particle::Density); put_dbl(out, "dDensity", dom, &SPH::Particle::dDensity);
    put_dbl(out, "Pressure", dom, &SPH::Particle::Pressure);
    put_mat3(out, "Sigma", dom, &SPH::Particle::Sigma); put_mat3(out, "ShearStress", dom, &SPH::Particle::ShearStress);
    put_mat3(out, "StrainRate", dom, &SPH::Particle::StrainRate); put_mat3(out, "RotationRate", dom, &SPH::Particle::RotationRate);
    put_mat3(out, "Strain", dom, &SPH::Particle::Strain); put_mat3(out, "Strain_pl", dom, &SPH::Particle::Strain_pl);
    put_mat3(out, "TIR", dom, &SPH::Particle::TIR);
    if (dom.gradKernelCorr) put_mat3(out, "gradCorrM", dom, &SPH::Particle::gradCorrM);
    put_dbl(out, "pl_strain", dom, &SPH::Particle::pl_strain); put_dbl(out, "Sigmay", dom, &SPH::Particle::Sigmay);
    put_dbl(out, "eff_strain_rate", dom, &SPH::Particle::eff_strain_rate);
    put_dbl(out, "T", dom, &SPH::Particle::T); put_dbl(out, "dTdt", dom, &SPH::Particle::dTdt);
    put_dbl(out, "q_plheat", dom, &SPH::Particle::q_plheat); put_dbl(out, "q_conv", dom, &SPH::Particle::q_conv);
    put_dbl(out, "delta_cont", dom, &SPH::Particle::delta_cont);
    put_dbl(out, "dkin_energy_dt", dom, &SPH::Particle::dkin_energy_dt); put_dbl(out, "dint_energy_dt", dom, &SPH::Particle::dint_energy_dt);
    put_int(out, "ID", dom, &SPH::Particle::ID); put_int(out, "Nb", dom, &SPH::Particle::Nb);
    dump_pairs(out, "pairs", dom.SMPairs, true);
    dump_pairs(out, "cont_pairs", dom.ContPairs, true);
    for (int m = 0; m < mesh_count; ++m) {
        char key[64]; TriMesh * tm = meshes[m];
        std::vector<double> nodes(3 * tm->node.Size()), nrm(3 * tm->element.Size()), pp(tm->element.Size());
        for (size_t n = 0; n < tm->node.Size(); ++n) for (int k = 0; k < 3; ++k) nodes[3 * n + k] = (*tm->node[n])(k);
        for (size_t e = 0; e < tm->element.Size(); ++e) { for (int k = 0; k < 3; ++k) nrm[3 * e + k] = tm->element[e]->normal(k); pp[e] = tm->element[e]->pplane; }
        std::snprintf(key, 64, "mesh%d_nodes", m); out.put_f(key, nodes.data(), nodes.size());
        std::snprintf(key, 64, "mesh%d_normals", m); out.put_f(key, nrm.data(), nrm.size());
        std::snprintf(key, 64, "mesh%d_pplane", m); out.put_f(key, pp.data(), pp.size());
    }
    if (!out.write(argv[2])) { std::fprintf(stderr, "wf_ref: cannot write %s\n", argv[2]); return 2; }
    return 0;
}
catch (Fatal * f) { f->Cout(); delete f; return 1; }
catch (char const * m) { std::fprintf(stderr, "wf_ref: fatal: %s\n", m); return 1; }
