// Host-side mirror of the reference's SPH::Domain source API (Source/Domain.h:96-470) on top of the weldform_b200 C ABI.
//
// The reference is a header-only unity build whose drivers (examples/*.cpp) construct an SPH::Domain, fill
// dom.Particles[a]->field, install a GeneralAfter callback and call dom.SolveDiffUpdateFraser(...). This header keeps
// that surface -- same class, member and file names, same argument meaning, errors as `throw new Fatal(...)` -- so those
// drivers compile and run unchanged with   -I weldform_b200/host -I include -L weldform_b200 -lweldform_b200.
// The per-timestep work (Source/SolverFraser.cpp:133-505) is NOT done here: every phase is a call into
// libweldform_b200.so (include/weldform_b200.h), i.e. hand-written sm_100a CUDA kernels. There is no CPU fallback: without
// the library or a CUDA device, Solve throws.
//
// What stays on the host: lattice generation (through the library's wf_gen_* helpers), packing the per-particle mirrors
// into structure-of-arrays uploads, the solver loop's control flow, the GeneralAfter/GeneralBefore callback bridge, output.
//
// Callback bridge. GeneralAfter runs twice per step on host Particle objects (SolverFraser.cpp:332,360). Environment
// variable WELDFORM_SYNC selects what is mirrored around each call:
//   full    (default) x, v, a, Sigma, T, pl_strain, Displacement, Density, Pressure, Sigmay, ID of ALL particles are downloaded before the
//           callback. After it v and a are uploaded, and so is every other member the device accepts (x, T, Density, Sigmay, ID) that the
//           callback changed (compared with the downloaded values); a change to a member the device cannot take back (Sigma, pl_strain,
//           Displacement, Pressure) throws a Fatal instead of being dropped silently;
//   tagged  the same, for the particles whose ID differs from the most common ID only (the boundary-condition rows of the usual drivers,
//           e.g. ID 2/3 in examples/Compression_thermal.cpp:30-66) -- O(surface) traffic instead of O(N);
//   push    like tagged, but nothing is read back before the callback: for callbacks that only ASSIGN whole v and a vectors of their
//           boundary rows (examples/Compression_thermal.cpp:30-66 does); v and a of the tagged rows are uploaded after it (one vector and a
//           fill kernel per ID group when the group's rows all got the same values). The callback does not wait for the device: getTime()
//           is advanced on the host (Time += deltat, the device's own arithmetic), max_disp / contact_force_sum / the energy sums are
//           refreshed every 32 steps and before every output, so the callback overlaps the GPU's work on the step;
//   none    the callback only sees scalars (time, max_disp, contact_force_sum) and trimesh[m]->SetVel.
// Rigid-mesh velocities set from the callback (trimesh[m]->SetVel / SetRotAxisVel) are always forwarded.
// Two more environment knobs, for driver tests that cannot afford a driver's full run: WELDFORM_MAX_STEPS=n and WELDFORM_MAX_OUTPUTS=k end
// the solver loop after n steps / after the k-th output table, as if tf had been reached.
#ifndef WF_HOST_DOMAIN_H
#define WF_HOST_DOMAIN_H

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <new>
#include <sys/mman.h>
#include <utility>
#include <sstream>
#include <string>
#include <vector>

#include "weldform_b200.h"
#include "Backend.h"
#include "matvec.h"
#include "Particle.h"
#include "Mesh.h"

using std::cout;
using std::endl;
using std::string;

// the AoS <-> SoA loops of the callback bridge run on Domain::Nproc host threads when the driver is compiled with OpenMP (the reference's
// drivers are: their callbacks carry `#pragma omp parallel for ... num_threads(domi.Nproc)`); plain loops otherwise
#ifdef _OPENMP
#define WF_HOST_PRAGMA(x) _Pragma(#x)
#define WF_HOST_PARALLEL_FOR WF_HOST_PRAGMA(omp parallel for schedule(static) num_threads((int)(Nproc > 0 ? Nproc : 1)))
#else
#define WF_HOST_PARALLEL_FOR
#endif

enum domain_bid_type { PlaneStress = 0, PlaneStrain = 1, AxiSymmetric = 2, AxiSymm_3D = 2 };
enum Kernels_Type { Qubic_Spline = 0, Quintic = 1, Quintic_Spline = 2, Hyperbolic_Spline = 3 };
enum Viscosity_Eq_Type { Morris = 0, Shao = 1, Incompressible_Full = 2, Takeda = 3 };
enum Gradient_Type { Squared_density = 0, Multiplied_density = 1, Randles_Libersky = 2 };
enum Damage_Model { None = 0, Rankine = 1 };
enum Function_Type { Constant = 0, Linear = 1, Multilinear = 2 };
enum Friction_Type { Fr_Sta = 0, Fr_Dyn, Fr_StaDyn, Fr_Bound };
enum Contact_Alg { Fraser = 0, Wang, Seo, Zhan, LSDyna };

namespace SPH {

class Domain;
typedef void (*PtDom)(Domain & dom);

struct Boundary { int InOutFlow = 0; bool Periodic[3] = { false, false, false }; };
struct iKernel { iKernel() {} iKernel(size_t, double) {} };     // the device evaluates the kernel; kept so that `dom.m_kernel = SPH::iKernel(dim, h)` compiles

class Domain {
public:
    typedef void (*PtVel)(Vec3_t & position, Vec3_t & Vel, double & Den, Boundary & bdry);

    Domain();
    ~Domain();

    // ---- geometry (Source/Domain.cpp:326-455, 708-860; Source/Contact.cpp:10-38)
    void AddBoxLength(int tag, Vec3_t const & V, double Lx, double Ly, double Lz, double r, double Density, double h, int type, int rotation, bool random, bool Fixed);
    void AddCylinderLength(int tag, Vec3_t const & V, double Rxy, double Lz, double r, double Density, double h, bool Fixed, bool ghost = false);
    void AddTrimeshParticles(TriMesh * mesh, const float & hfac, const int & id);
    void AddSingleParticle(int tag, Vec3_t const & x, double Mass, double Density, double h, bool Fixed);
    void Add3DCubicBoxParticles(int tag, Vec3_t const & V, double Lx, double Ly, double Lz, double r, double Density, double h);
    void Calculate3DMass(double Density);
    void AddTractionProbeLength(int tag, Vec3_t const & V, double Rxy, double Lz_side, double Lz_neckmin, double Lz_necktot, double Rxy_center,
                                double r, double Density, double h, bool Fixed);
    void AddCylUniformLength(int tag, double Rxy, double Lz, double r, double Density, double h, double ang = 2 * M_PI, int rows = 1, double r_i = 0.0, bool ghost = true);
    void AddXYSymCylinderLength(int tag, double Rxy, double Lz, double r, double Density, double h, bool Fixed, bool symlength = false);
    void Kernel_Set(Kernels_Type const & KT) { KernelType = KT; Cellfac = (KT == Quintic_Spline) ? 3.0 : 2.0; }
    void Gradient_Approach_Set(Gradient_Type const & GT) { GradientType = GT; }
    void Viscosity_Eq_Set(Viscosity_Eq_Type const & VQ) { VisEq = VQ; }

    // ---- solvers (Source/Domain.h:197-210): the Fraser, KickDrift and LeapFrog loops and ThermalSolve run on the GPU
    void SolveDiffUpdateFraser(double tf, double dt, double dtOut, char const * TheFileKey, size_t maxidx);
    void SolveDiffUpdateKickDrift(double tf, double dt, double dtOut, char const * k, size_t m) { solve_kick_drift_leap_frog(1, tf, dt, dtOut, k, m); }
    void SolveDiffUpdateLeapFrog(double tf, double dt, double dtOut, char const * k, size_t m) { solve_kick_drift_leap_frog(2, tf, dt, dtOut, k, m); }
    void solve_kick_drift_leap_frog(int solver, double tf, double dt, double dtOut, char const * TheFileKey, size_t maxidx);
    void SolveDiffUpdateLeapfrog(double tf, double dt, double dtOut, char const * k, size_t m) { SolveDiffUpdateLeapFrog(tf, dt, dtOut, k, m); }
    // the reference's Verlet loop calls its reductions under `#ifdef NONLOCK_SUM`, a macro nothing defines (SolverVerlet.cpp:176-178,
    // 206-208): with the default nonlock_sum = true no pair sum ever reaches the particles there. Not restated.
    void SolveDiffUpdateVerlet(double, double, double, char const *, size_t) { not_implemented("SolveDiffUpdateVerlet"); }
    void ThermalSolve(double tf, double dt, double dtOut, char const * TheFileKey, size_t maxidx);
    void ThermalStructSolve(double, double, double, char const *, size_t) { not_implemented("ThermalStructSolve"); }
    void Solve(double, double, double, char const *, size_t) { throw new Fatal("Domain::Solve: DEPRECATED (the reference throws here too, Source/Domain.cpp:2130); use SolveDiffUpdateFraser"); }

    // ---- neighbour search as the reference's component drivers call it (examples/Neighbour_Test.cpp:84-118): CellInitiate uploads the
    // particles and sizes the grid on the device (Neighbour.cpp:10-139), MainNeighbourSearch builds the lists there and exports the pair set
    // into SMPairs[0] as (i, j) with i < j, sorted -- the reference fills one list per thread in discovery order; the SET is the same
    // (tests/test_parity_gpu.py::test_pairs_bit_exact_12k). ListGenerate / CellReset are part of the device build.
    void CellInitiate();
    void ListGenerate() {}
    void CellReset() {}
    void MainNeighbourSearch();
    void SaveNeighbourData();
    void ClearNbData() { for (size_t k = 0; k < SMPairs.Size(); ++k) SMPairs[k].Clear(); }
    int AvgNeighbourCount();
    Array<Array<std::pair<size_t, size_t> > > SMPairs, NSMPairs, FSMPairs;     // Domain.h:405-407
    Vec3_t BLPF, TRPR, CellSize; int CellNo[3] = { 0, 0, 0 };                  // Domain.h:323-330, valid after CellInitiate

    // ---- output
    void WriteXDMF(char const * FileKey);
    void WriteCSV(char const * FileKey);
    double getTime() { return Time; }

    // ---- data members drivers set / read (defaults: Source/Domain.cpp:78-175)
    int Dimension = 3;
    size_t Nproc = 1;                          // host threads of the reference; unused (the step runs on the GPU)
    size_t Scheme = 0;
    double CFL = 0.7, XSPH = 0.0, deltat = 0.0, deltatmin = 0.0, Alpha = 0.0, Beta = 0.0;
    bool auto_ts = false, auto_ts_acc = false; int solver_kind = 0;
    int ts_nb_inc = 5;
    Vec3_t DomMax, DomMin, Gravity;
    PtDom GeneralBefore = nullptr, GeneralAfter = nullptr;
    Boundary BC;
    iKernel m_kernel;
    size_t KernelType = 0, GradientType = 0, VisEq = 0;
    double Cellfac = 2.0, hmax = 0.0;
    bool thermal_solver = false, pl_heating = false, gradKernelCorr = false, nonlock_sum = true;
    double pl_work_heat_frac = 0.9;
    bool contact = false, update_contact_surface = true, cont_heat_gen = false, cont_heat_cond = false, cont_heat_fric = false, model_damage = false;
    double PFAC = 0.4, DFAC = 0.0, friction_sta = 0.0, friction_dyn = 0.0, friction_m = 0.0, friction_b = 0.0, contact_hc = 0.0;
    double friction = 0.0;                     // Domain.h:291: legacy coefficient of the Fraser / Seo contact laws; CalcContactForcesWang reads friction_sta / friction_dyn only
    int friction_function = Constant, fric_type = Fr_Dyn, contact_alg = Wang;
    int id_free_surf = 1;
    std::vector<int> contact_surf_id;
    std::vector<TriMesh *> trimesh;
    std::vector<Vec3_t> m_contact_force;
    int meshcount = 0, solid_part_count = 0;
    std::vector<int> first_fem_particle_idx;                       // Domain.h:348: first element particle of every rigid mesh (Contact.cpp:12)
    Array<Particle *> Particles;
    Array<std::pair<size_t, size_t> > GhostPairs;                  // Domain.h:473 (inner particle, ghost particle)
    double m_scalar_prop = 0.0, Vol = 0.0, R = 0.0;          // Vol: accumulated by Add3DCubicBoxParticles for Calculate3DMass (Domain.h:337); R: particle radius of the last lattice
    double sqrt_h_a = 0.0025;                                // Domain.cpp:109 (CheckMinTSAccel)
    Vec3_t max_disp;
    double contact_force_sum = 0.0, contact_reaction_sum = 0.0, ext_forces_work = 0.0, plastic_work = 0.0, accum_cont_heat_cond = 0.0,
           contact_friction_work = 0.0, kin_energy_sum = 0.0, int_energy_sum = 0.0, max_contact_force = 0.0, m_maxT = 0.0, m_minT = 0.0;
    size_t steps = 0;

    // ---- weldform_b200 specifics
    wf_ctx * backend = nullptr;                // rank 0's context; live between the start and the end of a Solve call
    wf_host::Backend bk;                       // all GPUs of the run (host/Backend.h)
    int device = 0;                            // first CUDA device (WELDFORM_DEVICE)
    int gpus = 1;                              // GPUs of one node this Domain runs on (WELDFORM_GPUS): slab decomposition, one host thread per GPU
    double wall_seconds = 0.0;                 // wall time of the last Solve's step loop
    // benchmark hook of the Fraser loop (bench.py's e2e leg, examples/bench_domain.cpp): once the body yields (the rebuild cadence is then the
    // reference's every-ts_nb_inc-steps one) and at least bench_spin_min steps have run, bench_warmup untimed steps, then bench_steps steps
    // timed on the host clock between two wf_sync calls -> bench_seconds; the loop ends there. bench_steps = 0: no effect.
    size_t bench_steps = 0, bench_warmup = 0, bench_spin_min = 20; double bench_seconds = 0.0; size_t bench_first_step = 0; bool isyielding = false, bench_wait_yield = true;

private:
    double Time = 0.0;
    enum SyncMode { SYNC_FULL = 0, SYNC_TAGGED = 1, SYNC_NONE = 2, SYNC_PUSH = 3 } sync_mode = SYNC_FULL;
    std::vector<int32_t> sync_idx;             // creation indices mirrored around callbacks (empty + SYNC_FULL: everything)
    struct PushGroup { int id; std::vector<int32_t> idx; int handle; std::vector<double> v, a; };   // SYNC_PUSH: the tagged rows by ID, registered on the device(s)
    int sync_group = -1;                       // SYNC_TAGGED: sync_idx as a registered group
    std::vector<PushGroup> push_groups;
    std::vector<double> buf3, buf9, buf1;
    // values the bridge handed to the last callback (FULL / TAGGED): what sync_up compares against
    std::vector<double> snap_x, snap_T, snap_dens, snap_sigy, snap_sigma, snap_pls, snap_disp, snap_pres; std::vector<int32_t> snap_id;
    bool have_snap = false;
    long long max_steps_env = -1, max_outputs_env = -1;
    std::ofstream history_of, prop_of;
    std::vector<Vec3_t> sent_mv, sent_mw;
    struct MeshRec { TriMesh * mesh; float hfac; int id; };
    std::vector<MeshRec> mesh_recs;

    // Particles created by the lattice generators live in one arena, contiguous in creation order: a driver's `for (i...) Particles[i]->ID`
    // loop then walks memory with a constant stride instead of chasing ten million heap pointers
    // (2 MB-aligned blocks advised as transparent huge pages: ten million particles are 11 GB, and with 4 kB pages a stride-1.1-kB walk over them
    // pays a TLB miss every fourth particle)
    std::vector<void *> pool_blocks; size_t pool_used = 0; static const size_t POOL_BLOCK = 16384;
    static void * pool_block_alloc()
    {
        const size_t huge = (size_t)2 << 20, bytes = ((POOL_BLOCK * sizeof(Particle) + huge - 1) / huge) * huge;
        void * p = nullptr;
        if (posix_memalign(&p, huge, bytes) != 0 || !p) throw std::bad_alloc();
#ifdef MADV_HUGEPAGE
        madvise(p, bytes, MADV_HUGEPAGE);
#endif
        return p;
    }
    // one arena for a whole lattice whose size the generator knows, FIRST TOUCHED IN PARALLEL with the static partition the callbacks' own
    // `omp parallel for schedule(static)` loops use: on a two-socket host each thread's share of the particles then sits in its own socket's memory
    char * pool_arena = nullptr; size_t pool_arena_cap = 0, pool_arena_used = 0;
    void pool_reserve(size_t n)
    {
        if (n < 65536 || pool_arena_cap - pool_arena_used >= n) return;
        const size_t huge = (size_t)2 << 20, bytes = ((n * sizeof(Particle) + huge - 1) / huge) * huge;
        void * p = nullptr;
        if (posix_memalign(&p, huge, bytes) != 0 || !p) return;
#ifdef MADV_HUGEPAGE
        madvise(p, bytes, MADV_HUGEPAGE);
#endif
        char * base = (char *)p;
        WF_HOST_PARALLEL_FOR
        for (size_t k = 0; k < n; ++k) std::memset(base + k * sizeof(Particle), 0, sizeof(Particle));
        pool_blocks.push_back(p); pool_arena = base; pool_arena_cap = n; pool_arena_used = 0;
    }
    template <typename... A> Particle * pool_make(A && ... a)
    {
        void * at;
        if (pool_arena_used < pool_arena_cap) at = pool_arena + (pool_arena_used++) * sizeof(Particle);
        else {
            if (pool_small == nullptr || pool_used == POOL_BLOCK) { pool_small = (char *)pool_block_alloc(); pool_blocks.push_back(pool_small); pool_used = 0; }
            at = pool_small + (pool_used++) * sizeof(Particle);
        }
        Particle * p = new (at) Particle(std::forward<A>(a)...);
        p->wf_pooled = true;
        return p;
    }
    char * pool_small = nullptr;
    void not_implemented(const char * what) { throw new Fatal("Domain::%s is not implemented in weldform_b200 (SURVEY.md section 8f, next rows); the shipped configurations use SolveDiffUpdateFraser", what); }
    void ck(int rc, const char * what) { if (rc != 0) throw new Fatal("%s failed (%d) at step %d, time %g: %s", what, rc, (int)steps, Time, wf_last_error(backend)); }
    void backend_begin(double dt);
    void backend_end();
    void choose_sync();
    void get3(wf_field f, Vec3_t Particle::*mem, bool subset);
    void get9(wf_field f, Mat3_t Particle::*mem, bool subset);
    void get1(wf_field f, double Particle::*mem, bool subset);
    void put3(wf_field f, Vec3_t Particle::*mem, bool subset);
    void put1_if_changed(wf_field f, double Particle::*mem, std::vector<double> & snap, bool subset);
    void refuse_if_changed(const char * what, const std::vector<double> & snap, int nc, bool subset, double (*get)(const Particle *, int));
    double max_pl_final() { wf_scalars s; ck(wf_reduce(backend, &s), "wf_reduce"); return s.max_pl_strain; }
    void open_history();
    void write_history();
    bool env_stop(size_t idx_out) const { return (max_steps_env >= 0 && (long long)steps >= max_steps_env) || (max_outputs_env >= 0 && (long long)idx_out > max_outputs_env); }
    void sync_down(bool everything);
    void sync_up();
    void forward_mesh_motion();
    void refresh_scalars();
    void run_callback(PtDom cb);
};

// =====================================================================================================================
inline Domain::Domain()
{
    DomMax = Vec3_t(-100000000000.0); DomMin = Vec3_t(100000000000.0);       // Source/Domain.cpp:118-119
    const char * dv = std::getenv("WELDFORM_DEVICE"); if (dv) device = std::atoi(dv);
    const char * ng = std::getenv("WELDFORM_GPUS"); if (ng && std::atoi(ng) > 0) gpus = std::atoi(ng);
}
inline Domain::~Domain()
{
    bk.destroy(); backend = nullptr;
    for (size_t i = 0; i < Particles.Size(); ++i) { if (Particles[i]->wf_pooled) Particles[i]->~Particle(); else delete Particles[i]; }   // Domain owns its particles (Source/Domain.cpp:177-181)
    for (void * b : pool_blocks) std::free(b);
}

// Source/Domain.cpp:708-860 (3-D). Points come from the library's generator, mass rule :839-841. ghost = true (:795-826): three mirror rows
// below the bottom plane, GhostPairs (inner, ghost), ghost_plane_axis 2, not_write_surf_ID; Domain::MoveGhost runs on the device.
inline void Domain::AddCylinderLength(int tag, Vec3_t const & V, double Rxy, double Lz, double r, double Density, double h, bool Fixed, bool ghost)
{
    if (Dimension != 3) throw new Fatal("AddCylinderLength: only Dimension = 3 is implemented in weldform_b200");
    const double Vd[3] = { V(0), V(1), V(2) };
    int64_t npairs = 0;
    const int64_t n = ghost ? wf_gen_cylinder_ghost(Vd, Rxy, Lz, r, nullptr, 0, nullptr, 0, &npairs) : wf_gen_cylinder(Vd, Rxy, Lz, r, nullptr, 0);
    if (n < 0) throw new Fatal("AddCylinderLength: bad arguments");
    std::vector<double> pts(3 * (size_t)n); std::vector<int32_t> pairs(2 * (size_t)npairs);
    if (ghost) wf_gen_cylinder_ghost(Vd, Rxy, Lz, r, pts.data(), n, pairs.data(), npairs, &npairs); else wf_gen_cylinder(Vd, Rxy, Lz, r, pts.data(), n);
    const size_t first = Particles.Size();
    if (ghost && first != 0) throw new Fatal("AddCylinderLength: ghost rows need the cylinder to be the first particle set (the reference indexes the first rows from particle 0, Source/Domain.cpp:801-806)");
    std::cout << "\n--------------Generating particles by AddCylinderLength with defined length of particles-----------" << std::endl;
    pool_reserve((size_t)n);
    for (int64_t i = 0; i < n; ++i) Particles.Push(pool_make(tag, Vec3_t(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), Vec3_t(0, 0, 0), 0.0, Density, h, Fixed));
    for (int64_t g = 0; g < npairs; ++g) {
        Particle * P = Particles[pairs[2 * g + 1]];
        P->inner_mirr_part = pairs[2 * g]; P->ghost_plane_axis = 2; P->not_write_surf_ID = true;
        GhostPairs.Push(std::make_pair((size_t)pairs[2 * g], (size_t)pairs[2 * g + 1]));
    }
    const double Vol = M_PI * Rxy * Rxy * Lz;
    const double Mass = Vol * Density / (double)(Particles.Size() - first);
    std::cout << "Total Particle count: " << Particles.Size() << std::endl;
    std::cout << "Particle mass: " << Mass << std::endl;
    for (size_t i = first; i < Particles.Size(); ++i) Particles[i]->Mass = Mass;
    solid_part_count = (int)Particles.Size();
}

// Source/Domain.cpp:321-324 (solid_part_count is left alone there; the mirror's upload needs it, so it follows the particle count)
inline void Domain::AddSingleParticle(int tag, Vec3_t const & x, double Mass, double Density, double h, bool Fixed)
{
    Particles.Push(pool_make(tag, x, Vec3_t(0, 0, 0), Mass, Density, h, Fixed));
    solid_part_count = (int)Particles.Size();
}

// Source/Domain.cpp:601-655: the cubic lattice of AddBoxLength without the mass rule; the box volume accumulates in Vol for Calculate3DMass
inline void Domain::Add3DCubicBoxParticles(int tag, Vec3_t const & V, double Lx, double Ly, double Lz, double r, double Density, double h)
{
    if (Dimension != 3) throw new Fatal("Add3DCubicBoxParticles: only Dimension = 3 is implemented in weldform_b200");
    const double Vd[3] = { V(0), V(1), V(2) }; double dims[3];
    const int64_t n = wf_gen_box(Vd, Lx, Ly, Lz, r, nullptr, 0, dims);
    if (n < 0) throw new Fatal("Add3DCubicBoxParticles: bad arguments");
    std::vector<double> pts(3 * (size_t)n);
    wf_gen_box(Vd, Lx, Ly, Lz, r, pts.data(), n, dims);
    std::cout << "\n--------------Generating particles by AddBoxLength with defined length of particles-----------" << std::endl;
    for (int64_t i = 0; i < n; ++i) Particles.Push(pool_make(tag, Vec3_t(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), Vec3_t(0, 0, 0), 0.0, Density, h, false));
    std::cout << "BoxDimensions: " << dims[0] << ", " << dims[1] << ", " << dims[2] << ", " << std::endl;
    Vol += dims[0] * dims[1] * dims[2];
    solid_part_count = (int)Particles.Size();
}
// Source/Domain.cpp:658-672
inline void Domain::Calculate3DMass(double Density)
{
    double Mass = Vol * Density / Particles.Size();
    std::cout << "Particle Mass: " << Mass << std::endl;
    for (size_t i = 0; i < Particles.Size(); ++i) Particles[i]->Mass = Mass;
}

namespace detail {
// Source/Domain.cpp:674-689 calcHalfPartCount
inline int calc_half_part_count(const double & r, const double & R, const int xinc)
{
    int ypartcount = -1;
    if (xinc > 0) {
        ypartcount = 1;
        double yp = r;
        double xp = r + (double)(xinc - 1) * 2. * r;
        double rad = sqrt(yp * yp + xp * xp);
        while (rad <= R - r) { yp += 2. * r; rad = sqrt(yp * yp + xp * xp); ypartcount++; }
        ypartcount -= 1;
    }
    return ypartcount;
}
}

// Source/Domain.cpp:1290-1412: tensile probe -- two grip cylinders of radius Rxy, two cones and a neck of radius Rxy_center. The first
// row sits AT z = V(2), the following ones at V(2) + (2k+1) r, as in the reference; every particle of the domain gets the probe's mass.
inline void Domain::AddTractionProbeLength(int tag, Vec3_t const & V, double Rxy, double Lz_side, double Lz_neckmin, double Lz_necktot, double Rxy_center,
                                           double r, double Density, double h, bool Fixed)
{
    if (Dimension != 3) throw new Fatal("AddTractionProbeLength: only Dimension = 3 is implemented in weldform_b200");
    std::cout << "\n--------------Generating particles by CylinderBoxLength with defined length of particles-----------" << std::endl;
    size_t PrePS = Particles.Size();
    double xp, yp;
    double Lz = 2. * Lz_side + Lz_necktot;
    double z_radiusreduction = (Lz_necktot - Lz_neckmin) / 2.;
    double tan = (Rxy - Rxy_center) / (z_radiusreduction);
    double R;
    double z1 = V(2) + Lz_side - r;
    double z2 = V(2) + Lz_side + z_radiusreduction - r;
    double z3 = V(2) + Lz_side + z_radiusreduction + Lz_neckmin - r;
    double z4 = V(2) + Lz_side + Lz_necktot - r;
    size_t k = 0;
    double zp = V(2);
    while (zp <= (V(2) + Lz - r)) {
        if (zp <= z1 || zp > z4) R = Rxy;
        else if (zp > z1 && zp <= z2) R = Rxy - (zp - z1) * tan;
        else if (zp >= z2 && zp < z3) R = Rxy_center;
        else if (zp >= z3 && zp < z4) R = Rxy_center + (zp - z3) * tan;
        int numpartxy = detail::calc_half_part_count(r, R, 1);
        yp = V(1) - r - (2. * r * (numpartxy - 1));
        int numypart = 2 * numpartxy, yinc = numpartxy, yinc_sign = -1;
        for (int j = 0; j < numypart; j++) {
            int numxpart = detail::calc_half_part_count(r, R, yinc);
            xp = V(0) - r - (2. * r * (numxpart - 1));
            for (int i = 0; i < 2 * numxpart; i++) {
                Particles.Push(pool_make(tag, Vec3_t(xp, yp, zp), Vec3_t(0, 0, 0), 0.0, Density, h, Fixed));
                xp += 2. * r;
            }
            yp += 2. * r;
            yinc += yinc_sign;
            if (yinc < 1) { yinc = 1; yinc_sign = 1; }
        }
        k++;
        zp = V(2) + (2.0 * k + 1) * r;
    }
    double L_cone = (Lz_necktot - Lz_neckmin) / 2.;
    double Vol_ = 2.0 * Lz_side * M_PI * Rxy * Rxy + Lz_neckmin * M_PI * Rxy_center * Rxy_center +
                  2.0 * L_cone * M_PI * (Rxy_center * Rxy_center + Rxy * Rxy + Rxy * Rxy_center) / 3.0;
    double Mass = Vol_ * Density / (Particles.Size() - PrePS);
    std::cout << "Particle mass: " << Mass << std::endl;
    for (size_t i = 0; i < Particles.Size(); ++i) Particles[i]->Mass = Mass;
    solid_part_count = (int)Particles.Size();
    this->R = r;
}

// Source/Domain.cpp:2605-2720: polar ("uniform") cylinder lattice from z = r upwards, rings every 2r from r_i, ceil(ang * ri / 2r) particles per
// ring. The slice bookkeeping of the reference (boundary zones, correct_vel_acc) feeds only code the new solvers have commented out
// (CorrectVelAcc, SolverFraser.cpp:333,357) and is not mirrored; `ghost` is unused there as well.
inline void Domain::AddCylUniformLength(int tag, double Rxy, double Lz, double r, double Density, double h, double ang, int rows, double r_i, bool ghost)
{
    if (Dimension != 3) throw new Fatal("AddCylUniformLength: only Dimension = 3 is implemented in weldform_b200");
    (void)rows; (void)ghost;
    std::cout << "\n--------------Generating particles by CylinderBoxLength with defined length of particles-----------" << std::endl;
    double xp, yp;
    double z0 = r;
    int tgcount;
    double zp = z0;
    while (zp <= (z0 + Lz - r)) {
        for (double ri = r_i; ri < Rxy; ri += 2. * r) {
            double dalpha;
            if (ri == 0.) { tgcount = 1; dalpha = 0.; }
            else {
                tgcount = (ceil)((ang * ri) / (2. * r));
                dalpha = ang / (tgcount);
            }
            for (int alphai = 0; alphai < tgcount; alphai++) {
                xp = ri * cos(alphai * dalpha);
                yp = ri * sin(alphai * dalpha);
                Particles.Push(pool_make(tag, Vec3_t(xp, yp, zp), Vec3_t(0, 0, 0), 0.0, Density, h, false));
            }
        }
        zp += 2.0 * r;
    }
    double Vol_ = M_PI * Rxy * Rxy * Lz;
    double Mass = Vol_ * Density / Particles.Size();
    std::cout << "Total Particle count: " << Particles.Size() << std::endl;
    std::cout << "Particle mass: " << Mass << std::endl;
    solid_part_count = (int)Particles.Size();
    for (size_t i = 0; i < Particles.Size(); ++i) Particles[i]->Mass = Mass;
    this->R = r;
}

// Source/Domain.cpp:880-1140 (3-D): a QUARTER cylinder in x, y > 0 with three mirror rows behind the x = 0 and y = 0 planes (Mirror_XYZ ghosts, axes 0 and 1)
// and, with symlength, three mirror rows below z = 0 (axis 2). The creation order is the API (GhostPairs index particles): first the body row by row
// (z, then y, then x), then per z row the y-plane ghost and its transposed x-plane twin alternately, then the z ghosts. Domain::MoveGhost runs on the device.
inline void Domain::AddXYSymCylinderLength(int tag, double Rxy, double Lz, double r, double Density, double h, bool Fixed, bool symlength)
{
    if (Dimension != 3) throw new Fatal("AddXYSymCylinderLength: only Dimension = 3 is implemented in weldform_b200");
    if (Particles.Size() != 0) throw new Fatal("AddXYSymCylinderLength: the quarter cylinder must be the first particle set (its GhostPairs count particles from 0, Source/Domain.cpp:899)");
    std::cout << "\n--------------Generating particles by CylinderBoxLength with XY SYMM AND GHOST part defined length of particles-----------" << std::endl;
    const int ghost_rows = 3;
    const int edge = detail::calc_half_part_count(r, Rxy, 1);
    const double z0 = symlength ? r : -Lz / 2. + r;
    std::vector<int> mirror_x, mirror_y;                 // body particles within three columns of x = 0 / three rows of y = 0, in creation order
    int per_layer = 0, x_cols_first_layer = 0, made = 0;
    size_t layer = 0;
    for (double zp = z0; zp <= (z0 + Lz - r); zp += 2.0 * r, ++layer) {
        double yp = r;
        for (int row = 0; row < edge; ++row, yp += 2. * r) {
            const int cols = detail::calc_half_part_count(r, Rxy, row + 1);
            double xp = r;
            for (int col = 0; col < cols; ++col, xp += 2. * r) {
                Particles.Push(pool_make(tag, Vec3_t(xp, yp, zp), Vec3_t(0, 0, 0), 0.0, Density, h, Fixed));
                if (col < ghost_rows) { mirror_x.push_back(made); if (layer == 0) x_cols_first_layer++; }
                if (row < ghost_rows) mirror_y.push_back(made);
                if (zp == z0) per_layer++;
                made++;
            }
        }
    }
    std::cout << "Particles per row: " << per_layer << std::endl;
    std::cout << " symmetric particles x, y :" << mirror_x.size() << ", " << mirror_y.size() << std::endl;
    std::cout << "Creating ghost particles" << std::endl;
    auto make_ghost = [&](const Vec3_t & x, int inner, int axis) {
        Particle * P = pool_make(tag, x, Vec3_t(0, 0, 0), 0.0, Density, h, Fixed);
        P->inner_mirr_part = inner; P->ghost_plane_axis = axis; P->is_ghost = true; P->not_write_surf_ID = true;
        Particles.Push(P);
        GhostPairs.Push(std::make_pair((size_t)inner, Particles.Size() - 1));
    };
    int next_y = 0;
    layer = 0;
    for (double zp = z0; zp <= (z0 + Lz - r); zp += 2.0 * r, ++layer) {
        for (int row = 0; row < ghost_rows; ++row) {
            const double yp = -r - 2 * r * row;
            const int cols = detail::calc_half_part_count(r, Rxy, row + 1);
            int next_x = x_cols_first_layer * (int)layer + row;
            double xp = r;
            for (int col = 0; col < cols; ++col, xp += 2. * r) {
                make_ghost(Vec3_t(xp, yp, zp), mirror_y[next_y], 1);          // behind the y = 0 plane
                make_ghost(Vec3_t(yp, xp, zp), mirror_x[next_x], 0);          // its transpose behind the x = 0 plane
                next_y++; next_x += ghost_rows;
            }
        }
    }
    std::cout << "inserting z particles" << std::endl;
    if (symlength) {
        std::cout << "Inserting z ghost particles at z bottom..." << std::endl;
        double zp = z0 - 2.0 * r; int inner = 0;
        for (int zr = 0; zr < ghost_rows; ++zr, zp -= 2.0 * r)
            for (int k = 0; k < per_layer; ++k, ++inner) make_ghost(Vec3_t(Particles[inner]->x(0), Particles[inner]->x(1), zp), inner, 2);
    }
    const double Vol_ = M_PI * Rxy * Rxy * Lz;                                  // the FULL cylinder's volume over every particle made, ghosts included (:1121-1122)
    const double Mass = Vol_ * Density / Particles.Size();
    std::cout << "Total Particle count: " << Particles.Size() << std::endl;
    std::cout << "Particle mass: " << Mass << std::endl;
    for (size_t i = 0; i < Particles.Size(); ++i) Particles[i]->Mass = Mass;
    solid_part_count = (int)Particles.Size();
    R = r;
}

// Source/Domain.cpp:326-455 (3-D cubic packing, type 1, rotation 0, random = false); mass rule :447-455
inline void Domain::AddBoxLength(int tag, Vec3_t const & V, double Lx, double Ly, double Lz, double r, double Density, double h, int type, int rotation, bool random, bool Fixed)
{
    if (Dimension != 3) throw new Fatal("AddBoxLength: only Dimension = 3 is implemented in weldform_b200");
    if (type != 1 || random) throw new Fatal("AddBoxLength: only cubic packing (type 1, random = false) is implemented in weldform_b200");
    (void)rotation;
    const double Vd[3] = { V(0), V(1), V(2) }; double dims[3];
    const int64_t n = wf_gen_box(Vd, Lx, Ly, Lz, r, nullptr, 0, dims);
    if (n <= 0) throw new Fatal("AddBoxLength: empty lattice");
    std::vector<double> pts(3 * (size_t)n);
    wf_gen_box(Vd, Lx, Ly, Lz, r, pts.data(), n, dims);
    const size_t first = Particles.Size();
    std::cout << "\n--------------Generating particles by AddBoxLength with defined length of particles-----------" << std::endl;
    pool_reserve((size_t)n);
    for (int64_t i = 0; i < n; ++i) Particles.Push(pool_make(tag, Vec3_t(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), Vec3_t(0, 0, 0), 0.0, Density, h, Fixed));
    const double Vol = dims[0] * dims[1] * dims[2];
    const double Mass = Vol * Density / (double)(Particles.Size() - first);
    std::cout << "Total Particle count: " << Particles.Size() << std::endl;
    for (size_t i = first; i < Particles.Size(); ++i) Particles[i]->Mass = Mass;
    solid_part_count = (int)Particles.Size();
}

// Source/Contact.cpp:10-38: one massless particle per element at its centroid, h = hfac * radius, ID = id
inline void Domain::AddTrimeshParticles(TriMesh * mesh, const float & hfac, const int & id)
{
    first_fem_particle_idx.push_back((int)Particles.Size());
    mesh->id = id;
    for (size_t e = 0; e < mesh->element.Size(); ++e) {
        Particle * p = pool_make(id, mesh->element[e]->centroid, Vec3_t(0, 0, 0), 0.0, 0.0, (double)hfac * mesh->element[e]->radius, false);
        p->normal = mesh->element[e]->normal; p->element = (int)e; p->mesh = meshcount;
        Particles.Push(p);
    }
    trimesh.push_back(mesh); contact_surf_id.push_back(id); m_contact_force.push_back(Vec3_t(0., 0., 0.));
    MeshRec r; r.mesh = mesh; r.hfac = hfac; r.id = id; mesh_recs.push_back(r);
    meshcount++;
    std::cout << "Particle count " << Particles.Size() << std::endl;
}

inline void Domain::choose_sync()
{
    const char * m = std::getenv("WELDFORM_SYNC");
    sync_mode = SYNC_FULL;
    if (m && std::strcmp(m, "tagged") == 0) sync_mode = SYNC_TAGGED;
    if (m && std::strcmp(m, "none") == 0) sync_mode = SYNC_NONE;
    if (m && std::strcmp(m, "push") == 0) sync_mode = SYNC_PUSH;
    const char * ms = std::getenv("WELDFORM_MAX_STEPS"); max_steps_env = ms ? std::atoll(ms) : -1;
    const char * mo = std::getenv("WELDFORM_MAX_OUTPUTS"); max_outputs_env = mo ? std::atoll(mo) : -1;
    have_snap = false;
    sync_idx.clear();
    if (sync_mode == SYNC_TAGGED || sync_mode == SYNC_PUSH) {
        std::map<int, size_t> hist;
        for (int i = 0; i < solid_part_count; ++i) hist[Particles[i]->ID]++;
        int bulk = 0; size_t best = 0;
        for (auto & kv : hist) if (kv.second > best) { best = kv.second; bulk = kv.first; }
        for (int i = 0; i < solid_part_count; ++i) if (Particles[i]->ID != bulk) sync_idx.push_back(i);
        push_groups.clear();
        if (sync_mode == SYNC_PUSH) {
            std::map<int, size_t> slot;
            for (size_t k = 0; k < sync_idx.size(); ++k) {
                const int id = Particles[sync_idx[k]]->ID;
                if (!slot.count(id)) { slot[id] = push_groups.size(); PushGroup g; g.id = id; g.handle = -1; push_groups.push_back(g); }
                push_groups[slot[id]].idx.push_back(sync_idx[k]);
            }
            for (auto & g : push_groups) g.handle = bk.add_group(g.idx);
        }
        else sync_group = bk.add_group(sync_idx);
        std::cout << "BRIDGE mode=" << (sync_mode == SYNC_PUSH ? "push" : "tagged") << " tagged=" << sync_idx.size() << " of " << solid_part_count << " particles" << std::endl;
    }
}

inline void Domain::backend_begin(double dt)
{
    bk.create(device, gpus); backend = bk.c0();
    std::cout << "weldform_b200: " << gpus << " GPU(s), first device " << device << std::endl;
    const size_t ntot = Particles.Size();
    const size_t nsph = !first_fem_particle_idx.empty() ? (size_t)first_fem_particle_idx[0] : ntot;
    if (nsph == 0) throw new Fatal("Domain: no particles");
    // ---- parameters
    wf_params p; wf_default_params(&p);
    p.dimension = Dimension; p.kernel_type = (int)KernelType; p.gradient_type = (int)GradientType; p.ts_nb_inc = ts_nb_inc;
    for (int k = 0; k < 3; ++k) { p.dom_min[k] = DomMin(k); p.dom_max[k] = DomMax(k); }
    p.thermal_solver = thermal_solver; p.pl_heating = pl_heating; p.pl_work_heat_frac = pl_work_heat_frac;
    p.contact = contact; p.friction_sta = friction_sta; p.friction_dyn = friction_dyn; p.pfac = PFAC; p.dfac = DFAC;
    p.cont_heat_fric = cont_heat_fric; p.cont_heat_cond = cont_heat_cond; p.contact_hc = contact_hc;
    p.solver = solver_kind; p.auto_ts = (solver_kind != 0 && auto_ts) ? 1 : 0; p.auto_ts_acc = (solver_kind != 0 && auto_ts_acc) ? 1 : 0; p.cfl = CFL;   // auto_ts is dead code in the Fraser loop (SolverFraser.cpp:500-502)
    p.grad_kernel_corr = gradKernelCorr ? 1 : 0;     // CalcGradCorrMatrix + the locking summation the reference switches to (SolverFraser.cpp:112-115)
    p.sqrt_h_a = sqrt_h_a;
    if (contact && contact_alg != Wang) throw new Fatal("Domain::contact_alg = %d: only CalcContactForcesWang (the default, Source/Domain.cpp:169) is implemented in weldform_b200", contact_alg);
    if (XSPH != 0.0) throw new Fatal("Domain::XSPH != 0 is not implemented in weldform_b200 (the reference's new solvers never compute VXSPH either)");
    { const char * tn = std::getenv("WELDFORM_TUNING"); if (tn) p.tuning = std::atoi(tn); }      // wf_params::tuning (A/B and test switches of the library)
    p.friction_function = (friction_function == Linear) ? 1 : 0; p.friction_m = friction_m; p.friction_b = friction_b;
    bk.each("wf_set_params", [&](int r) { return wf_set_params(bk.ctx[r], &p); });
    // ---- materials: distinct Material_ objects
    std::map<Material_ *, int> mat_ids; std::vector<wf_material> mats;
    for (size_t i = 0; i < nsph; ++i) {
        Material_ * m = Particles[i]->mat;
        if (m && !mat_ids.count(m)) {
            wf_material wm; std::memset(&wm, 0, sizeof(wm));
            wm.model = m->Material_model; wm.E = m->Elastic().E(); wm.nu = m->Elastic().Poisson();
            for (int k = 0; k < 16; ++k) wm.p[k] = m->params[k];
            mat_ids[m] = (int)mats.size(); mats.push_back(wm);
        }
    }
    bk.each("wf_set_materials", [&](int r) { return wf_set_materials(bk.ctx[r], mats.empty() ? nullptr : mats.data(), (int)mats.size()); });
    // ---- rigid meshes
    for (auto & mr : mesh_recs) {
        TriMesh * tm = mr.mesh; const int nn = (int)tm->node.Size(), ne = (int)tm->element.Size();
        std::vector<double> nodes(3 * nn), normals(3 * ne), radius(ne); std::vector<int32_t> elems(3 * ne), nfar(ne);
        for (int n = 0; n < nn; ++n) for (int k = 0; k < 3; ++k) nodes[3 * n + k] = (*tm->node[n])(k);
        for (int e = 0; e < ne; ++e) { for (int k = 0; k < 3; ++k) { elems[3 * e + k] = tm->element[e]->node[k]; normals[3 * e + k] = tm->element[e]->normal(k); }
                                       nfar[e] = tm->element[e]->nfar; radius[e] = tm->element[e]->radius; }
        wf_mesh wm; wm.n_nodes = nn; wm.n_elems = ne; wm.nodes = nodes.data(); wm.elems = elems.data(); wm.normals = normals.data(); wm.nfar = nfar.data();
        wm.radius = radius.data(); wm.hfac = (double)mr.hfac; wm.id = mr.id;
        bk.each("wf_add_mesh", [&](int r) { return wf_add_mesh(bk.ctx[r], &wm); });      // every rank holds every rigid mesh
    }
    // ---- particles: AoS mirrors -> SoA upload
    std::vector<double> x(3 * nsph), v(3 * nsph), mass(nsph), dens(nsph), hh(nsph), refd(nsph), G(nsph), Cs(nsph), P0(nsph), al(nsph), be(nsph), TI(nsph), TIn(nsph),
        TId(nsph), Ep(nsph), kT(nsph), cp(nsph), hc(nsph), Tinf(nsph), thx(nsph), qs(nsph), sy(nsph), T(nsph);
    std::vector<int32_t> id(nsph), isfree(nsph), nowr(nsph), preseq(nsph), fail(nsph), matl(nsph), mmodel(nsph), matid(nsph), thbc(nsph);
    for (size_t i = 0; i < nsph; ++i) {
        const Particle * P = Particles[i];
        for (int k = 0; k < 3; ++k) { x[3 * i + k] = P->x(k); v[3 * i + k] = P->v(k); }
        mass[i] = P->Mass; dens[i] = P->Density; hh[i] = P->h; refd[i] = P->RefDensity; G[i] = P->G; Cs[i] = P->Cs; P0[i] = P->P0; al[i] = P->Alpha; be[i] = P->Beta;
        TI[i] = P->TI; TIn[i] = P->TIn; TId[i] = P->TIInitDist; Ep[i] = P->Ep; kT[i] = P->k_T; cp[i] = P->cp_T; hc[i] = P->h_conv; Tinf[i] = P->T_inf; thx[i] = P->th_exp;
        qs[i] = P->q_source; sy[i] = P->Sigmay; T[i] = P->T;
        id[i] = P->ID; isfree[i] = P->IsFree ? 1 : 0; nowr[i] = P->not_write_surf_ID ? 1 : 0; preseq[i] = (int)P->PresEq; fail[i] = (int)P->Fail; matl[i] = P->Material;
        mmodel[i] = P->Material_model; matid[i] = P->mat ? mat_ids[P->mat] : -1; thbc[i] = P->Thermal_BC;
    }
    wf_particles wp; std::memset(&wp, 0, sizeof(wp));
    wp.n_sph = (int64_t)nsph; wp.x = x.data(); wp.v = v.data(); wp.mass = mass.data(); wp.density = dens.data(); wp.h = hh.data(); wp.ref_density = refd.data();
    wp.id = id.data(); wp.is_free = isfree.data(); wp.not_write_surf_id = nowr.data();
    wp.G = G.data(); wp.Cs = Cs.data(); wp.P0 = P0.data(); wp.alpha = al.data(); wp.beta = be.data(); wp.TI = TI.data(); wp.TIn = TIn.data(); wp.TIInitDist = TId.data();
    wp.Ep = Ep.data(); wp.k_T = kT.data(); wp.cp_T = cp.data(); wp.h_conv = hc.data(); wp.T_inf = Tinf.data(); wp.th_exp = thx.data(); wp.q_source = qs.data();
    wp.pres_eq = preseq.data(); wp.fail = fail.data(); wp.material = matl.data(); wp.material_model = mmodel.data(); wp.mat_id = matid.data(); wp.thermal_bc = thbc.data();
    wp.sigmay = sy.data(); wp.T = T.data();
    {
        std::vector<double> px(3 * ntot), ph(ntot);          // every particle shapes the cell grid the slabs are cut from, rigid-mesh element particles included
        for (size_t i = 0; i < ntot; ++i) { for (int k = 0; k < 3; ++k) px[3 * i + k] = Particles[i]->x(k); ph[i] = Particles[i]->h; }
        bk.set_particles(wp, px.data(), ph.data(), (int64_t)ntot, (int)KernelType, p.dom_min, p.dom_max);
    }
    if (ntot > nsph) {
        // rigid-mesh element particles: the driver's set-up loop runs over ALL particles (examples/FSW.cpp:180-199 gives the tool's particles
        // T = 20 like the plate's); the device keeps their temperature for the contact heat conduction term (Contact_Wang.cpp:323-327)
        std::vector<double> Tall(ntot); for (size_t i = 0; i < ntot; ++i) Tall[i] = Particles[i]->T;
        bk.upload_all(WF_F_T, 1, false, Tall.data());
    }
    if (GhostPairs.Size() > 0) {                                          // Domain::MoveGhost runs on the device (Mirror_XYZ ghosts)
        std::vector<int32_t> gp(2 * GhostPairs.Size()), ax(GhostPairs.Size());
        for (size_t g = 0; g < GhostPairs.Size(); ++g) { gp[2 * g] = (int32_t)GhostPairs[g].first; gp[2 * g + 1] = (int32_t)GhostPairs[g].second; ax[g] = Particles[GhostPairs[g].second]->ghost_plane_axis; }
        bk.each("wf_set_ghost_pairs", [&](int r) { return wf_set_ghost_pairs(bk.ctx[r], gp.data(), ax.data(), (int64_t)GhostPairs.Size()); });    // every rank gets every pair
    }
    sent_mv.assign(trimesh.size(), Vec3_t(0.)); sent_mw.assign(trimesh.size(), Vec3_t(0.));
    forward_mesh_motion();
    bk.each("wf_solver_init", [&](int r) { return wf_solver_init(bk.ctx[r], dt); });
    bk.refresh_ownership();
    choose_sync();
}

inline void Domain::backend_end() { bk.destroy(); backend = nullptr; }

// Source/Neighbour.cpp:10-139 (+ the upload): the cell grid of the reference, computed by the library
inline void Domain::CellInitiate()
{
    if (gpus > 1) throw new Fatal("Domain::CellInitiate / MainNeighbourSearch as stand-alone calls run on one GPU (WELDFORM_GPUS = %d)", gpus);
    backend_begin(deltat > 0.0 ? deltat : 1.0);
    double b[3], t[3], cs[3]; int32_t cn[3];
    ck(wf_get_grid(backend, b, t, cs, cn), "wf_get_grid");
    for (int k = 0; k < 3; ++k) { BLPF(k) = b[k]; TRPR(k) = t[k]; CellSize(k) = cs[k]; CellNo[k] = cn[k]; }
    const size_t np = Nproc > 0 ? Nproc : 1;
    while (SMPairs.Size() < np) { SMPairs.Push(Array<std::pair<size_t, size_t> >()); NSMPairs.Push(Array<std::pair<size_t, size_t> >()); FSMPairs.Push(Array<std::pair<size_t, size_t> >()); }
}
// Source/Neighbour.cpp:346-398: the search runs on the device; the same-material pair set comes back sorted, in creation numbering
inline void Domain::MainNeighbourSearch()
{
    if (!backend) CellInitiate();
    ck(wf_nb_build(backend), "wf_nb_build");
    int64_t np = 0;
    ck(wf_nb_export_pairs(backend, nullptr, 0, &np), "wf_nb_export_pairs");
    std::vector<int64_t> ij(2 * (size_t)np);
    ck(wf_nb_export_pairs(backend, ij.data(), np, &np), "wf_nb_export_pairs");
    ClearNbData();
    for (int64_t p = 0; p < np; ++p) SMPairs[0].Push(std::make_pair((size_t)ij[2 * p], (size_t)ij[2 * p + 1]));
}
// Source/Neighbour.cpp:499-532: Particle::Nb
inline void Domain::SaveNeighbourData()
{
    if (!backend) throw new Fatal("Domain::SaveNeighbourData: no neighbour search has run");
    const size_t n = Particles.Size(); std::vector<int32_t> ib(n);
    bk.download_all(WF_F_NB, 1, true, ib.data());
    for (size_t i = 0; i < n; ++i) Particles[i]->Nb = ib[i];
}
// Source/Neighbour.cpp:553-570
inline int Domain::AvgNeighbourCount()
{
    std::vector<int> nbcount(Particles.Size());
    for (size_t p = 0; p < SMPairs.Size(); ++p)
        for (size_t i = 0; i < SMPairs[p].Size(); ++i) { nbcount[SMPairs[p][i].first]++; nbcount[SMPairs[p][i].second]++; }
    int tot = 0;
    for (size_t p = 0; p < nbcount.size(); ++p) tot += nbcount[p];
    return nbcount.empty() ? 0 : tot / (int)nbcount.size();
}

// ---- field mirrors -------------------------------------------------------------------------------------------------
inline void Domain::get3(wf_field f, Vec3_t Particle::*mem, bool subset)
{
    const size_t n = subset ? sync_idx.size() : Particles.Size(); if (!n) return;
    buf3.resize(3 * n);
    if (subset) bk.download_group(sync_group, f, 3, false, buf3.data()); else bk.download_all(f, 3, false, buf3.data());
    WF_HOST_PARALLEL_FOR
    for (size_t k = 0; k < n; ++k) { Particle * P = Particles[subset ? sync_idx[k] : k]; for (int d = 0; d < 3; ++d) (P->*mem)(d) = buf3[3 * k + d]; }
}
inline void Domain::get9(wf_field f, Mat3_t Particle::*mem, bool subset)
{
    const size_t n = subset ? sync_idx.size() : Particles.Size(); if (!n) return;
    buf9.resize(9 * n);
    if (subset) bk.download_group(sync_group, f, 9, false, buf9.data()); else bk.download_all(f, 9, false, buf9.data());
    WF_HOST_PARALLEL_FOR
    for (size_t k = 0; k < n; ++k) { Particle * P = Particles[subset ? sync_idx[k] : k]; for (int d = 0; d < 9; ++d) (P->*mem).d[d] = buf9[9 * k + d]; }
}
inline void Domain::get1(wf_field f, double Particle::*mem, bool subset)
{
    const size_t n = subset ? sync_idx.size() : Particles.Size(); if (!n) return;
    buf1.resize(n);
    if (subset) bk.download_group(sync_group, f, 1, false, buf1.data()); else bk.download_all(f, 1, false, buf1.data());
    WF_HOST_PARALLEL_FOR
    for (size_t k = 0; k < n; ++k) Particles[subset ? sync_idx[k] : k]->*mem = buf1[k];
}
inline void Domain::put3(wf_field f, Vec3_t Particle::*mem, bool subset)
{
    const size_t n = subset ? sync_idx.size() : Particles.Size(); if (!n) return;
    buf3.resize(3 * n);
    WF_HOST_PARALLEL_FOR
    for (size_t k = 0; k < n; ++k) { const Particle * P = Particles[subset ? sync_idx[k] : k]; for (int d = 0; d < 3; ++d) buf3[3 * k + d] = (P->*mem)(d); }
    if (subset) bk.upload_group(sync_group, f, 3, false, buf3.data()); else bk.upload_all(f, 3, false, buf3.data());
}

inline void Domain::sync_down(bool everything)
{
    const bool subset = !everything && sync_mode == SYNC_TAGGED;
    if (!everything && (sync_mode == SYNC_NONE || sync_mode == SYNC_PUSH)) return;
    get3(WF_F_X, &Particle::x, subset); get3(WF_F_V, &Particle::v, subset); get3(WF_F_A, &Particle::a, subset);
    get3(WF_F_DISPLACEMENT, &Particle::Displacement, subset); get9(WF_F_SIGMA, &Particle::Sigma, subset);
    get1(WF_F_T, &Particle::T, subset); get1(WF_F_PL_STRAIN, &Particle::pl_strain, subset); get1(WF_F_DENSITY, &Particle::Density, subset);
    get1(WF_F_PRESSURE, &Particle::Pressure, subset); get1(WF_F_SIGMAY, &Particle::Sigmay, subset);
    {
        const size_t n = subset ? sync_idx.size() : Particles.Size(); std::vector<int32_t> ib(n);
        if (n) { if (subset) bk.download_group(sync_group, WF_F_ID, 1, true, ib.data()); else bk.download_all(WF_F_ID, 1, true, ib.data()); }
        for (size_t k = 0; k < n; ++k) Particles[subset ? sync_idx[k] : k]->ID = ib[k];
    }
    if (!everything) {            // what the callback is about to see: sync_up uploads / refuses what it changed
        const size_t n = subset ? sync_idx.size() : Particles.Size();
        snap_x.resize(3 * n); snap_disp.resize(3 * n); snap_sigma.resize(9 * n); snap_T.resize(n); snap_dens.resize(n); snap_sigy.resize(n); snap_pls.resize(n); snap_pres.resize(n); snap_id.resize(n);
        WF_HOST_PARALLEL_FOR
        for (size_t k = 0; k < n; ++k) {
            const Particle * P = Particles[subset ? sync_idx[k] : k];
            for (int d = 0; d < 3; ++d) { snap_x[3 * k + d] = P->x(d); snap_disp[3 * k + d] = P->Displacement(d); }
            for (int d = 0; d < 9; ++d) snap_sigma[9 * k + d] = P->Sigma.d[d];
            snap_T[k] = P->T; snap_dens[k] = P->Density; snap_sigy[k] = P->Sigmay; snap_pls[k] = P->pl_strain; snap_pres[k] = P->Pressure; snap_id[k] = P->ID;
        }
        have_snap = true;
    }
    if (everything) {
        get3(WF_F_CONTFORCE, &Particle::contforce, false); get3(WF_F_NORMAL, &Particle::normal, false);
        get9(WF_F_SHEARSTRESS, &Particle::ShearStress, false); get9(WF_F_STRAINRATE, &Particle::StrainRate, false); get9(WF_F_ROTATIONRATE, &Particle::RotationRate, false);
        get9(WF_F_STRAIN, &Particle::Strain, false); get9(WF_F_STRAIN_PL, &Particle::Strain_pl, false); get9(WF_F_TIR, &Particle::TIR, false);
        get1(WF_F_SIGMAY, &Particle::Sigmay, false); get1(WF_F_DDENSITY, &Particle::dDensity, false); get1(WF_F_DTDT, &Particle::dTdt, false);
        get1(WF_F_EFF_STRAIN_RATE, &Particle::eff_strain_rate, false); get1(WF_F_Q_PLHEAT, &Particle::q_plheat, false); get1(WF_F_DELTA_CONT, &Particle::delta_cont, false);
        const size_t n = Particles.Size(); std::vector<int32_t> ib(n);
        bk.download_all(WF_F_ID, 1, true, ib.data()); for (size_t i = 0; i < n; ++i) Particles[i]->ID = ib[i];
        bk.download_all(WF_F_NB, 1, true, ib.data()); for (size_t i = 0; i < n; ++i) Particles[i]->Nb = ib[i];
        // Particle::CalculateEquivalentStress (Source/Particle.cpp)
        for (size_t i = 0; i < n; ++i) { const Mat3_t & S = Particles[i]->Sigma;
            Particles[i]->Sigma_eq = std::sqrt(0.5 * ((S(0, 0) - S(1, 1)) * (S(0, 0) - S(1, 1)) + (S(1, 1) - S(2, 2)) * (S(1, 1) - S(2, 2)) + (S(2, 2) - S(0, 0)) * (S(2, 2) - S(0, 0))) +
                                               3.0 * (S(0, 1) * S(0, 1) + S(1, 2) * S(1, 2) + S(2, 0) * S(2, 0))); }
    }
}
// bitwise comparison: a callback that re-assigns the value it was shown changes nothing
static inline bool wf_same_bits(double a, double b) { return std::memcmp(&a, &b, sizeof(double)) == 0; }
inline void Domain::put1_if_changed(wf_field f, double Particle::*mem, std::vector<double> & snap, bool subset)
{
    const size_t n = subset ? sync_idx.size() : Particles.Size(); bool changed = false;
    for (size_t k = 0; k < n && !changed; ++k) changed = !wf_same_bits(Particles[subset ? sync_idx[k] : k]->*mem, snap[k]);
    if (!changed) return;
    buf1.resize(n);
    for (size_t k = 0; k < n; ++k) buf1[k] = Particles[subset ? sync_idx[k] : k]->*mem;
    if (subset) bk.upload_group(sync_group, f, 1, false, buf1.data()); else bk.upload_all(f, 1, false, buf1.data());
}
inline void Domain::refuse_if_changed(const char * what, const std::vector<double> & snap, int nc, bool subset, double (*get)(const Particle *, int))
{
    const size_t n = subset ? sync_idx.size() : Particles.Size();
    for (size_t k = 0; k < n; ++k) for (int d = 0; d < nc; ++d)
        if (!wf_same_bits(get(Particles[subset ? sync_idx[k] : k], d), snap[(size_t)nc * k + d]))
            throw new Fatal("weldform_b200: the GeneralAfter / GeneralBefore callback assigned Particle::%s of particle %d; the device state takes v, a, x, T, Density, Sigmay and ID "
                            "back from a callback, %s is computed there", what, (int)(subset ? sync_idx[k] : (int)k), what);
}
inline void Domain::sync_up()
{
    if (sync_mode == SYNC_NONE) return;
    const bool subset = sync_mode == SYNC_TAGGED || sync_mode == SYNC_PUSH;
    if (have_snap && sync_mode != SYNC_PUSH) {
        // everything else the device accepts, when the callback touched it (a Dirichlet temperature, a moved or re-tagged particle ...)
        const size_t n = subset ? sync_idx.size() : Particles.Size(); bool xch = false, idch = false;
        for (size_t k = 0; k < n && !xch; ++k) { const Particle * P = Particles[subset ? sync_idx[k] : k]; for (int d = 0; d < 3; ++d) xch = xch || !wf_same_bits(P->x(d), snap_x[3 * k + d]); }
        if (xch) put3(WF_F_X, &Particle::x, subset);
        put1_if_changed(WF_F_T, &Particle::T, snap_T, subset);
        put1_if_changed(WF_F_DENSITY, &Particle::Density, snap_dens, subset);
        put1_if_changed(WF_F_SIGMAY, &Particle::Sigmay, snap_sigy, subset);
        for (size_t k = 0; k < n && !idch; ++k) idch = Particles[subset ? sync_idx[k] : k]->ID != snap_id[k];
        if (idch) {
            std::vector<int32_t> ib(n); for (size_t k = 0; k < n; ++k) ib[k] = Particles[subset ? sync_idx[k] : k]->ID;
            if (subset) bk.upload_group(sync_group, WF_F_ID, 1, true, ib.data()); else bk.upload_all(WF_F_ID, 1, true, ib.data());
        }
        refuse_if_changed("Sigma", snap_sigma, 9, subset, [](const Particle * P, int d) { return P->Sigma.d[d]; });
        refuse_if_changed("pl_strain", snap_pls, 1, subset, [](const Particle * P, int) { return P->pl_strain; });
        refuse_if_changed("Displacement", snap_disp, 3, subset, [](const Particle * P, int d) { return P->Displacement(d); });
        refuse_if_changed("Pressure", snap_pres, 1, subset, [](const Particle * P, int) { return P->Pressure; });
    }
    if (sync_mode == SYNC_PUSH) {
        // write-only bridge: the rows of an ID group usually all get the SAME v and a (examples/Compression_thermal.cpp:30-66): then one
        // 3-vector and a fill kernel per group and field replace a vector per particle over PCIe (wf_fill_registered); otherwise the
        // group's values are uploaded through its registered index list
        for (auto & g : push_groups) {
            const size_t n = g.idx.size(); if (!n) continue;
            const Particle * P0 = Particles[g.idx[0]];
            const Vec3_t v0 = P0->v, a0 = P0->a;
            int differ = 0;
            #ifdef _OPENMP
            #pragma omp parallel for schedule(static) num_threads((int)(Nproc > 0 ? Nproc : 1)) reduction(|:differ)
            #endif
            for (size_t k = 0; k < n; ++k) {
                const Particle * P = Particles[g.idx[k]];
                for (int d = 0; d < 3; ++d) if (!wf_same_bits(P->v(d), v0(d)) || !wf_same_bits(P->a(d), a0(d))) differ |= 1;
            }
            if (!differ) {
                bk.fill_group(g.handle, WF_F_V, v0.d); bk.fill_group(g.handle, WF_F_A, a0.d);
            } else {
                g.v.resize(3 * n); g.a.resize(3 * n);
                WF_HOST_PARALLEL_FOR
                for (size_t k = 0; k < n; ++k) { const Particle * P = Particles[g.idx[k]]; for (int d = 0; d < 3; ++d) { g.v[3 * k + d] = P->v(d); g.a[3 * k + d] = P->a(d); } }
                bk.upload_group(g.handle, WF_F_V, 3, false, g.v.data()); bk.upload_group(g.handle, WF_F_A, 3, false, g.a.data());
            }
        }
        return;
    }
    put3(WF_F_V, &Particle::v, subset); put3(WF_F_A, &Particle::a, subset);
}
inline void Domain::forward_mesh_motion()
{
    for (size_t m = 0; m < trimesh.size(); ++m) {
        const Vec3_t & mv = trimesh[m]->m_v, & mw = trimesh[m]->m_w;
        bool ch = false; for (int k = 0; k < 3; ++k) if (mv(k) != sent_mv[m](k) || mw(k) != sent_mw[m](k)) ch = true;
        if (ch) { bk.each("wf_mesh_set_motion", [&](int r) { return wf_mesh_set_motion(bk.ctx[r], (int)m, mv.d, mw.d); }); sent_mv[m] = mv; sent_mw[m] = mw; }
    }
}
inline void Domain::refresh_scalars()
{
    wf_scalars s; ck(wf_reduce(backend, &s), "wf_reduce");
    Time = s.time; for (int k = 0; k < 3; ++k) max_disp(k) = s.max_disp[k];
    contact_force_sum = s.contact_force_sum; contact_reaction_sum = s.contact_reaction_sum; ext_forces_work = s.ext_forces_work; plastic_work = s.plastic_work;
    accum_cont_heat_cond = s.accum_cont_heat_cond; contact_friction_work = s.contact_friction_work; kin_energy_sum = s.kin_energy_sum; int_energy_sum = s.int_energy_sum;
    m_maxT = s.max_T; m_minT = s.min_T; steps = (size_t)s.steps; isyielding = s.is_yielding != 0;
    bk.steps_hint = (int)steps; bk.time_hint = Time;
    if (!m_contact_force.empty() && m_contact_force.size() <= 6) {                 // Domain::m_contact_force (Contact_Wang.cpp:312)
        double mf[18]; ck(wf_mesh_forces(backend, mf, (int)m_contact_force.size()), "wf_mesh_forces");
        for (size_t m = 0; m < m_contact_force.size(); ++m) m_contact_force[m] = Vec3_t(mf[3 * m], mf[3 * m + 1], mf[3 * m + 2]);
    }
}
// Source/SolverFraser.cpp:48,97-106: Prop.csv and History.csv (header: Time + the contact force of every rigid mesh), a row with every output
// table (:466,:476-482; the reference also adds a row whenever 60 s of CPU time pass without a table)
inline void Domain::open_history()
{
    prop_of.open("Prop.csv", std::ios::out);
    history_of.open("History.csv", std::ios::out);
    history_of << "Time, ";
    for (size_t cf = 0; cf < m_contact_force.size(); ++cf) {
        const int sid = contact_surf_id[cf];
        if (cf > 0) history_of << ", ";
        history_of << "cf" << sid << "x, " << " cf" << sid << "y, " << " cf" << sid << "z";
    }
    history_of << endl;
}
inline void Domain::write_history()
{
    prop_of << getTime() << ", " << m_scalar_prop << endl;
    history_of << Time << ", ";
    for (size_t cf = 0; cf < m_contact_force.size(); ++cf) {
        if (cf > 0) history_of << ", ";
        history_of << m_contact_force[cf](0) << ", " << m_contact_force[cf](1) << ", " << m_contact_force[cf](2);
    }
    history_of << endl;
}
inline void Domain::run_callback(PtDom cb)
{
    if (!cb) return;
    sync_down(false);
    cb(*this);
    sync_up();
    forward_mesh_motion();
}

// Source/Thermal.cpp:567-719: heat conduction only; one neighbour search (wf_solver_init with solver 3), then per step ThermalCalcs
// (wf_thermal), GeneralAfter on the host, Time += dt
inline void Domain::ThermalSolve(double tf, double dt, double dtOut, char const * TheFileKey, size_t maxidx)
{
    std::cout << "\n--------------Solving (weldform_b200, ThermalSolve on the GPU)----------------------------" << std::endl;
    solver_kind = 3;
    deltat = dt; deltatmin = dt;
    backend_begin(dt);
    size_t idx_out = 1; double tout = Time;
    auto t0 = std::chrono::steady_clock::now();
    refresh_scalars();
    while (Time < tf && idx_out <= maxidx) {
        if (GeneralAfter) { bk.phase("wf_thermal", wf_thermal); run_callback(GeneralAfter); bk.phase("wf_step_end", wf_step_end); }   // :657-674
        else bk.each("wf_run", [&](int r) { return wf_run(bk.ctx[r], 1); });
        refresh_scalars();
        if (Time >= tout) {
            if (TheFileKey != NULL) { String fn; fn.Printf("%s_%04d", TheFileKey, (int)idx_out); sync_down(true); WriteCSV(fn.CStr()); }
            idx_out++; tout += dtOut;
            std::cout << "Output No. " << idx_out << " at " << Time << " has been generated\nMax, Min, Avg temps: " << m_maxT << ", " << m_minT << ", " << (m_maxT + m_minT) / 2. << "\n";
        }
    }
    bk.phase("wf_sync", wf_sync);
    wall_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    sync_down(true);
    std::cout << "\n--------------Solving is finished: " << steps << " steps, " << wall_seconds << " s" << std::endl;
}

// Source/SolverKickDrift.cpp:21-400 / Source/SolverLeapfrog.cpp:24-450: the loops' control flow with the GeneralAfter calls at the
// reference's points; every stage is a C-ABI call (wf_solver_stage) into the CUDA library
inline void Domain::solve_kick_drift_leap_frog(int solver, double tf, double dt, double dtOut, char const * TheFileKey, size_t maxidx)
{
    std::cout << "\n--------------Solving (weldform_b200, " << (solver == 1 ? "SolveDiffUpdateKickDrift" : "SolveDiffUpdateLeapFrog") << " on the GPU)----------------------------" << std::endl;
    solver_kind = solver;
    deltat = dt; deltatmin = dt;
    backend_begin(dt);
    size_t idx_out = 1; double tout = Time;
    // SolverKickDrift.cpp:50,100-101 / SolverLeapfrog.cpp: Prop.csv gets (time, m_scalar_prop) and History.csv a row per print_history particle with every table
    prop_of.open("Prop.csv", std::ios::out);
    history_of.open("History.csv", std::ios::out);
    history_of << "Displacement, pl_strain, eff_strain_rate, sigma_eq, sigmay, contforcesum" << endl;
    auto t0 = std::chrono::steady_clock::now();
    refresh_scalars();
    while (Time <= tf && idx_out <= maxidx && !env_stop(idx_out)) {
        if (GeneralBefore) run_callback(GeneralBefore);
        bk.phase_i("wf_solver_stage 0", wf_solver_stage, 0);
        if (GeneralAfter) { refresh_scalars(); run_callback(GeneralAfter); }     // fix free accel
        bk.phase_i("wf_solver_stage 1", wf_solver_stage, 1);
        if (GeneralAfter && solver == 2) run_callback(GeneralAfter);            // SolverLeapfrog.cpp:254
        bk.phase_i("wf_solver_stage 2", wf_solver_stage, 2);
        if (GeneralAfter) run_callback(GeneralAfter);                            // reinforce BC vel
        bk.phase_i("wf_solver_stage 3", wf_solver_stage, 3);
        if (GeneralAfter && solver == 1) run_callback(GeneralAfter);            // SolverKickDrift.cpp:260
        bk.phase_i("wf_solver_stage 4", wf_solver_stage, 4);
        refresh_scalars();
        { wf_scalars s; ck(wf_reduce(backend, &s), "wf_reduce"); deltat = s.deltat; }
        if (Time >= tout) {
            if (TheFileKey != NULL) { String fn; fn.Printf("%s_%04d", TheFileKey, (int)idx_out); sync_down(true); WriteCSV(fn.CStr()); }
            idx_out++; tout += dtOut;
            std::cout << "Output No. " << idx_out << " at " << Time << " has been generated\nCurrent Time Step = " << deltat << "\n";
            prop_of << getTime() << ", " << m_scalar_prop << endl;                    // SolverKickDrift.cpp:379-387
            for (size_t p = 0; p < Particles.Size(); p++)
                if (Particles[p]->print_history) {
                    if (TheFileKey == NULL) sync_down(true);
                    history_of << Particles[p]->Displacement << ", " << Particles[p]->pl_strain << ", " << Particles[p]->eff_strain_rate << ", "
                               << Particles[p]->Sigma_eq << ", " << Particles[p]->Sigmay << ", " << contact_force_sum << endl;
                }
        }
    }
    history_of.close(); prop_of.close();
    bk.phase("wf_sync", wf_sync);
    wall_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    sync_down(true);
    for (int k = 0; k < 3; ++k) max_disp(k) = 0.0;       // those loops shadow Domain::max_disp with a local (SolverKickDrift.cpp:122)
    std::cout << "\n--------------Solving is finished: " << steps << " steps, " << wall_seconds << " s" << std::endl;
}

// Source/SolverFraser.cpp:21-513: the loop's control flow; every phase is a C-ABI call into the CUDA library
inline void Domain::SolveDiffUpdateFraser(double tf, double dt, double dtOut, char const * TheFileKey, size_t maxidx)
{
    std::cout << "\n--------------Solving (weldform_b200, SolveDiffUpdateFraser on the GPU)----------------------------" << std::endl;
    solver_kind = 0;
    deltat = dt; deltatmin = dt;
    backend_begin(dt);
    size_t idx_out = 1; double tout = Time; double max_pl = 0.0;
    auto t0 = std::chrono::steady_clock::now();
    refresh_scalars();
    open_history();
    int bench_state = 0; size_t bench_mark = 0; std::chrono::steady_clock::time_point bench_t0;
    while (Time <= tf && idx_out <= maxidx && !env_stop(idx_out)) {
        if (bench_steps > 0) {
            if (bench_state == 0 && (isyielding || !bench_wait_yield) && steps >= bench_spin_min) { bench_state = 1; bench_mark = steps + bench_warmup; }
            if (bench_state == 1 && steps == bench_mark) { bk.phase("wf_sync", wf_sync); bench_t0 = std::chrono::steady_clock::now(); bench_state = 2; bench_first_step = steps; bench_mark = steps + bench_steps; }
            else if (bench_state == 2 && steps == bench_mark) { bk.phase("wf_sync", wf_sync); bench_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - bench_t0).count(); break; }
        }
        if (GeneralBefore) run_callback(GeneralBefore);                          // :238
        bk.phase("wf_fraser_stage_a", wf_fraser_stage_a);                        // :137-312 rebuild policy, density, rates + stress, accel, contact
        // push mode (write-only callbacks): the callback does not wait for the device -- Time is advanced on the host with the device's own
        // arithmetic (Time += deltat) and the other scalars are as of the last completed step -- so it runs WHILE stage a executes on the GPU
        if (GeneralAfter) { if (sync_mode != SYNC_PUSH) refresh_scalars(); run_callback(GeneralAfter); }     // :332
        bk.phase("wf_fraser_stage_b1", wf_fraser_stage_b1);                      // :337-343 position update
        bk.phase("wf_fraser_stage_b2", wf_fraser_stage_b2);                      // :352-355 velocity update
        if (GeneralAfter) run_callback(GeneralAfter);                            // :360
        bk.phase("wf_fraser_stage_c", wf_fraser_stage_c);                        // :367-412 heat, energies, time, mesh update, cadence
        // push mode: no device round trip between outputs (the scalars a callback could read are refreshed every 32 steps and before every output)
        if (sync_mode == SYNC_PUSH && (steps + 1) % 32 != 0 && Time + deltat < tout) { Time += deltat; steps++; bk.steps_hint = (int)steps; }
        else refresh_scalars();
        if (Time >= tout) {                                                      // :420-499
            wf_scalars s; ck(wf_reduce(backend, &s), "wf_reduce"); max_pl = s.max_pl_strain;
            if (TheFileKey != NULL) { String fn; fn.Printf("%s_%04d", TheFileKey, (int)idx_out); sync_down(true); WriteCSV(fn.CStr()); }
            idx_out++; tout += dtOut;
            write_history();
            wall_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            std::ostringstream o;
            o << "\n---------------------------------------\n Total time: " << wall_seconds << " s (GPU backend)\nstep number: " << steps << "\n";
            o << "Output No. " << idx_out << " at " << Time << " has been generated\nCurrent Time Step = " << deltat << "\n";
            o << "Max plastic strain: " << max_pl << "\n";
            if (max_pl > 0.) o << "Plastic Work " << plastic_work << "\n";
            o << "Max Displacements (No Cont Surf): " << max_disp << "\n";
            if (contact) o << "Contact Force Sum " << contact_force_sum << ", Reaction Sum " << contact_reaction_sum << "\nContact Friction Work " << contact_friction_work
                           << "\nExternal Forces Work " << ext_forces_work << "\n";
            o << "Int Energy: " << int_energy_sum << ", Kin Energy: " << kin_energy_sum << "\n";
            if (thermal_solver) o << "Max, Min, Avg temps: " << m_maxT << ", " << m_minT << ", " << (m_maxT + m_minT) / 2. << "\n";
            o << "-----------------------------------------------------------------------------\n";
            std::cout << o.str();
        }
    }
    bk.phase("wf_sync", wf_sync);
    wall_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    sync_down(true);
    for (size_t m = 0; m < trimesh.size(); ++m) {                                // rigid-mesh node positions / normals back into the driver's TriMesh
        TriMesh * tm = trimesh[m]; const size_t nn = tm->node.Size(), ne = tm->element.Size();
        std::vector<double> nodes(3 * nn), nrm(3 * ne), pp(ne);
        ck(wf_mesh_download(backend, (int)m, nodes.data(), nrm.data(), pp.data()), "wf_mesh_download");
        for (size_t n = 0; n < nn; ++n) for (int k = 0; k < 3; ++k) (*tm->node[n])(k) = nodes[3 * n + k];
        for (size_t e = 0; e < ne; ++e) { for (int k = 0; k < 3; ++k) tm->element[e]->normal(k) = nrm[3 * e + k]; tm->element[e]->pplane = pp[e]; }
    }
    history_of.close(); prop_of.close();
    refresh_scalars();
    std::cout << "Final state: step " << steps << ", time " << Time << ", max displacements " << max_disp << ", max plastic strain " << max_pl_final() << std::endl;
    std::cout << "\n--------------Solving is finished: " << steps << " steps, " << wall_seconds << " s, "
              << (double)solid_part_count * (double)steps / (wall_seconds > 0 ? wall_seconds : 1.0) << " particle-steps/s--------------" << std::endl;
    backend_end();
}

// Source/Output.cpp:520-554 (same columns)
inline void Domain::WriteCSV(char const * FileKey)
{
    std::ostringstream oss;
    oss << "X, Y, Z, ID, Sigma_eq, Pl_Strain, vx, vy, vz, ax, ay, az, CFx, CFy, CZFz, p, Temp" << endl;
    for (size_t i = 0; i < Particles.Size(); i++) {
        const Particle * P = Particles[i];
        for (int j = 0; j < 3; j++) oss << P->x(j) << ", ";
        oss << P->ID << ", " << P->Sigma_eq << ", " << P->pl_strain << ", " << P->v(0) << ", " << P->v(1) << ", " << P->v(2) << ", "
            << P->a(0) << ", " << P->a(1) << ", " << P->a(2) << ", " << P->contforce(0) << ", " << P->contforce(1) << ", " << P->contforce(2) << ", "
            << P->Pressure << ", " << P->T << endl;
    }
    String fn(FileKey); fn.append(".csv");
    std::ofstream of(fn.CStr(), std::ios::out); of << oss.str(); of.close();
}
// Source/Output.cpp:75-518 writes XDMF + HDF5 (float32). HDF5 is outside the hot path: the same
// particle table is written as CSV next to the name the driver asked for.
inline void Domain::WriteXDMF(char const * FileKey) { WriteCSV(FileKey); }

} // namespace SPH
#endif
