// Internal declarations shared by the translation units of libweldform_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/weldform_b200.h"
#include "wf_math.cuh"

#define WF_MAX_MESH_FORCES 6      // rigid meshes whose summed contact force (Domain::m_contact_force) rides in the 32-double scalar record

namespace wf {

// ---- property classes: distinct tuples of the per-particle constants the drivers assign (Particle.h:53-231) ------
struct ClassProps {
    double G, Cs, P0, RefDensity, Alpha, Beta, TI, TIn, TIInitDist, Ep, k_T, cp_T, h_conv, T_inf, th_exp, q_source;
    int PresEq, Fail, Material, Material_model, mat_id, Thermal_BC, IsMesh, pad;
};
struct MatProps { int model, pad; double E, nu, p[16], eps0, eps1; };

// ---- own (non-gathered) per-particle doubles, structure of arrays: field f of slot s is F[f*cap + s] -------------
enum OwnField {
    F_AX = 0, F_AY, F_AZ, F_DX, F_DY, F_DZ, F_DDENS,
    F_SH,                 // ShearStress xx,yy,zz,xy,yz,xz
    F_SR = F_SH + 6,      // StrainRate 6
    F_RR = F_SR + 6,      // RotationRate 01,02,12
    F_ST = F_RR + 3,      // Strain 6
    F_SP = F_ST + 6,      // Strain_pl 6
    F_SI = F_SP + 6,      // Strain_pl_incr 6
    F_PLS = F_SI + 6, F_DPLS, F_SIGY, F_EFFSR, F_DTDT, F_QCONV, F_QPL, F_QFRIC, F_QCONT,
    F_CFX, F_CFY, F_CFZ, F_DCONT, F_NX, F_NY, F_NZ, F_DKIN, F_DINT,
    F_CSUM, F_CREAC, F_CEXT,          // per-particle partial sums of the Wang contact loop
    F_SG,                 // Sigma xx,yy,zz,xy,yz,xz (own copy; the gathered S arrays hold Sigma/rho^2 for GradientType 0)
    F_PRES = F_SG + 6,    // Pressure
    F_DDN,                // density rate prepared for the NEXT step by k_post (or by the stand-alone density pass)
    F_TN,                 // TIR prepared for the next step, 6
    F_GC = F_TN + 6,      // gradCorrM (Domain.cpp:1832-1940), full 3x3 row-major: the reference's matrix is not exactly symmetric
    F_COUNT = F_GC + 9
};
enum OwnInt { I_ID = 0, I_IDORIG, I_NB, I_NOWRITE, I_ORIG, I_OLDSURF, I_CLS, I_LID, I_FREE, I_COUNT };   // I_ORIG: global creation index; I_LID: local index (slot_of inverse)

struct DevMesh {      // concatenated over meshes
    int n_nodes = 0, n_elems = 0;
    double *node = nullptr, *node_v = nullptr;       // 3*n_nodes
    int *en = nullptr;                               // 3*n_elems (global node ids)
    int *nfar = nullptr, *emesh = nullptr;           // per element
    double *enormal = nullptr, *ev = nullptr, *pplane = nullptr;   // 3*,3*,1* n_elems
    double *motion = nullptr;                        // per mesh: m_v(3), m_w(3), vmax(3), tau, has_ramp  (11 doubles)
};
struct HostMesh { int node0, elem0, n_nodes, n_elems, id; int64_t first_particle; };

struct BCEntry { int id, set_a; double v[3]; double tau; };

// device-resident step scalars (advanced by kernels so that whole steps can be replayed without host involvement)
struct DevScalars {
    double time, dt;
    double kin_energy_sum, int_energy_sum, plastic_work, contact_force_sum, contact_reaction_sum, ext_forces_work,
           contact_friction_work, accum_cont_heat_cond;
    double max_pl_strain, max_disp2[3], max_disp2_prev[3], max_T, min_T;
    double mcf[WF_MAX_MESH_FORCES][3];      // Domain::m_contact_force of the last completed step
};

enum ProfKernel { PK_NB_KEYS = 0, PK_NB_SORT, PK_NB_PERMUTE, PK_NB_CELLS, PK_NB_LIST, PK_NB_SURFACE, PK_NB_CONTACT,
                  PK_BEGIN, PK_DENSITY, PK_RATE_STRESS, PK_ACCEL, PK_CONTACT, PK_BC, PK_MOVE, PK_THERMAL, PK_ENERGY,
                  PK_REDUCE, PK_END, PK_HALO, PK_COUNT };

} // namespace wf

struct wf_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    wf_params prm;
    bool have_particles = false, solver_ready = false;

    int64_t n = 0, nsph = 0, cap = 0;            // active slots (local + halo), GLOBAL SPH count, allocated slots
    int64_t n_loc = 0;                           // slots [0, n_loc) are this rank's own particles (== n on one GPU); halos follow
    int64_t nmeshp = 0;                          // rigid-mesh element particles (replicated on every rank of a multi-GPU run)
    bool mesh_tail = false;                      // mesh particles are kept out of the SPH cell table (tail of the local region with a cell table of their own, see wf_neighbour.cu)
    // ---- multi-GPU (wf_comm.cu): 1-D slab decomposition along the slowest cell axis, NCCL point-to-point halos
    struct Comm {
        void *nccl = nullptr;                    // ncclComm_t
        int rank = 0, nranks = 1;
        int layer_lo = 0, layer_hi = 0;          // owned slow-axis cell layers [layer_lo, layer_hi)
        int64_t send_lo_start = 0, send_lo_n = 0, send_hi_start = 0, send_hi_n = 0;   // boundary-layer slices of the local region
        int64_t halo_lo_start = 0, halo_lo_n = 0, halo_hi_start = 0, halo_hi_n = 0;   // halo regions (from rank-1 / rank+1)
        double *stage_send[2] = { nullptr, nullptr }, *stage_recv[2] = { nullptr, nullptr }; size_t stage_cap = 0;   // migration staging
        int *d_emig = nullptr; int *d_cnt = nullptr; int64_t emig_cap = 0;            // emigrant slot lists (2 x emig_cap), counters
        cudaStream_t stream = nullptr; cudaEvent_t ev_ready = nullptr, ev_done = nullptr; bool pending = false;   // halo exchange overlapped with interior work
        double *d_rank = nullptr, *d_all = nullptr;                                     // scalar partials: this rank / all ranks
        // end-of-step scalars gathered OFF the compute stream (two buffer sets by step parity): the all-gather and the accumulation kernel run on
        // the communication stream while the next step's first kernels are already executing
        double *d_red_rank[2] = { nullptr, nullptr }, *d_red_all[2] = { nullptr, nullptr };
        cudaEvent_t ev_red_ready = nullptr, ev_red_done[2] = { nullptr, nullptr }; bool red_pending[2] = { false, false }; int red_parity = 0;
        int64_t migrated_out = 0, migrated_in = 0, halo_bytes_step = 0;
        // time-based rebalancing of the slabs: event pairs around the interior momentum launch of every step since the last rebuild
        std::vector<std::pair<cudaEvent_t, cudaEvent_t> > t_events; int t_used = 0; int64_t rebalance_moves = 0; int rebuilds_seen = 0;
    } comm;
    // gathered arrays, double-buffered (index cur / 1-cur): Q0 (x,y,z,h) Q1 (vx,vy,vz,m/rho) Q2 (m,rho,T,P/rho^2)
    // S0 (pxx,pyy,pzz,pxy) S1 (pyz,pxz,Rxx,Ryy) S2 (Rzz,Rxy,Ryz,Rxz), p = Sigma/rho^2 (GradientType 0) or Sigma, R = TIR
    double4 *Q0[2] = { nullptr, nullptr }, *Q1[2] = { nullptr, nullptr }, *Q2[2] = { nullptr, nullptr },
            *S0[2] = { nullptr, nullptr }, *S1[2] = { nullptr, nullptr }, *S2[2] = { nullptr, nullptr };
    double *F[2] = { nullptr, nullptr };
    int *I[2] = { nullptr, nullptr };
    int cur = 0;
    int *slot_of = nullptr;                      // creation index -> slot
    // classes / materials
    std::vector<wf::ClassProps> h_classes; wf::ClassProps *d_classes = nullptr; int n_classes = 0;
    std::vector<wf::MatProps> h_mats; wf::MatProps *d_mats = nullptr;
    bool any_ti = false, uniform_h = false, single_class = false, any_fixed = false; double h_uniform = 0.0;
    bool pair_uniform = false;                   // every SPH class has the same PAIR-relevant constants (Cs, PresEq, RefDensity, Alpha, Beta, TI, TIn, TIInitDist, k_T, Material):
                                                 // classes then differ in own-particle properties only (cp_T, h_conv, T_inf, Thermal_BC, G, hardening ...), e.g. FSW's convection rows
    bool uniform_mass = false; double mass_min = 0.0, mass_max = 0.0;       // every SPH particle of the DOMAIN has the same mass (lattice generators): the contracted momentum kernel then skips the Q2 gather
    bool dens_valid = false;                     // F_DDN / F_TN hold the coming step's density rate and TIR (set by k_post)
    double totmass_mean = 0.0;                   // CalculateSurface's serial mean mass (Domain.cpp:1421-1425)
    // cell grid (Neighbour.cpp:10-139)
    double BLPF[3], TRPR[3], CellSize[3], hmax = 0.0, Cellfac = 2.0; int CellNo[3] = { 0, 0, 0 }; int64_t ncells = 0;
    int ax_fast = 0, ax_mid = 1, ax_slow = 2;
    unsigned long long *keys[2] = { nullptr, nullptr }; int *vals[2] = { nullptr, nullptr };
    int *cell_start = nullptr, *cell_end = nullptr; int64_t cell_cap = 0;
    int *mcell_start = nullptr, *mcell_end = nullptr;   // cell table of the rigid-mesh element particles (mesh_tail layout)
    void *cub_tmp = nullptr; size_t cub_tmp_bytes = 0;
    // neighbour lists (ELL, transposed): nbr[k*cap + slot]; bit31 = "slot is temp1 (first) of the reference pair"
    uint32_t *nbr = nullptr; int nbr_k = 0; int64_t nbr_alloc = 0;
    int *nb_cnt = nullptr, *nb_hi = nullptr, *d_maxcnt = nullptr;
    // contact candidate lists: cont[k*cap + slot] = global element index
    int *cont = nullptr; int cont_k = 0; int64_t cont_alloc = 0; int *cont_cnt = nullptr;
    // meshes
    wf::DevMesh dm; std::vector<wf::HostMesh> meshes; std::vector<int> mesh_ids; int *d_mesh_ids = nullptr;
    std::vector<double> h_motion;
    struct PendingMesh { std::vector<double> nodes, normals, radius; std::vector<int> elems, nfar; double hfac; int id; };
    std::vector<PendingMesh> pending_meshes;
    // boundary conditions
    std::vector<wf::BCEntry> h_bcs; wf::BCEntry *d_bcs = nullptr; int n_bcs = 0;
    // scalars
    wf::DevScalars *d_sc = nullptr; wf::DevScalars *h_sc = nullptr;   // h_sc pinned
    double *d_partial = nullptr; int partial_blocks = 0;
    double *d_mf_part = nullptr; int *d_mf_flag = nullptr; int mf_cap = 0, mf_blocks = 0;   // per-block m_contact_force partials of k_contact
    // solver state (SolverFraser.cpp:62-68)
    double prev_dt = 0.0;                        // deltat of the previous step (the kicks of the KickDrift / LeapFrog orders use both)
    double dt_init = 0.0, dt_min = 0.0;          // Domain::deltatint / deltatmin (adaptive time step, Domain.cpp:205-297)
    double dt = 0.0, time = 0.0; int ts_i = 0; bool isfirst = true, isyielding = false, nb_cleared = true;
    int64_t steps = 0, nb_rebuilds = 0, launches = 0; bool in_step = false;
    double step_max_pl = 0.0;                    // `max` of SolverFraser.cpp:139-149 for the step in flight
    int *d_ghost_loc = nullptr; bool ghost_loc_valid = false;   // several GPUs: the pairs as LOCAL indices of this rank (-1: not here), refreshed at every neighbour build
    int *d_ghost = nullptr; int64_t n_ghost = 0;         // Domain::GhostPairs (Domain.h:473) as (inner, ghost, ghost_plane_axis) creation indices, Mirror_XYZ ghosts
    std::vector<std::pair<int *, int64_t> > subsets;     // registered index lists of the callback bridge (device copies)
    // staging
    void *stage = nullptr; size_t stage_bytes = 0; void *h_stage = nullptr; size_t h_stage_bytes = 0;
    // profiling
    // profiling: event pairs are recorded on the ctx stream without synchronising and resolved in wf_prof_get
    bool prof = false; double prof_ms[wf::PK_COUNT]; int64_t prof_n[wf::PK_COUNT]; cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    struct ProfRec { cudaEvent_t a, b; int k; };
    std::vector<ProfRec> prof_pending; std::vector<std::pair<cudaEvent_t, cudaEvent_t> > prof_free;

    double *Fp(int f) { return F[cur] + (int64_t)f * cap; }
    int *Ip(int f) { return I[cur] + (int64_t)f * cap; }
};

namespace wf {

int fail(wf_ctx *c, int code, const char *fmt, ...);
#define WF_CUDA(c, expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return wf::fail((c), WF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define WF_CHECK(expr) do { int r_ = (expr); if (r_ != 0) return r_; } while (0)
#define WF_LAUNCH_CHECK(c) do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return wf::fail((c), WF_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

void prof_resolve(wf_ctx *c);
struct ProfScope {
    wf_ctx *c; int k; cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(wf_ctx *c_, int k_) : c(c_), k(k_) {
        if (!c->prof) return;
        if (c->prof_free.empty()) { if (c->prof_pending.size() >= 8192) prof_resolve(c); }
        if (c->prof_free.empty()) { cudaEventCreate(&a); cudaEventCreate(&b); } else { a = c->prof_free.back().first; b = c->prof_free.back().second; c->prof_free.pop_back(); }
        cudaEventRecord(a, c->stream);
    }
    ~ProfScope() { if (a) { cudaEventRecord(b, c->stream); wf_ctx::ProfRec r; r.a = a; r.b = b; r.k = k; c->prof_pending.push_back(r); } }
};

inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }
int ensure_stage(wf_ctx *c, size_t bytes);

// wf_neighbour.cu
int nb_cell_initiate(wf_ctx *c);
int nb_build_impl(wf_ctx *c);
int nb_gradcorr_fsm(wf_ctx *c);                         // CalcGradCorrMatrix: contributions of the pairs with one fixed particle
// wf_pair.cu
int pair_gradcorr(wf_ctx *c);
int pair_tir(wf_ctx *c);
int pair_temp_inc_pp(wf_ctx *c);                         // CalcTempIncPP: the initial dT/dt of ThermalSolve
int pair_min_ts(wf_ctx *c, double *test_min);
int pair_density(wf_ctx *c, bool with_tir = true);
int pair_rate_stress(wf_ctx *c, int64_t lo = 0, int64_t hi = -1);   // slots [lo, hi); hi < 0: n_loc
int pair_accel(wf_ctx *c, int64_t lo = 0, int64_t hi = -1);
int pair_thermal(wf_ctx *c);
int pair_energy(wf_ctx *c);
int pair_post(wf_ctx *c, bool skip_dens = false);
// wf_comm.cu
enum HaloSet { HALO_Q1 = 1, HALO_Q0 = 2, HALO_Q2 = 4, HALO_S = 8 };
int comm_halo_refresh(wf_ctx *c, int sets);             // no-op on one GPU
// asynchronous form: the exchange runs on the communication stream; [*i0, *i1) receives the interior slot range (particles whose
// neighbour lists cannot reach a halo slot) that may be evaluated meanwhile; comm_halo_end makes the compute stream wait for it
int comm_halo_begin(wf_ctx *c, int sets, int64_t *i0, int64_t *i1);
int comm_halo_end(wf_ctx *c);
int comm_time_begin(wf_ctx *c);                          // brackets the launch whose duration measures this rank's speed (no-op on one GPU)
int comm_time_end(wf_ctx *c);
int comm_rebalance(wf_ctx *c);                           // at a rebuild: move slab boundaries by at most one layer towards equal step time
void comm_destroy(wf_ctx *c);                           // ncclCommDestroy, no-op on one GPU
int comm_gather_scalars(wf_ctx *c, int nvals);          // d_rank -> d_all (all ranks), no-op on one GPU
int comm_gather_scalars_on(wf_ctx *c, const double *src, double *dst, int nvals, cudaStream_t st);
int comm_red_wait(wf_ctx *c);                            // compute stream waits for every end-of-step scalar gather still in flight
int comm_nb_exchange(wf_ctx *c);                        // migration + halo set-up inside the neighbour build
int comm_global_grid(wf_ctx *c, double BL[3], double TR[3], double *hmax);
int comm_global_minmax(wf_ctx *c, double *mn, double *mx);
int comm_agree(wf_ctx *c, int local_status, const char *what);
int comm_unify_classes(wf_ctx *c, std::vector<wf::ClassProps> &cls, std::vector<int> &remap);   // one class table for all ranks (halo neighbours carry class indices)
int comm_serial_mass_sum(wf_ctx *c, const double *mass, const int64_t *gidx, int64_t n_local, int64_t n_global, double *tot);
int comm_plan_layers(wf_ctx *c, const std::vector<long long> &hist);
int comm_migrate(wf_ctx *c, int64_t *n_sort, int64_t *n_out);
int comm_halo_setup(wf_ctx *c, const unsigned long long *sorted_keys);
int comm_local_index(wf_ctx *c);
// wf_step.cu
int step_begin_impl(wf_ctx *c);
int contact_impl(wf_ctx *c);
int apply_bcs_impl(wf_ctx *c);
int move_ghost_impl(wf_ctx *c);                          // Domain::MoveGhost; no-op without ghost pairs
int ghost_localize(wf_ctx *c);                           // several GPUs: ghost pairs -> this rank's local indices (after comm_local_index)
int move_pos_impl(wf_ctx *c);
int move_vel_impl(wf_ctx *c);
int step_end_impl(wf_ctx *c);
int reduce_impl(wf_ctx *c, bool to_host);

} // namespace wf
