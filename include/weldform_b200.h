/* weldform_b200 -- C ABI of the B200-native SPH hot path (libweldform_b200.so).
 *
 * The reference (luchete80/WeldForm) has no FFI boundary: it is a header-only unity build whose hot path is a set of
 * SPH::Domain member functions called from Domain::SolveDiffUpdateFraser (Source/SolverFraser.cpp:133-505). This
 * header is the boundary a maintainer would bind instead; each entry point names the reference member(s) it replaces.
 * The host-side mirror of the reference's source API (weldform_b200/host/Domain.h ...) and the Python harness
 * (weldform_b200/api.py, ctypes) both sit on top of exactly these symbols.
 *
 * Conventions: every function returns 0 on success or a negative wf_status; wf_last_error() gives the message.
 * No function throws or calls back into the host. One wf_ctx per GPU; calls on a ctx come from one host thread at a
 * time and are asynchronous on the ctx-owned stream unless stated; the caller owns host buffers, the ctx owns device
 * buffers. Host arrays are always in PARTICLE CREATION ORDER (SPH particles first, rigid-mesh particles after, as
 * Source/Contact.cpp:12 requires); the cell-sorted device order is internal. There is no CPU fallback.
 */
#ifndef WELDFORM_B200_H
#define WELDFORM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wf_ctx wf_ctx;

typedef enum wf_status {
    WF_OK = 0,
    WF_ERR_CUDA = -1,          /* a CUDA runtime call or kernel failed */
    WF_ERR_ARG = -2,           /* bad argument / call order */
    WF_ERR_UNSUPPORTED = -3,   /* feature of the reference outside the implemented scope (fails loudly) */
    WF_ERR_NOMEM = -4
} wf_status;

/* Domain-level switches; defaults are those of Domain::Domain (Source/Domain.cpp:78-175). */
typedef struct wf_params {
    int    dimension;          /* Domain::Dimension; only 3 is implemented */
    int    kernel_type;        /* Domain::KernelType 0 Qubic_Spline, 1 Quintic, 2 Quintic_Spline, 3 Hyperbolic (Functions.cpp:25) */
    int    gradient_type;      /* Domain::GradientType 0/1/2 (InteractionAlt.cpp:138-143) */
    int    ts_nb_inc;          /* Domain::ts_nb_inc: neighbour rebuild cadence once yielding (SolverFraser.cpp:397-404) */
    double dom_min[3];         /* Domain::DomMin / DomMax grid overrides (Neighbour.cpp:36-42) */
    double dom_max[3];
    int    thermal_solver;     /* Domain::thermal_solver */
    int    pl_heating;         /* Domain::pl_heating */
    double pl_work_heat_frac;  /* Domain::pl_work_heat_frac */
    int    contact;            /* Domain::contact (Wang algorithm, Contact_Wang.cpp) */
    double friction_sta, friction_dyn;
    int    friction_function;  /* Domain::friction_function: 0 Constant, 1 Linear: mu = friction_m*T + friction_b (Contact_Wang.cpp:243-244) */
    double friction_m, friction_b;
    double pfac, dfac;         /* Domain::PFAC / DFAC */
    double contact_force_factor;
    int    cont_heat_fric, cont_heat_cond;
    double contact_hc;
    int    nb_order_reference; /* 1: neighbour lists in the reference's Nproc=1 summation order (default) */
    int    grad_kernel_corr;   /* Domain::gradKernelCorr: Bonet-Lok corrected kernel gradient (Domain.cpp:1832-1940). As in the reference the Fraser
                                * solver then sums per pair with locks, in pair-discovery order (SolverFraser.cpp:112-115) */
    int    solver;             /* step order run by wf_run: 0 SolveDiffUpdateFraser (SolverFraser.cpp), 1 SolveDiffUpdateKickDrift
                                * (SolverKickDrift.cpp:21-400), 2 SolveDiffUpdateLeapFrog (SolverLeapfrog.cpp:24-450), 3 ThermalSolve (Thermal.cpp:567-719:
                                * heat conduction only, one neighbour search at wf_solver_init) */
    int    auto_ts;            /* Domain::auto_ts: CheckMinTSVel + AdaptiveTimeStep (Domain.cpp:205-297); honoured by solvers 1 and 2 only, like the reference */
    double cfl;                /* Domain::CFL (0.7) */
    int    auto_ts_acc;        /* Domain::auto_ts_acc: CheckMinTSAccel, sqrt_h_a*sqrt(h/|a|) (Domain.cpp:302-317, sqrt_h_a below); it replaces the
                                * velocity criterion when both are set, as the reference's call sequence does */
    int    tuning;             /* A/B measurement and test switches, default 0 (never needed for correct results):
                                *   bit 0 no software-pipelined gathers; bits 1-5 momentum-kernel variants (occupancy, general instead of lean
                                *   kernel, 2 x LDG.128 instead of LDG.256); bits 6-7 post-pass occupancy; bit 8 the multi-GPU layout of the
                                *   rigid-mesh element particles on one GPU (tests); bit 9 blocking halo refreshes instead of the overlapped
                                *   ones; bit 10 EXACT ARITHMETIC: the pair kernels evaluate every term with the reference's operand order and without FMA
                                *   contraction everywhere (bit-identical to the reference build; the default uses contracted arithmetic, a few ulp per pair
                                *   term, on the uniform-h single-class cubic-spline path, wf_pair_fm.cu); bit 11 end-of-step scalar gather in stream order;
                                *   bit 13 force one slab-rebalancing move at the third rebuild (tests); bit 14 slab rebalancing off; bit 15 keep the
                                *   Q2 gather in the contracted momentum kernel */
    double sqrt_h_a;           /* Domain::sqrt_h_a of CheckMinTSAccel; 0 = the reference's default 0.0025 (Domain.cpp:109) */
} wf_params;

/* Material_ subclasses (Source/Material.h:96-224). model: 0 bilinear (no object), 1 Hollomon, 2 Johnson-Cook, 3 GMT (Material.cpp:96-161). */
typedef struct wf_material {
    int    model;
    double E, nu;
    double p[16];              /* Hollomon: sy0, K, m.  Johnson-Cook: A, B, n, C, eps_0, m, T_m, T_t.
                                * GMT: n1, n2, C1, C2, m1, m2, I1, I2, e_min, e_max, er_min, er_max, T_min, T_max (the constructor's argument order) */
} wf_material;

/* Initial particle set, creation order, n = n_sph + total mesh particles. NULL pointers take the value the reference's
 * Particle constructor assigns (Source/Particle.cpp:28-158). */
typedef struct wf_particles {
    int64_t n_sph;
    const double *x;           /* 3*n_sph */
    const double *v;           /* 3*n_sph or NULL */
    const double *mass, *density, *h;   /* n_sph */
    const double *ref_density; /* NULL: = density */
    const int32_t *id;         /* Particle::ID or NULL (0) */
    const int32_t *is_free;    /* NULL: all free. Fixed particles get no SM pairs (Neighbour.cpp:400-411); their Shepard-smoothed TIR is not computed */
    const int32_t *not_write_surf_id;
    /* per-particle "class" properties; few distinct tuples are expected (they are deduplicated into a table) */
    const double *G, *Cs, *P0, *alpha, *beta, *TI, *TIn, *TIInitDist, *Ep, *k_T, *cp_T, *h_conv, *T_inf, *th_exp, *q_source;
    const int32_t *pres_eq, *fail, *material, *material_model, *mat_id, *thermal_bc;
    /* per-particle state */
    const double *sigmay, *T;
    /* multi-GPU (wf_comm_init with nranks > 1): the arrays above hold THIS RANK's particles; global_index gives their creation
     * index in the whole domain (NULL: 0..n_sph-1) and n_sph_global the domain's SPH particle count (0: n_sph) */
    const int64_t *global_index;
    int64_t n_sph_global;
} wf_particles;

/* Rigid tool mesh, replaces TriMesh (Source/Mesh.h:44-75) + Domain::AddTrimeshParticles (Source/Contact.cpp:10-38).
 * Element particles are appended after all SPH particles in call order. */
typedef struct wf_mesh {
    int32_t n_nodes, n_elems;
    const double *nodes;       /* 3*n_nodes */
    const int32_t *elems;      /* 3*n_elems node ids */
    const double *normals;     /* 3*n_elems initial normals (AxisPlaneMesh may have flipped them) */
    const int32_t *nfar;       /* n_elems, node slot used for pplane (Mesh.cpp:206-228) */
    const double *radius;      /* n_elems */
    double hfac;               /* h = (float)hfac * radius */
    int32_t id;                /* contact_surf_id */
} wf_mesh;

/* Scalars the solver loop and user callbacks read (SolverFraser.cpp:139-159,437-467; Domain.h). */
typedef struct wf_scalars {
    double time;
    double max_pl_strain;
    double max_disp[3];
    double kin_energy_sum, int_energy_sum, plastic_work;
    double contact_force_sum, contact_reaction_sum, ext_forces_work, contact_friction_work, accum_cont_heat_cond;
    double max_T, min_T;
    int64_t steps, nb_rebuilds;
    int32_t is_yielding;
    double deltat;             /* current time step (changes under auto_ts) */
} wf_scalars;

/* Field selectors for wf_download / wf_upload. Tensor fields use 9 doubles (row-major 3x3) per particle on the host side. */
typedef enum wf_field {
    WF_F_X = 0, WF_F_V, WF_F_A, WF_F_DISPLACEMENT, WF_F_NORMAL, WF_F_CONTFORCE,            /* 3 doubles */
    WF_F_DENSITY, WF_F_DDENSITY, WF_F_PRESSURE, WF_F_MASS, WF_F_H, WF_F_PL_STRAIN, WF_F_SIGMAY, WF_F_EFF_STRAIN_RATE,
    WF_F_T, WF_F_DTDT, WF_F_Q_PLHEAT, WF_F_Q_CONV, WF_F_DELTA_CONT, WF_F_DKIN_ENERGY_DT, WF_F_DINT_ENERGY_DT, /* 1 double */
    WF_F_SIGMA, WF_F_SHEARSTRESS, WF_F_STRAINRATE, WF_F_ROTATIONRATE, WF_F_STRAIN, WF_F_STRAIN_PL, WF_F_TIR, WF_F_GRADCORRM,   /* 9 doubles */
    WF_F_ID, WF_F_NB, WF_F_GLOBAL_INDEX,                                                                        /* 1 int32 */
    WF_F_COUNT
} wf_field;

/* ---- lifetime ------------------------------------------------------------------------------------------------ */
int  wf_create(int device, wf_ctx **out);
void wf_destroy(wf_ctx *ctx);
const char *wf_last_error(const wf_ctx *ctx);      /* ctx may be NULL: last error of wf_create */
/* sizeof of the structures that cross this boundary, in the order wf_params, wf_material, wf_particles, wf_mesh, wf_scalars: a binding in another
 * language (ctypes, cgo ...) checks its own declarations against the library it loaded. No device needed. */
int wf_abi_sizes(int64_t out[5]);
int  wf_sync(wf_ctx *ctx);
void wf_default_params(wf_params *p);               /* Domain::Domain defaults */

/* ---- setup (replaces the driver-side loops over dom.Particles[a]-> and the Domain data members) --------------- */
int wf_set_params(wf_ctx *ctx, const wf_params *p);
int wf_set_materials(wf_ctx *ctx, const wf_material *m, int n);
int wf_set_particles(wf_ctx *ctx, const wf_particles *p);
/* Domain::GhostPairs (Domain.h:473) for Domain::MoveGhost (Domain.cpp:1142-1236), ghost_type Mirror_XYZ: n x (inner, ghost) creation indices,
 * axis = ghost_plane_axis per pair (NULL: 2, what AddCylinderLength creates). After wf_set_particles, before wf_solver_init. On several GPUs every
 * rank is given ALL pairs (creation indices of the whole domain); a ghost and the particle it mirrors must end up on the same rank (they do unless the
 * mirror plane crosses a slab boundary), otherwise the neighbour build fails on every rank. The Symmetric-plane ghosts of AddCylSliceLength are not implemented. */
int wf_set_ghost_pairs(wf_ctx *ctx, const int32_t *pairs, const int32_t *axis, int64_t n);
int wf_add_mesh(wf_ctx *ctx, const wf_mesh *m);                       /* AddTrimeshParticles */
int wf_mesh_set_motion(wf_ctx *ctx, int mesh, const double v[3], const double w[3]);   /* TriMesh::SetVel / SetRotAxisVel */
int wf_mesh_set_ramp(wf_ctx *ctx, int mesh, const double vmax[3], double tau);         /* v(t) = vmax*min(t/tau,1), evaluated on device */
/* Device-side restatement of the usual GeneralAfter callback (examples/Compression_thermal.cpp:30-66): particles whose
 * ID == id get a = 0 (if set_a) and v = vmax*min(t/tau,1), at both callback points of the step. */
int wf_add_velocity_bc(wf_ctx *ctx, int id, const double vmax[3], double tau, int set_a);
int wf_clear_bcs(wf_ctx *ctx);

/* ---- data movement -------------------------------------------------------------------------------------------- */
int64_t wf_num_particles(const wf_ctx *ctx);        /* SPH + mesh particles */
int64_t wf_num_sph(const wf_ctx *ctx);
int wf_download(wf_ctx *ctx, wf_field f, void *host);                 /* blocking; creation order; all particles */
int wf_upload(wf_ctx *ctx, wf_field f, const void *host);             /* blocking; X,V,A,DENSITY,T,SIGMAY,ID only */
int wf_download_subset(wf_ctx *ctx, wf_field f, const int32_t *idx, int64_t n, void *host);   /* callback bridge */
int wf_upload_subset(wf_ctx *ctx, wf_field f, const int32_t *idx, int64_t n, const void *host);
/* Registered subsets: a GeneralAfter callback touches the same tagged particles every step; their index list (local creation order)
 * is kept on the device so that the per-step bridge moves field values only. Indices refer to this rank's current particles:
 * register again after particles migrated between ranks (wf_comm_info). */
int wf_subset_register(wf_ctx *ctx, const int32_t *idx, int64_t n, int *handle);
int wf_subset_release(wf_ctx *ctx, int handle);
int wf_download_registered(wf_ctx *ctx, wf_field f, int handle, void *host);
int wf_upload_registered(wf_ctx *ctx, wf_field f, int handle, const void *host);
/* same without waiting for the copy: `host` (pinned) must stay untouched until the next blocking call on the ctx (wf_sync, wf_reduce,
 * any download); the upload is ordered before every later phase call on the ctx stream */
int wf_upload_registered_async(wf_ctx *ctx, wf_field f, int handle, const void *host);
/* one value (3 doubles for X, V, A; 1 for T) for every particle of a registered subset: the GeneralAfter callbacks of the reference's drivers
 * assign the same v and a to a whole ID group (examples/Compression_thermal.cpp:30-66); asynchronous on the ctx stream */
int wf_fill_registered(wf_ctx *ctx, wf_field f, int handle, const double *value);
int wf_mesh_download(wf_ctx *ctx, int mesh, double *nodes, double *normals, double *pplane);

/* ---- hot path, one call per phase of the Fraser step (SolverFraser.cpp:133-505) -------------------------------- */
int wf_solver_init(wf_ctx *ctx, double dt);         /* InitialChecks, CellInitiate, ListGenerate, contact prologue (:41-94) */
int wf_nb_build(wf_ctx *ctx);                       /* ClearNbData+MainNeighbourSearch+CalcPairPosList+SaveNeighbourData(+ContactNbUpdate) */
int wf_nb_export_pairs(wf_ctx *ctx, int64_t *ij, int64_t cap, int64_t *n_pairs);      /* sorted (i<j), creation numbering */
int wf_get_grid(wf_ctx *ctx, double blpf[3], double trpr[3], double cell_size[3], int32_t cell_no[3]);   /* Domain::BLPF, TRPR, CellSize, CellNo after CellInitiate (Neighbour.cpp:10-139) */
int wf_cont_export_pairs(wf_ctx *ctx, int64_t *ij, int64_t cap, int64_t *n_pairs);
int wf_grad_corr(wf_ctx *ctx);                      /* CalcGradCorrMatrix (Domain.cpp:1832-1940) on the current lists; no-op unless grad_kernel_corr */
int wf_step_begin(wf_ctx *ctx);                     /* StartAcceleration: resets + TIR (Domain.cpp:1519-1597) */
int wf_density(wf_ctx *ctx);                        /* CalcDensInc + DensReduction + rho += dt*drho (:243-252) */
int wf_rate_stress(wf_ctx *ctx);                    /* CalcRateTensors + RateTensorsReduction + CalcStressStrain (:261-271) */
int wf_accel(wf_ctx *ctx);                          /* CalcAccel + AccelReduction (:280-289) */
int wf_contact(wf_ctx *ctx);                        /* CalcContactForcesWang (:307-312) */
int wf_apply_bcs(wf_ctx *ctx);                      /* device-side GeneralAfter (:332,:360) */
int wf_integrate_pos(wf_ctx *ctx);                  /* x_prev, Displacement, x update (:337-343) */
int wf_integrate_vel(wf_ctx *ctx);                  /* v += a*dt (:352-355) */
int wf_move_ghost(wf_ctx *ctx);                     /* MoveGhost (:335,:358); part of wf_fraser_stage_b1 / b2 and of wf_solver_stage 1..3; no-op without ghost pairs */
int wf_thermal(wf_ctx *ctx);                        /* CalcPlasticWorkHeat + ThermalCalcs (:369-370) */
int wf_energy(wf_ctx *ctx);                         /* CalcKinEnergyEqn + CalcIntEnergyEqn (:375-376) */
int wf_step_end(wf_ctx *ctx);                       /* Time += dt, TriMesh::Update, UpdateContactParticles, cadence (:384-412) */
int wf_reduce(wf_ctx *ctx, wf_scalars *out);        /* blocking: scalars */
/* Domain::m_contact_force (Domain.h:302; summed per rigid mesh in CalcContactForcesWang, Contact_Wang.cpp:312; the History.csv columns of
 * SolverFraser.cpp:476-482): out[3*m + k] for the first cap_meshes meshes, as of the last completed step. Blocking. At most 6 meshes. */
int wf_mesh_forces(wf_ctx *ctx, double *out, int cap_meshes);

/* Whole steps. stage_a = begin .. contact; [GeneralAfter]; stage_b = move x, v; [GeneralAfter]; stage_c = thermal .. end. */
int wf_fraser_stage_a(wf_ctx *ctx);
int wf_fraser_stage_b1(wf_ctx *ctx);
int wf_fraser_stage_b2(wf_ctx *ctx);
int wf_fraser_stage_c(wf_ctx *ctx);
int wf_fraser_run(wf_ctx *ctx, int64_t nsteps);     /* nsteps full steps with device-side BCs, no host callback */
/* Whole steps of the solver wf_params::solver selects (Fraser / KickDrift / LeapFrog step order, ThermalSolve), device-side BCs.
 * wf_run: nsteps steps. wf_run_until: while (Time <= tf) as the reference's loops do (the only stop rule that makes sense
 * with auto_ts; ThermalSolve's loop stops at Time < tf, Thermal.cpp:626); *steps_done receives the number of steps taken. */
int wf_run(wf_ctx *ctx, int64_t nsteps);
/* One KickDrift / LeapFrog step in five stages split at the reference loops' GeneralAfter calls (host-callback bridge): 0 rebuild
 * policy + TIR + CalcAccel | 1 contact (+ LeapFrog's first-step half kick) | 2 adaptive time step + velocity kick | 3 density + drift
 * (+ KickDrift's second half kick) | 4 rate tensors + stresses + heat + energies + end of step. No device-side BCs are applied. */
int wf_solver_stage(wf_ctx *ctx, int stage);
int wf_run_until(wf_ctx *ctx, double tf, int64_t *steps_done);

/* ---- host-side lattice generators (Domain::AddCylinderLength / AddBoxLength, Domain.cpp:708,326); 3-D, cubic packing.
 * Call with pts = NULL to get the count. Returns particle count or a negative wf_status. ------------------------ */
int64_t wf_gen_cylinder(const double V[3], double Rxy, double Lz, double r, double *pts, int64_t cap);
int64_t wf_gen_box(const double V[3], double Lx, double Ly, double Lz, double r, double *pts, int64_t cap, double box_dims[3]);
/* AddCylinderLength(..., ghost = true), Domain.cpp:795-826: the lattice plus three mirror rows below the bottom plane; pairs = (inner, ghost)
 * creation indices (Domain::GhostPairs). pts / pairs may be NULL to get the counts. The generator's particle mass divides by the count
 * including the ghosts (Domain.cpp:839-841). */
int64_t wf_gen_cylinder_ghost(const double V[3], double Rxy, double Lz, double r, double *pts, int64_t cap, int32_t *pairs, int64_t pair_cap, int64_t *n_pairs);

/* ---- multi-GPU: one process (and one wf_ctx) per GPU, 1-D slab decomposition of the cell grid along its longest axis, NCCL
 * point-to-point halos; the reference is single-process (no counterpart). Rank 0 obtains an id with wf_comm_unique_id and
 * hands it to the other ranks by any means (bench.py: torch.distributed broadcast); every rank calls wf_comm_init BEFORE
 * wf_set_particles and uploads only the particles wf_plan_slabs assigns to it. All per-phase entry points are then collective.
 * Host arrays of wf_download / wf_upload hold this rank's current particles in ascending creation index (WF_F_GLOBAL_INDEX). */
int wf_comm_unique_id(void *out128);
int wf_comm_init(wf_ctx *ctx, int rank, int nranks, const void *id128);
int wf_comm_info(wf_ctx *ctx, int64_t out[8]);      /* n_local, n_halo, migrated_out, migrated_in, layer_lo, layer_hi, halo bytes sent per step, layers */
int wf_plan_slabs(const double *x, const double *h, int64_t n, int kernel_type, const double dom_min[3], const double dom_max[3],
                  int nranks, int32_t *rank_of, int32_t *bounds);   /* host-only planner: same grid + balancing code as the library */
/* Same with rigid meshes: x, h = the n_sph SPH particles followed by the element particles (centroid, (float)hfac*radius), which
 * shape the cell grid like every particle does (Neighbour.cpp:10-139) but are not distributed: each rank is given ALL meshes
 * (wf_add_mesh) and keeps a replica of their element particles; rank_of = -1 for them. */
/* Host-only: the slab-boundary moves the library's time-based rebalancing (at every neighbour rebuild; each boundary by at most one
 * cell layer towards equal step time) would decide from per-rank statistics: 8 doubles per rank = seconds per particle, particles,
 * particles in the first layer, in the last layer, emigrant-list capacity, 3 unused. bounds (nranks + 1 layer indices) is updated in place. */
int wf_plan_rebalance(const double *stats, int nranks, int32_t *bounds);
int wf_plan_slabs_ex(const double *x, const double *h, int64_t n, int64_t n_sph, int kernel_type, const double dom_min[3], const double dom_max[3],
                     int nranks, int32_t *rank_of, int32_t *bounds);

/* timing helper for bench.py: CUDA-event time (ms) of each kernel family accumulated since the last reset */
int wf_prof_enable(wf_ctx *ctx, int on);
int wf_prof_get(wf_ctx *ctx, double *ms, int64_t *launches, int cap, int *n_used);
const char *wf_prof_name(int k);
int wf_launch_count(wf_ctx *ctx, int64_t *launches);
void *wf_get_stream(wf_ctx *ctx);                   /* the ctx-owned cudaStream_t every kernel is launched on (for external event timing) */
int wf_measure_fp64_peak(wf_ctx *ctx, double *tflops);   /* DFMA microbenchmark: the FP64-pipe roofline denominator (2 flop per DFMA) */

#ifdef __cplusplus
}
#endif
#endif
