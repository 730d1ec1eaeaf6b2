// Kernel-argument records of the pair passes, shared by wf_pair.cu (exact arithmetic, -fmad=false) and wf_pair_fm.cu (contracted arithmetic).
#pragma once
#include "wf_internal.cuh"

namespace wf {

// FAST path (3-D cubic spline, uniform h, one property class): every per-thread constant is uniform over the launch.
// They are evaluated once on the host with the same expressions (IEEE, no contraction) and travel as kernel parameters,
// i.e. constant-bank operands: no registers, no per-thread divisions.
struct FastConsts {
    double h, C3, C4, inv_h, W0, inv_W0, Alpha, Beta, TI, TIn, TId, Cs, RefDensity, kT, kk, ksum, inv_ksum, cpT, th_exp, q_source;
    double AlphaAvg, BetaAvg, AlphaCij, h2_001;    // lean momentum kernel: (Alpha+Alpha)/2, (Beta+Beta)/2, AlphaAvg*(0.5*(Cs+Cs)), 0.01*h*h
    int PresEq, fast_h, fast_W0, fast_k;
    int own_uniform;                                // one class: cp_T, th_exp, q_source of the particle itself are fc's; 0: read them from the particle's class
    double mass; int uniform_mass;                  // every particle of the domain has this mass (0: read m_j from Q2)
};
struct PairArgs {
    FastConsts fc;
    int64_t n, cap, s0;           // slots [s0, n) of this launch (s0 = 0, n = n_loc unless the launch covers a sub-range)
    int nsph, K, dim, kt, gt;
    const uint32_t *nbr; const int *nb_cnt, *nb_hi; const int *orig; const int *cls_of; const int *is_free;   // is_free: NULL when every particle is free
    double4 *Q0, *Q1, *Q2, *S0, *S1, *S2; double *F; const ClassProps *cls; const MatProps *mats;
    double dt; int any_ti, thermal, pl_heating, contact, cont_heat_fric; double pl_work_heat_frac;
    int gc;                       // Domain::gradKernelCorr: corrected gradients, pairs in stored (first, second) roles, lists in plain discovery order
};


// wf_pair_fm.cu: the contracted-arithmetic ("fm") variants of the lean pair kernels, selected by wf_params::tuning bit 10
int launch_accel_fm(wf_ctx *c, const PairArgs &a, int grid, bool ti_on);
int launch_rate_fm(wf_ctx *c, const PairArgs &a, int grid);
int launch_post_fm(wf_ctx *c, const PairArgs &a, int grid, bool therm, bool ekin, bool dens);   // -1: no such variant

} // namespace wf
