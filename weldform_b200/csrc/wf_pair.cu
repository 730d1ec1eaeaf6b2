// Pair loops of the Fraser step, gather form: one thread per particle walks its own neighbour list over cell-sorted,
// 32-byte packed arrays; no atomics, no pair tables. Replaces the per-pair passes + Nishiura reductions
//   CalcDensInc + DensReduction              Source/InteractionAlt.cpp:483-640
//   CalcRateTensors + RateTensorsReduction   Source/InteractionAlt.cpp:263-479   (+ Particle::CalcStressStrain, Particle.cpp:960-1067)
//   CalcAccel + AccelReduction               Source/InteractionAlt.cpp:12-260
//   CalcTempInc                              Source/Thermal.cpp:53-217
//   CalcKinEnergyEqn                         Source/Domain.cpp:1752-1808
// and Domain::StartAcceleration (Source/Domain.cpp:1519-1597).
//
// Three pair passes per step instead of the reference's five:
//   k_rate_stress   strain/rotation-rate sums + the whole constitutive update
//   k_accel         momentum
//   k_post          heat conduction + kinetic-energy rate + (for the NEXT step) density rate and tensile-instability tensor:
//                   between the end of step n and the density pass of step n+1 neither x, v nor rho change, so the
//                   three loops share one walk (x_ij, r, q, grad-W evaluated once). When the neighbour list is rebuilt
//                   or the host uploads state in between, the prologue falls back to the stand-alone pass (k_post<DENS only>).
//
// Gathered arrays (one 32-byte sector per neighbour and array):
//   Q0 = (x, y, z, h)   Q1 = (vx, vy, vz, m/rho)   Q2 = (m, rho, T, P/rho^2)
//   S0 = (pxx, pyy, pzz, pxy)  S1 = (pyz, pxz, Rxx, Ryy)  S2 = (Rzz, Rxy, Ryz, Rxz)
//   with p = Sigma/rho^2 for GradientType 0 (the reference's 1/(rho*rho)*Sigma factor, evaluated once per particle
//   instead of once per pair -- same expression, same bits) or Sigma itself for GradientType 1/2, and R = TIR.
//
// Each pair term is evaluated with the reference's operand order. The reference evaluates a pair once with
// P1 = min(i,j), P2 = max(i,j) and adds +m_j*t to P1, -m_i*t to P2; every term used here is exactly symmetric or
// antisymmetric under i<->j in IEEE arithmetic (x_ij, v_ij flip sign exactly; h, rho sums commute), so evaluating it
// from particle i's side gives the same bits. GradientType 2 (sigma_i - sigma_j) is the exception and is evaluated in
// min/max order explicitly. Sums run in the reference's order (see wf_neighbour.cu).
//
// Divisions by a quantity that is constant for the thread (h, W(dx0/h), k_i + k_j) use div_c(): a correctly rounded
// quotient from the correctly rounded reciprocal and two FMA residual corrections (Markstein), bit-identical to an IEEE
// division, cheaper than the generic division sequence. Guarded against the one divisor pattern (all-ones significand)
// for which the theorem does not hold.
#include <cstring>
#include <cub/cub.cuh>
#include "wf_internal.cuh"
#include "wf_pair_args.cuh"
#include "wf_constitutive.cuh"

using namespace wf;

namespace {

#define OWN(f) a.F[(int64_t)(f) * a.cap + s]

// One 256-bit load per 32-byte record (LDG.E.256, sm_100): a gathered double4 costs one L1 request instead of the two
// LDG.E.128 the compiler emits for a double4 (which is only 16-byte aligned by type). All gathered arrays are cudaMalloc'ed
// (256-byte aligned) arrays of 32-byte records.
template <bool W256> __device__ __forceinline__ double4 ldrec(const double4 *p)
{
    if (W256) {
        double4 r;
        asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
        return r;
    }
    return *p;
}

__device__ inline int nb_row(int n, int nhi, int K) { return n < nhi ? n : (K - 1 - (n - nhi)); }

// kernel constants for the 3-D cubic spline, same expressions as Functions.cpp:33,82
__device__ inline double c3_of(double h) { return 1.0 / (h * h * h * WF_PI); }
__device__ inline double c4_of(double h) { return 1.0 / (h * h * h * h * WF_PI); }

// ---- division by a per-thread constant -------------------------------------------------------------------------------
struct DivC { double b, y; int fast; };
__device__ inline DivC divc_make(double b)
{
    DivC d; d.b = b; d.y = 1.0 / b;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(b);
    const unsigned ex = (unsigned)((bits >> 52) & 0x7ff);
    // exclusions: all-ones significand (Markstein's theorem needs RN(1/b) with |b*y - 1| < 2^-53), and extreme exponents
    d.fast = ((bits & 0xFFFFFFFFFFFFFull) != 0xFFFFFFFFFFFFFull) && ex > 200 && ex < 1800;
    return d;
}
__device__ inline DivC divc_from(double b, double y, int fast) { DivC d; d.b = b; d.y = y; d.fast = fast; return d; }
__device__ inline double div_c(double a, const DivC &d)
{
    if (!d.fast) return a / d.b;
    const unsigned ea = (unsigned)(((unsigned long long)__double_as_longlong(a) >> 52) & 0x7ff);
    if (ea < 200 || ea > 1800) return a / d.b;                 // zero / subnormal / huge numerators: generic path
    const double q0 = a * d.y;
    const double r0 = __fma_rn(-q0, d.b, a);
    const double q1 = __fma_rn(r0, d.y, q0);                   // faithful
    const double r1 = __fma_rn(-q1, d.b, a);                   // exact
    return __fma_rn(r1, d.y, q1);                              // RN(a/b)
}

template <bool FAST> __device__ inline double GK_of(int dim, int kt, double q, double hh, double C4)
{ if (FAST) return cubicGW_c(C4, q, hh); else return gradW(dim, kt, q, hh); }
template <bool FAST> __device__ inline double W_of(int dim, int kt, double q, double hh, double C3)
{ if (FAST) return cubicW_c(C3, q); else return kernelW(dim, kt, q, hh); }

// x^4 rounded from a double-double product: reproduces a correctly rounded pow(x, 4.0) (InteractionAlt.cpp:108 with
// the default TIn = 4) without the cost of a general fp64 pow.
__device__ inline double pow4_cr(double r)
{
    double r2h = r * r, r2l = __fma_rn(r, r, -r2h);
    double r4h = r2h * r2h, e1 = __fma_rn(r2h, r2h, -r4h);
    return r4h + (e1 + 2.0 * r2h * r2l);
}

// Domain.cpp:1528-1570: tensile-instability tensor from Sigma and rho
__device__ inline void tir_of(const S6 &sg, double rho, double TI, double out[6])
{
    // No principal stress is positive <=> Sigma is negative semi-definite; then every Val(k,k) below is 0 and TIR = Vec*0*Vec^T = 0
    // whatever the eigenvectors are. Sylvester's criterion on -Sigma with margins nine orders of magnitude above the Jacobi
    // iteration's round-off (all leading minors > 1e-6 * scale^k  =>  every eigenvalue < -1e-7 * scale) decides the common
    // compressive state without the eigen-solve, which costs as much as a third of the pair walk of this pass.
    {
        const double sc = fmax(fmax(fabs(sg.xx), fabs(sg.yy)), fmax(fmax(fabs(sg.zz), fabs(sg.xy)), fmax(fabs(sg.yz), fabs(sg.xz))));
        const double m2 = sg.xx * sg.yy - sg.xy * sg.xy;
        const double det = sg.xx * (sg.yy * sg.zz - sg.yz * sg.yz) - sg.xy * (sg.xy * sg.zz - sg.yz * sg.xz) + sg.xz * (sg.xy * sg.yz - sg.yy * sg.xz);
        if (-sg.xx > 1e-6 * sc && m2 > 1e-6 * sc * sc && -det > 1e-6 * sc * sc * sc) { for (int k = 0; k < 6; ++k) out[k] = 0.0; return; }
    }
    const double rho2 = rho * rho;
    const double pc = TI / (rho2);
    double d[3]; M3 Vec; sym_eig3(m3sym(sg), d, Vec);
    M3 VecT = mtrans(Vec), Val = m3zero();
    for (int k = 0; k < 3; ++k) { if (d[k] > 0) Val(k, k) = -pc * d[k]; else Val(k, k) = 0.0; }
    M3 T = mmul(mmul(Vec, Val), VecT);
    // stored symmetric (upper triangle); the reference's TIR differs from its transpose by round-off only
    out[0] = T(0, 0); out[1] = T(1, 1); out[2] = T(2, 2); out[3] = T(0, 1); out[4] = T(1, 2); out[5] = T(0, 2);
}

// ---- Bonet-Lok gradient correction (Domain::gradKernelCorr) -------------------------------------------------------------
// Mult(GK * gradCorrM, xij) as InteractionAlt.cpp:118-131 evaluates it: the matrix is scaled first (matvec.h:547, element*GK),
// then B(r) = x0*M(r,0) + x1*M(r,1) + x2*M(r,2) (matvec.h:836-841)
__device__ inline V3 gc_scaled_mat_vec(const double M[9], double GK, const V3 &x)
{
    double g[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) g[k] = M[k] * GK;
    return v3(x.x * g[0] + x.y * g[1] + x.z * g[2], x.x * g[3] + x.y * g[4] + x.z * g[5], x.x * g[6] + x.y * g[7] + x.z * g[8]);
}
// Mult(gradCorrM, GK*xij) as Thermal.cpp:101 evaluates it (matrix unscaled, vector scaled)
__device__ inline V3 gc_mat_vec(const double M[9], const V3 &g)
{ return v3(g.x * M[0] + g.y * M[1] + g.z * M[2], g.x * M[3] + g.y * M[4] + g.z * M[5], g.x * M[6] + g.y * M[7] + g.z * M[8]); }

// Domain::CalcGradCorrMatrix (Domain.cpp:1832-1940), gather form: M_i = - sum_j (m_j/rho_j) (GK xij) (x) xij over the list in
// discovery order (k_gradcorr), plus the pairs with a fixed particle (k_gradcorr_fsm, wf_neighbour.cu), then InvLog (k_gradcorr_inv).
__global__ void __launch_bounds__(128) k_gradcorr(PairArgs a)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + a.s0;
    if (s >= a.n) return;
    if (a.orig[s] >= a.nsph) return;
    const double4 qi = a.Q0[s];
    const V3 xi = v3(qi.x, qi.y, qi.z);
    const int cnt = a.nb_cnt[s], nhi = a.nb_hi[s];
    double T[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) T[k] = 0.0;
    for (int n = 0; n < cnt; ++n) {
        const uint32_t e = a.nbr[(int64_t)nb_row(n, nhi, a.K) * a.cap + s];
        const int j = (int)(e & 0x7fffffffu); const bool ifirst = (e >> 31) != 0;
        const double4 qj = a.Q0[j], q2j = a.Q2[j];
        const V3 xj = v3(qj.x, qj.y, qj.z);
        const V3 xij = ifirst ? (xi - xj) : (xj - xi);
        const double hh = (qi.w + qj.w) / 2.0;
        const double q = norm(xij) / hh;
        const double GK = (a.kt == 0 && a.dim == 3) ? cubicGW_c(c4_of(hh), q, hh) : gradW(a.dim, a.kt, q, hh);
        const double f = q2j.x / q2j.y;                 // m/rho of the other particle of the pair (Domain.cpp:1859-1860)
        const V3 g = GK * xij;
        const double gx[3] = { g.x, g.y, g.z }, xx[3] = { xij.x, xij.y, xij.z };
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) T[3 * r + c] = T[3 * r + c] - (gx[r] * xx[c]) * f;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) OWN(F_GC + k) = T[k];        // raw sum; k_gradcorr_fsm (fixed neighbours) may add to it, k_gradcorr_inv inverts
}
// InvLog (matvec.h:673-693) of the summed matrix; identity when the determinant is below 1e-10 and for fixed particles (Domain.cpp:1916-1937)
__global__ void __launch_bounds__(128) k_gradcorr_inv(PairArgs a)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + a.s0;
    if (s >= a.n) return;
    if (a.orig[s] >= a.nsph) return;
    double T[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) T[k] = OWN(F_GC + k);
    const double det = T[0] * (T[4] * T[8] - T[5] * T[7]) - T[1] * (T[3] * T[8] - T[5] * T[6]) + T[2] * (T[3] * T[7] - T[4] * T[6]);
    double Mi[9];
    if (fabs(det) < 1.0e-10 || (a.is_free && !a.is_free[s])) { for (int k = 0; k < 9; ++k) Mi[k] = (k == 0 || k == 4 || k == 8) ? 1.0 : 0.0; }
    else {
        Mi[0] = (T[4] * T[8] - T[5] * T[7]) / det; Mi[1] = (T[2] * T[7] - T[1] * T[8]) / det; Mi[2] = (T[1] * T[5] - T[2] * T[4]) / det;
        Mi[3] = (T[5] * T[6] - T[3] * T[8]) / det; Mi[4] = (T[0] * T[8] - T[2] * T[6]) / det; Mi[5] = (T[2] * T[3] - T[0] * T[5]) / det;
        Mi[6] = (T[3] * T[7] - T[4] * T[6]) / det; Mi[7] = (T[1] * T[6] - T[0] * T[7]) / det; Mi[8] = (T[0] * T[4] - T[1] * T[3]) / det;
    }
#pragma unroll
#pragma unroll
    for (int k = 0; k < 9; ++k) OWN(F_GC + k) = Mi[k];
}

// ---- StartAcceleration's tensile-instability tensor (Domain.cpp:1528-1570) as a per-particle pass of its own. The eigen-solve is
// register-hungry and as long as a third of a pair walk: kept out of k_post it runs at full occupancy and leaves k_post its registers.
// HANDOVER = true (Fraser step): the tensor of the NEXT step, written to F_TN only (the prologue hands it over);
// HANDOVER = false (KickDrift / LeapFrog: momentum sums first): also written to the gathered arrays
template <bool HANDOVER> __global__ void __launch_bounds__(128) k_tir(PairArgs a)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + a.s0;
    if (s >= a.n) return;
    if (a.orig[s] >= a.nsph) return;
    const double TI = a.cls[a.cls_of[s]].TI;
    if (!(TI > 0.0) || (a.is_free && !a.is_free[s])) return;
    S6 sg; sg.xx = OWN(F_SG); sg.yy = OWN(F_SG + 1); sg.zz = OWN(F_SG + 2); sg.xy = OWN(F_SG + 3); sg.yz = OWN(F_SG + 4); sg.xz = OWN(F_SG + 5);
    double t[6]; tir_of(sg, a.Q2[s].y, TI, t);
    if (!HANDOVER) {            // StartAcceleration proper: the gathered copy is needed by the momentum pass that follows
        ((double2 *)&a.S1[s])[1] = make_double2(t[0], t[1]);
        a.S2[s] = make_double4(t[2], t[3], t[4], t[5]);
    }
    for (int k = 0; k < 6; ++k) OWN(F_TN + k) = t[k];
}

// ---- Domain::CalcTempIncPP / CalcTempIncPair (ThermalTest.cpp:8-49): the initial dT/dt of ThermalSolve, every pair from particle
// i's side over the Nishiura order (the list order), dot(xij, xij) in the denominator, times 1/(rho_i cp_i) at the end
__global__ void __launch_bounds__(128) k_temp_inc_pp(PairArgs a)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + a.s0;
    if (s >= a.n) return;
    double acc = 0.0;
    const double4 qi = a.Q0[s], q2i = a.Q2[s];
    const ClassProps cpi = a.cls[a.cls_of[s]];
    if (a.orig[s] < a.nsph) {
        const V3 xi = v3(qi.x, qi.y, qi.z);
        const int cnt = a.nb_cnt[s], nhi = a.nb_hi[s];
        for (int n = 0; n < cnt; ++n) {
            const int j = (int)(a.nbr[(int64_t)nb_row(n, nhi, a.K) * a.cap + s] & 0x7fffffffu);
            const double4 qj = a.Q0[j], q2j = a.Q2[j];
            const double kj = a.cls[a.cls_of[j]].k_T;
            const V3 xij = xi - v3(qj.x, qj.y, qj.z);
            const double hh = (qi.w + qj.w) / 2.0;
            const double q = norm(xij) / hh;
            const double GK = (a.kt == 0 && a.dim == 3) ? cubicGW_c(c4_of(hh), q, hh) : gradW(a.dim, a.kt, q, hh);
            acc += q2j.x / q2j.y * 4. * (cpi.k_T * kj) / (cpi.k_T + kj) * (q2i.z - q2j.z) * dot(xij, GK * xij) / (dot(xij, xij));
        }
    }
    OWN(F_DTDT) = acc * (1. / (q2i.y * cpi.cp_T));
}

// ---- Domain::CheckMinTSVel (Domain.cpp:267-297): min over all particles of CFL * (nearest listed neighbour) / (Cs + |v|);
// the CFL factor and the comparison with deltatint are applied by the caller
__global__ void __launch_bounds__(128) k_min_ts(PairArgs a, double cfl, double sqrt_h_a, double *partial)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + a.s0;
    double test = 1.0e300;
    if (s < a.n) {
        const double4 qi = a.Q0[s], wi = a.Q1[s];
        const V3 xi = v3(qi.x, qi.y, qi.z);
        double mn = 1000.;
        if (sqrt_h_a != 0.0) {      // Domain::CheckMinTSAccel (Domain.cpp:302-317): sqrt_h_a * sqrt(h / |a|); 0 selects the velocity criterion
            test = sqrt_h_a * sqrt(qi.w / norm(v3(OWN(F_AX), OWN(F_AY), OWN(F_AZ))));
        } else {
        if (a.orig[s] < a.nsph) {
            const int cnt = a.nb_cnt[s], nhi = a.nb_hi[s];
            for (int n = 0; n < cnt; ++n) {
                const int j = (int)(a.nbr[(int64_t)nb_row(n, nhi, a.K) * a.cap + s] & 0x7fffffffu);
                const double4 qj = a.Q0[j];
                const double d = norm(v3(qj.x, qj.y, qj.z) - xi);
                if (d < mn) mn = d;
            }
        }
        test = cfl * mn / (a.cls[a.cls_of[s]].Cs + norm(v3(wi.x, wi.y, wi.z)));
        }
    }
    typedef cub::BlockReduce<double, 128> BR; __shared__ typename BR::TempStorage tmp;
    const double r = BR(tmp).Reduce(test, cub::Min());
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}
__global__ void __launch_bounds__(256) k_min_final(const double *partial, int n, double *out)
{
    double v = 1.0e300;
    for (int b = threadIdx.x; b < n; b += blockDim.x) v = fmin(v, partial[b]);
    typedef cub::BlockReduce<double, 256> BR; __shared__ typename BR::TempStorage tmp;
    const double r = BR(tmp).Reduce(v, cub::Min());
    if (threadIdx.x == 0) out[0] = r;
}

// ---- step prologue per particle: rho += dt*drho (SolverFraser.cpp:248-252), m/rho, TIR <- the tensor prepared for this step
__global__ void k_prologue(PairArgs a)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + a.s0;
    if (s >= a.n) return;
    if (a.orig[s] >= a.nsph) return;
    const double dd = OWN(F_DDN);
    OWN(F_DDENS) = dd;
    double4 q2 = a.Q2[s];
    const double rho = q2.y + a.dt * dd;
    q2.y = rho;
    a.Q2[s] = q2;
    ((double *)&a.Q1[s])[3] = q2.x / rho;
    if (a.any_ti) {
        ((double2 *)&a.S1[s])[1] = make_double2(OWN(F_TN), OWN(F_TN + 1));
        a.S2[s] = make_double4(OWN(F_TN + 2), OWN(F_TN + 3), OWN(F_TN + 4), OWN(F_TN + 5));
    }
}

// ---- CalcRateTensors + RateTensorsReduction + CalcStressStrain -------------------------------------------------------
template <bool FAST, bool PF> __global__ void __launch_bounds__(128, 5) k_rate_stress(PairArgs a)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + a.s0;
    if (s >= a.n) return;
    if (a.orig[s] >= a.nsph) return;
    const double4 qi = a.Q0[s], wi = a.Q1[s];
    const V3 xi = v3(qi.x, qi.y, qi.z), vi = v3(wi.x, wi.y, wi.z);
    const double hi = FAST ? a.fc.h : qi.w, C4i = FAST ? a.fc.C4 : c4_of(hi);
    const DivC dh = FAST ? divc_from(a.fc.h, a.fc.inv_h, a.fc.fast_h) : divc_make(hi);
    const int cnt = a.nb_cnt[s], nhi = a.nb_hi[s];
    double sxx = 0., syy = 0., szz = 0., sxy = 0., syz = 0., sxz = 0., r01 = 0., r02 = 0., r12 = 0.;
    // software pipeline: the gathers of neighbour n+1 (and the list entry of n+2) are in flight while n is evaluated
    double4 qn = qi, wn = wi; uint32_t e2 = 0;
    const bool gc = !FAST && a.gc;
    double Mg[9];
    if (!FAST) { for (int k = 0; k < 9; ++k) Mg[k] = gc ? OWN(F_GC + k) : 0.0; }
    if (gc) {
        // CalcRateTensors, locking branch with corrected gradients (InteractionAlt.cpp:362-441): per pair, in discovery order,
        // StrainRate_i += (float)(m_j/rho_j) * (-0.5)(L + L^T), L = v_ab (x) (GK*gradCorrM_i) x_ij, roles as stored in the pair
        for (int n = 0; n < cnt; ++n) {
            const uint32_t e = a.nbr[(int64_t)nb_row(n, nhi, a.K) * a.cap + s];
            const int j = (int)(e & 0x7fffffffu); const bool ifirst = (e >> 31) != 0;
            const double4 qj = a.Q0[j], wj = a.Q1[j];
            const V3 xj = v3(qj.x, qj.y, qj.z), vj = v3(wj.x, wj.y, wj.z);
            const V3 xij = ifirst ? (xi - xj) : (xj - xi), vab = ifirst ? (vi - vj) : (vj - vi);
            const double hh = (hi + qj.w) / 2;
            const double q = norm(xij) / hh;
            const double GK = (a.kt == 0 && a.dim == 3) ? cubicGW_c(c4_of(hh), q, hh) : gradW(a.dim, a.kt, q, hh);
            const V3 gk = gc_scaled_mat_vec(Mg, GK, xij);
            const double mf = (double)(float)wj.w;        // float mj_dj = mj/dj (InteractionAlt.cpp:416,432)
            const double L00 = vab.x * gk.x, L01 = vab.x * gk.y, L02 = vab.x * gk.z, L10 = vab.y * gk.x, L11 = vab.y * gk.y, L12 = vab.y * gk.z,
                         L20 = vab.z * gk.x, L21 = vab.z * gk.y, L22 = vab.z * gk.z;
            sxx = sxx + ((L00 + L00) * -0.5) * mf; syy = syy + ((L11 + L11) * -0.5) * mf; szz = szz + ((L22 + L22) * -0.5) * mf;
            sxy = sxy + ((L01 + L10) * -0.5) * mf; syz = syz + ((L12 + L21) * -0.5) * mf; sxz = sxz + ((L02 + L20) * -0.5) * mf;
            r01 = r01 + ((L01 - L10) * -0.5) * mf; r02 = r02 + ((L02 - L20) * -0.5) * mf; r12 = r12 + ((L12 - L21) * -0.5) * mf;
        }
    }
    if (PF && cnt > 0 && !gc) { const int j0 = (int)(a.nbr[(int64_t)nb_row(0, nhi, a.K) * a.cap + s] & 0x7fffffffu); qn = ldrec<true>(a.Q0 + j0); wn = ldrec<true>(a.Q1 + j0);
                         if (cnt > 1) e2 = a.nbr[(int64_t)nb_row(1, nhi, a.K) * a.cap + s]; }
    for (int n = 0; n < (gc ? 0 : cnt); ++n) {
        double4 qj, wj;
        if (PF) {
            qj = qn; wj = wn;
            if (n + 1 < cnt) { const int j1 = (int)(e2 & 0x7fffffffu); qn = ldrec<true>(a.Q0 + j1); wn = ldrec<true>(a.Q1 + j1); }
            if (n + 2 < cnt) e2 = a.nbr[(int64_t)nb_row(n + 2, nhi, a.K) * a.cap + s];
        } else {
            const int j = (int)(a.nbr[(int64_t)nb_row(n, nhi, a.K) * a.cap + s] & 0x7fffffffu);
            qj = a.Q0[j]; wj = a.Q1[j];
        }
        const double mjdj = wj.w;                                    // m_j / rho_j (InteractionAlt.cpp:455), kept per particle
        const V3 xij = xi - v3(qj.x, qj.y, qj.z);
        const double rij = norm(xij);
        const V3 vab = vi - v3(wj.x, wj.y, wj.z);
        double GK;
        if (FAST) GK = cubicGW_c(C4i, div_c(rij, dh), hi);
        else { double hh, C4; if (qj.w == hi) { hh = hi; C4 = C4i; } else { hh = (hi + qj.w) / 2; C4 = c4_of(hh); }
               GK = (a.kt == 0 && a.dim == 3) ? cubicGW_c(C4, rij / hh, hh) : gradW(a.dim, a.kt, rij / hh, hh); }
        const double f = -0.5 * GK;
        // InteractionAlt.cpp:338-361, then :456-457 (StrainRate = StrainRate + mjdj * pair_StrainRate)
        sxx = sxx + ((2.0 * vab.x * xij.x) * f) * mjdj;
        sxy = sxy + ((vab.x * xij.y + vab.y * xij.x) * f) * mjdj;
        sxz = sxz + ((vab.x * xij.z + vab.z * xij.x) * f) * mjdj;
        syy = syy + ((2.0 * vab.y * xij.y) * f) * mjdj;
        syz = syz + ((vab.y * xij.z + vab.z * xij.y) * f) * mjdj;
        szz = szz + ((2.0 * vab.z * xij.z) * f) * mjdj;
        r01 = r01 + ((vab.x * xij.y - vab.y * xij.x) * f) * mjdj;
        r02 = r02 + ((vab.x * xij.z - vab.z * xij.x) * f) * mjdj;
        r12 = r12 + ((vab.y * xij.z - vab.z * xij.y) * f) * mjdj;
    }
    stress_update(a, s, sxx, syy, szz, sxy, syz, sxz, r01, r02, r12);
}

// ---- CalcAccel + AccelReduction ------------------------------------------------------------------------------------------
template <bool FAST, int GT, int MB, bool PF> __global__ void __launch_bounds__(128, MB) k_accel(PairArgs a)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + a.s0;
    if (s >= a.n) return;
    if (a.orig[s] >= a.nsph) return;
    const double4 qi = a.Q0[s], wi = a.Q1[s], q2i = a.Q2[s];
    const V3 xi = v3(qi.x, qi.y, qi.z), vi = v3(wi.x, wi.y, wi.z);
    const double hi = FAST ? a.fc.h : qi.w, C3i = FAST ? a.fc.C3 : c3_of(hi), C4i = FAST ? a.fc.C4 : c4_of(hi), di = q2i.y;
    const int ci = FAST ? 0 : a.cls_of[s];
    // property record: constant-bank fields on the FAST path, a per-thread copy of the class otherwise
    struct { double Alpha, Beta, TI, TIn, TIInitDist, Cs, RefDensity; int PresEq; } cpi;
    if (FAST) { cpi.Alpha = a.fc.Alpha; cpi.Beta = a.fc.Beta; cpi.TI = a.fc.TI; cpi.TIn = a.fc.TIn; cpi.TIInitDist = a.fc.TId; cpi.Cs = a.fc.Cs; cpi.RefDensity = a.fc.RefDensity; cpi.PresEq = a.fc.PresEq; }
    else { const ClassProps cp0 = a.cls[ci]; cpi.Alpha = cp0.Alpha; cpi.Beta = cp0.Beta; cpi.TI = cp0.TI; cpi.TIn = cp0.TIn; cpi.TIInitDist = cp0.TIInitDist; cpi.Cs = cp0.Cs; cpi.RefDensity = cp0.RefDensity; cpi.PresEq = cp0.PresEq; }
    const double4 s0i = a.S0[s], s1i = a.S1[s], s2i = a.S2[s];
    const int cnt = a.nb_cnt[s], nhi = a.nb_hi[s];
    const double Ci = soundSpeed(cpi.PresEq, cpi.Cs, di, cpi.RefDensity);
    // W(dx0/h) of the tensile-instability term for the common same-class, same-h neighbour
    const double W0i = FAST ? a.fc.W0 : W_of<FAST>(a.dim, a.kt, (cpi.TIInitDist + cpi.TIInitDist) / (2.0 * hi), hi, C3i);
    const DivC dh = FAST ? divc_from(a.fc.h, a.fc.inv_h, a.fc.fast_h) : divc_make(hi), dW0 = FAST ? divc_from(a.fc.W0, a.fc.inv_W0, a.fc.fast_W0) : divc_make(W0i);
    const bool ti_on = a.any_ti && (!FAST || cpi.TI > 0.0);
    double ax = 0., ay = 0., az = 0.;
    double4 qn = qi, wn = wi, q2n = q2i, s0n = s0i, s1n = s1i; uint32_t e2 = 0; int jcur = 0; bool ifirst_cur = true;
    const bool gc = !FAST && a.gc;
    double Mg[9];
    if (!FAST) { for (int k = 0; k < 9; ++k) Mg[k] = gc ? OWN(F_GC + k) : 0.0; }
    if (PF && cnt > 0) { const uint32_t e0 = a.nbr[(int64_t)nb_row(0, nhi, a.K) * a.cap + s]; jcur = (int)(e0 & 0x7fffffffu); ifirst_cur = (e0 >> 31) != 0; qn = a.Q0[jcur]; wn = a.Q1[jcur]; q2n = a.Q2[jcur];
                         if (MB == 2) { s0n = a.S0[jcur]; s1n = a.S1[jcur]; }
                         if (cnt > 1) e2 = a.nbr[(int64_t)nb_row(1, nhi, a.K) * a.cap + s]; }
    for (int n = 0; n < cnt; ++n) {
        int j; double4 qj, wj, q2j; bool ifirst_n = true;
        if (PF) {
            j = jcur; qj = qn; wj = wn; q2j = q2n; ifirst_n = ifirst_cur;
        } else {
            const uint32_t en = a.nbr[(int64_t)nb_row(n, nhi, a.K) * a.cap + s];
            j = (int)(en & 0x7fffffffu); ifirst_n = (en >> 31) != 0;
            qj = a.Q0[j]; wj = a.Q1[j]; q2j = a.Q2[j];
        }
        double4 s0j, s1j;
        if (PF && MB == 2) { s0j = s0n; s1j = s1n; }
        else { s0j = a.S0[j]; s1j = a.S1[j]; }            // consumed late in the iteration: their latency overlaps the kernel / viscosity math
        if (PF) {
            if (n + 1 < cnt) { jcur = (int)(e2 & 0x7fffffffu); ifirst_cur = (e2 >> 31) != 0; qn = a.Q0[jcur]; wn = a.Q1[jcur]; q2n = a.Q2[jcur]; if (MB == 2) { s0n = a.S0[jcur]; s1n = a.S1[jcur]; } }
            if (n + 2 < cnt) e2 = a.nbr[(int64_t)nb_row(n + 2, nhi, a.K) * a.cap + s];
        }
        const double mj = q2j.x, dj = q2j.y;
        // GradientType 2 is not symmetric under i<->j: P1 = smaller creation index. With gradKernelCorr the reference evaluates the
        // pair in its stored roles (P1 = first) and subtracts from P2 (InteractionAlt.cpp:26-28,216-225)
        const bool swap = gc ? !ifirst_n : ((GT == 2) && (n >= nhi));
        V3 xij = xi - v3(qj.x, qj.y, qj.z), vij = vi - v3(wj.x, wj.y, wj.z);
        if (swap) { xij = v3(-xij.x, -xij.y, -xij.z); vij = v3(-vij.x, -vij.y, -vij.z); }
        const double rij = norm(xij);
        double hh, C3, C4, q, Alpha, Beta, TIj, TInj, TId_j, Cj; bool same_w0;
        if (FAST) {
            hh = hi; C3 = C3i; C4 = C4i; q = div_c(rij, dh);
            Alpha = (cpi.Alpha + cpi.Alpha) / 2.0; Beta = (cpi.Beta + cpi.Beta) / 2.0; TIj = cpi.TI; TInj = cpi.TIn; TId_j = cpi.TIInitDist;
            Cj = soundSpeed(cpi.PresEq, cpi.Cs, dj, cpi.RefDensity); same_w0 = true;
        } else {
            if (qj.w == hi) { hh = hi; C3 = C3i; C4 = C4i; } else { hh = (hi + qj.w) / 2; C3 = c3_of(hh); C4 = c4_of(hh); }
            q = rij / hh;
            const int cj = a.cls_of[j];
            if (cj == ci) { Alpha = (cpi.Alpha + cpi.Alpha) / 2.0; Beta = (cpi.Beta + cpi.Beta) / 2.0; TIj = cpi.TI; TInj = cpi.TIn; TId_j = cpi.TIInitDist;
                            Cj = soundSpeed(cpi.PresEq, cpi.Cs, dj, cpi.RefDensity); }
            else { const ClassProps cpj = a.cls[cj]; Alpha = (cpi.Alpha + cpj.Alpha) / 2.0; Beta = (cpi.Beta + cpj.Beta) / 2.0; TIj = cpj.TI; TInj = cpj.TIn; TId_j = cpj.TIInitDist;
                   Cj = soundSpeed(cpj.PresEq, cpj.Cs, dj, cpj.RefDensity); }
            same_w0 = (cj == ci) && (qj.w == hi);
        }
        double GK;
        if (FAST) GK = cubicGW_c(C4, q, hh); else GK = (a.kt == 0 && a.dim == 3) ? cubicGW_c(C4, q, hh) : gradW(a.dim, a.kt, q, hh);
        // artificial viscosity (InteractionAlt.cpp:72-94)
        double pi = 0.0; bool pi_on = false;
        if (Alpha != 0.0 || Beta != 0.0) {
            const double vx = dot(vij, xij);
            if (vx < 0) {
                const double MUij = hh * vx / (rij * rij + 0.01 * hh * hh);
                const double Cij = 0.5 * (Ci + Cj);
                pi = (Alpha * Cij * MUij + Beta * MUij * MUij) / (0.5 * (di + dj)); pi_on = true;
            }
        }
        // momentum tensor (InteractionAlt.cpp:138-143)
        double mxx, myy, mzz, mxy, myz, mxz;
        if (GT == 0) {          // S holds 1.0/(rho*rho)*Sigma
            mxx = s0i.x + s0j.x; myy = s0i.y + s0j.y; mzz = s0i.z + s0j.z;
            mxy = s0i.w + s0j.w; myz = s1i.x + s1j.x; mxz = s1i.y + s1j.y;
        } else if (GT == 1) {
            const double f = 1.0 / (di * dj);
            mxx = (s0i.x + s0j.x) * f; myy = (s0i.y + s0j.y) * f; mzz = (s0i.z + s0j.z) * f;
            mxy = (s0i.w + s0j.w) * f; myz = (s1i.x + s1j.x) * f; mxz = (s1i.y + s1j.y) * f;
        } else {
            const double f = swap ? 1.0 / (dj * di) : 1.0 / (di * dj);
            const double sg = swap ? -1.0 : 1.0;     // (Sigma_P1 - Sigma_P2): exact negation when i is P2
            mxx = (sg * (s0i.x - s0j.x)) * f; myy = (sg * (s0i.y - s0j.y)) * f; mzz = (sg * (s0i.z - s0j.z)) * f;
            mxy = (sg * (s0i.w - s0j.w)) * f; myz = (sg * (s1i.x - s1j.x)) * f; mxz = (sg * (s1i.y - s1j.y)) * f;
        }
        if (pi_on) { mxx += pi; myy += pi; mzz += pi; }
        // tensile instability (InteractionAlt.cpp:103-110)
        if (ti_on && (cpi.TI > 0.0 || TIj > 0.0)) {
            const double K = W_of<FAST>(a.dim, a.kt, q, hh, C3);
            double r;
            if (same_w0) r = div_c(K, dW0); else r = K / W_of<FAST>(a.dim, a.kt, (cpi.TIInitDist + TId_j) / (2.0 * hh), hh, C3);
            const double nbar = (cpi.TIn + TInj) / 2.0;
            const double fti = (nbar == 4.0) ? pow4_cr(r) : pow(r, nbar);
            const double4 t2j = a.S2[j];
            mxx += (s1i.z + s1j.z) * fti; myy += (s1i.w + s1j.w) * fti; mzz += (s2i.x + t2j.x) * fti;
            mxy += (s2i.y + t2j.y) * fti; myz += (s2i.z + t2j.z) * fti; mxz += (s2i.w + t2j.w) * fti;
        }
        // temp = (GK*xij)^T M (matvec.h:829-834), a_i += m_j*temp (or -= when i is P2 of a GradientType-2 pair);
        // gradKernelCorr: the corrected vector (GK*gradCorrM_i) xij takes the place of GK*xij (InteractionAlt.cpp:129-133,170-183)
        V3 g = GK * xij;
        if (gc) g = gc_scaled_mat_vec(Mg, GK, xij);
        double tx = g.x * mxx + g.y * mxy + g.z * mxz;
        double ty = g.x * mxy + g.y * myy + g.z * myz;
        double tz = g.x * mxz + g.y * myz + g.z * mzz;
        if (a.dim == 2) tz = 0.0;
        if (swap) { ax -= mj * tx; ay -= mj * ty; az -= mj * tz; }
        else      { ax += mj * tx; ay += mj * ty; az += mj * tz; }
    }
    OWN(F_AX) = ax; OWN(F_AY) = ay; OWN(F_AZ) = az;
}


// ---- CalcAccel + AccelReduction, lean form of the common case --------------------------------------------------------
// FAST path (3-D cubic spline, uniform h, one property class) with GradientType 0, TIn = 4, PresEq = 0 (constant sound
// speed) and Markstein-safe divisors h and W0: every loop invariant is a kernel parameter (constant bank, no registers),
// no generic pow / sound-speed code in the loop, and the neighbour pipeline is unrolled by two with named buffers so that
// the prefetched sectors are consumed where they landed (no register rotation). Same expressions, same operand order
// and same summation order as k_accel<true, 0, ...>: bit-identical results.
struct NbBuf { double4 q, w; int j; };
template <bool TI_ON, bool W256> __device__ __forceinline__ void accel_term(const PairArgs &a, const NbBuf &nb, const V3 &xi, const V3 &vi, double di,
                                                              const double4 &s0i, const double4 &s1i, const double4 &s2i,
                                                              double &ax, double &ay, double &az)
{
    const int j = nb.j;
    const double4 q2j = ldrec<W256>(a.Q2 + j), s0j = ldrec<W256>(a.S0 + j), s1j = ldrec<W256>(a.S1 + j);      // consumed late: their latency overlaps the kernel / viscosity math
    double4 t2j = make_double4(0., 0., 0., 0.);
    if (TI_ON) t2j = ldrec<W256>(a.S2 + j);
    const V3 xij = xi - v3(nb.q.x, nb.q.y, nb.q.z), vij = vi - v3(nb.w.x, nb.w.y, nb.w.z);
    const double rij = norm(xij);
    const double h = a.fc.h;
    const double q = div_c(rij, divc_from(h, a.fc.inv_h, 1));
    const double GK = cubicGW_c(a.fc.C4, q, h);
    const double mj = q2j.x, dj = q2j.y;
    // artificial viscosity (InteractionAlt.cpp:72-94)
    double pi = 0.0; bool pi_on = false;
    if (a.fc.AlphaAvg != 0.0 || a.fc.BetaAvg != 0.0) {
        const double vx = dot(vij, xij);
        if (vx < 0) {
            const double MUij = h * vx / (rij * rij + a.fc.h2_001);
            pi = (a.fc.AlphaCij * MUij + a.fc.BetaAvg * MUij * MUij) / (0.5 * (di + dj)); pi_on = true;
        }
    }
    // momentum tensor, GradientType 0 (InteractionAlt.cpp:139): S holds 1.0/(rho*rho)*Sigma
    double mxx = s0i.x + s0j.x, myy = s0i.y + s0j.y, mzz = s0i.z + s0j.z;
    double mxy = s0i.w + s0j.w, myz = s1i.x + s1j.x, mxz = s1i.y + s1j.y;
    if (pi_on) { mxx += pi; myy += pi; mzz += pi; }
    // tensile instability (InteractionAlt.cpp:103-110), TIn = 4
    if (TI_ON) {
        const double K = cubicW_c(a.fc.C3, q);
        const double r = div_c(K, divc_from(a.fc.W0, a.fc.inv_W0, 1));
        const double fti = pow4_cr(r);
        mxx += (s1i.z + s1j.z) * fti; myy += (s1i.w + s1j.w) * fti; mzz += (s2i.x + t2j.x) * fti;
        mxy += (s2i.y + t2j.y) * fti; myz += (s2i.z + t2j.z) * fti; mxz += (s2i.w + t2j.w) * fti;
    }
    const V3 g = GK * xij;
    const double tx = g.x * mxx + g.y * mxy + g.z * mxz;
    const double ty = g.x * mxy + g.y * myy + g.z * myz;
    const double tz = g.x * mxz + g.y * myz + g.z * mzz;
    ax += mj * tx; ay += mj * ty; az += mj * tz;
}
template <bool TI_ON, int MB, bool W256> __global__ void __launch_bounds__(128, MB) k_accel_lean(PairArgs a)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + a.s0;
    if (s >= a.n) return;
    if (a.orig[s] >= a.nsph) return;
    const double4 qi = a.Q0[s], wi = a.Q1[s];
    const V3 xi = v3(qi.x, qi.y, qi.z), vi = v3(wi.x, wi.y, wi.z);
    const double di = a.Q2[s].y;
    const double4 s0i = a.S0[s], s1i = a.S1[s], s2i = TI_ON ? a.S2[s] : make_double4(0., 0., 0., 0.);
    const int cnt = a.nb_cnt[s], nhi = a.nb_hi[s];
    double ax = 0., ay = 0., az = 0.;
    const uint32_t *row = a.nbr + s;
    NbBuf A, B; uint32_t e = 0;
    A.q = qi; A.w = wi; A.j = 0; B = A;
    if (cnt > 0) { A.j = (int)(row[(int64_t)nb_row(0, nhi, a.K) * a.cap] & 0x7fffffffu); A.q = ldrec<W256>(a.Q0 + A.j); A.w = ldrec<W256>(a.Q1 + A.j);
                   if (cnt > 1) e = row[(int64_t)nb_row(1, nhi, a.K) * a.cap]; }
    for (int n = 0; n < cnt; n += 2) {
        if (n + 1 < cnt) { B.j = (int)(e & 0x7fffffffu); B.q = ldrec<W256>(a.Q0 + B.j); B.w = ldrec<W256>(a.Q1 + B.j); }
        if (n + 2 < cnt) e = row[(int64_t)nb_row(n + 2, nhi, a.K) * a.cap];
        accel_term<TI_ON, W256>(a, A, xi, vi, di, s0i, s1i, s2i, ax, ay, az);
        if (n + 1 >= cnt) break;
        if (n + 2 < cnt) { A.j = (int)(e & 0x7fffffffu); A.q = ldrec<W256>(a.Q0 + A.j); A.w = ldrec<W256>(a.Q1 + A.j); }
        if (n + 3 < cnt) e = row[(int64_t)nb_row(n + 3, nhi, a.K) * a.cap];
        accel_term<TI_ON, W256>(a, B, xi, vi, di, s0i, s1i, s2i, ax, ay, az);
    }
    OWN(F_AX) = ax; OWN(F_AY) = ay; OWN(F_AZ) = az;
}

// ---- CalcPlasticWorkHeat, T += dt*dTdt, CalcConvHeat (ThermalStuct.cpp:14-39, Thermal.cpp:721-732,339-389) ------------
__global__ void k_thermal_pre(PairArgs a)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + a.s0;
    if (s >= a.n) return;
    if (a.orig[s] >= a.nsph) return;
    const double dt = a.dt;
    if (a.pl_heating && OWN(F_DPLS) > 0.0) {
        const double f = 1. / dt;
        const double sxx = OWN(F_SG), syy = OWN(F_SG + 1), szz = OWN(F_SG + 2), sxy = OWN(F_SG + 3), syz = OWN(F_SG + 4), sxz = OWN(F_SG + 5);
        const double dxx = OWN(F_SI) * f, dyy = OWN(F_SI + 1) * f, dzz = OWN(F_SI + 2) * f, dxy = OWN(F_SI + 3) * f, dyz = OWN(F_SI + 4) * f, dxz = OWN(F_SI + 5) * f;
        OWN(F_QPL) = sxx * dxx + 2.0 * sxy * dxy + 2.0 * sxz * dxz + syy * dyy + 2.0 * syz * dyz + szz * dzz;
    }
    if (a.thermal) {
        double4 q2 = a.Q2[s];
        const double T = q2.z + dt * OWN(F_DTDT);
        ((double *)&a.Q2[s])[2] = T;
        const ClassProps cp = a.cls[a.cls_of[s]];
        if (cp.Thermal_BC == 1) {
            const double dens = q2.y;
            const double dS2 = pow(q2.x / dens, 2.0 / 3.0);
            OWN(F_QCONV) = dens * cp.h_conv * dS2 * (cp.T_inf - T) / q2.x;
        }
    }
}

// ---- k_post: CalcTempInc (Thermal.cpp:53-217) + CalcKinEnergyEqn / CalcIntEnergyEqn (Domain.cpp:1752-1829) + the
// density rate (InteractionAlt.cpp:483-640) and TIR (Domain.cpp:1528-1570) of the step that follows. One neighbour
// walk; each sum keeps its own reference order and operand order. Thermal: the m/rho factor is the SECOND particle's
// of the reference pair, and the sums run in list order (round-off level difference to the reference's pair-discovery
// order when several threads race in the reference; identical for Nproc = 1).
template <bool FAST, bool THERM, bool EKIN, bool DENS, bool PF, int MB = 4> __global__ void __launch_bounds__(128, MB) k_post(PairArgs a)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + a.s0;
    if (s >= a.n) return;
    if (a.orig[s] >= a.nsph) { if (THERM) OWN(F_DTDT) = 0.; if (EKIN) OWN(F_DKIN) = 0.; return; }
    const double4 qi = a.Q0[s], wi = a.Q1[s], q2i = a.Q2[s];
    const V3 xi = v3(qi.x, qi.y, qi.z), vi = v3(wi.x, wi.y, wi.z);
    const double hi = FAST ? a.fc.h : qi.w, C4i = FAST ? a.fc.C4 : c4_of(hi), di = q2i.y, Ti = q2i.z, mi = q2i.x, Pi2 = q2i.w, midi = wi.w;
    const int ci = FAST ? 0 : a.cls_of[s];
    struct { double k_T, cp_T, th_exp, q_source, TI; } cpi;
    if (FAST) { cpi.k_T = a.fc.kT; cpi.cp_T = a.fc.cpT; cpi.th_exp = a.fc.th_exp; cpi.q_source = a.fc.q_source; cpi.TI = a.fc.TI;
                if (!a.fc.own_uniform) { const ClassProps cp0 = a.cls[a.cls_of[s]]; cpi.cp_T = cp0.cp_T; cpi.th_exp = cp0.th_exp; cpi.q_source = cp0.q_source; } }
    else { const ClassProps cp0 = a.cls[ci]; cpi.k_T = cp0.k_T; cpi.cp_T = cp0.cp_T; cpi.th_exp = cp0.th_exp; cpi.q_source = cp0.q_source; cpi.TI = cp0.TI; }
    const int cnt = a.nb_cnt[s], nhi = a.nb_hi[s];
    const DivC dh = FAST ? divc_from(a.fc.h, a.fc.inv_h, a.fc.fast_h) : divc_make(hi), dkk = FAST ? divc_from(a.fc.ksum, a.fc.inv_ksum, a.fc.fast_k) : divc_make(cpi.k_T + cpi.k_T);
    const double kk_same = FAST ? a.fc.kk : cpi.k_T * cpi.k_T;
    double temp = 0., dk = 0., dd = 0.;
    const bool gc = !FAST && a.gc;
    double Mg[9];
    if (!FAST) { for (int k = 0; k < 9; ++k) Mg[k] = gc ? OWN(F_GC + k) : 0.0; }
    double4 qn = qi, wn = wi, q2n = q2i; uint32_t e1 = 0, e2 = 0;
    if (PF && cnt > 0) { e1 = a.nbr[(int64_t)nb_row(0, nhi, a.K) * a.cap + s]; const int j0 = (int)(e1 & 0x7fffffffu); qn = ldrec<true>(a.Q0 + j0); wn = ldrec<true>(a.Q1 + j0);
                         if (THERM || EKIN) q2n = ldrec<true>(a.Q2 + j0);
                         if (cnt > 1) e2 = a.nbr[(int64_t)nb_row(1, nhi, a.K) * a.cap + s]; }
    for (int n = 0; n < cnt; ++n) {
        uint32_t e; double4 qj, wj;
        if (PF) { e = e1; qj = qn; wj = wn; } else { e = a.nbr[(int64_t)nb_row(n, nhi, a.K) * a.cap + s]; }
        const int j = (int)(e & 0x7fffffffu); const bool ifirst = (e >> 31) != 0;
        if (!PF) { qj = a.Q0[j]; wj = a.Q1[j]; }
        double4 q2j_pf = make_double4(0., 0., 0., 0.);
        if (THERM || EKIN) { if (PF) q2j_pf = q2n; else q2j_pf = a.Q2[j]; }
        if (PF) {
            if (n + 1 < cnt) { e1 = e2; const int j1 = (int)(e1 & 0x7fffffffu); qn = ldrec<true>(a.Q0 + j1); wn = ldrec<true>(a.Q1 + j1); if (THERM || EKIN) q2n = ldrec<true>(a.Q2 + j1); }
            if (n + 2 < cnt) e2 = a.nbr[(int64_t)nb_row(n + 2, nhi, a.K) * a.cap + s];
        }
        const V3 xj = v3(qj.x, qj.y, qj.z), vj = v3(wj.x, wj.y, wj.z);
        // x_ij with the reference's sign for this pair (P1 - P2); r, q and grad-W are even in it
        const V3 xp = ifirst ? (xi - xj) : (xj - xi);
        const double nr = norm(xp);
        double hh, C4, q;
        if (FAST) { hh = hi; C4 = C4i; q = div_c(nr, dh); }
        else { if (qj.w == hi) { hh = hi; C4 = C4i; } else { hh = (hi + qj.w) / 2.0; C4 = c4_of(hh); } q = nr / hh; }
        double GK;
        if (FAST) GK = cubicGW_c(C4, q, hh); else GK = (a.kt == 0 && a.dim == 3) ? cubicGW_c(C4, q, hh) : gradW(a.dim, a.kt, q, hh);
        const V3 gx = GK * xp;
        if (DENS) {
            // CalcDensInc evaluates with x_i - x_j, v_i - v_j from particle i's side; dot() of two sign-flipped vectors is bit-identical
            const V3 vij = vi - vj;
            const V3 xij = xi - xj;
            if (gc) {
                // locking branch (InteractionAlt.cpp:569-582): dDensity_i += m_j * (rho_i/rho_j) * dot(v_ij, (GK*gradCorrM_i) x_ij), pair roles
                const double4 qd = (THERM || EKIN) ? q2j_pf : a.Q2[j];
                const V3 vp = ifirst ? vij : (vj - vi);
                dd += (1.0 * qd.x) * (di / qd.y) * dot(vp, gc_scaled_mat_vec(Mg, GK, xp));
            } else {
            const double t1 = dot(vij, GK * xij);
            dd += wj.w * t1;
            }
        }
        if (THERM || EKIN) {
            const double4 q2j = q2j_pf;
            if (THERM) {
                double kj, kki, num; bool same = true;
                if (!FAST) { const int cj = a.cls_of[j]; same = (cj == ci); kj = same ? cpi.k_T : a.cls[cj].k_T; } else kj = cpi.k_T;
                // gradKernelCorr (Thermal.cpp:90-111): dot(xij, gradCorrM_i (GK xij)) replaces dot(xij, GK xij)
                const double xgx = gc ? dot(xp, gc_mat_vec(Mg, gx)) : dot(xp, gx);
                if (ifirst) {      // i is P1: temp[i] += m, with m_j/rho_j
                    const double t0 = wj.w * 4. * (same ? kk_same : (cpi.k_T * kj));
                    const double t1 = same ? div_c(t0, dkk) : t0 / (cpi.k_T + kj);
                    temp += t1 * (Ti - q2j.z) * xgx / (nr * nr);
                } else {           // i is P2: temp[i] -= m, m built with P1 = j and m_i/rho_i
                    const double t0 = midi * 4. * (same ? kk_same : (kj * cpi.k_T));
                    const double t1 = same ? div_c(t0, dkk) : t0 / (kj + cpi.k_T);
                    temp -= t1 * (q2j.z - Ti) * xgx / (nr * nr);
                }
                (void)kki; (void)num;
            }
            if (EKIN) {
                const V3 vp = ifirst ? (vi - vj) : (vj - vi);
                if (ifirst) dk += 0.5 * q2j.x * (Pi2 + q2j.w) * dot(vp, gx);
                else        dk -= 0.5 * mi * (q2j.w + Pi2) * dot(vp, gx);
            }
        }
    }
    if (THERM) {
        const double f = 1. / (di * cpi.cp_T);
        double dTdt = f * (temp + OWN(F_QCONV) + cpi.q_source + OWN(F_QPL) * a.pl_work_heat_frac + OWN(F_QCONT));
        if (a.contact && a.cont_heat_fric) dTdt += f * OWN(F_QFRIC);
        OWN(F_DTDT) = dTdt;
        // CalcThermalExpStrainRate (Particle.cpp:953-957): StrainRate += th_exp*dTdt*I
        const double te = cpi.th_exp * dTdt;
        OWN(F_SR) = OWN(F_SR) + te; OWN(F_SR + 1) = OWN(F_SR + 1) + te; OWN(F_SR + 2) = OWN(F_SR + 2) + te;
    }
    if (EKIN) {
        OWN(F_DKIN) = dk;
        OWN(F_DINT) = mi / di * (OWN(F_SH) * OWN(F_SR) + 2.0 * OWN(F_SH + 3) * OWN(F_SR + 3) + 2.0 * OWN(F_SH + 5) * OWN(F_SR + 5) +
                                 OWN(F_SH + 1) * OWN(F_SR + 1) + 2.0 * OWN(F_SH + 4) * OWN(F_SR + 4) + OWN(F_SH + 2) * OWN(F_SR + 2));
    }
    if (DENS) {
        if (!gc) dd *= di;
        OWN(F_DDN) = dd;
    }
}

inline int divc_ok(double b)
{
    unsigned long long bits; std::memcpy(&bits, &b, 8);
    const unsigned ex = (unsigned)((bits >> 52) & 0x7ff);
    return ((bits & 0xFFFFFFFFFFFFFull) != 0xFFFFFFFFFFFFFull) && ex > 200 && ex < 1800;
}
// FAST path: 3-D cubic spline, uniform h, and classes that agree in every constant a PAIR term reads (own-particle properties may differ)
inline bool fast_path(const wf_ctx *c) { return c->prm.kernel_type == 0 && c->prm.dimension == 3 && c->uniform_h && c->pair_uniform && !c->prm.grad_kernel_corr; }

PairArgs make_args(wf_ctx *c)
{
    PairArgs a; std::memset(&a.fc, 0, sizeof(a.fc));
    if (fast_path(c)) {
        const ClassProps &cp = c->h_classes[0]; FastConsts &f = a.fc;
        f.h = c->h_uniform; f.C3 = 1.0 / (f.h * f.h * f.h * WF_PI); f.C4 = 1.0 / (f.h * f.h * f.h * f.h * WF_PI); f.inv_h = 1.0 / f.h; f.fast_h = divc_ok(f.h);
        f.W0 = cubicW_c(f.C3, (cp.TIInitDist + cp.TIInitDist) / (2.0 * f.h)); f.inv_W0 = 1.0 / f.W0; f.fast_W0 = divc_ok(f.W0);
        f.Alpha = cp.Alpha; f.Beta = cp.Beta; f.TI = cp.TI; f.TIn = cp.TIn; f.TId = cp.TIInitDist; f.Cs = cp.Cs; f.RefDensity = cp.RefDensity; f.PresEq = cp.PresEq;
        f.kT = cp.k_T; f.kk = cp.k_T * cp.k_T; f.ksum = cp.k_T + cp.k_T; f.inv_ksum = 1.0 / f.ksum; f.fast_k = divc_ok(f.ksum);
        f.cpT = cp.cp_T; f.th_exp = cp.th_exp; f.q_source = cp.q_source;
        f.AlphaAvg = (cp.Alpha + cp.Alpha) / 2.0; f.BetaAvg = (cp.Beta + cp.Beta) / 2.0;
        f.AlphaCij = f.AlphaAvg * (0.5 * (cp.Cs + cp.Cs)); f.h2_001 = 0.01 * f.h * f.h;
        f.own_uniform = c->single_class ? 1 : 0;
        f.mass = c->mass_min; f.uniform_mass = (c->uniform_mass && !(c->prm.tuning & 8192 * 4)) ? 1 : 0;     // tuning bit 15: keep the Q2 gather (A/B)
    }
    a.n = c->n_loc; a.s0 = 0; a.cap = c->cap; a.nsph = (int)c->nsph; a.K = c->nbr_k; a.dim = c->prm.dimension; a.kt = c->prm.kernel_type; a.gt = c->prm.gradient_type;
    a.nbr = c->nbr; a.nb_cnt = c->nb_cnt; a.nb_hi = c->nb_hi; a.orig = c->Ip(I_ORIG); a.cls_of = c->Ip(I_CLS); a.is_free = c->any_fixed ? c->Ip(I_FREE) : nullptr;
    a.Q0 = c->Q0[c->cur]; a.Q1 = c->Q1[c->cur]; a.Q2 = c->Q2[c->cur]; a.S0 = c->S0[c->cur]; a.S1 = c->S1[c->cur]; a.S2 = c->S2[c->cur];
    a.F = c->F[c->cur]; a.cls = c->d_classes; a.mats = c->d_mats; a.dt = c->dt; a.any_ti = c->any_ti ? 1 : 0;
    a.thermal = c->prm.thermal_solver; a.pl_heating = c->prm.pl_heating; a.contact = c->prm.contact; a.cont_heat_fric = c->prm.cont_heat_fric;
    a.pl_work_heat_frac = c->prm.pl_work_heat_frac; a.gc = c->prm.grad_kernel_corr ? 1 : 0;
    return a;
}

// lean momentum kernel: uniform class with PresEq 0 (constant sound speed), TIn = 4, Markstein-safe h and W0, 3-D
inline bool lean_ok(const wf_ctx *c, const PairArgs &a)
{
    const FastConsts &f = a.fc;
    // fixed particles keep an empty list and are in nobody's (wf_neighbour.cu): the pair walk needs no special case for them
    return a.dim == 3 && f.PresEq == 0 && f.fast_h && (!c->any_ti || (f.TI > 0.0 && f.TIn == 4.0 && f.fast_W0));
}
// contracted-arithmetic variants (wf_pair_fm.cu): the default on the lean FAST path; wf_params::tuning bit 10 selects the exact kernels
// (bit-identical to the reference's arithmetic) there too
inline bool use_fm(const wf_ctx *c, const PairArgs &a) { return !(c->prm.tuning & 1024) && fast_path(c) && lean_ok(c, a) && a.gt == 0; }
inline bool use_pf(const wf_ctx *c) { return (c->prm.tuning & 1) == 0; }     // bit 0 of wf_params::tuning switches the software pipeline OFF

} // namespace

namespace wf {

#define WF_PAIR_PRECHECK(c) if (!(c)->nbr) return fail((c), WF_ERR_ARG, "pair kernel called before the first neighbour build")
// launch KERNEL<targs..., PF> with PF = software pipeline on/off; grid of one thread per slot, 128 threads per block
#define WF_LAUNCH_PF(KERNEL, ...) do { if (pf) KERNEL<__VA_ARGS__, true><<<g, 128, 0, c->stream>>>(a); else KERNEL<__VA_ARGS__, false><<<g, 128, 0, c->stream>>>(a); } while (0)

// CalcGradCorrMatrix (Domain.cpp:1832-1940) on the current neighbour lists
int pair_gradcorr(wf_ctx *c)
{
    WF_PAIR_PRECHECK(c);
    if (!c->prm.grad_kernel_corr) return 0;
    PairArgs a = make_args(c);
    ProfScope ps(c, PK_NB_LIST);
    k_gradcorr<<<grid_for(c->n_loc, 128), 128, 0, c->stream>>>(a); WF_LAUNCH_CHECK(c); c->launches++;
    if (c->any_fixed) WF_CHECK(nb_gradcorr_fsm(c));           // FSM pairs (one particle fixed): a cell walk of its own, after the SM sums
    k_gradcorr_inv<<<grid_for(c->n_loc, 128), 128, 0, c->stream>>>(a); WF_LAUNCH_CHECK(c); c->launches++;
    c->dens_valid = false;
    return 0;
}
// StartAcceleration's TIR, stand-alone (KickDrift / LeapFrog step orders)
int pair_tir(wf_ctx *c)
{
    if (!c->any_ti) return 0;
    PairArgs a = make_args(c);
    ProfScope ps(c, PK_BEGIN);
    k_tir<false><<<grid_for(c->n_loc, 128), 128, 0, c->stream>>>(a); WF_LAUNCH_CHECK(c); c->launches++;
    return 0;
}
int pair_temp_inc_pp(wf_ctx *c)
{
    WF_PAIR_PRECHECK(c);
    PairArgs a = make_args(c);
    ProfScope ps(c, PK_THERMAL);
    k_temp_inc_pp<<<grid_for(c->n_loc, 128), 128, 0, c->stream>>>(a); WF_LAUNCH_CHECK(c); c->launches++;
    return 0;
}
// CheckMinTSVel: this rank's min of CFL*dmin/(Cs+|v|) (blocking read-back)
int pair_min_ts(wf_ctx *c, double *test_min)
{
    WF_PAIR_PRECHECK(c);
    PairArgs a = make_args(c);
    const int g = grid_for(c->n_loc, 128);
    WF_CHECK(ensure_stage(c, (size_t)(g + 8) * sizeof(double)));
    double *partial = (double *)c->stage;
    {   ProfScope ps(c, PK_REDUCE);
        k_min_ts<<<g, 128, 0, c->stream>>>(a, c->prm.cfl, c->prm.auto_ts_acc ? (c->prm.sqrt_h_a != 0.0 ? c->prm.sqrt_h_a : 0.0025) : 0.0, partial + 8); WF_LAUNCH_CHECK(c);
        k_min_final<<<1, 256, 0, c->stream>>>(partial + 8, g, partial); WF_LAUNCH_CHECK(c); c->launches += 2; }
    WF_CUDA(c, cudaMemcpyAsync(test_min, partial, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    WF_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
// CalcDensInc + DensReduction + rho update. If the previous step's k_post already holds this step's density rate and TIR
// (nothing moved in between) only the per-particle prologue runs. with_tir = false (KickDrift / LeapFrog: StartAcceleration's tensor was
// evaluated by pair_tir at the top of the step and sits in F_TN and in the gathered arrays already): the tensor is not evaluated again --
// the same values without ghosts, and with ghosts MoveGhost has changed Sigma and the density of the mirror rows since.
int pair_density(wf_ctx *c, bool with_tir)
{
    WF_PAIR_PRECHECK(c);
    PairArgs a = make_args(c); const int g = grid_for(c->n_loc, 128); const bool pf = use_pf(c);
    ProfScope ps(c, PK_DENSITY);
    if (!c->dens_valid) {
        if (use_fm(c, a) && launch_post_fm(c, a, g, false, false, true) == 0) {}
        else if (fast_path(c)) WF_LAUNCH_PF(k_post, true, false, false, true); else WF_LAUNCH_PF(k_post, false, false, false, true);
        WF_LAUNCH_CHECK(c); c->launches++;
        if (c->any_ti && with_tir) { k_tir<true><<<g, 128, 0, c->stream>>>(a); WF_LAUNCH_CHECK(c); c->launches++; }
    }
    k_prologue<<<grid_for(c->n_loc, 256), 256, 0, c->stream>>>(a); WF_LAUNCH_CHECK(c); c->launches++;
    c->dens_valid = false;
    return 0;
}
int pair_rate_stress(wf_ctx *c, int64_t lo, int64_t hi)
{
    WF_PAIR_PRECHECK(c);
    if (hi < 0) hi = c->n_loc;
    if (hi <= lo) return 0;
    PairArgs a = make_args(c); a.s0 = lo; a.n = hi; const int g = grid_for(hi - lo, 128); const bool pf = use_pf(c);
    ProfScope ps(c, PK_RATE_STRESS);
    if (use_fm(c, a)) { WF_CHECK(launch_rate_fm(c, a, g)); WF_LAUNCH_CHECK(c); c->launches++; return 0; }
    if (fast_path(c)) WF_LAUNCH_PF(k_rate_stress, true); else WF_LAUNCH_PF(k_rate_stress, false);
    WF_LAUNCH_CHECK(c); c->launches++;
    return 0;
}
int pair_accel(wf_ctx *c, int64_t lo, int64_t hi)
{
    WF_PAIR_PRECHECK(c);
    if (hi < 0) hi = c->n_loc;
    if (hi <= lo) return 0;
    PairArgs a = make_args(c); a.s0 = lo; a.n = hi; const int g = grid_for(hi - lo, 128); const bool pf = use_pf(c);
    ProfScope ps(c, PK_ACCEL);
    const bool fast = fast_path(c);
    const bool mb3 = (c->prm.tuning & 2) == 0;        // 3 resident blocks per SM (168 registers) by default; bit 1 of wf_params::tuning selects 4 (128 registers, spills)
    switch (a.gt) {
    case 0: if (fast && lean_ok(c, a) && !(c->prm.tuning & 8)) {        // bit 3 of wf_params::tuning: the general FAST kernel instead of the lean one
                const bool w256 = (c->prm.tuning & 32) == 0;           // bit 5: 2 x LDG.128 per gathered record instead of one LDG.256
                if (use_fm(c, a)) { WF_CHECK(launch_accel_fm(c, a, g, c->any_ti)); WF_LAUNCH_CHECK(c); c->launches++; return 0; }   // bit 10: contracted arithmetic (wf_pair_fm.cu)
                if (c->any_ti) {
                    if (c->prm.tuning & 16) { if (w256) k_accel_lean<true, 4, true><<<g, 128, 0, c->stream>>>(a); else k_accel_lean<true, 4, false><<<g, 128, 0, c->stream>>>(a); }
                    else { if (w256) k_accel_lean<true, 3, true><<<g, 128, 0, c->stream>>>(a); else k_accel_lean<true, 3, false><<<g, 128, 0, c->stream>>>(a); }
                }
                else { if (w256) k_accel_lean<false, 4, true><<<g, 128, 0, c->stream>>>(a); else k_accel_lean<false, 4, false><<<g, 128, 0, c->stream>>>(a); }
            }
            else if (fast) { if (c->prm.tuning & 4) WF_LAUNCH_PF(k_accel, true, 0, 2); else if (mb3) WF_LAUNCH_PF(k_accel, true, 0, 3); else WF_LAUNCH_PF(k_accel, true, 0, 4); }
            else WF_LAUNCH_PF(k_accel, false, 0, 4); break;
    case 1: if (fast) WF_LAUNCH_PF(k_accel, true, 1, 4); else WF_LAUNCH_PF(k_accel, false, 1, 4); break;
    default: if (fast) WF_LAUNCH_PF(k_accel, true, 2, 4); else WF_LAUNCH_PF(k_accel, false, 2, 4); break;
    }
    WF_LAUNCH_CHECK(c); c->launches++;
    return 0;
}
// CalcPlasticWorkHeat + ThermalCalcs, then CalcKinEnergyEqn + CalcIntEnergyEqn, then next step's density rate: one walk
// the per-particle prologue of the post pass: CalcPlasticWorkHeat, T += dt*dTdt, CalcConvHeat
int pair_post_pre(wf_ctx *c)
{
    WF_PAIR_PRECHECK(c);
    PairArgs a = make_args(c);
    ProfScope ps(c, PK_BEGIN);
    if (c->prm.thermal_solver != 0 || c->prm.pl_heating) { k_thermal_pre<<<grid_for(c->n_loc, 256), 256, 0, c->stream>>>(a); WF_LAUNCH_CHECK(c); c->launches++; }
    return 0;
}
// the pair walk of the post pass over slots [lo, hi)
int pair_post_range(wf_ctx *c, bool skip_dens, int64_t lo, int64_t hi)
{
    WF_PAIR_PRECHECK(c);
    if (hi < 0) hi = c->n_loc;
    if (hi <= lo) return 0;
    PairArgs a = make_args(c); a.s0 = lo; a.n = hi; const int g = grid_for(hi - lo, 128); const bool pf = use_pf(c);
    ProfScope ps(c, PK_THERMAL);
    const bool th = c->prm.thermal_solver != 0;
    const bool fast = fast_path(c);
    if (use_fm(c, a) && launch_post_fm(c, a, g, th, true, !skip_dens) == 0) {
        WF_LAUNCH_CHECK(c); c->launches++;
        if (!skip_dens && c->any_ti) { k_tir<true><<<g, 128, 0, c->stream>>>(a); WF_LAUNCH_CHECK(c); c->launches++; }
        return 0;
    }
    if (skip_dens) {
        // the coming step rebuilds its neighbour lists first (known from the cadence): a density rate and TIR summed over the
        // current lists would be discarded, so only the conduction and energy sums of this step are evaluated
        if (th) { if (fast) WF_LAUNCH_PF(k_post, true, true, true, false); else WF_LAUNCH_PF(k_post, false, true, true, false); }
        else    { if (fast) WF_LAUNCH_PF(k_post, true, false, true, false); else WF_LAUNCH_PF(k_post, false, false, true, false); }
        WF_LAUNCH_CHECK(c); c->launches++;
        return 0;
    }
    if (th && fast && pf && (c->prm.tuning & (64 | 128))) {      // bits 6 / 7 of wf_params::tuning: 5 / 3 resident blocks per SM for the fused pass (A/B measurements)
        if (c->prm.tuning & 64) k_post<true, true, true, true, true, 5><<<g, 128, 0, c->stream>>>(a); else k_post<true, true, true, true, true, 3><<<g, 128, 0, c->stream>>>(a);
    }
    else if (th) { if (fast) WF_LAUNCH_PF(k_post, true, true, true, true); else WF_LAUNCH_PF(k_post, false, true, true, true); }
    else    { if (fast) WF_LAUNCH_PF(k_post, true, false, true, true); else WF_LAUNCH_PF(k_post, false, false, true, true); }
    WF_LAUNCH_CHECK(c); c->launches++;
    if (c->any_ti) { k_tir<true><<<g, 128, 0, c->stream>>>(a); WF_LAUNCH_CHECK(c); c->launches++; }      // the next step's TIR, same slot range
    return 0;
}
// CalcPlasticWorkHeat + ThermalCalcs, then CalcKinEnergyEqn + CalcIntEnergyEqn, then next step's density rate: one walk.
// Multi-GPU: the halo refresh (x, v, T after the move / temperature update) travels on the communication stream while the
// interior slots -- whose neighbour lists cannot reach a halo slot -- are evaluated; the two boundary layers follow.
int pair_post(wf_ctx *c, bool skip_dens)
{
    WF_CHECK(pair_post_pre(c));
    int64_t i0 = 0, i1 = c->n_loc;
    WF_CHECK(comm_halo_begin(c, HALO_Q0 | HALO_Q1 | HALO_Q2, &i0, &i1));
    WF_CHECK(pair_post_range(c, skip_dens, i0, i1));
    WF_CHECK(comm_halo_end(c));
    WF_CHECK(pair_post_range(c, skip_dens, 0, i0));
    WF_CHECK(pair_post_range(c, skip_dens, i1, c->n_loc));
    c->dens_valid = !skip_dens;
    return 0;
}
// stand-alone phases of the C ABI (wf_thermal, wf_energy): same kernel, one family of sums each
int pair_thermal(wf_ctx *c)
{
    WF_PAIR_PRECHECK(c);
    if (!c->prm.thermal_solver && !c->prm.pl_heating) return 0;
    PairArgs a = make_args(c); const bool pf = use_pf(c);
    ProfScope ps(c, PK_THERMAL);
    k_thermal_pre<<<grid_for(c->n_loc, 256), 256, 0, c->stream>>>(a); WF_LAUNCH_CHECK(c); c->launches++;
    WF_CHECK(comm_halo_refresh(c, HALO_Q0 | HALO_Q1 | HALO_Q2));
    if (c->prm.thermal_solver) {
        const int g = grid_for(c->n_loc, 128);
        if (fast_path(c)) WF_LAUNCH_PF(k_post, true, true, false, false); else WF_LAUNCH_PF(k_post, false, true, false, false);
        WF_LAUNCH_CHECK(c); c->launches++;
    }
    return 0;
}
int pair_energy(wf_ctx *c)
{
    WF_PAIR_PRECHECK(c);
    PairArgs a = make_args(c); const int g = grid_for(c->n_loc, 128); const bool pf = use_pf(c);
    WF_CHECK(comm_halo_refresh(c, HALO_Q0 | HALO_Q1 | HALO_Q2));
    ProfScope ps(c, PK_ENERGY);
    if (fast_path(c)) WF_LAUNCH_PF(k_post, true, false, true, false); else WF_LAUNCH_PF(k_post, false, false, true, false);
    WF_LAUNCH_CHECK(c); c->launches++;
    return 0;
}

} // namespace wf
