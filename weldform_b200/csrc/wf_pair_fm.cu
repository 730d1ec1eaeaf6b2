// Contracted-arithmetic ("fm") variants of the lean pair kernels. This translation unit is compiled WITH FMA contraction (-fmad=true);
// everything else in the library is compiled -fmad=false (bit-identity with the reference's x86-64 build).
//
// Same physics, same neighbour lists and the same summation order over neighbours as the exact kernels of wf_pair.cu; what changes is the
// arithmetic INSIDE a pair term: fused multiply-adds, one reciprocal square root per pair instead of sqrt + divisions (the cubic-spline
// gradient needs only 1/r: (-3q + 9/4 q^2)/(q h) = (-3 + 9/4 q)/h and -(3/4)(2-q)^2/(q h) with 1/q = h/r), branch-free selects instead of
// the piecewise kernel / viscosity branches, and Newton-refined rcp / rsqrt approximations without slow-path guards. A pair term differs
// from the exact one by a few ulp (<= 1e-15 relative); BASELINE.json's bars are 1e-12 after one step and 1e-6 after 1000, and the parity
// tests run this path against the oracle at those bars, element-wise. The neighbour search (CheckRadius) is NOT here: the pair set stays
// bit-exact.
//
// Why: ncu on the exact kernel (profiles/r2_v0_ncu_pair_kernels.txt) shows 217 SASS instructions per neighbour, 114 of them FP64, in seven
// reconvergence regions (the guards of sqrt and of every division, the piecewise polynomial, the viscosity branch): the FP64 dependency
// chains of ONE neighbour cannot overlap those of the next, and 12 resident warps per SM cannot hide them (stall `wait` 3.4 warps per issue).
// One straight-line block per pair lets the two unrolled neighbours interleave.
#include "wf_pair_args.cuh"
#include "wf_constitutive.cuh"

using namespace wf;

namespace {

#define OWN(f) a.F[(int64_t)(f) * a.cap + s]

__device__ __forceinline__ double4 ld256(const double4 *p)
{
    double4 r;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ int nb_row(int n, int nhi, int K) { return n < nhi ? n : (K - 1 - (n - nhi)); }

// 1/x and 1/sqrt(x) for positive normal x: hardware approximation (MUFU.RCP64H / MUFU.RSQ64H, ~2^-22) + two Newton steps -> <= 2 ulp
__device__ __forceinline__ double rcp_fm(double x)
{
    double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0); y = fma(y, e, y);
    e = fma(-x, y, 1.0); y = fma(y, e, y);
    return y;
}
__device__ __forceinline__ double rsqrt_fm(double x)
{
    double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-(x * y), y, 1.0); y = fma(0.5 * y, e, y);
    e = fma(-(x * y), y, 1.0); y = fma(0.5 * y, e, y);
    return y;
}

struct NbBuf { double4 q, w; int j; };

// one neighbour of the momentum sum (InteractionAlt.cpp:12-235 for GradientType 0, 3-D cubic spline, uniform h and class, TIn = 4, PresEq 0)
// UM: every particle has the mass fc.mass -- m_j is a constant and rho_j = m / (m/rho)_j comes from Q1.w, which is gathered anyway: the Q2
// record is not read at all (five gathered sectors per neighbour instead of six; the kernel is bound by L1 wavefronts, profiles/README.md)
template <bool TI_ON, bool UM> __device__ __forceinline__ void accel_term_fm(const PairArgs &a, const NbBuf &nb, const V3 &xi, const V3 &vi, double di,
                                                                     const double4 &s0i, const double4 &s1i, const double4 &s2i, double c4h, double rW,
                                                                     double &ax, double &ay, double &az)
{
    const int j = nb.j;
    double4 q2j = make_double4(0., 0., 0., 0.);
    if (!UM) q2j = ld256(a.Q2 + j);
    const double4 s0j = ld256(a.S0 + j), s1j = ld256(a.S1 + j);
    double4 t2j = make_double4(0., 0., 0., 0.);
    if (TI_ON) t2j = ld256(a.S2 + j);
    const double xx = xi.x - nb.q.x, xy = xi.y - nb.q.y, xz = xi.z - nb.q.z;
    const double ux = vi.x - nb.w.x, uy = vi.y - nb.w.y, uz = vi.z - nb.w.z;
    const double r2 = fmax(fma(xx, xx, fma(xy, xy, xz * xz)), 1e-300);      // coincident particles: x_ij = 0 makes the term vanish, keep 1/r finite
    const double vx = fma(ux, xx, fma(uy, xy, uz * xz));
    const double inv_r = rsqrt_fm(r2);
    const double h = a.fc.h;
    const double q = (r2 * inv_r) * a.fc.inv_h;
    // gradient of the cubic spline over r (Functions.cpp:74-90): GK = C4/h * gk(q)
    const double t = 2.0 - q;
    const double gk1 = fma(2.25, q, -3.0);
    const double gk2 = (-0.75 * t) * (t * (h * inv_r));
    const double gk = (q < 1.0) ? gk1 : ((q < 2.0) ? gk2 : 0.0);
    const double mj = UM ? a.fc.mass : q2j.x;
    // artificial viscosity (InteractionAlt.cpp:72-94): one reciprocal serves MUij = h vx / (r^2 + 0.01 h^2) and the division by the mean density
    double pi = 0.0;
    if (a.fc.AlphaAvg != 0.0 || a.fc.BetaAvg != 0.0) {          // uniform over the launch
        const double D = r2 + a.fc.h2_001;
        double p;
        if (UM) {
            // rho_j = m / r_j with r_j = (m/rho)_j: rho_i + rho_j = (rho_i r_j + m) / r_j, so 1 / (D (rho_i r_j + m)) gives both quotients
            const double rj = nb.w.w, t2 = fma(di, rj, a.fc.mass);
            const double ip = rcp_fm(D * t2);
            const double MU = (h * vx) * (ip * t2);
            p = (MU * fma(a.fc.BetaAvg, MU, a.fc.AlphaCij)) * ((2.0 * rj) * (ip * D));
        } else {
            const double S = di + q2j.y;
            const double ip = rcp_fm(D * S);
            const double MU = (h * vx) * (ip * S);
            p = (MU * fma(a.fc.BetaAvg, MU, a.fc.AlphaCij)) * ((2.0 * ip) * D);
        }
        pi = (vx < 0.0) ? p : 0.0;
    }
    // momentum tensor, GradientType 0 (InteractionAlt.cpp:139): S holds Sigma / rho^2
    double mxx = (s0i.x + s0j.x) + pi, myy = (s0i.y + s0j.y) + pi, mzz = (s0i.z + s0j.z) + pi;
    double mxy = s0i.w + s0j.w, myz = s1i.x + s1j.x, mxz = s1i.y + s1j.y;
    if (TI_ON) {
        // tensile instability (InteractionAlt.cpp:103-110), TIn = 4: (W(q) / W(dx0/h))^4, the kernel constant cancels
        const double q2 = q * q;
        const double w1 = fma(q2, fma(0.75, q, -1.5), 1.0);
        const double w2 = 0.25 * t * (t * t);
        const double R = ((q < 1.0) ? w1 : ((q < 2.0) ? w2 : 0.0)) * rW;
        const double R2 = R * R, fti = R2 * R2;
        mxx = fma(s1i.z + s1j.z, fti, mxx); myy = fma(s1i.w + s1j.w, fti, myy); mzz = fma(s2i.x + t2j.x, fti, mzz);
        mxy = fma(s2i.y + t2j.y, fti, mxy); myz = fma(s2i.z + t2j.z, fti, myz); mxz = fma(s2i.w + t2j.w, fti, mxz);
    }
    // a_i += m_j (GK x_ij)^T M
    const double gm = (c4h * gk) * mj;
    const double gx = gm * xx, gy = gm * xy, gz = gm * xz;
    ax = fma(gx, mxx, fma(gy, mxy, fma(gz, mxz, ax)));
    ay = fma(gx, mxy, fma(gy, myy, fma(gz, myz, ay)));
    az = fma(gx, mxz, fma(gy, myz, fma(gz, mzz, az)));
}

template <bool TI_ON, int MB, bool UM> __global__ void __launch_bounds__(128, MB) k_accel_fm(PairArgs a)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + a.s0;
    if (s >= a.n) return;
    if (a.orig[s] >= a.nsph) return;
    const double4 qi = a.Q0[s], wi = a.Q1[s];
    const V3 xi = v3(qi.x, qi.y, qi.z), vi = v3(wi.x, wi.y, wi.z);
    const double di = a.Q2[s].y;
    const double4 s0i = a.S0[s], s1i = a.S1[s], s2i = TI_ON ? a.S2[s] : make_double4(0., 0., 0., 0.);
    const int cnt = a.nb_cnt[s], nhi = a.nb_hi[s];
    const double c4h = a.fc.C4 * a.fc.inv_h, rW = a.fc.C3 * a.fc.inv_W0;
    double ax = 0., ay = 0., az = 0.;
    const uint32_t *row = a.nbr + s;
    NbBuf A, B; uint32_t e = 0;
    A.q = qi; A.w = wi; A.j = 0; B = A;
    if (cnt > 0) { A.j = (int)(row[(int64_t)nb_row(0, nhi, a.K) * a.cap] & 0x7fffffffu); A.q = ld256(a.Q0 + A.j); A.w = ld256(a.Q1 + A.j);
                   if (cnt > 1) e = row[(int64_t)nb_row(1, nhi, a.K) * a.cap]; }
    for (int n = 0; n < cnt; n += 2) {
        if (n + 1 < cnt) { B.j = (int)(e & 0x7fffffffu); B.q = ld256(a.Q0 + B.j); B.w = ld256(a.Q1 + B.j); }
        if (n + 2 < cnt) e = row[(int64_t)nb_row(n + 2, nhi, a.K) * a.cap];
        accel_term_fm<TI_ON, UM>(a, A, xi, vi, di, s0i, s1i, s2i, c4h, rW, ax, ay, az);
        if (n + 1 >= cnt) break;
        if (n + 2 < cnt) { A.j = (int)(e & 0x7fffffffu); A.q = ld256(a.Q0 + A.j); A.w = ld256(a.Q1 + A.j); }
        if (n + 3 < cnt) e = row[(int64_t)nb_row(n + 3, nhi, a.K) * a.cap];
        accel_term_fm<TI_ON, UM>(a, B, xi, vi, di, s0i, s1i, s2i, c4h, rW, ax, ay, az);
    }
    OWN(F_AX) = ax; OWN(F_AY) = ay; OWN(F_AZ) = az;
}


// r^2 -> (q, gradient factor gk with GK = C4/h * gk) of the 3-D cubic spline; one rsqrt, no division, no branch (Functions.cpp:74-90)
__device__ __forceinline__ void spline_grad_fm(double r2, double h, double inv_h, double &q, double &gk, double &tq)
{
    const double inv_r = rsqrt_fm(r2);
    q = (r2 * inv_r) * inv_h;
    tq = 2.0 - q;
    const double gk1 = fma(2.25, q, -3.0);
    const double gk2 = (-0.75 * tq) * (tq * (h * inv_r));
    gk = (q < 1.0) ? gk1 : ((q < 2.0) ? gk2 : 0.0);
}

// ---- CalcRateTensors + RateTensorsReduction (InteractionAlt.cpp:263-479), FAST path, + the shared constitutive update ----------------
__global__ void __launch_bounds__(128, 5) k_rate_fm(PairArgs a)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + a.s0;
    if (s >= a.n) return;
    if (a.orig[s] >= a.nsph) return;
    const double4 qi = a.Q0[s], wi = a.Q1[s];
    const int cnt = a.nb_cnt[s], nhi = a.nb_hi[s];
    const double h = a.fc.h, inv_h = a.fc.inv_h, c4h = -0.5 * (a.fc.C4 * a.fc.inv_h);
    double sxx = 0., syy = 0., szz = 0., sxy = 0., syz = 0., sxz = 0., r01 = 0., r02 = 0., r12 = 0.;
    const uint32_t *row = a.nbr + s;
    double4 qn = qi, wn = wi; uint32_t e2 = 0;
    if (cnt > 0) { const int j0 = (int)(row[(int64_t)nb_row(0, nhi, a.K) * a.cap] & 0x7fffffffu); qn = ld256(a.Q0 + j0); wn = ld256(a.Q1 + j0);
                   if (cnt > 1) e2 = row[(int64_t)nb_row(1, nhi, a.K) * a.cap]; }
#pragma unroll 2
    for (int n = 0; n < cnt; ++n) {
        const double4 qj = qn, wj = wn;
        if (n + 1 < cnt) { const int j1 = (int)(e2 & 0x7fffffffu); qn = ld256(a.Q0 + j1); wn = ld256(a.Q1 + j1); }
        if (n + 2 < cnt) e2 = row[(int64_t)nb_row(n + 2, nhi, a.K) * a.cap];
        const double xx = qi.x - qj.x, xy = qi.y - qj.y, xz = qi.z - qj.z;
        const double ux = wi.x - wj.x, uy = wi.y - wj.y, uz = wi.z - wj.z;
        const double r2 = fmax(fma(xx, xx, fma(xy, xy, xz * xz)), 1e-300);
        double q, gk, tq; spline_grad_fm(r2, h, inv_h, q, gk, tq);
        const double f = (c4h * gk) * wj.w;                    // -0.5 GK m_j/rho_j (InteractionAlt.cpp:338-361,455-457)
        const double axy = ux * xy, ayx = uy * xx, axz = ux * xz, azx = uz * xx, ayz = uy * xz, azy = uz * xy;
        sxx = fma(2.0 * (ux * xx), f, sxx); syy = fma(2.0 * (uy * xy), f, syy); szz = fma(2.0 * (uz * xz), f, szz);
        sxy = fma(axy + ayx, f, sxy); sxz = fma(axz + azx, f, sxz); syz = fma(ayz + azy, f, syz);
        r01 = fma(axy - ayx, f, r01); r02 = fma(axz - azx, f, r02); r12 = fma(ayz - azy, f, r12);
    }
    stress_update(a, s, sxx, syy, szz, sxy, syz, sxz, r01, r02, r12);
}

// ---- k_post, FAST path: CalcTempInc (Thermal.cpp:53-217) + CalcKinEnergyEqn (Domain.cpp:1752-1808) + next step's CalcDensInc ----------
// (InteractionAlt.cpp:483-640) in one walk. The conduction term's dot(x, GK x) / |x|^2 is GK; the pair roles (first / second of the
// reference pair, bit 31 of the list entry) only select whose m/rho and whose mass multiplies the term.
template <bool THERM, bool EKIN, bool DENS> __global__ void __launch_bounds__(128, 4) k_post_fm(PairArgs a)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + a.s0;
    if (s >= a.n) return;
    if (a.orig[s] >= a.nsph) { if (THERM) OWN(F_DTDT) = 0.; if (EKIN) OWN(F_DKIN) = 0.; return; }
    const double4 qi = a.Q0[s], wi = a.Q1[s], q2i = a.Q2[s];
    const double di = q2i.y, Ti = q2i.z, mi = q2i.x, Pi2 = q2i.w, midi = wi.w;
    const int cnt = a.nb_cnt[s], nhi = a.nb_hi[s];
    const double h = a.fc.h, inv_h = a.fc.inv_h, c4h = a.fc.C4 * a.fc.inv_h;
    const double ck = 4.0 * a.fc.kk * a.fc.inv_ksum;          // 4 k_i k_j / (k_i + k_j), one class
    double temp = 0., dk = 0., dd = 0.;
    const uint32_t *row = a.nbr + s;
    double4 qn = qi, wn = wi, q2n = q2i; uint32_t e1 = 0, e2 = 0;
    if (cnt > 0) { e1 = row[(int64_t)nb_row(0, nhi, a.K) * a.cap]; const int j0 = (int)(e1 & 0x7fffffffu); qn = ld256(a.Q0 + j0); wn = ld256(a.Q1 + j0);
                   if (THERM || EKIN) q2n = ld256(a.Q2 + j0);
                   if (cnt > 1) e2 = row[(int64_t)nb_row(1, nhi, a.K) * a.cap]; }
#pragma unroll 2
    for (int n = 0; n < cnt; ++n) {
        const bool ifirst = (e1 >> 31) != 0;
        const double4 qj = qn, wj = wn, q2j = q2n;
        if (n + 1 < cnt) { e1 = e2; const int j1 = (int)(e1 & 0x7fffffffu); qn = ld256(a.Q0 + j1); wn = ld256(a.Q1 + j1); if (THERM || EKIN) q2n = ld256(a.Q2 + j1); }
        if (n + 2 < cnt) e2 = row[(int64_t)nb_row(n + 2, nhi, a.K) * a.cap];
        const double xx = qi.x - qj.x, xy = qi.y - qj.y, xz = qi.z - qj.z;
        const double r2 = fmax(fma(xx, xx, fma(xy, xy, xz * xz)), 1e-300);
        double q, gk, tq; spline_grad_fm(r2, h, inv_h, q, gk, tq);
        const double GK = c4h * gk;
        double gvx = 0.0;
        if (DENS || EKIN) {
            const double ux = wi.x - wj.x, uy = wi.y - wj.y, uz = wi.z - wj.z;
            gvx = GK * fma(ux, xx, fma(uy, xy, uz * xz));       // dot(v_ij, GK x_ij)
        }
        if (DENS) dd = fma(wj.w, gvx, dd);
        if (THERM) temp = fma(((ifirst ? wj.w : midi) * ck) * (Ti - q2j.z), GK, temp);
        if (EKIN) dk = fma((ifirst ? (0.5 * q2j.x) : (-0.5 * mi)) * (Pi2 + q2j.w), gvx, dk);
    }
    if (THERM) {
        double cpT = a.fc.cpT, th_exp = a.fc.th_exp, q_source = a.fc.q_source;
        if (!a.fc.own_uniform) { const ClassProps cp0 = a.cls[a.cls_of[s]]; cpT = cp0.cp_T; th_exp = cp0.th_exp; q_source = cp0.q_source; }
        const double f = 1. / (di * cpT);
        double dTdt = f * (temp + OWN(F_QCONV) + q_source + OWN(F_QPL) * a.pl_work_heat_frac + OWN(F_QCONT));
        if (a.contact && a.cont_heat_fric) dTdt += f * OWN(F_QFRIC);
        OWN(F_DTDT) = dTdt;
        const double te = th_exp * dTdt;                       // CalcThermalExpStrainRate (Particle.cpp:953-957)
        OWN(F_SR) = OWN(F_SR) + te; OWN(F_SR + 1) = OWN(F_SR + 1) + te; OWN(F_SR + 2) = OWN(F_SR + 2) + te;
    }
    if (EKIN) {
        OWN(F_DKIN) = dk;
        OWN(F_DINT) = mi / di * (OWN(F_SH) * OWN(F_SR) + 2.0 * OWN(F_SH + 3) * OWN(F_SR + 3) + 2.0 * OWN(F_SH + 5) * OWN(F_SR + 5) +
                                 OWN(F_SH + 1) * OWN(F_SR + 1) + 2.0 * OWN(F_SH + 4) * OWN(F_SR + 4) + OWN(F_SH + 2) * OWN(F_SR + 2));
    }
    if (DENS) OWN(F_DDN) = dd * di;
}

} // namespace

namespace wf {

int launch_accel_fm(wf_ctx *c, const PairArgs &a, int grid, bool ti_on)
{
    const bool mb4 = (c->prm.tuning & 16) != 0;
    if (a.fc.uniform_mass) {
        if (ti_on) { if (mb4) k_accel_fm<true, 4, true><<<grid, 128, 0, c->stream>>>(a); else k_accel_fm<true, 3, true><<<grid, 128, 0, c->stream>>>(a); }
        else k_accel_fm<false, 4, true><<<grid, 128, 0, c->stream>>>(a);
    } else {
        if (ti_on) { if (mb4) k_accel_fm<true, 4, false><<<grid, 128, 0, c->stream>>>(a); else k_accel_fm<true, 3, false><<<grid, 128, 0, c->stream>>>(a); }
        else k_accel_fm<false, 4, false><<<grid, 128, 0, c->stream>>>(a);
    }
    return 0;
}

int launch_rate_fm(wf_ctx *c, const PairArgs &a, int grid) { k_rate_fm<<<grid, 128, 0, c->stream>>>(a); return 0; }

int launch_post_fm(wf_ctx *c, const PairArgs &a, int grid, bool therm, bool ekin, bool dens)
{
    if (therm && ekin && dens) k_post_fm<true, true, true><<<grid, 128, 0, c->stream>>>(a);
    else if (therm && ekin) k_post_fm<true, true, false><<<grid, 128, 0, c->stream>>>(a);
    else if (ekin && dens) k_post_fm<false, true, true><<<grid, 128, 0, c->stream>>>(a);
    else if (ekin) k_post_fm<false, true, false><<<grid, 128, 0, c->stream>>>(a);
    else if (dens && !therm) k_post_fm<false, false, true><<<grid, 128, 0, c->stream>>>(a);
    else return -1;
    return 0;
}

} // namespace wf
