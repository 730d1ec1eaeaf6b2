// Pair loops of the Fraser step, gather form: one thread per particle walks its own neighbour list over cell-sorted,
// 32-byte packed arrays; no atomics, no pair tables. Replaces the per-pair passes + Nishiura reductions
//   CalcDensInc + DensReduction              Source/InteractionAlt.cpp:483-640
//   CalcRateTensors + RateTensorsReduction   Source/InteractionAlt.cpp:263-479   (+ Particle::CalcStressStrain, Particle.cpp:960-1067)
//   CalcAccel + AccelReduction               Source/InteractionAlt.cpp:12-260
//   CalcTempInc                              Source/Thermal.cpp:53-217
//   CalcKinEnergyEqn                         Source/Domain.cpp:1752-1808
// and Domain::StartAcceleration (Source/Domain.cpp:1519-1597), which is fused into the density kernel.
//
// Each pair term is evaluated with the reference's operand order. The reference evaluates a pair once with
// P1 = min(i,j), P2 = max(i,j) and adds +m_j*t to P1, -m_i*t to P2; every term used here is exactly symmetric or
// antisymmetric under i<->j in IEEE arithmetic (x_ij, v_ij flip sign exactly; h, rho sums commute), so evaluating it
// from particle i's side gives the same bits. GradientType 2 (sigma_i - sigma_j) is the exception and is evaluated in
// min/max order explicitly. Sums run in the reference's order (see wf_neighbour.cu).
#include "wf_internal.cuh"

using namespace wf;

namespace {

struct PairArgs {
    int64_t n, cap; int nsph, K, dim, kt, gt;
    const uint32_t *nbr; const int *nb_cnt, *nb_hi; const int *orig;
    double4 *Q0, *Q1, *Q2, *S0, *S1, *S2; double *F; const ClassProps *cls; const MatProps *mats;
    double dt; int any_ti, thermal, pl_heating, contact, cont_heat_fric; double pl_work_heat_frac;
};

#define OWN(f) a.F[(int64_t)(f) * a.cap + s]

__device__ inline int nb_row(int n, int nhi, int K) { return n < nhi ? n : (K - 1 - (n - nhi)); }

// kernel constants for the 3-D cubic spline, same expressions as Functions.cpp:33,82
__device__ inline double c3_of(double h) { return 1.0 / (h * h * h * WF_PI); }
__device__ inline double c4_of(double h) { return 1.0 / (h * h * h * h * WF_PI); }

template <int KT> __device__ inline double GK_of(int dim, int kt, double q, double hh, double C4)
{ if (KT == 0) return cubicGW_c(C4, q, hh); else return gradW(dim, kt, q, hh); }
template <int KT> __device__ inline double W_of(int dim, int kt, double q, double hh, double C3)
{ if (KT == 0) return cubicW_c(C3, q); else return kernelW(dim, kt, q, hh); }

// x^4 rounded from a double-double product: reproduces a correctly rounded pow(x, 4.0) (InteractionAlt.cpp:108 with
// the default TIn = 4) without the cost of a general fp64 pow.
__device__ inline double pow4_cr(double r)
{
    double r2h = r * r, r2l = __fma_rn(r, r, -r2h);
    double r4h = r2h * r2h, e1 = __fma_rn(r2h, r2h, -r4h);
    return r4h + (e1 + 2.0 * r2h * r2l);
}

// ---- StartAcceleration (TIR) + CalcDensInc + DensReduction ------------------------------------------------------------
template <int KT> __global__ void __launch_bounds__(128) k_density(PairArgs a)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n) return;
    if (a.orig[s] >= a.nsph) return;
    const double4 qi = a.Q0[s], wi = a.Q1[s];
    const int ci = (int)a.Q2[s].z;
    // Domain.cpp:1528-1570: tensile-instability tensor from the stress and density the step starts with
    if (a.any_ti) {
        const double TI = a.cls[ci].TI;
        if (TI > 0.0) {
            const double rho2 = wi.w * wi.w;
            const double pc = TI / (rho2);
            double4 s0 = a.S0[s]; double2 s1 = *(const double2 *)&a.S1[s];
            S6 sg; sg.xx = s0.x; sg.yy = s0.y; sg.zz = s0.z; sg.xy = s0.w; sg.yz = s1.x; sg.xz = s1.y;
            double d[3]; M3 Vec; sym_eig3(m3sym(sg), d, Vec);
            M3 VecT = mtrans(Vec), Val = m3zero();
            for (int k = 0; k < 3; ++k) { if (d[k] > 0) Val(k, k) = -pc * d[k]; else Val(k, k) = 0.0; }
            M3 T = mmul(mmul(Vec, Val), VecT);
            // stored symmetric (upper triangle); the reference's TIR differs from its transpose by round-off only
            ((double2 *)&a.S1[s])[1] = make_double2(T(0, 0), T(1, 1));
            a.S2[s] = make_double4(T(2, 2), T(0, 1), T(1, 2), T(0, 2));
        }
    }
    const V3 xi = v3(qi.x, qi.y, qi.z), vi = v3(wi.x, wi.y, wi.z);
    const double hi = qi.w, C4i = c4_of(hi);
    const int cnt = a.nb_cnt[s], nhi = a.nb_hi[s];
    double dd = 0.;
    for (int n = 0; n < cnt; ++n) {
        const int j = (int)(a.nbr[(int64_t)nb_row(n, nhi, a.K) * a.cap + s] & 0x7fffffffu);
        const double4 qj = a.Q0[j], wj = a.Q1[j];
        const double mj = a.Q2[j].x;
        double hh, C4;
        if (qj.w == hi) { hh = hi; C4 = C4i; } else { hh = (hi + qj.w) / 2; C4 = c4_of(hh); }
        const V3 xij = xi - v3(qj.x, qj.y, qj.z);
        const double rij = norm(xij);
        const V3 vij = vi - v3(wj.x, wj.y, wj.z);
        const double GK = GK_of<KT>(a.dim, a.kt, rij / hh, hh, C4);
        const double t1 = dot(vij, GK * xij);
        dd += mj / wj.w * t1;
    }
    dd *= wi.w;
    OWN(F_DDENS) = dd;
}

// rho += dt*drho (SolverFraser.cpp:248-252)
__global__ void k_density_update(PairArgs a)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n) return;
    double *rho = &((double *)&a.Q1[s])[3];
    *rho = *rho + a.dt * OWN(F_DDENS);
}

// Material.cpp:6-25,59-64
__device__ inline double yield_stress(const MatProps &m, double strain, double strain_rate, double temp)
{
    if (m.model == 1) { if (strain + m.eps0 > m.eps1) return m.p[1] * pow(strain + m.eps0, m.p[2]); else return m.p[0]; }
    const double A = m.p[0], B = m.p[1], n = m.p[2], C = m.p[3], e0 = m.p[4], mm = m.p[5], Tm = m.p[6], Tt = m.p[7];
    const double T_h = (temp - Tt) / (Tm - Tt);
    double f = 1.0;
    if (strain_rate > e0) f = (1.0 + C * log(strain_rate / e0));
    return (A + B * pow(strain, n)) * f * (1.0 - pow(T_h, mm));
}

// ---- CalcRateTensors + RateTensorsReduction + CalcStressStrain -------------------------------------------------------
template <int KT> __global__ void __launch_bounds__(128) k_rate_stress(PairArgs a)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n) return;
    if (a.orig[s] >= a.nsph) return;
    const double4 qi = a.Q0[s], wi = a.Q1[s];
    const V3 xi = v3(qi.x, qi.y, qi.z), vi = v3(wi.x, wi.y, wi.z);
    const double hi = qi.w, C4i = c4_of(hi);
    const int cnt = a.nb_cnt[s], nhi = a.nb_hi[s];
    double sxx = 0., syy = 0., szz = 0., sxy = 0., syz = 0., sxz = 0., r01 = 0., r02 = 0., r12 = 0.;
    for (int n = 0; n < cnt; ++n) {
        const int j = (int)(a.nbr[(int64_t)nb_row(n, nhi, a.K) * a.cap + s] & 0x7fffffffu);
        const double4 qj = a.Q0[j], wj = a.Q1[j];
        const double mjdj = a.Q2[j].x / wj.w;
        double hh, C4;
        if (qj.w == hi) { hh = hi; C4 = C4i; } else { hh = (hi + qj.w) / 2; C4 = c4_of(hh); }
        const V3 xij = xi - v3(qj.x, qj.y, qj.z);
        const double rij = norm(xij);
        const V3 vab = vi - v3(wj.x, wj.y, wj.z);
        const double GK = GK_of<KT>(a.dim, a.kt, rij / hh, hh, C4);
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
    // ---------------- Particle::CalcStressStrain (Particle.cpp:960-1067) ----------------
    const int ci = (int)a.Q2[s].z; const ClassProps cp = a.cls[ci];
    const double dt = a.dt;
    const double rho = wi.w;
    const double Pressure = eos(cp.PresEq, cp.Cs, cp.P0, rho, cp.RefDensity);
    S6 srs; srs.xx = sxx; srs.yy = syy; srs.zz = szz; srs.xy = sxy; srs.yz = syz; srs.xz = sxz;
    const M3 SR = m3sym(srs);
    M3 RR = m3zero(); RR(0, 1) = r01; RR(0, 2) = r02; RR(1, 2) = r12; RR(1, 0) = -r01; RR(2, 0) = -r02; RR(2, 1) = -r12;
    S6 shs; shs.xx = OWN(F_SH); shs.yy = OWN(F_SH + 1); shs.zz = OWN(F_SH + 2); shs.xy = OWN(F_SH + 3); shs.yz = OWN(F_SH + 4); shs.xz = OWN(F_SH + 5);
    M3 SS = m3sym(shs);
    const M3 RRT = mtrans(RR), SRT = mmul(SS, RRT), RS = mmul(RR, SS);
    M3 I = m3zero(); I(0, 0) = 1.0; I(1, 1) = 1.0; I(2, 2) = 1.0;
    const double tr = SR(0, 0) + SR(1, 1) + SR(2, 2);
    const M3 dev = SR - (1.0 / 3.0 * tr) * I;
    SS = dt * (((2.0 * cp.G) * dev + SRT) + RS) + SS;
    const double effsr = sqrt(0.5 * ((SR(0, 0) - SR(1, 1)) * (SR(0, 0) - SR(1, 1)) + (SR(1, 1) - SR(2, 2)) * (SR(1, 1) - SR(2, 2)) +
                                     (SR(2, 2) - SR(0, 0)) * (SR(2, 2) - SR(0, 0))) +
                              3.0 * (SR(0, 1) * SR(0, 1) + SR(1, 2) * SR(1, 2) + SR(2, 0) * SR(2, 0)));
    double dep = 0., Sigmay = OWN(F_SIGY);
    if (cp.Fail == 1) {
        const double J2 = 0.5 * (SS(0, 0) * SS(0, 0) + 2.0 * SS(0, 1) * SS(1, 0) + 2.0 * SS(0, 2) * SS(2, 0) + SS(1, 1) * SS(1, 1) +
                                 2.0 * SS(1, 2) * SS(2, 1) + SS(2, 2) * SS(2, 2));
        const double sig_trial = sqrt(3.0 * J2);
        const double rs = Sigmay / sig_trial;
        SS = ((1.0 < rs) ? 1.0 : rs) * SS;            // std::min(Sigmay/sig_trial, 1.0), NaN-propagating like the reference
        double pl = OWN(F_PLS);
        if (cp.Material_model == 1) Sigmay = yield_stress(a.mats[cp.mat_id], pl, 0.0, 0.0);
        else if (cp.Material_model == 2) Sigmay = yield_stress(a.mats[cp.mat_id], pl, effsr, a.Q2[s].y);
        if (sig_trial > Sigmay) {
            double Ep = cp.Ep;
            if (Ep < 0) Ep = 1. * a.mats[cp.mat_id].E;
            dep = (sig_trial - Sigmay) / (3. * cp.G + Ep);
            pl += dep;
            OWN(F_PLS) = pl; OWN(F_DPLS) = dep;
            if (cp.Material_model == 0) Sigmay += dep * Ep;
        }
        OWN(F_SIGY) = Sigmay;
    }
    const M3 Sg = (-Pressure) * I + SS;
    OWN(F_SH) = SS(0, 0); OWN(F_SH + 1) = SS(1, 1); OWN(F_SH + 2) = SS(2, 2); OWN(F_SH + 3) = SS(0, 1); OWN(F_SH + 4) = SS(1, 2); OWN(F_SH + 5) = SS(0, 2);
    a.S0[s] = make_double4(Sg(0, 0), Sg(1, 1), Sg(2, 2), Sg(0, 1));
    ((double2 *)&a.S1[s])[0] = make_double2(Sg(1, 2), Sg(0, 2));
    ((double *)&a.Q2[s])[3] = Pressure;
    if (dep > 0.0) {
        const double f = dep / Sigmay;
        const double ixx = f * (Sg(0, 0) - 0.5 * (Sg(1, 1) + Sg(2, 2)));
        const double iyy = f * (Sg(1, 1) - 0.5 * (Sg(0, 0) + Sg(2, 2)));
        const double izz = f * (Sg(2, 2) - 0.5 * (Sg(0, 0) + Sg(1, 1)));
        const double ixy = 1.5 * f * (Sg(0, 1)), ixz = 1.5 * f * (Sg(0, 2)), iyz = 1.5 * f * (Sg(1, 2));
        OWN(F_SI) = ixx; OWN(F_SI + 1) = iyy; OWN(F_SI + 2) = izz; OWN(F_SI + 3) = ixy; OWN(F_SI + 4) = iyz; OWN(F_SI + 5) = ixz;
        OWN(F_SP) = OWN(F_SP) + ixx; OWN(F_SP + 1) = OWN(F_SP + 1) + iyy; OWN(F_SP + 2) = OWN(F_SP + 2) + izz;
        OWN(F_SP + 3) = OWN(F_SP + 3) + ixy; OWN(F_SP + 4) = OWN(F_SP + 4) + iyz; OWN(F_SP + 5) = OWN(F_SP + 5) + ixz;
    }
    OWN(F_ST) = dt * sxx + OWN(F_ST); OWN(F_ST + 1) = dt * syy + OWN(F_ST + 1); OWN(F_ST + 2) = dt * szz + OWN(F_ST + 2);
    OWN(F_ST + 3) = dt * sxy + OWN(F_ST + 3); OWN(F_ST + 4) = dt * syz + OWN(F_ST + 4); OWN(F_ST + 5) = dt * sxz + OWN(F_ST + 5);
    OWN(F_SR) = sxx; OWN(F_SR + 1) = syy; OWN(F_SR + 2) = szz; OWN(F_SR + 3) = sxy; OWN(F_SR + 4) = syz; OWN(F_SR + 5) = sxz;
    OWN(F_RR) = r01; OWN(F_RR + 1) = r02; OWN(F_RR + 2) = r12;
    OWN(F_EFFSR) = effsr;
}

// ---- CalcAccel + AccelReduction ------------------------------------------------------------------------------------------
template <int KT, int GT> __global__ void __launch_bounds__(128) k_accel(PairArgs a)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n) return;
    if (a.orig[s] >= a.nsph) return;
    const double4 qi = a.Q0[s], wi = a.Q1[s];
    const V3 xi = v3(qi.x, qi.y, qi.z), vi = v3(wi.x, wi.y, wi.z);
    const double hi = qi.w, C3i = c3_of(hi), C4i = c4_of(hi), di = wi.w;
    const int ci = (int)a.Q2[s].z; const ClassProps cpi = a.cls[ci];
    const double4 s0i = a.S0[s], s1i = a.S1[s], s2i = a.S2[s];
    const int cnt = a.nb_cnt[s], nhi = a.nb_hi[s];
    const double inv_di2 = 1.0 / (di * di);
    const double Ci = soundSpeed(cpi.PresEq, cpi.Cs, di, cpi.RefDensity);
    // W(dx0/h) of the tensile-instability term for the common same-class, same-h neighbour
    const double W0i = W_of<KT>(a.dim, a.kt, (cpi.TIInitDist + cpi.TIInitDist) / (2.0 * hi), hi, C3i);
    double ax = 0., ay = 0., az = 0.;
    for (int n = 0; n < cnt; ++n) {
        const int j = (int)(a.nbr[(int64_t)nb_row(n, nhi, a.K) * a.cap + s] & 0x7fffffffu);
        const double4 qj = a.Q0[j], wj = a.Q1[j], q2j = a.Q2[j];
        const double4 s0j = a.S0[j]; const double2 s1j = *(const double2 *)&a.S1[j];
        const double mj = q2j.x, dj = wj.w; const int cj = (int)q2j.z;
        const bool same = (cj == ci);
        const bool sameh = (qj.w == hi);
        double hh, C3, C4;
        if (sameh) { hh = hi; C3 = C3i; C4 = C4i; } else { hh = (hi + qj.w) / 2; C3 = c3_of(hh); C4 = c4_of(hh); }
        // GradientType 2 is not symmetric under i<->j: evaluate it the way the reference does, P1 = smaller creation index
        const bool swap = (GT == 2) && (n >= nhi);
        V3 xij = xi - v3(qj.x, qj.y, qj.z), vij = vi - v3(wj.x, wj.y, wj.z);
        if (swap) { xij = v3(-xij.x, -xij.y, -xij.z); vij = v3(-vij.x, -vij.y, -vij.z); }
        const double rij = norm(xij);
        const double q = rij / hh;
        const double GK = GK_of<KT>(a.dim, a.kt, q, hh, C4);
        double Alpha, Beta, TIj, TInj, TId_j, Cj;
        if (same) { Alpha = (cpi.Alpha + cpi.Alpha) / 2.0; Beta = (cpi.Beta + cpi.Beta) / 2.0; TIj = cpi.TI; TInj = cpi.TIn; TId_j = cpi.TIInitDist; Cj = soundSpeed(cpi.PresEq, cpi.Cs, dj, cpi.RefDensity); }
        else { const ClassProps cpj = a.cls[cj]; Alpha = (cpi.Alpha + cpj.Alpha) / 2.0; Beta = (cpi.Beta + cpj.Beta) / 2.0; TIj = cpj.TI; TInj = cpj.TIn; TId_j = cpj.TIInitDist;
               Cj = soundSpeed(cpj.PresEq, cpj.Cs, dj, cpj.RefDensity); }
        // artificial viscosity (InteractionAlt.cpp:72-94)
        double pi = 0.0; bool pi_on = false;
        if (Alpha != 0.0 || Beta != 0.0) {
            const double vx = dot(vij, xij);
            const double MUij = hh * vx / (rij * rij + 0.01 * hh * hh);
            const double Cij = 0.5 * (Ci + Cj);
            if (vx < 0) { pi = (Alpha * Cij * MUij + Beta * MUij * MUij) / (0.5 * (di + dj)); pi_on = true; }
        }
        // momentum tensor (InteractionAlt.cpp:138-143)
        double mxx, myy, mzz, mxy, myz, mxz;
        if (GT == 0) {
            const double inv_dj2 = 1.0 / (dj * dj);
            mxx = s0i.x * inv_di2 + s0j.x * inv_dj2; myy = s0i.y * inv_di2 + s0j.y * inv_dj2; mzz = s0i.z * inv_di2 + s0j.z * inv_dj2;
            mxy = s0i.w * inv_di2 + s0j.w * inv_dj2; myz = s1i.x * inv_di2 + s1j.x * inv_dj2; mxz = s1i.y * inv_di2 + s1j.y * inv_dj2;
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
        if (a.any_ti && (cpi.TI > 0.0 || TIj > 0.0)) {
            const double K = W_of<KT>(a.dim, a.kt, q, hh, C3);
            const double W0 = (same && sameh) ? W0i : W_of<KT>(a.dim, a.kt, (cpi.TIInitDist + TId_j) / (2.0 * hh), hh, C3);
            const double nbar = (cpi.TIn + TInj) / 2.0;
            const double r = K / W0;
            const double fti = (nbar == 4.0) ? pow4_cr(r) : pow(r, nbar);
            const double2 t1j = ((const double2 *)&a.S1[j])[1]; const double4 t2j = a.S2[j];
            mxx += (s1i.z + t1j.x) * fti; myy += (s1i.w + t1j.y) * fti; mzz += (s2i.x + t2j.x) * fti;
            mxy += (s2i.y + t2j.y) * fti; myz += (s2i.z + t2j.z) * fti; mxz += (s2i.w + t2j.w) * fti;
        }
        // temp = (GK*xij)^T M (matvec.h:829-834), a_i += m_j*temp (or -= when i is P2 of a GradientType-2 pair)
        const V3 g = GK * xij;
        double tx = g.x * mxx + g.y * mxy + g.z * mxz;
        double ty = g.x * mxy + g.y * myy + g.z * myz;
        double tz = g.x * mxz + g.y * myz + g.z * mzz;
        if (a.dim == 2) tz = 0.0;
        if (swap) { ax -= mj * tx; ay -= mj * ty; az -= mj * tz; }
        else      { ax += mj * tx; ay += mj * ty; az += mj * tz; }
    }
    OWN(F_AX) = ax; OWN(F_AY) = ay; OWN(F_AZ) = az;
}

// ---- CalcPlasticWorkHeat, T += dt*dTdt, CalcConvHeat (ThermalStuct.cpp:14-39, Thermal.cpp:721-732,339-389) ------------
__global__ void k_thermal_pre(PairArgs a)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n) return;
    if (a.orig[s] >= a.nsph) return;
    const double dt = a.dt;
    if (a.pl_heating && OWN(F_DPLS) > 0.0) {
        const double f = 1. / dt;
        const double4 s0 = a.S0[s]; const double2 s1 = *(const double2 *)&a.S1[s];
        const double dxx = OWN(F_SI) * f, dyy = OWN(F_SI + 1) * f, dzz = OWN(F_SI + 2) * f, dxy = OWN(F_SI + 3) * f, dyz = OWN(F_SI + 4) * f, dxz = OWN(F_SI + 5) * f;
        OWN(F_QPL) = s0.x * dxx + 2.0 * s0.w * dxy + 2.0 * s1.y * dxz + s0.y * dyy + 2.0 * s1.x * dyz + s0.z * dzz;
    }
    if (a.thermal) {
        double4 q2 = a.Q2[s];
        const double T = q2.y + dt * OWN(F_DTDT);
        ((double *)&a.Q2[s])[1] = T;
        const ClassProps cp = a.cls[(int)q2.z];
        if (cp.Thermal_BC == 1) {
            const double dens = a.Q1[s].w;
            const double dS2 = pow(q2.x / dens, 2.0 / 3.0);
            OWN(F_QCONV) = dens * cp.h_conv * dS2 * (cp.T_inf - T) / q2.x;
        }
    }
}

// ---- CalcTempInc (Thermal.cpp:53-217); sums run in list order (round-off level difference to the reference's
// pair-discovery order); the m_j/rho_j factor is the SECOND particle's, as in the reference ---------------------------
template <int KT> __global__ void __launch_bounds__(128) k_thermal(PairArgs a)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n) return;
    if (a.orig[s] >= a.nsph) { OWN(F_DTDT) = 0.; return; }
    const double4 qi = a.Q0[s], q2i = a.Q2[s];
    const V3 xi = v3(qi.x, qi.y, qi.z);
    const double hi = qi.w, C4i = c4_of(hi), di = a.Q1[s].w, Ti = q2i.y, mi = q2i.x;
    const int ci = (int)q2i.z; const ClassProps cpi = a.cls[ci];
    const int cnt = a.nb_cnt[s], nhi = a.nb_hi[s];
    double temp = 0.;
    for (int n = 0; n < cnt; ++n) {
        const uint32_t e = a.nbr[(int64_t)nb_row(n, nhi, a.K) * a.cap + s];
        const int j = (int)(e & 0x7fffffffu); const bool ifirst = (e >> 31) != 0;
        const double4 qj = a.Q0[j], q2j = a.Q2[j];
        const double dj = a.Q1[j].w;
        const int cj = (int)q2j.z; const double kj = (cj == ci) ? cpi.k_T : a.cls[cj].k_T;
        double hh, C4;
        if (qj.w == hi) { hh = hi; C4 = C4i; } else { hh = (hi + qj.w) / 2.0; C4 = c4_of(hh); }
        const V3 xj = v3(qj.x, qj.y, qj.z);
        if (ifirst) {      // i is P1: temp[i] += m
            const V3 xij = xi - xj; const double nr = norm(xij);
            const double GK = GK_of<KT>(a.dim, a.kt, nr / hh, hh, C4);
            const double m = q2j.x / dj * 4. * (cpi.k_T * kj) / (cpi.k_T + kj) * (Ti - q2j.y) * dot(xij, GK * xij) / (nr * nr);
            temp += m;
        } else {           // i is P2: temp[i] -= m, m built with P1 = j and m_i/rho_i
            const V3 xij = xj - xi; const double nr = norm(xij);
            const double GK = GK_of<KT>(a.dim, a.kt, nr / hh, hh, C4);
            const double m = mi / di * 4. * (kj * cpi.k_T) / (kj + cpi.k_T) * (q2j.y - Ti) * dot(xij, GK * xij) / (nr * nr);
            temp -= m;
        }
    }
    const double f = 1. / (di * cpi.cp_T);
    double dTdt = f * (temp + OWN(F_QCONV) + cpi.q_source + OWN(F_QPL) * a.pl_work_heat_frac + OWN(F_QCONT));
    if (a.contact && a.cont_heat_fric) dTdt += f * OWN(F_QFRIC);
    OWN(F_DTDT) = dTdt;
    // CalcThermalExpStrainRate (Particle.cpp:953-957): StrainRate += th_exp*dTdt*I
    const double te = cpi.th_exp * dTdt;
    OWN(F_SR) = OWN(F_SR) + te; OWN(F_SR + 1) = OWN(F_SR + 1) + te; OWN(F_SR + 2) = OWN(F_SR + 2) + te;
}

// ---- CalcKinEnergyEqn pair loop + CalcIntEnergyEqn (Domain.cpp:1752-1829, Particle.cpp:941-950) ---------------------
template <int KT> __global__ void __launch_bounds__(128) k_energy(PairArgs a)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n) return;
    if (a.orig[s] >= a.nsph) { OWN(F_DKIN) = 0.; return; }
    const double4 qi = a.Q0[s], wi = a.Q1[s], q2i = a.Q2[s];
    const V3 xi = v3(qi.x, qi.y, qi.z), vi = v3(wi.x, wi.y, wi.z);
    const double hi = qi.w, C4i = c4_of(hi), di = wi.w, Pi = q2i.w, mi = q2i.x;
    const int cnt = a.nb_cnt[s], nhi = a.nb_hi[s];
    double dk = 0.;
    for (int n = 0; n < cnt; ++n) {
        const uint32_t e = a.nbr[(int64_t)nb_row(n, nhi, a.K) * a.cap + s];
        const int j = (int)(e & 0x7fffffffu); const bool ifirst = (e >> 31) != 0;
        const double4 qj = a.Q0[j], wj = a.Q1[j], q2j = a.Q2[j];
        const double dj = wj.w;
        double hh, C4;
        if (qj.w == hi) { hh = hi; C4 = C4i; } else { hh = (hi + qj.w) / 2.0; C4 = c4_of(hh); }
        const V3 xj = v3(qj.x, qj.y, qj.z), vj = v3(wj.x, wj.y, wj.z);
        if (ifirst) {
            const V3 xij = xi - xj, vij = vi - vj;
            const double GK = GK_of<KT>(a.dim, a.kt, norm(xij) / hh, hh, C4);
            dk += 0.5 * q2j.x * (Pi / (di * di) + q2j.w / (dj * dj)) * dot(vij, GK * xij);
        } else {
            const V3 xij = xj - xi, vij = vj - vi;
            const double GK = GK_of<KT>(a.dim, a.kt, norm(xij) / hh, hh, C4);
            dk -= 0.5 * mi * (q2j.w / (dj * dj) + Pi / (di * di)) * dot(vij, GK * xij);
        }
    }
    OWN(F_DKIN) = dk;
    OWN(F_DINT) = mi / di * (OWN(F_SH) * OWN(F_SR) + 2.0 * OWN(F_SH + 3) * OWN(F_SR + 3) + 2.0 * OWN(F_SH + 5) * OWN(F_SR + 5) +
                             OWN(F_SH + 1) * OWN(F_SR + 1) + 2.0 * OWN(F_SH + 4) * OWN(F_SR + 4) + OWN(F_SH + 2) * OWN(F_SR + 2));
}

PairArgs make_args(wf_ctx *c)
{
    PairArgs a; a.n = c->n; a.cap = c->cap; a.nsph = (int)c->nsph; a.K = c->nbr_k; a.dim = c->prm.dimension; a.kt = c->prm.kernel_type; a.gt = c->prm.gradient_type;
    a.nbr = c->nbr; a.nb_cnt = c->nb_cnt; a.nb_hi = c->nb_hi; a.orig = c->Ip(I_ORIG);
    a.Q0 = c->Q0[c->cur]; a.Q1 = c->Q1[c->cur]; a.Q2 = c->Q2[c->cur]; a.S0 = c->S0[c->cur]; a.S1 = c->S1[c->cur]; a.S2 = c->S2[c->cur];
    a.F = c->F[c->cur]; a.cls = c->d_classes; a.mats = c->d_mats; a.dt = c->dt; a.any_ti = c->any_ti ? 1 : 0;
    a.thermal = c->prm.thermal_solver; a.pl_heating = c->prm.pl_heating; a.contact = c->prm.contact; a.cont_heat_fric = c->prm.cont_heat_fric;
    a.pl_work_heat_frac = c->prm.pl_work_heat_frac;
    return a;
}

} // namespace

namespace wf {

#define WF_PAIR_PRECHECK(c) if (!(c)->nbr) return fail((c), WF_ERR_ARG, "pair kernel called before the first neighbour build")

int pair_density(wf_ctx *c)
{
    WF_PAIR_PRECHECK(c);
    PairArgs a = make_args(c); const int g = grid_for(c->n, 128);
    {   ProfScope ps(c, PK_DENSITY);
        if (a.kt == 0 && a.dim == 3) k_density<0><<<g, 128, 0, c->stream>>>(a); else k_density<1><<<g, 128, 0, c->stream>>>(a);
        WF_LAUNCH_CHECK(c);
        k_density_update<<<grid_for(c->n, 256), 256, 0, c->stream>>>(a); WF_LAUNCH_CHECK(c); }
    c->launches += 2;
    return 0;
}
int pair_rate_stress(wf_ctx *c)
{
    WF_PAIR_PRECHECK(c);
    PairArgs a = make_args(c); const int g = grid_for(c->n, 128);
    ProfScope ps(c, PK_RATE_STRESS);
    if (a.kt == 0 && a.dim == 3) k_rate_stress<0><<<g, 128, 0, c->stream>>>(a); else k_rate_stress<1><<<g, 128, 0, c->stream>>>(a);
    WF_LAUNCH_CHECK(c); c->launches++;
    return 0;
}
int pair_accel(wf_ctx *c)
{
    WF_PAIR_PRECHECK(c);
    PairArgs a = make_args(c); const int g = grid_for(c->n, 128);
    ProfScope ps(c, PK_ACCEL);
    const bool fast = (a.kt == 0 && a.dim == 3);
    switch (a.gt) {
    case 0: if (fast) k_accel<0, 0><<<g, 128, 0, c->stream>>>(a); else k_accel<1, 0><<<g, 128, 0, c->stream>>>(a); break;
    case 1: if (fast) k_accel<0, 1><<<g, 128, 0, c->stream>>>(a); else k_accel<1, 1><<<g, 128, 0, c->stream>>>(a); break;
    default: if (fast) k_accel<0, 2><<<g, 128, 0, c->stream>>>(a); else k_accel<1, 2><<<g, 128, 0, c->stream>>>(a); break;
    }
    WF_LAUNCH_CHECK(c); c->launches++;
    return 0;
}
int pair_thermal(wf_ctx *c)
{
    WF_PAIR_PRECHECK(c);
    if (!c->prm.thermal_solver && !c->prm.pl_heating) return 0;
    PairArgs a = make_args(c);
    ProfScope ps(c, PK_THERMAL);
    k_thermal_pre<<<grid_for(c->n, 256), 256, 0, c->stream>>>(a); WF_LAUNCH_CHECK(c); c->launches++;
    if (c->prm.thermal_solver) {
        const int g = grid_for(c->n, 128);
        if (a.kt == 0 && a.dim == 3) k_thermal<0><<<g, 128, 0, c->stream>>>(a); else k_thermal<1><<<g, 128, 0, c->stream>>>(a);
        WF_LAUNCH_CHECK(c); c->launches++;
    }
    return 0;
}
int pair_energy(wf_ctx *c)
{
    WF_PAIR_PRECHECK(c);
    PairArgs a = make_args(c); const int g = grid_for(c->n, 128);
    ProfScope ps(c, PK_ENERGY);
    if (a.kt == 0 && a.dim == 3) k_energy<0><<<g, 128, 0, c->stream>>>(a); else k_energy<1><<<g, 128, 0, c->stream>>>(a);
    WF_LAUNCH_CHECK(c); c->launches++;
    return 0;
}

} // namespace wf
