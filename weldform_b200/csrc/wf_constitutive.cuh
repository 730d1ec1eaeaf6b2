// The per-particle constitutive update that closes the rate-tensor pass, shared by the exact kernel (wf_pair.cu) and the contracted one
// (wf_pair_fm.cu). Every expression keeps the reference's operand order; compiled -fmad=false it is bit-identical to the reference.
#pragma once
#include "wf_pair_args.cuh"

namespace wf {

#define WF_OWN(f) a.F[(int64_t)(f) * a.cap + s]

// Material.cpp:6-25,59-64
__device__ __forceinline__ double yield_stress(const MatProps &m, double strain, double strain_rate, double temp)
{
    if (m.model == 1) { if (strain + m.eps0 > m.eps1) return m.p[1] * pow(strain + m.eps0, m.p[2]); else return m.p[0]; }
    if (m.model == 3) {        // GMT::CalcYieldStress (Material.cpp:96-134): clamped strain, strain rate and temperature; p = n1, n2, C1, C2, m1, m2, I1, I2, limits
        const double n1 = m.p[0], n2 = m.p[1], C1 = m.p[2], C2 = m.p[3], m1 = m.p[4], m2 = m.p[5], I1 = m.p[6], I2 = m.p[7];
        double e = strain, er = strain_rate, T = temp;
        if (e < m.p[8]) e = m.p[8]; else if (e > m.p[9]) e = m.p[9];
        if (er < m.p[10]) er = m.p[10]; else if (er > m.p[11]) er = m.p[11];
        if (T < m.p[12]) T = m.p[12]; else if (T > m.p[13]) T = m.p[13];
        return C1 * exp(C2 * T) * pow(e, n1 * T + n2) * exp((I1 * T + I2) / e) * pow(er, m1 * T + m2);
    }
    const double A = m.p[0], B = m.p[1], n = m.p[2], C = m.p[3], e0 = m.p[4], mm = m.p[5], Tm = m.p[6], Tt = m.p[7];
    const double T_h = (temp - Tt) / (Tm - Tt);
    double f = 1.0;
    if (strain_rate > e0) f = (1.0 + C * log(strain_rate / e0));
    return (A + B * pow(strain, n)) * f * (1.0 - pow(T_h, mm));
}


// s.. = StrainRate (6), r.. = RotationRate (01, 02, 12) of particle s as summed over its neighbours
__device__ __forceinline__ void stress_update(const PairArgs &a, int64_t s, double sxx, double syy, double szz, double sxy, double syz, double sxz,
                                              double r01, double r02, double r12)
{
    // ---------------- Particle::CalcStressStrain (Particle.cpp:960-1067) ----------------
    const int ci = a.cls_of[s]; const ClassProps cp = a.cls[ci];
    const double dt = a.dt;
    double4 q2 = a.Q2[s];
    const double rho = q2.y;
    const double Pressure = eos(cp.PresEq, cp.Cs, cp.P0, rho, cp.RefDensity);
    S6 srs; srs.xx = sxx; srs.yy = syy; srs.zz = szz; srs.xy = sxy; srs.yz = syz; srs.xz = sxz;
    const M3 SR = m3sym(srs);
    M3 RR = m3zero(); RR(0, 1) = r01; RR(0, 2) = r02; RR(1, 2) = r12; RR(1, 0) = -r01; RR(2, 0) = -r02; RR(2, 1) = -r12;
    S6 shs; shs.xx = WF_OWN(F_SH); shs.yy = WF_OWN(F_SH + 1); shs.zz = WF_OWN(F_SH + 2); shs.xy = WF_OWN(F_SH + 3); shs.yz = WF_OWN(F_SH + 4); shs.xz = WF_OWN(F_SH + 5);
    // StartAcceleration zeroes Sigma / ShearStress of fixed particles every step (Domain.cpp:1572-1578)
    if (a.is_free && !a.is_free[s]) { shs.xx = shs.yy = shs.zz = shs.xy = shs.yz = shs.xz = 0.0; }
    M3 SS = m3sym(shs);
    const M3 RRT = mtrans(RR), SRT = mmul(SS, RRT), RS = mmul(RR, SS);
    M3 I = m3zero(); I(0, 0) = 1.0; I(1, 1) = 1.0; I(2, 2) = 1.0;
    const double tr = SR(0, 0) + SR(1, 1) + SR(2, 2);
    const M3 dev = SR - (1.0 / 3.0 * tr) * I;
    SS = dt * (((2.0 * cp.G) * dev + SRT) + RS) + SS;
    const double effsr = sqrt(0.5 * ((SR(0, 0) - SR(1, 1)) * (SR(0, 0) - SR(1, 1)) + (SR(1, 1) - SR(2, 2)) * (SR(1, 1) - SR(2, 2)) +
                                     (SR(2, 2) - SR(0, 0)) * (SR(2, 2) - SR(0, 0))) +
                              3.0 * (SR(0, 1) * SR(0, 1) + SR(1, 2) * SR(1, 2) + SR(2, 0) * SR(2, 0)));
    double dep = 0., Sigmay = WF_OWN(F_SIGY);
    if (cp.Fail == 1) {
        const double J2 = 0.5 * (SS(0, 0) * SS(0, 0) + 2.0 * SS(0, 1) * SS(1, 0) + 2.0 * SS(0, 2) * SS(2, 0) + SS(1, 1) * SS(1, 1) +
                                 2.0 * SS(1, 2) * SS(2, 1) + SS(2, 2) * SS(2, 2));
        const double sig_trial = sqrt(3.0 * J2);
        const double rs = Sigmay / sig_trial;
        SS = ((1.0 < rs) ? 1.0 : rs) * SS;            // std::min(Sigmay/sig_trial, 1.0), NaN-propagating like the reference
        double pl = WF_OWN(F_PLS);
        if (cp.Material_model == 1) Sigmay = yield_stress(a.mats[cp.mat_id], pl, 0.0, 0.0);
        else if (cp.Material_model == 2 || cp.Material_model == 3) Sigmay = yield_stress(a.mats[cp.mat_id], pl, effsr, q2.z);     // JOHNSON_COOK || _GMT_ (Particle.cpp:999)
        if (sig_trial > Sigmay) {
            double Ep = cp.Ep;
            if (Ep < 0) Ep = 1. * a.mats[cp.mat_id].E;
            dep = (sig_trial - Sigmay) / (3. * cp.G + Ep);
            pl += dep;
            WF_OWN(F_PLS) = pl; WF_OWN(F_DPLS) = dep;
            if (cp.Material_model == 0) Sigmay += dep * Ep;
        }
        WF_OWN(F_SIGY) = Sigmay;
    }
    const M3 Sg = (-Pressure) * I + SS;
    WF_OWN(F_SH) = SS(0, 0); WF_OWN(F_SH + 1) = SS(1, 1); WF_OWN(F_SH + 2) = SS(2, 2); WF_OWN(F_SH + 3) = SS(0, 1); WF_OWN(F_SH + 4) = SS(1, 2); WF_OWN(F_SH + 5) = SS(0, 2);
    WF_OWN(F_SG) = Sg(0, 0); WF_OWN(F_SG + 1) = Sg(1, 1); WF_OWN(F_SG + 2) = Sg(2, 2); WF_OWN(F_SG + 3) = Sg(0, 1); WF_OWN(F_SG + 4) = Sg(1, 2); WF_OWN(F_SG + 5) = Sg(0, 2);
    WF_OWN(F_PRES) = Pressure;
    // gathered copies for the momentum pass: 1.0/(rho*rho)*Sigma (InteractionAlt.cpp:139) or Sigma (GradientType 1, 2);
    // and P/(rho*rho) for the kinetic-energy pass (Domain.cpp:1784)
    const double inv_d2 = 1.0 / (rho * rho);
    if (a.gt == 0) {
        a.S0[s] = make_double4(Sg(0, 0) * inv_d2, Sg(1, 1) * inv_d2, Sg(2, 2) * inv_d2, Sg(0, 1) * inv_d2);
        ((double2 *)&a.S1[s])[0] = make_double2(Sg(1, 2) * inv_d2, Sg(0, 2) * inv_d2);
    } else {
        a.S0[s] = make_double4(Sg(0, 0), Sg(1, 1), Sg(2, 2), Sg(0, 1));
        ((double2 *)&a.S1[s])[0] = make_double2(Sg(1, 2), Sg(0, 2));
    }
    ((double *)&a.Q2[s])[3] = Pressure / (rho * rho);
    if (dep > 0.0) {
        const double f = dep / Sigmay;
        const double ixx = f * (Sg(0, 0) - 0.5 * (Sg(1, 1) + Sg(2, 2)));
        const double iyy = f * (Sg(1, 1) - 0.5 * (Sg(0, 0) + Sg(2, 2)));
        const double izz = f * (Sg(2, 2) - 0.5 * (Sg(0, 0) + Sg(1, 1)));
        const double ixy = 1.5 * f * (Sg(0, 1)), ixz = 1.5 * f * (Sg(0, 2)), iyz = 1.5 * f * (Sg(1, 2));
        WF_OWN(F_SI) = ixx; WF_OWN(F_SI + 1) = iyy; WF_OWN(F_SI + 2) = izz; WF_OWN(F_SI + 3) = ixy; WF_OWN(F_SI + 4) = iyz; WF_OWN(F_SI + 5) = ixz;
        WF_OWN(F_SP) = WF_OWN(F_SP) + ixx; WF_OWN(F_SP + 1) = WF_OWN(F_SP + 1) + iyy; WF_OWN(F_SP + 2) = WF_OWN(F_SP + 2) + izz;
        WF_OWN(F_SP + 3) = WF_OWN(F_SP + 3) + ixy; WF_OWN(F_SP + 4) = WF_OWN(F_SP + 4) + iyz; WF_OWN(F_SP + 5) = WF_OWN(F_SP + 5) + ixz;
    }
    WF_OWN(F_ST) = dt * sxx + WF_OWN(F_ST); WF_OWN(F_ST + 1) = dt * syy + WF_OWN(F_ST + 1); WF_OWN(F_ST + 2) = dt * szz + WF_OWN(F_ST + 2);
    WF_OWN(F_ST + 3) = dt * sxy + WF_OWN(F_ST + 3); WF_OWN(F_ST + 4) = dt * syz + WF_OWN(F_ST + 4); WF_OWN(F_ST + 5) = dt * sxz + WF_OWN(F_ST + 5);
    WF_OWN(F_SR) = sxx; WF_OWN(F_SR + 1) = syy; WF_OWN(F_SR + 2) = szz; WF_OWN(F_SR + 3) = sxy; WF_OWN(F_SR + 4) = syz; WF_OWN(F_SR + 5) = sxz;
    WF_OWN(F_RR) = r01; WF_OWN(F_RR + 1) = r02; WF_OWN(F_RR + 2) = r12;
    WF_OWN(F_EFFSR) = effsr;
}

} // namespace wf
