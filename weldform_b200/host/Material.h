// Host-side mirror of the reference's material classes (Source/Material.h:64-224): constructors and accessors the
// drivers use. The yield-stress evaluation itself runs on the device (wf_pair.cu yield_stress, Material.cpp:6-73); these
// objects only carry the parameters across the API (Domain packs them into wf_material records).
#ifndef WF_HOST_MATERIAL_H
#define WF_HOST_MATERIAL_H
#include <cmath>
#include "fatal.h"
using namespace std;      // Source/Material.h:10 -- the reference drivers rely on it (unqualified ofstream, cout, vector)

#define BILINEAR      0
#define HOLLOMON      1
#define JOHNSON_COOK  2
#define _GMT_          3

class Elastic_ {
    double E_m = 0.0, nu_m = 0.0;
public:
    Elastic_() {}
    Elastic_(const double & e, const double & nu) : E_m(e), nu_m(nu) {}
    const double & E() const { return E_m; }
    const double & Poisson() const { return nu_m; }
};

class Material_ {
protected:
    Elastic_ elastic_m;
    double T_t = 0.0, T_m = 0.0;
public:
    int Material_model = BILINEAR;        // which of the subclasses this is (the device needs a plain tag, not a vtable)
    double params[16] = { 0 };            // Hollomon: sy0, K, m.  Johnson-Cook: A, B, n, C, eps_0, m, T_m, T_t.  GMT: n1, n2, C1, C2, m1, m2, I1, I2 + six limits  (wf_material::p)
    Material_() {}
    Material_(const Elastic_ el) : elastic_m(el) {}
    virtual ~Material_() {}
    double & getT_t() { return T_t; }
    double & getT_m() { return T_m; }
    const Elastic_ & Elastic() const { return elastic_m; }
    // Source/Material.h:96-101: the base class answers 0; drivers call these on the host to seed Particle::Sigmay before Solve
    virtual double CalcTangentModulus() { return 0.0; }
    virtual double CalcTangentModulus(const double &, const double &, const double &) { return 0.0; }
    virtual double CalcTangentModulus(const double &) { return 0.0; }
    virtual double CalcYieldStress() { return 0.0; }
    virtual double CalcYieldStress(const double &) { return 0.0; }
    virtual double CalcYieldStress(const double &, const double &, const double &) { return 0.0; }
};

// Source/Material.h:117-136: JohnsonCook(el, a, b, n, c, eps_0, m, T_m, T_t) -- argument ORDER is the contract (examples/FSW.cpp:113-115
// passes its values in a different order; the outcome of that call is reproduced because the signature is the same)
class JohnsonCook : public Material_ {
public:
    JohnsonCook() {}
    JohnsonCook(const Elastic_ & el, const double & a, const double & b, const double & n_, const double & c, const double & eps_0_,
                const double & m_, const double & T_m_, const double & T_t_) : Material_(el)
    {
        Material_model = JOHNSON_COOK; T_m = T_m_; T_t = T_t_;
        params[0] = a; params[1] = b; params[2] = n_; params[3] = c; params[4] = eps_0_; params[5] = m_; params[6] = T_m_; params[7] = T_t_;
    }
    double & getRefStrainRate() { return params[4]; }
    // Source/Material.cpp:6-47 and Material.h:137-146, same operand order (host-side seed values; the step evaluates the law on the device)
    double CalcYieldStress() { return 0.0; }
    double CalcYieldStress(const double & plstrain)          // Material.h:138-146: despite its name this overload returns a tangent modulus
    {
        const double B = params[1], n = params[2];
        double Et = 0.;
        if (plstrain > 0.) Et = n * B * pow(plstrain, n - 1.);
        else Et = Elastic().E() * 0.1;
        return Et;
    }
    double CalcYieldStress(const double & strain, const double & strain_rate, const double & temp)
    {
        const double A = params[0], B = params[1], n = params[2], C = params[3], eps_0 = params[4], m = params[5];
        double T_h = (temp - T_t) / (T_m - T_t);
        double f = 1.0;
        if (strain_rate > eps_0) f = (1.0 + C * log(strain_rate / eps_0));
        double sy = (A + B * pow(strain, n)) * f * (1.0 - pow(T_h, m));
        return sy;
    }
    double CalcTangentModulus(const double & plstrain, const double & strain_rate, const double & temp)
    {
        const double B = params[1], n = params[2], C = params[3], eps_0 = params[4], m = params[5];
        double T_h = (temp - T_t) / (T_m - T_t);
        double Et = 0.;
        double f = 1.0;
        if (strain_rate > eps_0) f = (1.0 + C * log(strain_rate / eps_0));
        if (plstrain > 0.) Et = n * B * pow(plstrain, n - 1.) * f * (1.0 - pow(T_h, m));
        else Et = Elastic().E();
        return Et;
    }
};

// Source/Material.cpp:48-57: Hollomon(el, sy0, K, m); eps0 = sy0/E, eps1 = (sy0/K)^(1/m), "bad definition" if eps0 > eps1
class Hollomon : public Material_ {
public:
    Hollomon() {}
    Hollomon(const Elastic_ & el, const double sy0_, const double & k_, const double & m_) : Material_(el)
    {
        Material_model = HOLLOMON; params[0] = sy0_; params[1] = k_; params[2] = m_;
        const double eps0 = sy0_ / el.E(), eps1 = std::pow(sy0_ / k_, 1. / m_);
        if (eps0 > eps1) throw new Fatal("ERROR, Hollomon material bad definition, please correct Yield Stress, Elastic Modulus or Material Constants.");
        eps0_m = eps0; eps1_m = eps1;
    }
    // Source/Material.cpp:59-73
    double CalcYieldStress() { return 0.0; }
    double CalcYieldStress(const double & strain)
    {
        const double sy0 = params[0], K = params[1], m = params[2];
        double sy;
        if (strain + eps0_m > eps1_m) sy = K * pow(strain + eps0_m, m);
        else sy = sy0;
        return sy;
    }
    double CalcTangentModulus(const double & strain)
    {
        const double K = params[1], m = params[2];
        double Et;
        if (strain + eps0_m > eps1_m) Et = K * m * pow(strain + eps0_m, (m - 1.0));
        else Et = 0.;
        return Et;
    }
private:
    double eps0_m = 0.0, eps1_m = 0.0;
};

// Source/Material.h:154-222, Material.cpp:96-161: GMT(el, n1, n2, C1, C2, m1, m2, I1, I2, e_min, e_max, er_min, er_max, T_min, T_max):
// sy = C1 exp(C2 T) e^(n1 T + n2) exp((I1 T + I2) / e) er^(m1 T + m2) with strain, strain rate and temperature clamped to their limits
class GMT : public Material_ {
public:
    GMT() {}
    GMT(const Elastic_ & el, const double & n1_, const double & n2_, const double & C1_, const double & C2_, const double & m1_, const double & m2_,
        const double & I1_, const double & I2_, const double & e_min_ = 0.0, const double & e_max_ = 1.0e10, const double & er_min_ = 0.0, const double & er_max_ = 1.0e10,
        const double & T_min_ = 0.0, const double & T_max_ = 1.0e10) : Material_(el)
    {
        Material_model = _GMT_;
        params[0] = n1_; params[1] = n2_; params[2] = C1_; params[3] = C2_; params[4] = m1_; params[5] = m2_; params[6] = I1_; params[7] = I2_;
        params[8] = e_min_; params[9] = e_max_; params[10] = er_min_; params[11] = er_max_; params[12] = T_min_; params[13] = T_max_;
    }
    double CalcYieldStress() { return 0.0; }
    double CalcYieldStress(const double &) { return 0.0; }
    double CalcYieldStress(const double & strain, const double & strain_rate, const double & temp)      // Material.cpp:96-134, same operand order
    {
        const double n1 = params[0], n2 = params[1], C1 = params[2], C2 = params[3], m1 = params[4], m2 = params[5], I1 = params[6], I2 = params[7];
        double e = strain, er = strain_rate, T = temp;
        if (e < params[8]) e = params[8]; else if (e > params[9]) e = params[9];
        if (er < params[10]) er = params[10]; else if (er > params[11]) er = params[11];
        if (T < params[12]) T = params[12]; else if (T > params[13]) T = params[13];
        return C1 * exp(C2 * T) * pow(e, n1 * T + n2) * exp((I1 * T + I2) / e) * pow(er, m1 * T + m2);
    }
    double CalcTangentModulus(const double & plstrain, const double & strain_rate, const double & temp)  // Material.cpp:136-161
    {
        const double n1 = params[0], n2 = params[1], C1 = params[2], C2 = params[3], m1 = params[4], m2 = params[5], I1 = params[6], I2 = params[7];
        double e = plstrain, er = strain_rate, T = temp;
        if (plstrain < params[8]) e = params[8]; else if (plstrain > params[9]) e = params[9];
        if (strain_rate < params[10]) er = params[10]; else if (strain_rate > params[11]) er = params[11];
        if (temp < params[12]) T = params[12]; else if (temp > params[13]) T = params[13];
        return C1 * exp(C2 * T) * pow(er, m1 * T + m2) * pow(e, T * n1 + n2 - 2.0) * (-I1 * T - I2 + e * (n1 * T + n2)) * exp((I1 * T + I2) / e);
    }
};

#endif
