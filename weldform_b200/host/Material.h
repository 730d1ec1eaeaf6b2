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
    double params[13] = { 0 };            // Hollomon: sy0, K, m.  Johnson-Cook: A, B, n, C, eps_0, m, T_m, T_t  (wf_material::p)
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

#endif
