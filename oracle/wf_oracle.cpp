// TEST INFRASTRUCTURE -- the CPU restatement ("port") of the reference's per-timestep SPH hot path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the product
// (weldform_b200/) never does. Every function cites the reference code it restates (paths relative to
// /root/reference). It is a restatement, not a copy: structure-of-arrays storage, one flat pair list, explicit
// loops; the arithmetic (operand order, association, branches) follows the reference expression by expression so
// that with FMA contraction off it reproduces oracle/_ref/wf_ref (the reference itself, Nproc=1) bit for bit.
//
// Parity status: PINNED against the reference run here (tests/test_oracle_vs_ref.py, tests/golden/*.wfc generated
// by tests/golden/make_golden.py from oracle/_ref/wf_ref). The reference ships no golden vectors of its own
// (SURVEY.md section 4). The 3x3 symmetric eigen-solve (GSL in the reference, External/matvec.h:776-804) is
// PARITY UNPINNED: GSL is not installed, oracle/shim/gsl/gsl_eigen.h substitutes a cyclic Jacobi iteration, and
// sym_eig3() below restates that same iteration.
//
// Scope: CellInitiate / ListGenerate / MainNeighbourSearch / AllocateNbPair,
// CalcPairPosList ordering, StartAcceleration (TIR), CalcDensInc+DensReduction, CalcRateTensors+RateTensorsReduction,
// Particle::CalcStressStrain (+Hollomon / Johnson-Cook), CalcAccel+AccelReduction, CalculateSurface,
// ContactNbSearch, CalcContactForcesWang, TriMesh::Update, position/velocity update, CalcPlasticWorkHeat,
// ThermalCalcs (CalcConvHeat, CalcTempInc), CalcKinEnergyEqn/CalcIntEnergyEqn, and the Fraser step order with its
// neighbour-rebuild cadence (Source/SolverFraser.cpp:133-505). Fixed particles (IsFree = false) are routed as the reference
// does (no SM pairs); their Shepard-smoothed TIR (PrimaryComputeAcceleration, Domain.cpp:1599-1698, read by nothing) is not
// restated. gradKernelCorr (Domain.cpp:1832-1940 + the locking branches of InteractionAlt.cpp) is restated, including the FSM pairs of fixed particles in CalcGradCorrMatrix. Ghost particles: the Mirror_XYZ rows of AddCylinderLength(..., ghost = true) with MoveGhost (Domain.cpp:795-826, 1142-1236). Not restated: the Symmetric-plane ghosts of AddCylSliceLength, damage, 2-D axisymmetric branches (all "next" or out of scope in SURVEY.md section 8).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <utility>
#include <vector>
#include <chrono>

#include "wfc_io.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace {

typedef std::vector<double> dvec;
typedef std::vector<int> ivec;

struct M3 { double m[9]; double & operator()(int r, int c) { return m[3 * r + c]; } double operator()(int r, int c) const { return m[3 * r + c]; } };
struct V3 { double v[3]; double & operator()(int i) { return v[i]; } double operator()(int i) const { return v[i]; } };

inline M3 zero3() { M3 r; for (int k = 0; k < 9; ++k) r.m[k] = 0.0; return r; }
// External/matvec.h:547 (double * Mat3_t): element * a
inline M3 smul(double a, const M3 & A) { M3 r; for (int k = 0; k < 9; ++k) r.m[k] = A.m[k] * a; return r; }
inline M3 madd(const M3 & A, const M3 & B) { M3 r; for (int k = 0; k < 9; ++k) r.m[k] = A.m[k] + B.m[k]; return r; }
inline M3 msub(const M3 & A, const M3 & B) { M3 r; for (int k = 0; k < 9; ++k) r.m[k] = A.m[k] - B.m[k]; return r; }
// External/matvec.h:845-850 Mult(Mat3,Mat3,Mat3): k runs 2,1,0
inline M3 mmul(const M3 & A, const M3 & B)
{ M3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = A(i, 2) * B(2, j) + A(i, 1) * B(1, j) + A(i, 0) * B(0, j); return r; }
inline M3 trans(const M3 & A) { M3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = A(j, i); return r; }
// External/matvec.h:829-834 Mult(Vec3,Mat3,Vec3): B = A^T M
inline V3 vmul(const V3 & A, const M3 & M) { V3 b; for (int j = 0; j < 3; ++j) b(j) = A(0) * M(0, j) + A(1) * M(1, j) + A(2) * M(2, j); return b; }
// Blitz dot meta-program: a0*b0 + (a1*b1 + a2*b2)   (oracle/shim/blitz/tinyvec-et.h)
inline double dot3(const V3 & a, const V3 & b) { return a(0) * b(0) + (a(1) * b(1) + a(2) * b(2)); }
inline double norm3(const V3 & a) { return std::sqrt(dot3(a, a)); }
inline V3 vsub(const V3 & a, const V3 & b) { V3 r; for (int k = 0; k < 3; ++k) r(k) = a(k) - b(k); return r; }
inline V3 vadd(const V3 & a, const V3 & b) { V3 r; for (int k = 0; k < 3; ++k) r(k) = a(k) + b(k); return r; }
inline V3 vscale(double s, const V3 & a) { V3 r; for (int k = 0; k < 3; ++k) r(k) = s * a(k); return r; }   // s*V and V*s commute bitwise
inline V3 vdiv(const V3 & a, double s) { V3 r; for (int k = 0; k < 3; ++k) r(k) = a(k) / s; return r; }
// External/matvec.h:1967-1973
inline V3 cross3(const V3 & a, const V3 & b) { V3 r; r(0) = a(1) * b(2) - a(2) * b(1); r(1) = a(2) * b(0) - a(0) * b(2); r(2) = a(0) * b(1) - a(1) * b(0); return r; }

// ---- Source/Functions.cpp:25-72 Kernel ------------------------------------------------------------------------
double Kernel(int Dim, int KT, double q, double h)
{
    double C;
    switch (KT) {
    case 0:
        Dim == 2 ? C = 10.0 / (7.0 * h * h * M_PI) : C = 1.0 / (h * h * h * M_PI);
        if (q < 1.0) return C * (1.0 - (3.0 / 2.0) * q * q + (3.0 / 4.0) * q * q * q);
        else if (q < 2.0) return C * ((1.0 / 4.0) * (2.0 - q) * (2.0 - q) * (2.0 - q));
        else return 0.0;
    case 1:
        Dim == 2 ? C = 7.0 / (4.0 * h * h * M_PI) : C = 7.0 / (8.0 * h * h * h * M_PI);
        if (q < 2.0) return C * std::pow((1.0 - q / 2.0), 4.0) * (2.0 * q + 1.0);
        else return 0.0;
    case 2:
        Dim == 2 ? C = 7.0 / (478.0 * h * h * M_PI) : C = 1.0 / (120.0 * h * h * h * M_PI);
        if (q < 1.0) return C * (std::pow((3.0 - q), 5.0) - 6.0 * std::pow((2.0 - q), 5.0) + 15.0 * std::pow((1.0 - q), 5.0));
        else if (q < 2.0) return C * (std::pow((3.0 - q), 5.0) - 6.0 * std::pow((2.0 - q), 5.0));
        else if (q < 3.0) return C * (std::pow((3.0 - q), 5.0));
        else return 0.0;
    case 3:
        Dim == 2 ? C = 1.0 / (3.0 * h * h * M_PI) : C = 15.0 / (62.0 * h * h * h * M_PI);
        if (q < 1.0) return C * (std::pow(q, 3.0) - 6.0 * q + 6.);
        else if (q < 2.0) return C * (std::pow((2.0 - q), 3.0));
        else return 0.0;
    }
    return 0.0;
}
// ---- Source/Functions.cpp:74-123 GradKernel -------------------------------------------------------------------
double GradKernel(int Dim, int KT, double q, double h)
{
    double C;
    switch (KT) {
    case 0:
        Dim == 2 ? C = 10.0 / (7.0 * h * h * h * M_PI) : C = 1.0 / (h * h * h * h * M_PI);
        if (q == 0.0) return C / h * (-3.0 + (9.0 / 2.0) * q);
        else if (q < 1.0) return C / (q * h) * (-3.0 * q + (9.0 / 4.0) * q * q);
        else if (q < 2.0) return C / (q * h) * ((-3.0 / 4.0) * (2.0 - q) * (2.0 - q));
        else return 0.0;
    case 1:
        Dim == 2 ? C = 7.0 / (4.0 * h * h * h * M_PI) : C = 7.0 / (8.0 * h * h * h * h * M_PI);
        if (q == 0.0) return C * -5.0 / h * (std::pow((1.0 - q / 2.0), 3.0) - 3.0 * q / 2.0 * std::pow((1.0 - q / 2.0), 2.0));
        else if (q < 2.0) return C / (q * h) * -5.0 * q * std::pow((1.0 - q / 2.0), 3.0);
        else return 0.0;
    case 2:
        Dim == 2 ? C = 7.0 / (478.0 * h * h * h * M_PI) : C = 1.0 / (120.0 * h * h * h * h * M_PI);
        if (q == 0.0) return C / h * (20.0 * std::pow((3.0 - q), 3.0) - 120.0 * std::pow((2.0 - q), 3.0) + 300.0 * std::pow((1.0 - q), 3.0));
        else if (q < 1.0) return C / (q * h) * (-5.0 * std::pow((3.0 - q), 4.0) + 30.0 * std::pow((2.0 - q), 4.0) - 75.0 * std::pow((1.0 - q), 4.0));
        else if (q < 2.0) return C / (q * h) * (-5.0 * std::pow((3.0 - q), 4.0) + 30.0 * std::pow((2.0 - q), 4.0));
        else if (q < 3.0) return C / (q * h) * (-5.0 * std::pow((3.0 - q), 4.0));
        else return 0.0;
    case 3:
        Dim == 2 ? C = 1.0 / (3.0 * h * h * h * M_PI) : C = 15.0 / (62.0 * h * h * h * h * M_PI);
        if (q < 1.0) return C / (q * h) * (3.0 * std::pow(q, 2.0) - 6.0);
        else if (q < 2.0) return C / (q * h) * (3.0 * std::pow((2.0 - q), 2.0) - 1.0);
        else return 0.0;
    }
    return 0.0;
}
// ---- Source/Functions.cpp:206-230 EOS, :232-256 SoundSpeed ----------------------------------------------------
double EOS(int EQ, double Cs0, double P00, double Density, double Density0)
{
    switch (EQ) {
    case 0: return P00 + (Cs0 * Cs0) * (Density - Density0);
    case 1: return P00 + (Density0 * Cs0 * Cs0 / 7.0) * (std::pow(Density / Density0, 7.0) - 1);
    case 2: return (Cs0 * Cs0) * Density;
    }
    return 0.0;
}
double SoundSpeed(int EQ, double Cs0, double Density, double Density0)
{
    switch (EQ) {
    case 0: return Cs0;
    case 1: return std::sqrt((Cs0 * Cs0) * std::pow(Density / Density0, 6.0));
    case 2: return Cs0;
    }
    return 0.0;
}

// ---- cyclic Jacobi, same operation sequence as oracle/shim/gsl/gsl_eigen.h (stand-in for gsl_eigen_symmv) -----
void sym_eig3(M3 a, double d[3], M3 & v)
{
    for (int i = 0; i < 3; ++i) for (int j = i + 1; j < 3; ++j) a(i, j) = a(j, i);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) v(i, j) = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 3; ++p) for (int q = p + 1; q < 3; ++q) off += std::fabs(a(p, q));
        if (off == 0.0) break;
        for (int p = 0; p < 3; ++p) for (int q = p + 1; q < 3; ++q) {
            double apq = a(p, q);
            if (apq == 0.0) continue;
            double app = a(p, p), aqq = a(q, q);
            double g = 100.0 * std::fabs(apq);
            if (std::fabs(app) + g == std::fabs(app) && std::fabs(aqq) + g == std::fabs(aqq)) { a(p, q) = 0.0; a(q, p) = 0.0; continue; }
            double theta = (aqq - app) / (2.0 * apq);
            double t = 1.0 / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
            if (theta < 0.0) t = -t;
            double c = 1.0 / std::sqrt(t * t + 1.0);
            double s = t * c;
            a(p, p) = app - t * apq; a(q, q) = aqq + t * apq; a(p, q) = 0.0; a(q, p) = 0.0;
            for (int r = 0; r < 3; ++r) {
                if (r == p || r == q) continue;
                double arp = a(r, p), arq = a(r, q);
                double nrp = c * arp - s * arq, nrq = s * arp + c * arq;
                a(r, p) = nrp; a(p, r) = nrp; a(r, q) = nrq; a(q, r) = nrq;
            }
            for (int r = 0; r < 3; ++r) {
                double vrp = v(r, p), vrq = v(r, q);
                v(r, p) = c * vrp - s * vrq; v(r, q) = s * vrp + c * vrq;
            }
        }
    }
    for (int i = 0; i < 3; ++i) d[i] = a(i, i);
    // ascending selection sort (Functions.cpp:384 calls Eig(...,SortAsc=true); shim gsl_eigen_symmv_sort)
    for (int i = 0; i < 2; ++i) {
        int k = i; double ek = d[i];
        for (int j = i + 1; j < 3; ++j) if (d[j] < ek) { k = j; ek = d[j]; }
        if (k != i) { std::swap(d[i], d[k]); for (int r = 0; r < 3; ++r) std::swap(v(r, i), v(r, k)); }
    }
}

// ---- materials (Source/Material.cpp:6-73) ---------------------------------------------------------------------
struct Mat { int model; double E, nu; double p[16]; double eps0, eps1; };
enum { BILINEAR = 0, HOLLOMON = 1, JOHNSON_COOK = 2, GMT_MODEL = 3 };

double mat_yield(const Mat & m, double strain, double strain_rate, double temp)
{
    if (m.model == HOLLOMON) {                         // Material.cpp:59-64; p = sy0, K, m
        double sy;
        if (strain + m.eps0 > m.eps1) sy = m.p[1] * std::pow(strain + m.eps0, m.p[2]);
        else sy = m.p[0];
        return sy;
    }
    if (m.model == GMT_MODEL) {                        // GMT::CalcYieldStress Material.cpp:96-134; p = n1,n2,C1,C2,m1,m2,I1,I2,e_min,e_max,er_min,er_max,T_min,T_max
        double n1 = m.p[0], n2 = m.p[1], C1 = m.p[2], C2 = m.p[3], m1 = m.p[4], m2 = m.p[5], I1 = m.p[6], I2 = m.p[7];
        double e = strain, er = strain_rate, T = temp;
        if (e < m.p[8]) e = m.p[8]; else if (e > m.p[9]) e = m.p[9];
        if (er < m.p[10]) er = m.p[10]; else if (er > m.p[11]) er = m.p[11];
        if (T < m.p[12]) T = m.p[12]; else if (T > m.p[13]) T = m.p[13];
        return C1 * std::exp(C2 * T) * std::pow(e, n1 * T + n2) * std::exp((I1 * T + I2) / e) * std::pow(er, m1 * T + m2);
    }
    // JohnsonCook::CalcYieldStress Material.cpp:6-25; p = A,B,n,C,eps_0,m,T_m,T_t
    double A = m.p[0], B = m.p[1], n = m.p[2], C = m.p[3], e0 = m.p[4], mm = m.p[5], Tm = m.p[6], Tt = m.p[7];
    double T_h = (temp - Tt) / (Tm - Tt);
    double f = 1.0;
    if (strain_rate > e0) f = (1.0 + C * std::log(strain_rate / e0));
    return (A + B * std::pow(strain, n)) * f * (1.0 - std::pow(T_h, mm));
}

// ---- rigid tool mesh (Source/Mesh.cpp) ------------------------------------------------------------------------
struct Mesh {
    std::vector<V3> node, node_v; std::vector<int> en;   // 3 node ids per element
    std::vector<V3> centroid, normal; dvec radius, pplane; ivec nfar;
    V3 m_v, m_w; int id; int first_particle; double hfac;
    double vel[3], tau;
    size_t nelem() const { return en.size() / 3; }
    // Mesh.cpp:254-271 CalcNormals (dimension 3)
    void calc_normals() {
        for (size_t e = 0; e < nelem(); ++e) {
            V3 u = vsub(node[en[3 * e + 1]], node[en[3 * e]]), v = vsub(node[en[3 * e + 2]], node[en[3 * e]]);
            V3 w = cross3(u, v); normal[e] = vdiv(w, norm3(w));
        }
    }
    // Mesh.cpp:70-80 CalcCentroids: sum of nodes / dimension
    void calc_centroids() {
        for (size_t e = 0; e < nelem(); ++e) {
            V3 c; c(0) = c(1) = c(2) = 0.0;
            for (int i = 0; i < 3; ++i) c = vadd(c, node[en[3 * e + i]]);
            centroid[e] = vdiv(c, 3.0);
        }
    }
    void update_pplane() { for (size_t e = 0; e < nelem(); ++e) pplane[e] = dot3(node[en[3 * e + nfar[e]]], normal[e]); }   // Mesh.cpp:222-228
    // Mesh.cpp:206-220 CalcSpheres (nfar always ends as the last node: reference quirk)
    void calc_spheres() {
        for (size_t e = 0; e < nelem(); ++e) {
            double mx = 0.0;
            for (int n = 0; n < 3; ++n) { V3 rv = vsub(node[en[3 * e + n]], centroid[e]); if (norm3(rv) > mx) mx = norm3(rv); nfar[e] = n; }
            radius[e] = mx;
        }
        update_pplane();
    }
    // Mesh.cpp:82-202 AxisPlaneMesh (dimension 3)
    void axis_plane(int axis, bool posit, V3 p1, V3 p2, int dens) {
        V3 p = vsub(p2, p1); int dir[3];
        if (axis == 0) { dir[0] = 1; dir[1] = 2; } else if (axis == 1) { dir[0] = 0; dir[1] = 2; } else { dir[0] = 0; dir[1] = 1; }
        dir[2] = axis;
        double x1, x2, x3 = p1(dir[2]); x2 = p1(dir[1]);
        double dl = p(dir[0]) / dens;
        for (int j = 0; j < dens + 1; ++j) {
            x1 = p1(dir[0]);
            for (int i = 0; i < dens + 1; ++i) {
                V3 v; v(dir[0]) = x1; v(dir[1]) = x2; v(dir[2]) = x3;
                node.push_back(v); V3 z; z(0) = z(1) = z(2) = 0.0; node_v.push_back(z);
                x1 += dl;
            }
            x2 += dl;
        }
        for (int j = 0; j < dens; ++j) for (int i = 0; i < dens; ++i) {
            int n0 = (dens + 1) * j + i, n1 = n0 + 1, n2 = (dens + 1) * (j + 1) + i, n3 = n2 + 1;
            int elcon[2][3];
            if (posit) { elcon[0][0] = n0; elcon[0][1] = n1; elcon[0][2] = n2; elcon[1][0] = n1; elcon[1][1] = n3; elcon[1][2] = n2; }
            else       { elcon[0][0] = n0; elcon[0][1] = n2; elcon[0][2] = n1; elcon[1][0] = n1; elcon[1][1] = n2; elcon[1][2] = n3; }
            for (int e = 0; e < 2; ++e) {
                for (int k = 0; k < 3; ++k) en.push_back(elcon[e][k]);
                V3 c = vdiv(vadd(vadd(node[elcon[e][0]], node[elcon[e][1]]), node[elcon[e][2]]), 3.);
                centroid.push_back(c);
            }
        }
        size_t ne = nelem(); normal.resize(ne); radius.resize(ne); pplane.resize(ne); nfar.resize(ne);
        calc_normals();
        if (posit) for (size_t e = 0; e < ne; ++e) normal[e] = vscale(-1.0, normal[e]);   // Mesh.cpp:186-189 (unary minus)
    }
    // Mesh.cpp:273-293 Update
    void update(double dt) {
        for (size_t n = 0; n < node.size(); ++n) {
            V3 vr = cross3(m_w, node[n]);
            node_v[n] = vadd(m_v, vr);
            node[n] = vadd(node[n], vscale(dt, node_v[n]));
        }
        calc_centroids(); calc_normals(); update_pplane();
    }
};

struct BC { int id; double v[3]; double tau; int set_a; };
inline double ramp(double t, double tau) { if (tau <= 0.0) return 1.0; return t < tau ? t / tau : 1.0; }

// ================================================================================================================
struct Dom {
    // parameters (Domain ctor defaults, Source/Domain.cpp:78-175)
    int Dimension = 3, KernelType = 0, GradientType = 0, ts_nb_inc = 5;
    double Cellfac = 2.0, deltat = 0.0, Time = 0.0;
    bool thermal_solver = false, pl_heating = false, contact = false, cont_heat_fric = false, cont_heat_cond = false;
    std::vector<std::pair<int, int> > GhostPairs;   // Domain.h:473 (inner particle, ghost particle)
    int solver = 0;                          // 0 SolveDiffUpdateFraser, 1 SolveDiffUpdateKickDrift, 2 SolveDiffUpdateLeapFrog, 3 ThermalSolve
    bool auto_ts_acc = false; double sqrt_h_a = 0.0025;   // Domain.cpp:109,127
    bool auto_ts = false; double CFL = 0.7, deltatint = 0.0, deltatmin = 0.0, manual_min_ts = 1.e-10;   // Domain.cpp:126,153,170
    long steps_done = 0;
    bool gradKernelCorr = false;             // Domain::gradKernelCorr: Bonet-Lok corrected gradient; the Fraser solver then sums with locks, in pair order (SolverFraser.cpp:112-115)
    int friction_function = 0; double friction_m = 0.0, friction_b = 0.0;   // Domain::friction_function (0 Constant, 1 Linear in T), Contact_Wang.cpp:243-244
    double pl_work_heat_frac = 0.9, friction_sta = 0.0, friction_dyn = 0.0, PFAC = 0.4, DFAC = 0.0, contact_force_factor = 1.0, contact_hc = 0.0;
    double DomMax[3] = { -100000000000.0, -100000000000.0, -100000000000.0 }, DomMin[3] = { 100000000000.0, 100000000000.0, 100000000000.0 };
    int solid_part_count = 0; int id_free_surf = 1;
    // per-particle state
    size_t N = 0;
    dvec x, v, a, Disp, normal, contforce;                       // 3N
    dvec Mass, Density, dDensity, RefDensity, h, Pressure, G, Cs, P0, Alpha, Beta, TI, TIn, TIInitDist, Sigmay, Ep;
    dvec pl_strain, delta_pl_strain, eff_strain_rate, T, dTdt, k_T, cp_T, h_conv, T_inf, th_exp, q_source, q_conv, q_plheat, q_fric_work, q_cont_conv;
    dvec delta_cont, dkin, dint;
    dvec Sigma, Shear, SRate, RRate, Strain, Strain_pl, Strain_pl_incr, TIR, gradCorrM;   // 9N
    ivec ID, ID_orig, IsFree, PresEq, Fail, Material, Material_model, mat_id, Thermal_BC, nowrite, Nb, pmesh, pelem;
    std::vector<Mat> mats; std::vector<Mesh> meshes; std::vector<BC> bcs;
    // cell grid
    double BLPF[3], TRPR[3], CellSize[3], hmax = 0.0; int CellNo[3]; ivec HOC, LL; bool grid_ready = false;
    // pair lists (Nproc = 1 discovery order)
    std::vector<std::pair<int, int> > SM, FSM, RIG, CONT;   // FSM: one particle fixed (Neighbour.cpp:410); read by CalcGradCorrMatrix only
    std::vector<ivec> ipairs, jpairs;   // per particle: pair indices with the particle as min / as max, discovery order
    bool nb_cleared = true;
    // scalars
    double kin_energy_sum = 0, int_energy_sum = 0, plastic_work = 0, contact_force_sum = 0, contact_reaction_sum = 0,
           ext_forces_work = 0, contact_friction_work = 0, accum_cont_heat_cond = 0, max_disp[3] = { 0, 0, 0 };
    std::vector<V3> m_contact_force;

    V3 X(size_t i) const { V3 r; r(0) = x[3 * i]; r(1) = x[3 * i + 1]; r(2) = x[3 * i + 2]; return r; }
    V3 Vv(size_t i) const { V3 r; r(0) = v[3 * i]; r(1) = v[3 * i + 1]; r(2) = v[3 * i + 2]; return r; }
    V3 A(size_t i) const { V3 r; r(0) = a[3 * i]; r(1) = a[3 * i + 1]; r(2) = a[3 * i + 2]; return r; }
    V3 Nrm(size_t i) const { V3 r; r(0) = normal[3 * i]; r(1) = normal[3 * i + 1]; r(2) = normal[3 * i + 2]; return r; }
    M3 getM(const dvec & f, size_t i) const { M3 r; std::memcpy(r.m, &f[9 * i], 72); return r; }
    void setM(dvec & f, size_t i, const M3 & m) { std::memcpy(&f[9 * i], m.m, 72); }
    void set3(dvec & f, size_t i, const V3 & q) { f[3 * i] = q(0); f[3 * i + 1] = q(1); f[3 * i + 2] = q(2); }

    void resize(size_t n) {
        N = n;
        dvec * v3[] = { &x, &v, &a, &Disp, &normal, &contforce };
        for (auto p : v3) p->assign(3 * n, 0.0);
        dvec * v1[] = { &Mass, &Density, &dDensity, &RefDensity, &h, &Pressure, &G, &Cs, &P0, &Alpha, &Beta, &TI, &TIn, &TIInitDist, &Sigmay, &Ep,
                        &pl_strain, &delta_pl_strain, &eff_strain_rate, &T, &dTdt, &k_T, &cp_T, &h_conv, &T_inf, &th_exp, &q_source, &q_conv, &q_plheat,
                        &q_fric_work, &q_cont_conv, &delta_cont, &dkin, &dint };
        for (auto p : v1) p->assign(n, 0.0);
        dvec * v9[] = { &Sigma, &Shear, &SRate, &RRate, &Strain, &Strain_pl, &Strain_pl_incr, &TIR, &gradCorrM };
        for (auto p : v9) p->assign(9 * n, 0.0);
        ivec * vi[] = { &ID, &ID_orig, &IsFree, &PresEq, &Fail, &Material, &Material_model, &mat_id, &Thermal_BC, &nowrite, &Nb, &pmesh, &pelem };
        for (auto p : vi) p->assign(n, 0);
        for (size_t i = 0; i < n; ++i) { IsFree[i] = 1; TIn[i] = 4.0; mat_id[i] = -1; pmesh[i] = -1; pelem[i] = -1; }   // Particle ctor, Particle.cpp:28-158
    }

    // ---- Source/Neighbour.cpp:10-139 CellInitiate (3-D, non-periodic) -----------------------------------------
    void CellInitiate() {
        for (int k = 0; k < 3; ++k) { BLPF[k] = x[k]; TRPR[k] = x[k]; }
        hmax = h[0];
        for (size_t i = 0; i < N; ++i) {
            for (int k = 0; k < 3; ++k) { if (x[3 * i + k] > TRPR[k]) TRPR[k] = x[3 * i + k]; if (x[3 * i + k] < BLPF[k]) BLPF[k] = x[3 * i + k]; }
            if (h[i] > hmax) hmax = h[i];
        }
        for (int k = 0; k < 3; ++k) { if (DomMax[k] > TRPR[k]) TRPR[k] = DomMax[k]; if (DomMin[k] < BLPF[k]) BLPF[k] = DomMin[k]; }
        for (int k = 0; k < 3; ++k) { TRPR[k] += hmax / 2; BLPF[k] -= hmax / 2; }
        for (int k = 0; k < 3; ++k) {
            double q = (TRPR[k] - BLPF[k]) / (Cellfac * hmax);
            if ((double)(std::ceil(q) - q) < (hmax / 10.0)) CellNo[k] = (int)std::ceil(q); else CellNo[k] = (int)std::floor(q);
            CellSize[k] = (TRPR[k] - BLPF[k]) / CellNo[k];
        }
        HOC.assign((size_t)CellNo[0] * CellNo[1] * CellNo[2], -1); LL.assign(N, -1);
        grid_ready = true;
    }
    size_t cidx(int i, int j, int k) const { return ((size_t)i * CellNo[1] + j) * CellNo[2] + k; }
    // ---- Neighbour.cpp:141-257 ListGenerate (3-D) ------------------------------------------------------------
    void ListGenerate() {
        for (size_t p = 0; p < N; ++p) {
            int c[3];
            for (int k = 0; k < 3; ++k) {
                c[k] = (int)(std::floor((x[3 * p + k] - BLPF[k]) / CellSize[k]));
                if (c[k] < 0) c[k] = 0;
                if (c[k] >= CellNo[k]) c[k] = CellNo[k] - 1;
            }
            size_t ci = cidx(c[0], c[1], c[2]);
            int temp = HOC[ci]; HOC[ci] = (int)p; LL[p] = temp;
        }
    }
    void CellReset() { std::fill(HOC.begin(), HOC.end(), -1); std::fill(LL.begin(), LL.end(), -1); }   // Neighbour.cpp:259-284
    // ---- Neighbour.cpp:361-372 CheckRadius -------------------------------------------------------------------
    bool CheckRadius(int p1, int p2) const {
        double hh = (h[p1] + h[p2]) / 2;
        V3 xij = vsub(X(p1), X(p2));
        double rij = norm3(xij);
        return (rij / hh) <= Cellfac;
    }
    // ---- Neighbour.cpp:374-417 AllocateNbPair ----------------------------------------------------------------
    void AllocateNbPair(int t1, int t2) {
        if (contact) {
            bool is_contact = false;
            for (size_t m = 0; m < meshes.size(); ++m) if (ID[t1] == meshes[m].id || ID[t2] == meshes[m].id) is_contact = true;
            if (is_contact) {
                if (ID[t1] == id_free_surf || ID[t2] == id_free_surf) RIG.push_back(std::make_pair(t1, t2));
                return;
            }
        }
        if (CheckRadius(t1, t2)) {
            if (IsFree[t1] || IsFree[t2]) {
                if (Material[t1] == Material[t2]) { if (IsFree[t1] && IsFree[t2]) SM.push_back(std::make_pair(t1, t2)); else FSM.push_back(std::make_pair(t1, t2)); }
            }
        }
    }
    // ---- Neighbour.cpp:346-359, 418-477 MainNeighbourSearch / YZPlaneCellsNeighbourSearch (Nproc = 1) ----------
    void MainNeighbourSearch() {
        for (int q1 = 0; q1 < CellNo[0]; ++q1)
        for (int q3 = 0; q3 < CellNo[2]; ++q3)
        for (int q2 = 0; q2 < CellNo[1]; ++q2) {
            if (HOC[cidx(q1, q2, q3)] == -1) continue;
            int t1 = HOC[cidx(q1, q2, q3)];
            while (t1 != -1) {
                int t2 = LL[t1];
                while (t2 != -1) { AllocateNbPair(t1, t2); t2 = LL[t2]; }
                if (q1 + 1 < CellNo[0]) { t2 = HOC[cidx(q1 + 1, q2, q3)]; while (t2 != -1) { AllocateNbPair(t1, t2); t2 = LL[t2]; } }
                if (q2 + 1 < CellNo[1]) for (int i = q1 - 1; i <= q1 + 1; ++i) if (i < CellNo[0] && i >= 0) {
                    t2 = HOC[cidx(i, q2 + 1, q3)]; while (t2 != -1) { AllocateNbPair(t1, t2); t2 = LL[t2]; }
                }
                if (q3 + 1 < CellNo[2]) for (int j = q2 - 1; j <= q2 + 1; ++j) for (int i = q1 - 1; i <= q1 + 1; ++i)
                    if (i < CellNo[0] && i >= 0 && j < CellNo[1] && j >= 0) {
                        t2 = HOC[cidx(i, j, q3 + 1)]; while (t2 != -1) { AllocateNbPair(t1, t2); t2 = LL[t2]; }
                    }
                t1 = LL[t1];
            }
        }
        nb_cleared = false;
    }
    void ClearNbData() { SM.clear(); FSM.clear(); RIG.clear(); CONT.clear(); CellReset(); ListGenerate(); nb_cleared = true; }   // Neighbour.cpp:479-497
    // ---- Neighbour.cpp:617-694 CalcPairPosList: per particle, pairs as min (front) and as max (back), discovery order
    void CalcPairPosList() {
        ipairs.assign(N, ivec()); jpairs.assign(N, ivec());
        for (size_t p = 0; p < SM.size(); ++p) {
            int i = std::min(SM[p].first, SM[p].second), j = std::max(SM[p].first, SM[p].second);
            ipairs[i].push_back((int)p); jpairs[j].push_back((int)p);
        }
    }
    void SaveNeighbourData() { std::fill(Nb.begin(), Nb.end(), 0); for (size_t p = 0; p < SM.size(); ++p) { Nb[SM[p].first] += 1; Nb[SM[p].second] += 1; } }   // Neighbour.cpp:499-532

    int first_fem() const { return meshes.empty() ? (int)N : meshes[0].first_particle; }

    // ---- Source/Domain.cpp:1415-1491 CalculateSurface -----------------------------------------------------------
    void CalculateSurface(int id) {
        id_free_surf = id;
        double totmass = 0.;
        for (size_t i = 0; i < N; ++i) totmass += Mass[i];
        totmass /= N;
        int maxid = contact ? first_fem() : (int)N;
        for (int i = 0; i < maxid; ++i) { normal[3 * i] = normal[3 * i + 1] = normal[3 * i + 2] = 0.; ID[i] = ID_orig[i]; }
        for (size_t p = 0; p < SM.size(); ++p) {
            int P1 = SM[p].first, P2 = SM[p].second;
            V3 xij = vsub(X(P1), X(P2));
            double mi = Mass[P1], mj = Mass[P2];
            for (int k = 0; k < 3; ++k) normal[3 * P1 + k] += mj * xij(k);
            for (int k = 0; k < 3; ++k) normal[3 * P2 + k] -= mi * xij(k);
        }
        int max_nb = 46; if (Dimension == 2) max_nb = 12;
        for (int i = 0; i < maxid; ++i) {
            for (int k = 0; k < 3; ++k) normal[3 * i + k] *= 1. / totmass;
            if (norm3(Nrm(i)) >= 0.25 * h[i] && Nb[i] <= max_nb) { if (!nowrite[i]) ID[i] = id; }
        }
    }
    // ---- Source/Contact.cpp:57-132 ContactNbSearch ------------------------------------------------------------
    void ContactNbSearch() {
        for (size_t p = 0; p < RIG.size(); ++p) {
            int P1 = RIG[p].first, P2 = RIG[p].second;
            bool is_contact = false;
            for (size_t m = 0; m < meshes.size(); ++m) if (ID[P1] == meshes[m].id || ID[P2] == meshes[m].id) is_contact = true;
            if (is_contact) if (ID[P1] == id_free_surf || ID[P2] == id_free_surf)
                if (norm3(vsub(X(P1), X(P2))) < (h[P1] + h[P2])) CONT.push_back(RIG[p]);
        }
        RIG.clear();
    }
    void ContactNbUpdate() { CalculateSurface(1); ContactNbSearch(); }   // Domain.cpp:2481-2487

    // ---- Source/Domain.cpp:1519-1597 StartAcceleration (3-D) ---------------------------------------------------
    void StartAcceleration() {
        for (size_t i = 0; i < N; ++i) {
            double rho2 = Density[i] * Density[i];
            if (IsFree[i]) {
                if (TI[i] > 0.0) {
                    double pc = TI[i] / (rho2);
                    double d[3]; M3 Vec; sym_eig3(getM(Sigma, i), d, Vec);
                    M3 VecT = trans(Vec), Val = zero3();
                    for (int k = 0; k < 3; ++k) Val(k, k) = d[k];
                    for (int k = 0; k < 3; ++k) { if (Val(k, k) > 0) Val(k, k) = -pc * Val(k, k); else Val(k, k) = 0.0; }
                    M3 temp = mmul(Vec, Val);
                    setM(TIR, i, mmul(temp, VecT));
                }
            }
            a[3 * i] = a[3 * i + 1] = a[3 * i + 2] = 0.0;
            dDensity[i] = 0.0;
            setM(SRate, i, zero3()); setM(RRate, i, zero3());
        }
    }

    // common pair geometry: P1 = min, P2 = max as the nonlock_sum branch does (InteractionAlt.cpp:31-32)
    // ---- Source/InteractionAlt.cpp:483-611 CalcDensInc (per-pair term) + :613-640 DensReduction ------------------
    void Density_pass() {
        dvec pd(SM.size());
        for (size_t p = 0; p < SM.size(); ++p) {
            int P1 = std::min(SM[p].first, SM[p].second), P2 = std::max(SM[p].first, SM[p].second);
            double hh = (h[P1] + h[P2]) / 2;
            V3 xij = vsub(X(P1), X(P2)); double rij = norm3(xij);
            V3 vij = vsub(Vv(P1), Vv(P2));
            double GK = GradKernel(Dimension, KernelType, rij / hh, hh);
            pd[p] = dot3(vij, vscale(GK, xij));
        }
        for (int i = 0; i < solid_part_count; ++i) {
            dDensity[i] = 0.;
            for (size_t n = 0; n < ipairs[i].size(); ++n) { int p = ipairs[i][n]; int j = std::max(SM[p].first, SM[p].second); dDensity[i] += Mass[j] / Density[j] * pd[p]; }
            for (size_t n = 0; n < jpairs[i].size(); ++n) { int p = jpairs[i][n]; int j = std::min(SM[p].first, SM[p].second); dDensity[i] += Mass[j] / Density[j] * pd[p]; }
            dDensity[i] *= Density[i];
        }
    }
    // ---- InteractionAlt.cpp:263-446 CalcRateTensors + :450-479 RateTensorsReduction ---------------------------
    void RateTensors_pass() {
        std::vector<M3> pS(SM.size()), pR(SM.size());
        for (size_t p = 0; p < SM.size(); ++p) {
            int P1 = std::min(SM[p].first, SM[p].second), P2 = std::max(SM[p].first, SM[p].second);
            double hh = (h[P1] + h[P2]) / 2;
            V3 xij = vsub(X(P1), X(P2)); double rij = norm3(xij);
            V3 vab = vsub(Vv(P1), Vv(P2));
            double GK = GradKernel(Dimension, KernelType, rij / hh, hh);
            M3 S = zero3(), R = zero3();
            S(0, 0) = 2.0 * vab(0) * xij(0);
            S(0, 1) = vab(0) * xij(1) + vab(1) * xij(0);
            S(0, 2) = vab(0) * xij(2) + vab(2) * xij(0);
            S(1, 0) = S(0, 1);
            S(1, 1) = 2.0 * vab(1) * xij(1);
            S(1, 2) = vab(1) * xij(2) + vab(2) * xij(1);
            S(2, 0) = S(0, 2); S(2, 1) = S(1, 2);
            S(2, 2) = 2.0 * vab(2) * xij(2);
            S = smul(-0.5 * GK, S);
            R(0, 1) = vab(0) * xij(1) - vab(1) * xij(0);
            R(0, 2) = vab(0) * xij(2) - vab(2) * xij(0);
            R(1, 2) = vab(1) * xij(2) - vab(2) * xij(1);
            R(1, 0) = -R(0, 1); R(2, 0) = -R(0, 2); R(2, 1) = -R(1, 2);
            R = smul(-0.5 * GK, R);
            pS[p] = S; pR[p] = R;
        }
        for (int i = 0; i < solid_part_count; ++i) {
            M3 S = getM(SRate, i), R = getM(RRate, i);
            for (size_t n = 0; n < ipairs[i].size(); ++n) {
                int p = ipairs[i][n]; int j = std::max(SM[p].first, SM[p].second); double mjdj = Mass[j] / Density[j];
                S = madd(S, smul(mjdj, pS[p])); R = madd(R, smul(mjdj, pR[p]));
            }
            for (size_t n = 0; n < jpairs[i].size(); ++n) {
                int p = jpairs[i][n]; int j = std::min(SM[p].first, SM[p].second); double mjdj = Mass[j] / Density[j];
                S = madd(S, smul(mjdj, pS[p])); R = madd(R, smul(mjdj, pR[p]));
            }
            setM(SRate, i, S); setM(RRate, i, R);
        }
    }
    // ================= gradKernelCorr = true: locking summation, pairs as stored (first, second), pair order =================
    // External/matvec.h:836-841 Mult(Mat3,Vec3,Vec3): B = M A
    static V3 mvmul(const M3 & M, const V3 & A) { V3 b; for (int i = 0; i < 3; ++i) b(i) = A(0) * M(i, 0) + A(1) * M(i, 1) + A(2) * M(i, 2); return b; }
    // ---- Source/Domain.cpp:1832-1940 CalcGradCorrMatrix (SM pairs, then the FSM pairs with one fixed particle)
    void CalcGradCorrMatrix() {
        std::vector<M3> temp(N, zero3());
        for (size_t p = 0; p < SM.size(); ++p) {
            int P1 = SM[p].first, P2 = SM[p].second;
            V3 xij = vsub(X(P1), X(P2));
            double hh = (h[P1] + h[P2]) / 2.0;
            double GK = GradKernel(Dimension, KernelType, norm3(xij) / hh, hh);
            double di = Density[P1], mi = Mass[P1], dj = Density[P2], mj = Mass[P2];
            V3 gx = vscale(GK, xij);
            M3 m;                                     // Dyad(GK*xij, xij), matvec.h:813-818
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) m(r, c) = gx(r) * xij(c);
            temp[P1] = msub(temp[P1], smul(mj / dj, m));
            temp[P2] = msub(temp[P2], smul(mi / di, m));
        }
        for (size_t p = 0; p < FSM.size(); ++p) {          // Domain.cpp:1870-1898: the pairs with one fixed particle, same term
            int P1 = FSM[p].first, P2 = FSM[p].second;
            V3 xij = vsub(X(P1), X(P2));
            double hh = (h[P1] + h[P2]) / 2.0;
            double GK = GradKernel(Dimension, KernelType, norm3(xij) / hh, hh);
            double di = Density[P1], mi = Mass[P1], dj = Density[P2], mj = Mass[P2];
            V3 gx = vscale(GK, xij);
            M3 m;
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) m(r, c) = gx(r) * xij(c);
            temp[P1] = msub(temp[P1], smul(mj / dj, m));
            temp[P2] = msub(temp[P2], smul(mi / di, m));
        }
        int max_id = contact ? first_fem() : (int)N;
        M3 I = zero3(); I(0, 0) = I(1, 1) = I(2, 2) = 1.0;
        for (int i = 0; i < max_id; ++i) {
            if (IsFree[i]) {
                const M3 & M = temp[i]; M3 Mi;       // InvLog, matvec.h:673-693
                double det = M(0, 0) * (M(1, 1) * M(2, 2) - M(1, 2) * M(2, 1)) - M(0, 1) * (M(1, 0) * M(2, 2) - M(1, 2) * M(2, 0)) + M(0, 2) * (M(1, 0) * M(2, 1) - M(1, 1) * M(2, 0));
                bool ok = !(std::fabs(det) < 1.0e-10);
                Mi(0, 0) = (M(1, 1) * M(2, 2) - M(1, 2) * M(2, 1)) / det; Mi(0, 1) = (M(0, 2) * M(2, 1) - M(0, 1) * M(2, 2)) / det; Mi(0, 2) = (M(0, 1) * M(1, 2) - M(0, 2) * M(1, 1)) / det;
                Mi(1, 0) = (M(1, 2) * M(2, 0) - M(1, 0) * M(2, 2)) / det; Mi(1, 1) = (M(0, 0) * M(2, 2) - M(0, 2) * M(2, 0)) / det; Mi(1, 2) = (M(0, 2) * M(1, 0) - M(0, 0) * M(1, 2)) / det;
                Mi(2, 0) = (M(1, 0) * M(2, 1) - M(1, 1) * M(2, 0)) / det; Mi(2, 1) = (M(0, 1) * M(2, 0) - M(0, 0) * M(2, 1)) / det; Mi(2, 2) = (M(0, 0) * M(1, 1) - M(0, 1) * M(1, 0)) / det;
                setM(gradCorrM, i, ok ? Mi : I);
            } else setM(gradCorrM, i, I);
        }
    }
    // ---- InteractionAlt.cpp:483-611 CalcDensInc, !nonlock_sum && gradKernelCorr branch ----------------------------
    void Density_lock() {
        for (size_t p = 0; p < SM.size(); ++p) {
            int P1 = SM[p].first, P2 = SM[p].second;
            double hh = (h[P1] + h[P2]) / 2;
            V3 xij = vsub(X(P1), X(P2)); double rij = norm3(xij);
            double di = Density[P1], mi = Mass[P1], dj = Density[P2], mj = Mass[P2];
            V3 vij = vsub(Vv(P1), Vv(P2));
            double GK = GradKernel(Dimension, KernelType, rij / hh, hh);
            V3 vc0 = mvmul(smul(GK, getM(gradCorrM, P1)), xij), vc1 = mvmul(smul(GK, getM(gradCorrM, P2)), xij);
            double t0 = dot3(vij, vc0), t1 = dot3(vij, vc1);
            dDensity[P1] += 1.0 * mj * (di / dj) * t0;
            dDensity[P2] += 1.0 * mi * (dj / di) * t1;
        }
    }
    // ---- InteractionAlt.cpp:263-446 CalcRateTensors, locking branch with corrected gradients (float m/rho, :416,432) ---
    void RateTensors_lock() {
        for (size_t p = 0; p < SM.size(); ++p) {
            int P1 = SM[p].first, P2 = SM[p].second;
            double hh = (h[P1] + h[P2]) / 2;
            V3 xij = vsub(X(P1), X(P2)); double rij = norm3(xij);
            double di = Density[P1], mi = Mass[P1], dj = Density[P2], mj = Mass[P2];
            V3 vab = vsub(Vv(P1), Vv(P2));
            double GK = GradKernel(Dimension, KernelType, rij / hh, hh);
            M3 Sc[2], Rc[2];
            for (int k = 0; k < 2; ++k) {
                V3 gradK = mvmul(smul(GK, getM(gradCorrM, k == 0 ? P1 : P2)), xij);
                M3 gv, gvT;                           // Dyad(vab, gradK)
                for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) gv(r, c) = vab(r) * gradK(c);
                gvT = trans(gv);
                Sc[k] = smul(-0.5, madd(gv, gvT)); Rc[k] = smul(-0.5, msub(gv, gvT));
            }
            float mj_dj = mj / dj;
            setM(SRate, P1, madd(getM(SRate, P1), smul(mj_dj, Sc[0]))); setM(RRate, P1, madd(getM(RRate, P1), smul(mj_dj, Rc[0])));
            float mi_di = mi / di;
            setM(SRate, P2, madd(getM(SRate, P2), smul(mi_di, Sc[1]))); setM(RRate, P2, madd(getM(RRate, P2), smul(mi_di, Rc[1])));
        }
    }
    // ---- InteractionAlt.cpp:12-235 CalcAccel, locking branch with corrected gradients -------------------------------
    void Accel_lock() {
        M3 I = zero3(); I(0, 0) = I(1, 1) = I(2, 2) = 1.0;
        for (size_t p = 0; p < SM.size(); ++p) {
            int P1 = SM[p].first, P2 = SM[p].second;
            double hh = (h[P1] + h[P2]) / 2;
            V3 xij = vsub(X(P1), X(P2)); double rij = norm3(xij);
            double Alpha_ = (Alpha[P1] + Alpha[P2]) / 2.0, Beta_ = (Beta[P1] + Beta[P2]) / 2.0;
            double di = Density[P1], mi = Mass[P1], dj = Density[P2], mj = Mass[P2];
            V3 vij = vsub(Vv(P1), Vv(P2));
            double GK = GradKernel(Dimension, KernelType, rij / hh, hh);
            double K = Kernel(Dimension, KernelType, rij / hh, hh);
            M3 PIij = zero3();
            if (Alpha_ != 0.0 || Beta_ != 0.0) {
                double MUij = hh * dot3(vij, xij) / (rij * rij + 0.01 * hh * hh);
                double Ci = SoundSpeed(PresEq[P1], Cs[P1], di, RefDensity[P1]);
                double Cj = SoundSpeed(PresEq[P2], Cs[P2], dj, RefDensity[P2]);
                double Cij = 0.5 * (Ci + Cj);
                if (dot3(vij, xij) < 0) PIij = smul((Alpha_ * Cij * MUij + Beta_ * MUij * MUij) / (0.5 * (di + dj)), I);
            }
            M3 Sigmai = getM(Sigma, P1), Sigmaj = getM(Sigma, P2);
            M3 TIij = zero3();
            if (TI[P1] > 0.0 || TI[P2] > 0.0)
                TIij = smul(std::pow((K / Kernel(Dimension, KernelType, (TIInitDist[P1] + TIInitDist[P2]) / (2.0 * hh), hh)), (TIn[P1] + TIn[P2]) / 2.0),
                            madd(getM(TIR, P1), getM(TIR, P2)));
            M3 Msum;
            if (GradientType == 0) Msum = madd(madd(madd(smul(1.0 / (di * di), Sigmai), smul(1.0 / (dj * dj), Sigmaj)), PIij), TIij);
            else if (GradientType == 1) Msum = madd(madd(smul(1.0 / (di * dj), madd(Sigmai, Sigmaj)), PIij), TIij);
            else Msum = madd(madd(smul(1.0 / (di * dj), msub(Sigmai, Sigmaj)), PIij), TIij);
            V3 vc0 = mvmul(smul(GK, getM(gradCorrM, P1)), xij), vc1 = mvmul(smul(GK, getM(gradCorrM, P2)), xij);
            V3 t0 = vmul(vc0, Msum), t1 = vmul(vc1, Msum);
            for (int k = 0; k < 3; ++k) { a[3 * P1 + k] += (1.0 * mj) * t0(k); a[3 * P2 + k] -= (1.0 * mi) * t1(k); }
        }
    }
    // ---- Source/Particle.cpp:960-1067 CalcStressStrain --------------------------------------------------------
    void CalcStressStrain(size_t i, double dt) {
        double rho = Density[i];
        Pressure[i] = EOS(PresEq[i], Cs[i], P0[i], rho, RefDensity[i]);
        M3 RR = getM(RRate, i), SS = getM(Shear, i), SR = getM(SRate, i);
        M3 RRT = trans(RR), SRT = mmul(SS, RRT), RS = mmul(RR, SS);
        double dep = 0., sig_trial = 0.;
        M3 I = zero3(); I(0, 0) = I(1, 1) = I(2, 2) = 1.0;
        // ShearStress = dt*(2.0*G*(StrainRate-1.0/3.0*(tr)*I)+SRT+RS) + ShearStress
        double tr = SR(0, 0) + SR(1, 1) + SR(2, 2);
        M3 dev = msub(SR, smul(1.0 / 3.0 * tr, I));
        SS = madd(smul(dt, madd(madd(smul(2.0 * G[i], dev), SRT), RS)), SS);
        eff_strain_rate[i] = std::sqrt(0.5 * ((SR(0, 0) - SR(1, 1)) * (SR(0, 0) - SR(1, 1)) +
                                               (SR(1, 1) - SR(2, 2)) * (SR(1, 1) - SR(2, 2)) +
                                               (SR(2, 2) - SR(0, 0)) * (SR(2, 2) - SR(0, 0))) +
                                        3.0 * (SR(0, 1) * SR(0, 1) + SR(1, 2) * SR(1, 2) + SR(2, 0) * SR(2, 0)));
        if (Fail[i] == 1) {
            double J2 = 0.5 * (SS(0, 0) * SS(0, 0) + 2.0 * SS(0, 1) * SS(1, 0) + 2.0 * SS(0, 2) * SS(2, 0) + SS(1, 1) * SS(1, 1) +
                               2.0 * SS(1, 2) * SS(2, 1) + SS(2, 2) * SS(2, 2));
            sig_trial = std::sqrt(3.0 * J2);
            SS = smul(std::min((Sigmay[i] / sig_trial), 1.0), SS);
            if (Material_model[i] == HOLLOMON) Sigmay[i] = mat_yield(mats[mat_id[i]], pl_strain[i], 0.0, 0.0);
            else if (Material_model[i] == JOHNSON_COOK || Material_model[i] == GMT_MODEL) Sigmay[i] = mat_yield(mats[mat_id[i]], pl_strain[i], eff_strain_rate[i], T[i]);
            if (sig_trial > Sigmay[i]) {
                // Et is computed and discarded for Hollomon/JC; Ep keeps its initial value
                if (Ep[i] < 0) Ep[i] = 1. * mats[mat_id[i]].E;
                dep = (sig_trial - Sigmay[i]) / (3. * G[i] + Ep[i]);
                pl_strain[i] += dep;
                delta_pl_strain[i] = dep;
                if (Material_model[i] == BILINEAR) Sigmay[i] += dep * Ep[i];
            }
        }
        M3 Sg = madd(smul(-Pressure[i], I), SS);
        setM(Shear, i, SS); setM(Sigma, i, Sg);
        if (dep > 0.0) {
            double f = dep / Sigmay[i];
            M3 inc = getM(Strain_pl_incr, i);
            inc(0, 0) = f * (Sg(0, 0) - 0.5 * (Sg(1, 1) + Sg(2, 2)));
            inc(1, 1) = f * (Sg(1, 1) - 0.5 * (Sg(0, 0) + Sg(2, 2)));
            inc(2, 2) = f * (Sg(2, 2) - 0.5 * (Sg(0, 0) + Sg(1, 1)));
            inc(0, 1) = inc(1, 0) = 1.5 * f * (Sg(0, 1));
            inc(0, 2) = inc(2, 0) = 1.5 * f * (Sg(0, 2));
            inc(1, 2) = inc(2, 1) = 1.5 * f * (Sg(1, 2));
            setM(Strain_pl_incr, i, inc);
            setM(Strain_pl, i, madd(getM(Strain_pl, i), inc));
        }
        setM(Strain, i, madd(smul(dt, SR), getM(Strain, i)));
    }
    // ---- InteractionAlt.cpp:12-235 CalcAccel + :237-260 AccelReduction ------------------------------------------
    void Accel_pass() {
        std::vector<V3> pf(SM.size());
        M3 I = zero3(); I(0, 0) = I(1, 1) = I(2, 2) = 1.0;
        for (size_t p = 0; p < SM.size(); ++p) {
            int P1 = std::min(SM[p].first, SM[p].second), P2 = std::max(SM[p].first, SM[p].second);
            double hh = (h[P1] + h[P2]) / 2;
            V3 xij = vsub(X(P1), X(P2)); double rij = norm3(xij);
            double Alpha_ = (Alpha[P1] + Alpha[P2]) / 2.0, Beta_ = (Beta[P1] + Beta[P2]) / 2.0;
            double di = Density[P1], dj = Density[P2];
            V3 vij = vsub(Vv(P1), Vv(P2));
            double GK = GradKernel(Dimension, KernelType, rij / hh, hh);
            double K = Kernel(Dimension, KernelType, rij / hh, hh);
            M3 PIij = zero3();
            if (Alpha_ != 0.0 || Beta_ != 0.0) {
                double MUij = hh * dot3(vij, xij) / (rij * rij + 0.01 * hh * hh);
                double Ci = SoundSpeed(PresEq[P1], Cs[P1], di, RefDensity[P1]);
                double Cj = SoundSpeed(PresEq[P2], Cs[P2], dj, RefDensity[P2]);
                double Cij = 0.5 * (Ci + Cj);
                if (dot3(vij, xij) < 0) PIij = smul((Alpha_ * Cij * MUij + Beta_ * MUij * MUij) / (0.5 * (di + dj)), I);
            }
            M3 Sigmai = getM(Sigma, P1), Sigmaj = getM(Sigma, P2);
            M3 TIij = zero3();
            if (TI[P1] > 0.0 || TI[P2] > 0.0)
                TIij = smul(std::pow((K / Kernel(Dimension, KernelType, (TIInitDist[P1] + TIInitDist[P2]) / (2.0 * hh), hh)), (TIn[P1] + TIn[P2]) / 2.0),
                            madd(getM(TIR, P1), getM(TIR, P2)));
            V3 gx = vscale(GK, xij), temp;
            if (GradientType == 0)
                temp = vmul(gx, madd(madd(madd(smul(1.0 / (di * di), Sigmai), smul(1.0 / (dj * dj), Sigmaj)), PIij), TIij));
            else if (GradientType == 1)
                temp = vmul(gx, madd(madd(smul(1.0 / (di * dj), madd(Sigmai, Sigmaj)), PIij), TIij));
            else
                temp = vmul(gx, madd(madd(smul(1.0 / (di * dj), msub(Sigmai, Sigmaj)), PIij), TIij));
            if (Dimension == 2) temp(2) = 0.0;
            pf[p] = temp;
        }
        for (int i = 0; i < solid_part_count; ++i) { a[3 * i] = a[3 * i + 1] = a[3 * i + 2] = 0.; }
        for (int i = 0; i < solid_part_count; ++i) {
            for (size_t n = 0; n < ipairs[i].size(); ++n) { int p = ipairs[i][n]; int j = std::max(SM[p].first, SM[p].second); for (int k = 0; k < 3; ++k) a[3 * i + k] += Mass[j] * pf[p](k); }
            for (size_t n = 0; n < jpairs[i].size(); ++n) { int p = jpairs[i][n]; int j = std::min(SM[p].first, SM[p].second); for (int k = 0; k < 3; ++k) a[3 * i + k] -= Mass[j] * pf[p](k); }
        }
    }
    // ---- Source/Contact_Wang.cpp:15-348 CalcContactForcesWang (3-D mesh) ----------------------------------------
    void ContactWang() {
        for (size_t i = 0; i < N; ++i) {
            contforce[3 * i] = contforce[3 * i + 1] = contforce[3 * i + 2] = 0.; delta_cont[i] = 0.; q_fric_work[i] = 0.; q_cont_conv[i] = 0.;
        }
        for (size_t m = 0; m < meshes.size(); ++m) { m_contact_force[m](0) = m_contact_force[m](1) = m_contact_force[m](2) = 0.0; }
        double ext_forces_work_step = 0.;
        contact_force_sum = 0.; contact_reaction_sum = 0.;
        for (size_t c = 0; c < CONT.size(); ++c) {
            bool is_first = false;
            for (size_t m = 0; m < meshes.size(); ++m) if (ID[CONT[c].first] == meshes[m].id) is_first = true;
            int P1, P2;
            if (is_first) { P1 = CONT[c].second; P2 = CONT[c].first; } else { P1 = CONT[c].first; P2 = CONT[c].second; }
            V3 vr = vsub(Vv(P1), Vv(P2));
            V3 n2 = Nrm(P2);
            double delta_ = -dot3(n2, vr);
            int m = pmesh[P2]; const Mesh & ms = meshes[m]; int e = pelem[P2];
            V3 x_pred = vadd(vadd(X(P1), vscale(deltat, Vv(P1))), vdiv(vscale(deltat, vscale(deltat, A(P1))), 2.0));
            double dist = dot3(n2, x_pred) - ms.pplane[e];
            if (dist < h[P1]) {
                V3 qj = vsub(X(P1), vscale(dist, n2));
                bool inside = true; int i = 0;
                while (i < 3 && inside) {
                    int j = i + 1; if (j > 2) j = 0;
                    double crit = dot3(cross3(vsub(ms.node[ms.en[3 * e + j]], ms.node[ms.en[3 * e + i]]), vsub(qj, ms.node[ms.en[3 * e + i]])), n2);
                    if (crit < 0.0) inside = false;
                    i++;
                }
                if (inside) {
                    double delta = h[P1] - dist;
                    double kij = 2.0 * Mass[P1] / (deltat * deltat);
                    double omega = std::sqrt(kij / Mass[P1]);
                    double psi_cont = 2. * Mass[P1] * omega * DFAC;
                    V3 cf = vscale((kij * delta - psi_cont * delta_), n2);
                    set3(contforce, P1, cf); delta_cont[P1] = delta;
                    V3 aP = vadd(A(P1), vdiv(cf, Mass[P1])); set3(a, P1, aP);
                    double fr_sta = friction_sta, fr_dyn = friction_dyn;
                    if (friction_function == 1) fr_sta = fr_dyn = friction_m * T[P1] + friction_b;
                    double dens = Density[P1];
                    double dS2 = std::pow(Mass[P1] / dens, 2.0 / 3.0);
                    if (fr_sta > 0.) {
                        V3 du = vsub(vsub(x_pred, X(P1)), vscale(deltat, Vv(P2)));
                        V3 delta_tg = vsub(du, vscale(dot3(du, n2), n2));
                        V3 tgforce = vscale(kij, delta_tg);
                        if (norm3(tgforce) < fr_sta * norm3(cf)) {
                            cf = vsub(cf, vdiv(tgforce, Mass[P1])); set3(contforce, P1, cf);       // unit quirk kept
                            aP = vsub(aP, vdiv(tgforce, Mass[P1])); set3(a, P1, aP);
                        } else {
                            V3 tgforce_dyn = vdiv(vscale(fr_dyn * norm3(cf), tgforce), norm3(tgforce));
                            cf = vsub(cf, tgforce_dyn); set3(contforce, P1, cf);
                            aP = vsub(aP, vdiv(tgforce_dyn, Mass[P1])); set3(a, P1, aP);
                            if (cont_heat_fric) {
                                double abs_fv = std::fabs(dot3(tgforce_dyn, vr));
                                q_fric_work[P1] = abs_fv * dens / Mass[P1];
                            }
                        }
                    }
                    contact_force_sum += norm3(cf);
                    for (int k = 0; k < 3; ++k) m_contact_force[m](k) += cf(k);
                    contact_reaction_sum += dot3(aP, n2) * Mass[P1];
                    ext_forces_work_step += dot3(cf, Vv(P2));
                    if (cont_heat_cond) q_cont_conv[P1] = dens * contact_hc * dS2 * (T[P2] - T[P1]) / Mass[P1];
                }
            }
        }
        ext_forces_work += ext_forces_work_step * deltat;
    }
    // ---- Source/Contact.cpp:41-55 UpdateContactParticles --------------------------------------------------------
    void UpdateContactParticles() {
        for (size_t m = 0; m < meshes.size(); ++m) for (size_t e = 0; e < meshes[m].nelem(); ++e) {
            V3 vv; vv(0) = vv(1) = vv(2) = 0.;
            for (int en = 0; en < 3; ++en) vv = vadd(vv, meshes[m].node_v[meshes[m].en[3 * e + en]]);
            size_t p = meshes[m].first_particle + e;
            set3(v, p, vdiv(vv, 3)); a[3 * p] = a[3 * p + 1] = a[3 * p + 2] = 0.; set3(normal, p, meshes[m].normal[e]);
        }
    }
    // ---- Source/ThermalStuct.cpp:14-39 + Particle.cpp:922-939 CalcPlasticWorkHeat -------------------------------
    void CalcPlasticWorkHeat(double dt) {
        if (!pl_heating) return;
        double pl_sum = 0.;
        for (int p = 0; p < solid_part_count; ++p) {
            if (delta_pl_strain[p] > 0.0) {
                M3 depdt = smul(1. / dt, getM(Strain_pl_incr, p)); M3 S = getM(Sigma, p);
                q_plheat[p] = S(0, 0) * depdt(0, 0) + 2.0 * S(0, 1) * depdt(1, 0) + 2.0 * S(0, 2) * depdt(2, 0) + S(1, 1) * depdt(1, 1) +
                              2.0 * S(1, 2) * depdt(2, 1) + S(2, 2) * depdt(2, 2);
            }
        }
        for (int p = 0; p < solid_part_count; ++p) pl_sum += Mass[p] / Density[p] * q_plheat[p];
        plastic_work += pl_sum * dt;
    }
    // ---- Source/Thermal.cpp:721-748 ThermalCalcs, :339-389 CalcConvHeat, :53-217 CalcTempInc ----------------------
    void ThermalCalcs(double dt) {
        if (!thermal_solver) return;
        for (int i = 0; i < solid_part_count; ++i) T[i] += dt * dTdt[i];
        for (size_t i = 0; i < N; ++i) {
            double dens = Density[i];
            double dS2 = std::pow(Mass[i] / dens, 2.0 / 3.0);
            if (Thermal_BC[i] == 1) q_conv[i] = dens * h_conv[i] * dS2 * (T_inf[i] - T[i]) / Mass[i];
        }
        dvec temp(N, 0.0);
        for (size_t p = 0; p < SM.size(); ++p) {
            int P1 = SM[p].first, P2 = SM[p].second;   // discovery roles, not min/max (Thermal.cpp:71-72)
            V3 xij = vsub(X(P1), X(P2));
            double hh = (h[P1] + h[P2]) / 2.0;
            double GK = GradKernel(Dimension, KernelType, norm3(xij) / hh, hh);
            double dj = Density[P2], mj = Mass[P2];
            if (gradKernelCorr) {      // Thermal.cpp:90-111: v = gradCorrM_i (GK xij); both halves carry m_j/rho_j
                double mc[2];
                for (int k = 0; k < 2; ++k) {
                    V3 vv = mvmul(getM(gradCorrM, k == 0 ? P1 : P2), vscale(GK, xij));
                    mc[k] = mj / dj * 4. * (k_T[P1] * k_T[P2]) / (k_T[P1] + k_T[P2]) * (T[P1] - T[P2]) * dot3(xij, vv) / (norm3(xij) * norm3(xij));
                }
                temp[P1] += mc[0]; temp[P2] -= mc[1];
                continue;
            }
            double m = mj / dj * 4. * (k_T[P1] * k_T[P2]) / (k_T[P1] + k_T[P2]) * (T[P1] - T[P2]) * dot3(xij, vscale(GK, xij)) / (norm3(xij) * norm3(xij));
            temp[P1] += m; temp[P2] -= m;
        }
        double fr_temp = 0., cond_temp = 0.;
        for (int i = 0; i < solid_part_count; ++i) {
            double d = Density[i];
            double f = 1. / (d * cp_T[i]);
            dTdt[i] = f * (temp[i] + q_conv[i] + q_source[i] + q_plheat[i] * pl_work_heat_frac + q_cont_conv[i]);
            if (contact && cont_heat_fric) dTdt[i] += f * q_fric_work[i];
        }
        if (contact) for (int i = 0; i < solid_part_count; ++i) {
            double d = Density[i];
            fr_temp += q_fric_work[i] * Mass[i] / d;
            if (cont_heat_cond) cond_temp += q_cont_conv[i] * Mass[i] / d;
        }
        for (size_t i = solid_part_count; i < N; ++i) dTdt[i] = 0.;
        contact_friction_work += fr_temp * deltat;
        accum_cont_heat_cond += cond_temp * deltat;
        // CalcThermalExpStrainRate (ThermalStuct.cpp:7-12, Particle.cpp:953-957): StrainRate += th_exp*dTdt*I
        M3 I = zero3(); I(0, 0) = I(1, 1) = I(2, 2) = 1.0;
        for (size_t i = 0; i < N; ++i) setM(SRate, i, madd(getM(SRate, i), smul(th_exp[i] * dTdt[i], I)));
    }
    // ---- Source/Domain.cpp:1752-1829 CalcKinEnergyEqn / CalcIntEnergyEqn -----------------------------------------
    void Energies() {
        std::fill(dkin.begin(), dkin.end(), 0.);
        for (size_t p = 0; p < SM.size(); ++p) {
            int P1 = SM[p].first, P2 = SM[p].second;
            V3 xij = vsub(X(P1), X(P2)), vij = vsub(Vv(P1), Vv(P2));
            double mj = Mass[P2], di = Density[P1], dj = Density[P2];
            double hh = (h[P1] + h[P2]) / 2.0;
            double GK = GradKernel(Dimension, KernelType, norm3(xij) / hh, hh);
            double temp = 0.5 * mj * (Pressure[P1] / (di * di) + Pressure[P2] / (dj * dj)) * dot3(vij, vscale(GK, xij));
            dkin[P1] += temp; dkin[P2] -= temp;
        }
        double inc = 0.;
        for (size_t i = 0; i < N; ++i) inc += dkin[i];
        kin_energy_sum += inc * deltat;
        inc = 0.;
        int mx = contact ? first_fem() : (int)N;
        for (int i = 0; i < mx; ++i) {
            M3 S = getM(Shear, i), R = getM(SRate, i);
            dint[i] = Mass[i] / Density[i] * (S(0, 0) * R(0, 0) + 2.0 * S(0, 1) * R(1, 0) + 2.0 * S(0, 2) * R(2, 0) + S(1, 1) * R(1, 1) + 2.0 * S(1, 2) * R(2, 1) + S(2, 2) * R(2, 2));
        }
        for (int i = 0; i < mx; ++i) inc += dint[i];
        int_energy_sum += inc * deltat;
    }
    // GeneralAfter callback as the oracle driver installs it (oracle/ref_driver.cpp UserBC; examples/Compression_thermal.cpp:30-66)
    void UserBC() {
        for (size_t b = 0; b < bcs.size(); ++b) {
            double f = ramp(Time, bcs[b].tau);
            for (size_t i = 0; i < N; ++i) if (ID[i] == bcs[b].id) {
                if (bcs[b].set_a) a[3 * i] = a[3 * i + 1] = a[3 * i + 2] = 0.0;
                for (int k = 0; k < 3; ++k) v[3 * i + k] = bcs[b].v[k] * f;
            }
        }
        for (size_t m = 0; m < meshes.size(); ++m) { double f = ramp(Time, meshes[m].tau); for (int k = 0; k < 3; ++k) meshes[m].m_v(k) = meshes[m].vel[k] * f; }
    }

    // ---- Domain.cpp:267-297 CheckMinTSVel (over the Nishiura lists: pairs as min first, then as max) + :205-239 AdaptiveTimeStep
    void CheckMinTSVel() {
        deltatmin = deltatint;
        for (size_t i = 0; i < N; ++i) {
            double mn = 1000.;
            if (i < ipairs.size()) {
                for (size_t n = 0; n < ipairs[i].size(); ++n) { int p = ipairs[i][n]; int j = std::max(SM[p].first, SM[p].second); double d = norm3(vsub(X(j), X(i))); if (d < mn) mn = d; }
                for (size_t n = 0; n < jpairs[i].size(); ++n) { int p = jpairs[i][n]; int j = std::min(SM[p].first, SM[p].second); double d = norm3(vsub(X(j), X(i))); if (d < mn) mn = d; }
            }
            double test = CFL * mn / (Cs[i] + norm3(Vv(i)));
            if (deltatmin > test) deltatmin = test;
        }
    }
    // ---- Domain.cpp:302-317 CheckMinTSAccel
    void CheckMinTSAccel() {
        deltatmin = deltatint;
        for (size_t i = 0; i < N; ++i) {
            double test = sqrt_h_a * std::sqrt(h[i] / norm3(A(i)));
            if (deltatmin > test) deltatmin = test;
        }
    }
    void AdaptiveTimeStep() {
        if (deltatint > deltatmin) {
            if (deltat < deltatmin) deltat = 2.0 * deltat * deltatmin / (deltat + deltatmin);
            else deltat = deltatmin;
        } else {
            if (deltatint != deltat) deltat = 2.0 * deltat * deltatint / (deltat + deltatint);
            else deltat = deltatint;
        }
        if (deltat < manual_min_ts) deltat = manual_min_ts;
    }
    // ---- Source/Domain.cpp:1142-1236 MoveGhost, ghost_type Mirror_XYZ with ghost_plane_axis = 2 (the rows AddCylinderLength(..., ghost = true)
    //      creates, Domain.cpp:795-826): the ghost takes v, a (mirrored across the plane), Sigma, Strain, ShearStress, StrainRate, RotationRate
    //      and Density of its inner particle; va / vb are copied too but nothing on this path reads them
    void MoveGhost() {
        const int axis = 2;
        for (size_t gp = 0; gp < GhostPairs.size(); ++gp) {
            const size_t i = GhostPairs[gp].first, gi = GhostPairs[gp].second;
            for (int k = 0; k < 3; ++k) v[3 * gi + k] = v[3 * i + k];
            v[3 * gi + axis] = -v[3 * i + axis];
            for (int k = 0; k < 3; ++k) a[3 * gi + k] = a[3 * i + k];
            a[3 * gi + axis] = -a[3 * i + axis];
            setM(Sigma, gi, getM(Sigma, i)); setM(Strain, gi, getM(Strain, i)); setM(Shear, gi, getM(Shear, i));
            setM(SRate, gi, getM(SRate, i)); setM(RRate, gi, getM(RRate, i));
            Density[gi] = Density[i];
        }
    }
    // ---- Source/ThermalTest.cpp:8-49 CalcTempIncPair / CalcTempIncPP: the initial dT/dt of ThermalSolve, per particle over the
    //      Nishiura lists (pairs as min first, then as max), always from particle i's side
    void CalcTempIncPP() {
        for (size_t i = 0; i < N; ++i) dTdt[i] = 0.;
        for (size_t i = 0; i < N; ++i) {
            double k = 1. / (Density[i] * cp_T[i]);
            auto pair = [&](int P1, int P2) {
                V3 xij = vsub(X(P1), X(P2));
                double hh = (h[P1] + h[P2]) / 2.0;
                double GK = GradKernel(Dimension, KernelType, norm3(xij) / hh, hh);
                double dj = Density[P2], mj = Mass[P2];
                return mj / dj * 4. * (k_T[P1] * k_T[P2]) / (k_T[P1] + k_T[P2]) * (T[P1] - T[P2]) * dot3(xij, vscale(GK, xij)) / (dot3(xij, xij));
            };
            if (i < ipairs.size()) {
                for (size_t n = 0; n < ipairs[i].size(); ++n) { int p = ipairs[i][n]; dTdt[i] += pair((int)i, std::max(SM[p].first, SM[p].second)); }
                for (size_t n = 0; n < jpairs[i].size(); ++n) { int p = jpairs[i][n]; dTdt[i] += pair((int)i, std::min(SM[p].first, SM[p].second)); }
            }
            dTdt[i] *= k;
        }
    }
    // ---- Source/Thermal.cpp:567-719 ThermalSolve: one neighbour search, then ThermalCalcs + GeneralAfter per step
    void ThermalSolve(int nsteps, double dt) {
        deltat = deltatint = deltatmin = dt;
        for (size_t i = 0; i < N; ++i) Pressure[i] = EOS(PresEq[i], Cs[i], P0[i], Density[i], RefDensity[i]);
        CellInitiate(); ListGenerate();
        MainNeighbourSearch(); SaveNeighbourData();
        if (gradKernelCorr) CalcGradCorrMatrix();
        CalcPairPosList();
        CalcTempIncPP();
        for (int step = 0; step < nsteps; ++step) {
            ThermalCalcs(dt);
            UserBC();
            Time += deltat;
            steps_done++;
        }
    }
    // ---- Source/SolverFraser.cpp:21-513 SolveDiffUpdateFraser, SolverKickDrift.cpp:21-400 SolveDiffUpdateKickDrift,
    //      SolverLeapfrog.cpp:24-450 SolveDiffUpdateLeapFrog: one neighbour policy, three step orders.
    //      nsteps > 0: that many iterations; else while (Time <= tf) as the reference loops.
    void Solve(int nsteps, double dt, double tf = 0.0) {
        deltat = deltatint = deltatmin = dt;
        for (size_t i = 0; i < N; ++i) Pressure[i] = EOS(PresEq[i], Cs[i], P0[i], Density[i], RefDensity[i]);   // InitialChecks, Domain.cpp:2097-2113
        CellInitiate(); ListGenerate();
        int ts_i = 0; bool isfirst = true, isyielding = false;
        if (contact) {
            for (size_t i = 0; i < N; ++i) ID_orig[i] = ID[i];
            MainNeighbourSearch(); SaveNeighbourData(); CalculateSurface(1);
        }
        ClearNbData();
        const double MIN_PS_FOR_NBSEARCH = 1.e-6;   // Domain.cpp:28
        double prev_deltat = deltat;
        for (int step = 0; nsteps > 0 ? step < nsteps : Time <= tf; ++step) {
            StartAcceleration();
            double mx = 0;
            for (size_t i = 0; i < N; ++i) if (pl_strain[i] > mx) mx = pl_strain[i];
            double mdl[3] = { 0., 0., 0. };
            for (int i = 0; i < solid_part_count; ++i) for (int j = 0; j < 3; ++j)
                if (Disp[3 * i + j] * Disp[3 * i + j] > mdl[j]) mdl[j] = Disp[3 * i + j] * Disp[3 * i + j];
            for (int j = 0; j < 3; ++j) mdl[j] = std::sqrt(mdl[j]);
            // the KickDrift / LeapFrog loops keep max_disp in a LOCAL that shadows Domain::max_disp (SolverKickDrift.cpp:122,
            // SolverLeapfrog.cpp:138): the member a driver reads stays at its constructor value there
            if (solver == 0) for (int j = 0; j < 3; ++j) max_disp[j] = mdl[j];
            V3 md; md(0) = mdl[0]; md(1) = mdl[1]; md(2) = mdl[2];
            bool check_nb_every_time = norm3(md) > 0.1 * hmax;
            if (mx > MIN_PS_FOR_NBSEARCH && !isyielding) {
                ClearNbData(); MainNeighbourSearch(); CalcPairPosList(); SaveNeighbourData();
                if (contact) ContactNbUpdate();
                isyielding = true;
            }
            if (mx > MIN_PS_FOR_NBSEARCH || isfirst || check_nb_every_time) {
                if (ts_i == 0 && nb_cleared) {
                    MainNeighbourSearch(); CalcPairPosList(); SaveNeighbourData();
                    // SolverFraser.cpp:203-204, SolverLeapfrog.cpp:196-197: at cadence rebuilds, not at the first-yield search; KickDrift: never here
                    if (gradKernelCorr && solver != 1) CalcGradCorrMatrix();
                    if (contact) ContactNbUpdate();
                }
            }
            if (isfirst && gradKernelCorr) CalcGradCorrMatrix();     // SolverFraser.cpp:228-231, SolverKickDrift.cpp:187-190, SolverLeapfrog.cpp:208-211
            if (solver == 0) {
                if (gradKernelCorr) Density_lock(); else Density_pass();
                for (size_t i = 0; i < N; ++i) Density[i] += deltat * dDensity[i];
                if (gradKernelCorr) RateTensors_lock(); else RateTensors_pass();
                for (size_t i = 0; i < N; ++i) CalcStressStrain(i, deltat);
                if (gradKernelCorr) Accel_lock(); else Accel_pass();
                if (contact) ContactWang();
                UserBC();
                MoveGhost();                                   // SolverFraser.cpp:335
                for (size_t i = 0; i < N; ++i) for (int k = 0; k < 3; ++k) {
                    double du = (v[3 * i + k] + 0.0) * deltat + 0.5 * a[3 * i + k] * deltat * deltat;
                    Disp[3 * i + k] += du; x[3 * i + k] += du;
                }
                for (size_t i = 0; i < N; ++i) for (int k = 0; k < 3; ++k) v[3 * i + k] += a[3 * i + k] * deltat;
                MoveGhost();                                   // :358
                UserBC();
            } else {
                // SolverKickDrift.cpp:195-271 / SolverLeapfrog.cpp:213-290
                if (gradKernelCorr) Accel_lock(); else Accel_pass();
                UserBC();
                if (contact) ContactWang();
                if (solver == 2) {
                    if (isfirst) for (size_t i = 0; i < 3 * N; ++i) v[i] -= a[i] * 0.5 * deltat;
                    MoveGhost();                               // SolverLeapfrog.cpp:253
                    UserBC();
                }
                prev_deltat = deltat;
                if (auto_ts) CheckMinTSVel();
                if (auto_ts_acc) CheckMinTSAccel();
                if (auto_ts || auto_ts_acc) AdaptiveTimeStep();
                if (solver == 1) for (size_t i = 0; i < 3 * N; ++i) v[i] += a[i] * 0.5 * prev_deltat * 1.;
                else for (size_t i = 0; i < 3 * N; ++i) v[i] += a[i] * 0.5 * (prev_deltat + deltat);
                MoveGhost();                                   // SolverKickDrift.cpp:230, SolverLeapfrog.cpp:272
                UserBC();
                if (gradKernelCorr) Density_lock(); else Density_pass();
                if (solver == 1) for (size_t i = 0; i < N; ++i) Density[i] += deltat * dDensity[i] * 1.;
                else for (size_t i = 0; i < N; ++i) Density[i] += deltat * dDensity[i];
                for (size_t i = 0; i < 3 * N; ++i) {
                    double du = solver == 1 ? (v[i] + 0.0) * deltat * 1. : (v[i] + 0.0) * deltat;
                    Disp[i] += du; x[i] += du;
                }
                if (solver == 1) { for (size_t i = 0; i < 3 * N; ++i) v[i] += a[i] * deltat / 2. * 1.; MoveGhost(); UserBC(); }   // SolverKickDrift.cpp:266
                if (gradKernelCorr) RateTensors_lock(); else RateTensors_pass();
                for (size_t i = 0; i < N; ++i) CalcStressStrain(i, deltat);
            }
            CalcPlasticWorkHeat(deltat);
            ThermalCalcs(deltat);
            Energies();
            Time += deltat;
            if (contact) { for (size_t m = 0; m < meshes.size(); ++m) meshes[m].update(deltat); UpdateContactParticles(); }
            if (mx > MIN_PS_FOR_NBSEARCH) {
                if (ts_i == (ts_nb_inc - 1)) ClearNbData();
                ts_i++;
                if (ts_i > (ts_nb_inc - 1)) ts_i = 0;
            }
            if (isfirst) isfirst = false;
            steps_done++;
        }
    }
};

// ---- generators: Source/Domain.cpp:674-689 calcHalfPartCount, :708-860 AddCylinderLength (3-D; ghost = the three mirror rows at the bottom) -------
int calcHalfPartCount(double r, double R, int xinc)
{
    int ypartcount = -1;
    if (xinc > 0) {
        ypartcount = 1;
        double yp = r, xp = r + (double)(xinc - 1) * 2. * r;
        double rad = std::sqrt(yp * yp + xp * xp);
        while (rad <= R - r) { yp += 2. * r; rad = std::sqrt(yp * yp + xp * xp); ypartcount++; }
        ypartcount -= 1;
    }
    return ypartcount;
}
void gen_cylinder(const double * V, double Rxy, double Lz, double r, dvec & pts, bool ghost = false, std::vector<std::pair<int, int> > * ghost_pairs = nullptr)
{
    int numpartxy = calcHalfPartCount(r, Rxy, 1);
    double zp = V[2] + r;
    int part_per_row = 0;
    while (zp <= (V[2] + Lz - r)) {
        double yp = V[1] - r - (2. * r * (numpartxy - 1));
        int numypart = 2 * numpartxy, yinc = numpartxy, yinc_sign = -1;
        for (int j = 0; j < numypart; ++j) {
            int numxpart = calcHalfPartCount(r, Rxy, yinc);
            double xp = V[0] - r - (2. * r * (numxpart - 1));
            for (int i = 0; i < 2 * numxpart; ++i) { pts.push_back(xp); pts.push_back(yp); pts.push_back(zp); if (zp == V[2] + r) part_per_row++; xp += 2. * r; }
            yp += 2. * r; yinc += yinc_sign;
            if (yinc < 1) { yinc = 1; yinc_sign = 1; }
        }
        zp += 2. * r;
    }
    if (ghost) {
        // Domain.cpp:795-826: three mirror rows below the bottom plane; ghost particle id_part copies particle part (rows 0..2 from the bottom)
        const int ghost_rows = 3;
        zp = V[2] - r;
        int part = 0, id_part = (int)(pts.size() / 3);
        for (int zinc = 0; zinc < ghost_rows; ++zinc) {
            for (int xy = 0; xy < part_per_row; ++xy) {
                const double xp = pts[3 * part], yp = pts[3 * part + 1];
                pts.push_back(xp); pts.push_back(yp); pts.push_back(zp);
                if (ghost_pairs) ghost_pairs->push_back(std::make_pair(part, id_part));
                id_part++; part++;
            }
            zp -= 2.0 * r;
        }
    }
}
// ---- Source/Domain.cpp:326-470 AddBoxLength, 3-D cubic packing (type 1, no random) -----------------------------
void gen_box_cubic(const double * V, double Lx, double Ly, double Lz, double r, dvec & pts)
{
    size_t k = 0; double zp = V[2];
    while (zp <= (V[2] + Lz - r)) {
        size_t j = 0; double yp = V[1];
        while (yp <= (V[1] + Ly - r)) {
            size_t i = 0; double xp = V[0];
            while (xp <= (V[0] + Lx - r)) {
                pts.push_back(V[0] + (2.0 * i + 1) * r); pts.push_back(V[1] + (2.0 * j + 1) * r); pts.push_back(V[2] + (2.0 * k + 1) * r);
                i++; xp = V[0] + (2 * i + 1) * r;
            }
            j++; yp = V[1] + (2.0 * j + 1) * r;
        }
        k++; zp = V[2] + (2.0 * k + 1) * r;
    }
}

void load_case(const wfc::File & in, Dom & d)
{
    d.Dimension = (int)in.scalar("Dimension", 3);
    d.KernelType = (int)in.scalar("KernelType", 0);
    d.Cellfac = (d.KernelType == 2) ? 3.0 : 2.0;    // Domain::Kernel_Set, Domain.cpp:191-195
    d.GradientType = (int)in.scalar("GradientType", 0);
    if (in.has("DomMax")) for (int k = 0; k < 3; ++k) d.DomMax[k] = in.f("DomMax")[k];
    if (in.has("DomMin")) for (int k = 0; k < 3; ++k) d.DomMin[k] = in.f("DomMin")[k];
    const int gen_kind = (int)in.scalar("gen_kind", 0);
    dvec pts; double gmass = 0.0, grho = 0.0, gh = 0.0; int gtag = 0;
    if (gen_kind == 1) {
        const dvec & g = in.f("gen_args"); gen_cylinder(&g[1], g[4], g[5], g[6], pts, g.size() > 10 && g[10] != 0.0, &d.GhostPairs);
        gtag = (int)g[0]; grho = g[7]; gh = g[8];
        double Vol = M_PI * g[4] * g[4] * g[5]; gmass = Vol * grho / (pts.size() / 3);   // Domain.cpp:839-841
    } else if (gen_kind == 2) {
        const dvec & g = in.f("gen_args"); gen_box_cubic(&g[1], g[4], g[5], g[6], g[7], pts);
        gtag = (int)g[0]; grho = g[8]; gh = g[9];
        double Max[3] = { g[1], g[2], g[3] };
        for (size_t i = 0; i < pts.size() / 3; ++i) for (int k = 0; k < 3; ++k) if (pts[3 * i + k] > Max[k]) Max[k] = pts[3 * i + k];
        double t[3]; for (int k = 0; k < 3; ++k) { Max[k] += g[7]; t[k] = Max[k] - g[1 + k]; }
        gmass = t[0] * t[1] * t[2] * grho / (pts.size() / 3);                            // Domain.cpp:447-455
    } else pts = in.f("x");
    const size_t nsph = pts.size() / 3;
    // rigid meshes first so the particle arrays can be sized once
    const int mesh_count = (int)in.scalar("mesh_count", 0);
    size_t ntot = nsph;
    for (int m = 0; m < mesh_count; ++m) {
        char key[64]; std::snprintf(key, 64, "mesh%d_plane", m); const dvec & pl = in.f(key);
        Mesh ms; for (int k = 0; k < 3; ++k) { ms.m_v(k) = 0.; ms.m_w(k) = 0.; ms.vel[k] = 0.; } ms.tau = 0.;
        V3 p1, p2; for (int k = 0; k < 3; ++k) { p1(k) = pl[2 + k]; p2(k) = pl[5 + k]; }
        ms.axis_plane((int)pl[0], pl[1] != 0.0, p1, p2, (int)pl[8]); ms.calc_spheres();
        std::snprintf(key, 64, "mesh%d_hfac", m); ms.hfac = (float)in.scalar(key, 1.1);
        std::snprintf(key, 64, "mesh%d_id", m); ms.id = (int)in.scalar(key, 10 + m);
        std::snprintf(key, 64, "mesh%d_w", m); if (in.has(key)) { const dvec & w_ = in.f(key); for (int k = 0; k < 3; ++k) ms.m_w(k) = w_[k]; }   // TriMesh::SetRotAxisVel
        std::snprintf(key, 64, "mesh%d_vel", m); if (in.has(key)) { const dvec & a_ = in.f(key); for (int k = 0; k < 3; ++k) ms.vel[k] = a_[k]; ms.tau = a_[3]; }
        ms.first_particle = (int)ntot; ntot += ms.nelem();
        d.meshes.push_back(ms);
    }
    d.resize(ntot); d.solid_part_count = (int)nsph;
    for (size_t i = 0; i < 3 * nsph; ++i) d.x[i] = pts[i];
    if (gen_kind != 0) for (size_t i = 0; i < nsph; ++i) { d.Mass[i] = gmass; d.Density[i] = d.RefDensity[i] = grho; d.h[i] = gh; d.ID[i] = gtag; }
    for (size_t g = 0; g < d.GhostPairs.size(); ++g) d.nowrite[d.GhostPairs[g].second] = 1;     // Domain.cpp:814 not_write_surf_ID
#define LD_D(NAME, MEM) if (in.has(NAME)) { const dvec & a_ = in.f(NAME); for (size_t i = 0; i < nsph; ++i) d.MEM[i] = a_[i]; }
#define LD_I(NAME, MEM) if (in.has(NAME)) { const std::vector<int32_t> & a_ = in.i(NAME); for (size_t i = 0; i < nsph; ++i) d.MEM[i] = a_[i]; }
    LD_D("Mass", Mass) LD_D("h", h) LD_I("ID", ID) LD_I("IsFree", IsFree)
    if (in.has("Density")) { const dvec & a_ = in.f("Density"); for (size_t i = 0; i < nsph; ++i) { d.Density[i] = a_[i]; if (gen_kind == 0) d.RefDensity[i] = a_[i]; } }
    if (in.has("v")) { const dvec & a_ = in.f("v"); for (size_t i = 0; i < 3 * nsph; ++i) d.v[i] = a_[i]; }
    LD_D("RefDensity", RefDensity) LD_D("G", G) LD_D("Cs", Cs) LD_D("P0", P0) LD_I("PresEq", PresEq) LD_D("Alpha", Alpha) LD_D("Beta", Beta)
    LD_D("TI", TI) LD_D("TIn", TIn) LD_D("TIInitDist", TIInitDist) LD_D("Sigmay", Sigmay) LD_D("Ep", Ep) LD_I("Fail", Fail) LD_I("Material", Material)
    LD_I("Material_model", Material_model) LD_D("k_T", k_T) LD_D("cp_T", cp_T) LD_D("T", T) LD_D("h_conv", h_conv) LD_D("T_inf", T_inf)
    LD_I("Thermal_BC", Thermal_BC) LD_D("th_exp", th_exp) LD_D("q_source", q_source) LD_I("not_write_surf_ID", nowrite) LD_I("mat_id", mat_id)
    if (in.has("materials")) {
        const dvec & mt = in.f("materials");
        for (size_t r = 0; r + 16 <= mt.size(); r += 16) {
            Mat m; m.model = (int)mt[r]; m.E = mt[r + 1]; m.nu = mt[r + 2]; for (int k = 0; k < 16; ++k) m.p[k] = 0.0; for (int k = 0; k < 13; ++k) m.p[k] = mt[r + 3 + k];
            if (in.has("materials_ext")) { const dvec & me = in.f("materials_ext"); const size_t q = (r / 16) * 3; for (int k = 0; k < 3 && q + k < me.size(); ++k) m.p[13 + k] = me[q + k]; }
            m.eps0 = m.eps1 = 0.0;
            if (m.model == HOLLOMON) { m.eps0 = m.p[0] / m.E; m.eps1 = std::pow(m.p[0] / m.p[1], 1. / m.p[2]); }   // Material.cpp:48-57
            d.mats.push_back(m);
        }
    }
    // mesh particles: Contact.cpp:10-38 AddTrimeshParticles
    for (size_t m = 0; m < d.meshes.size(); ++m) {
        Mesh & ms = d.meshes[m];
        for (size_t e = 0; e < ms.nelem(); ++e) {
            size_t p = ms.first_particle + e;
            for (int k = 0; k < 3; ++k) { d.x[3 * p + k] = ms.centroid[e](k); d.normal[3 * p + k] = ms.normal[e](k); }
            d.h[p] = (float)ms.hfac * ms.radius[e];
            d.ID[p] = ms.id; d.pelem[p] = (int)e; d.pmesh[p] = (int)m;
        }
        V3 z; z(0) = z(1) = z(2) = 0.0; d.m_contact_force.push_back(z);
    }
    d.contact = in.scalar("contact", mesh_count > 0 ? 1 : 0) != 0.0;
    d.friction_sta = in.scalar("friction_sta", 0.0); d.friction_dyn = in.scalar("friction_dyn", 0.0);
    d.friction_function = (int)in.scalar("friction_function", 0); d.friction_m = in.scalar("friction_m", 0.0); d.friction_b = in.scalar("friction_b", 0.0);
    d.PFAC = in.scalar("PFAC", 0.4); d.DFAC = in.scalar("DFAC", 0.0);
    d.cont_heat_fric = in.scalar("cont_heat_fric", 0) != 0.0; d.cont_heat_cond = in.scalar("cont_heat_cond", 0) != 0.0;
    d.ts_nb_inc = (int)in.scalar("ts_nb_inc", 5);
    d.gradKernelCorr = in.scalar("gradKernelCorr", 0) != 0.0;
    d.solver = (int)in.scalar("solver", 0); d.auto_ts = in.scalar("auto_ts", 0) != 0.0; d.auto_ts_acc = in.scalar("auto_ts_acc", 0) != 0.0; d.CFL = in.scalar("CFL", 0.7);
    d.thermal_solver = in.scalar("thermal_solver", 0) != 0.0; d.pl_heating = in.scalar("pl_heating", 0) != 0.0;
    d.pl_work_heat_frac = in.scalar("pl_work_heat_frac", 0.9);
    if (in.has("bc")) { const dvec & b = in.f("bc"); for (size_t r = 0; r + 8 <= b.size(); r += 8) { BC s; s.id = (int)b[r]; s.v[0] = b[r + 1]; s.v[1] = b[r + 2]; s.v[2] = b[r + 3]; s.tau = b[r + 4]; s.set_a = (int)b[r + 5]; d.bcs.push_back(s); } }
}

void put9(wfc::File & out, const char * n, const dvec & f) { out.put_f(n, f.data(), f.size()); }

void dump_pairs(wfc::File & out, const char * name, const std::vector<std::pair<int, int> > & P, bool sorted)
{
    std::vector<std::pair<int64_t, int64_t> > v;
    for (size_t a = 0; a < P.size(); ++a) { int64_t i = P[a].first, j = P[a].second; if (sorted) v.push_back(std::make_pair(std::min(i, j), std::max(i, j))); else v.push_back(std::make_pair(i, j)); }
    if (sorted) std::sort(v.begin(), v.end());
    std::vector<int64_t> flat(2 * v.size());
    for (size_t a = 0; a < v.size(); ++a) { flat[2 * a] = v[a].first; flat[2 * a + 1] = v[a].second; }
    out.put_l(name, flat.data(), flat.size());
}

} // namespace

// Same file interface as oracle/_ref/wf_ref: case file in, state dump out. Returns 0 on success.
extern "C" int wfo_run_file(const char * case_path, const char * out_path)
{
    wfc::File in; if (!in.read(case_path)) return 2;
    Dom d; load_case(in, d);
    const int mode = (int)in.scalar("mode", 0);
    const double dt = in.scalar("dt", 1.0e-7); const int nsteps = (int)in.scalar("nsteps", 1);
    wfc::File out; double solve_seconds = 0.0;
    if (mode == 1) {
        d.CellInitiate(); d.ListGenerate(); d.MainNeighbourSearch(); d.SaveNeighbourData();
        double cell[9] = { d.BLPF[0], d.BLPF[1], d.BLPF[2], d.CellSize[0], d.CellSize[1], d.CellSize[2], (double)d.CellNo[0], (double)d.CellNo[1], (double)d.CellNo[2] };
        out.put_f("cell_grid", cell, 9);
        dump_pairs(out, "pairs_discovery", d.SM, false);
    } else {
        auto t0 = std::chrono::steady_clock::now();
        const double tf = in.scalar("tf", 0.0);          // > 0: loop while (Time <= tf) as the reference does (needed with auto_ts)
        if (d.solver == 3) d.ThermalSolve(nsteps, dt);
        else if (tf > 0.0) d.Solve(0, dt, tf); else d.Solve(nsteps, dt);
        solve_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    out.put_scalar("n_particles", (double)d.N); out.put_scalar("solid_part_count", (double)d.solid_part_count);
    out.put_scalar("Time", d.Time); out.put_scalar("solve_seconds", solve_seconds); out.put_scalar("deltat", d.deltat);
    out.put_scalar("kin_energy_sum", d.kin_energy_sum); out.put_scalar("int_energy_sum", d.int_energy_sum);
    out.put_scalar("plastic_work", d.plastic_work); out.put_scalar("contact_force_sum", d.contact_force_sum);
    out.put_scalar("contact_reaction_sum", d.contact_reaction_sum); out.put_scalar("ext_forces_work", d.ext_forces_work);
    out.put_scalar("contact_friction_work", d.contact_friction_work);
    out.put_f("max_disp", d.max_disp, 3);
    put9(out, "x", d.x); put9(out, "v", d.v); put9(out, "a", d.a); put9(out, "Displacement", d.Disp); put9(out, "normal", d.normal); put9(out, "contforce", d.contforce);
    put9(out, "Mass", d.Mass); put9(out, "h", d.h); put9(out, "Density", d.Density); put9(out, "dDensity", d.dDensity); put9(out, "Pressure", d.Pressure);
    put9(out, "Sigma", d.Sigma); put9(out, "ShearStress", d.Shear); put9(out, "StrainRate", d.SRate); put9(out, "RotationRate", d.RRate);
    put9(out, "Strain", d.Strain); put9(out, "Strain_pl", d.Strain_pl); put9(out, "TIR", d.TIR);
    if (d.gradKernelCorr) put9(out, "gradCorrM", d.gradCorrM);
    put9(out, "pl_strain", d.pl_strain); put9(out, "Sigmay", d.Sigmay); put9(out, "eff_strain_rate", d.eff_strain_rate);
    put9(out, "T", d.T); put9(out, "dTdt", d.dTdt); put9(out, "q_plheat", d.q_plheat); put9(out, "q_conv", d.q_conv); put9(out, "delta_cont", d.delta_cont);
    put9(out, "dkin_energy_dt", d.dkin); put9(out, "dint_energy_dt", d.dint);
    out.put_i("ID", d.ID.data(), d.ID.size()); out.put_i("Nb", d.Nb.data(), d.Nb.size());
    dump_pairs(out, "pairs", d.SM, true); dump_pairs(out, "cont_pairs", d.CONT, true);
    for (size_t m = 0; m < d.meshes.size(); ++m) {
        char key[64]; const Mesh & ms = d.meshes[m];
        dvec nodes(3 * ms.node.size()), nrm(3 * ms.nelem());
        for (size_t n = 0; n < ms.node.size(); ++n) for (int k = 0; k < 3; ++k) nodes[3 * n + k] = ms.node[n](k);
        for (size_t e = 0; e < ms.nelem(); ++e) for (int k = 0; k < 3; ++k) nrm[3 * e + k] = ms.normal[e](k);
        std::snprintf(key, 64, "mesh%d_nodes", (int)m); out.put_f(key, nodes.data(), nodes.size());
        std::snprintf(key, 64, "mesh%d_normals", (int)m); out.put_f(key, nrm.data(), nrm.size());
        std::snprintf(key, 64, "mesh%d_pplane", (int)m); out.put_f(key, ms.pplane.data(), ms.pplane.size());
    }
    return out.write(out_path) ? 0 : 2;
}
