// Device math for the SPH hot path. Every expression keeps the operand order and association of the reference
// expression it replaces (file:line given per function) and the library is compiled with -fmad=false, so a pair term
// computed here is bit-identical to the reference's (x86-64 -O3 without FMA). Paths relative to /root/reference.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define WF_PI 3.14159265358979323846

namespace wf {

struct V3 { double x, y, z; };
struct S6 { double xx, yy, zz, xy, yz, xz; };     // symmetric 3x3
struct M3 {                                        // full 3x3, row-major
    double m[9];
    __host__ __device__ double & operator()(int r, int c) { return m[3 * r + c]; }
    __host__ __device__ double operator()(int r, int c) const { return m[3 * r + c]; }
};

__host__ __device__ inline V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__host__ __device__ inline V3 operator-(const V3 & a, const V3 & b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ inline V3 operator+(const V3 & a, const V3 & b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ inline V3 operator*(double s, const V3 & a) { return v3(s * a.x, s * a.y, s * a.z); }
__host__ __device__ inline V3 operator/(const V3 & a, double s) { return v3(a.x / s, a.y / s, a.z / s); }
// Blitz++ dot meta-program order a0*b0 + (a1*b1 + a2*b2); norm = sqrt(dot(a,a)) (External/matvec.h:807-810)
__host__ __device__ inline double dot(const V3 & a, const V3 & b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
__host__ __device__ inline double norm(const V3 & a) { return sqrt(dot(a, a)); }
// External/matvec.h:1967-1973
__host__ __device__ inline V3 cross(const V3 & a, const V3 & b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

__host__ __device__ inline M3 m3zero() { M3 r; for (int k = 0; k < 9; ++k) r.m[k] = 0.0; return r; }
__host__ __device__ inline M3 m3sym(const S6 & s) { M3 r; r.m[0] = s.xx; r.m[1] = s.xy; r.m[2] = s.xz; r.m[3] = s.xy; r.m[4] = s.yy; r.m[5] = s.yz; r.m[6] = s.xz; r.m[7] = s.yz; r.m[8] = s.zz; return r; }
__host__ __device__ inline S6 s6of(const M3 & a) { S6 s; s.xx = a.m[0]; s.yy = a.m[4]; s.zz = a.m[8]; s.xy = a.m[1]; s.yz = a.m[5]; s.xz = a.m[2]; return s; }
// External/matvec.h:547-554 (double * Mat3_t): element * a
__host__ __device__ inline M3 operator*(double a, const M3 & A) { M3 r; for (int k = 0; k < 9; ++k) r.m[k] = A.m[k] * a; return r; }
__host__ __device__ inline M3 operator+(const M3 & A, const M3 & B) { M3 r; for (int k = 0; k < 9; ++k) r.m[k] = A.m[k] + B.m[k]; return r; }
__host__ __device__ inline M3 operator-(const M3 & A, const M3 & B) { M3 r; for (int k = 0; k < 9; ++k) r.m[k] = A.m[k] - B.m[k]; return r; }
// External/matvec.h:845-850 Mult(Mat3,Mat3,Mat3): inner index runs 2,1,0
__host__ __device__ inline M3 mmul(const M3 & A, const M3 & B)
{ M3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = A(i, 2) * B(2, j) + A(i, 1) * B(1, j) + A(i, 0) * B(0, j); return r; }
__host__ __device__ inline M3 mtrans(const M3 & A) { M3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = A(j, i); return r; }

// ---- Source/Functions.cpp:25-72 Kernel ------------------------------------------------------------------------
__host__ __device__ inline double kernelW(int Dim, int KT, double q, double h)
{
    double C;
    switch (KT) {
    case 0:
        C = (Dim == 2) ? 10.0 / (7.0 * h * h * WF_PI) : 1.0 / (h * h * h * WF_PI);
        if (q < 1.0) return C * (1.0 - (3.0 / 2.0) * q * q + (3.0 / 4.0) * q * q * q);
        else if (q < 2.0) return C * ((1.0 / 4.0) * (2.0 - q) * (2.0 - q) * (2.0 - q));
        else return 0.0;
    case 1:
        C = (Dim == 2) ? 7.0 / (4.0 * h * h * WF_PI) : 7.0 / (8.0 * h * h * h * WF_PI);
        if (q < 2.0) return C * pow((1.0 - q / 2.0), 4.0) * (2.0 * q + 1.0);
        else return 0.0;
    case 2:
        C = (Dim == 2) ? 7.0 / (478.0 * h * h * WF_PI) : 1.0 / (120.0 * h * h * h * WF_PI);
        if (q < 1.0) return C * (pow((3.0 - q), 5.0) - 6.0 * pow((2.0 - q), 5.0) + 15.0 * pow((1.0 - q), 5.0));
        else if (q < 2.0) return C * (pow((3.0 - q), 5.0) - 6.0 * pow((2.0 - q), 5.0));
        else if (q < 3.0) return C * (pow((3.0 - q), 5.0));
        else return 0.0;
    default:
        C = (Dim == 2) ? 1.0 / (3.0 * h * h * WF_PI) : 15.0 / (62.0 * h * h * h * WF_PI);
        if (q < 1.0) return C * (pow(q, 3.0) - 6.0 * q + 6.);
        else if (q < 2.0) return C * (pow((2.0 - q), 3.0));
        else return 0.0;
    }
}
// ---- Source/Functions.cpp:74-123 GradKernel -------------------------------------------------------------------
__host__ __device__ inline double gradW(int Dim, int KT, double q, double h)
{
    double C;
    switch (KT) {
    case 0:
        C = (Dim == 2) ? 10.0 / (7.0 * h * h * h * WF_PI) : 1.0 / (h * h * h * h * WF_PI);
        if (q == 0.0) return C / h * (-3.0 + (9.0 / 2.0) * q);
        else if (q < 1.0) return C / (q * h) * (-3.0 * q + (9.0 / 4.0) * q * q);
        else if (q < 2.0) return C / (q * h) * ((-3.0 / 4.0) * (2.0 - q) * (2.0 - q));
        else return 0.0;
    case 1:
        C = (Dim == 2) ? 7.0 / (4.0 * h * h * h * WF_PI) : 7.0 / (8.0 * h * h * h * h * WF_PI);
        if (q == 0.0) return C * -5.0 / h * (pow((1.0 - q / 2.0), 3.0) - 3.0 * q / 2.0 * pow((1.0 - q / 2.0), 2.0));
        else if (q < 2.0) return C / (q * h) * -5.0 * q * pow((1.0 - q / 2.0), 3.0);
        else return 0.0;
    case 2:
        C = (Dim == 2) ? 7.0 / (478.0 * h * h * h * WF_PI) : 1.0 / (120.0 * h * h * h * h * WF_PI);
        if (q == 0.0) return C / h * (20.0 * pow((3.0 - q), 3.0) - 120.0 * pow((2.0 - q), 3.0) + 300.0 * pow((1.0 - q), 3.0));
        else if (q < 1.0) return C / (q * h) * (-5.0 * pow((3.0 - q), 4.0) + 30.0 * pow((2.0 - q), 4.0) - 75.0 * pow((1.0 - q), 4.0));
        else if (q < 2.0) return C / (q * h) * (-5.0 * pow((3.0 - q), 4.0) + 30.0 * pow((2.0 - q), 4.0));
        else if (q < 3.0) return C / (q * h) * (-5.0 * pow((3.0 - q), 4.0));
        else return 0.0;
    default:
        C = (Dim == 2) ? 1.0 / (3.0 * h * h * h * WF_PI) : 15.0 / (62.0 * h * h * h * h * WF_PI);
        if (q < 1.0) return C / (q * h) * (3.0 * pow(q, 2.0) - 6.0);
        else if (q < 2.0) return C / (q * h) * (3.0 * pow((2.0 - q), 2.0) - 1.0);
        else return 0.0;
    }
}
// Cubic-spline (KT 0, 3-D) fast forms: identical expressions with the constant C hoisted by the caller
// (C3 = 1.0/(h*h*h*PI), C4 = 1.0/(h*h*h*h*PI), evaluated with the same operand order -> same bits).
__host__ __device__ inline double cubicW_c(double C3, double q)
{
    if (q < 1.0) return C3 * (1.0 - (3.0 / 2.0) * q * q + (3.0 / 4.0) * q * q * q);
    else if (q < 2.0) return C3 * ((1.0 / 4.0) * (2.0 - q) * (2.0 - q) * (2.0 - q));
    else return 0.0;
}
__host__ __device__ inline double cubicGW_c(double C4, double q, double h)
{
    if (q == 0.0) return C4 / h * (-3.0 + (9.0 / 2.0) * q);
    else if (q < 1.0) return C4 / (q * h) * (-3.0 * q + (9.0 / 4.0) * q * q);
    else if (q < 2.0) return C4 / (q * h) * ((-3.0 / 4.0) * (2.0 - q) * (2.0 - q));
    else return 0.0;
}
// ---- Source/Functions.cpp:206-230 EOS, :232-256 SoundSpeed ----------------------------------------------------
__host__ __device__ inline double eos(int EQ, double Cs0, double P00, double Density, double Density0)
{
    switch (EQ) {
    case 0: return P00 + (Cs0 * Cs0) * (Density - Density0);
    case 1: return P00 + (Density0 * Cs0 * Cs0 / 7.0) * (pow(Density / Density0, 7.0) - 1);
    default: return (Cs0 * Cs0) * Density;
    }
}
__host__ __device__ inline double soundSpeed(int EQ, double Cs0, double Density, double Density0)
{
    if (EQ == 1) return sqrt((Cs0 * Cs0) * pow(Density / Density0, 6.0));
    return Cs0;
}

// ---- 3x3 symmetric eigen-solve replacing GSL gsl_eigen_symmv behind Rotation() (Source/Functions.cpp:381-395,
// External/matvec.h:776-804). Cyclic Jacobi; same operation sequence as oracle/shim/gsl/gsl_eigen.h and
// oracle/wf_oracle.cpp::sym_eig3, eigenpairs sorted ascending as Eig(...,SortAsc=true) does.
__host__ __device__ inline void sym_eig3(M3 a, double d[3], M3 & v)
{
    for (int i = 0; i < 3; ++i) for (int j = i + 1; j < 3; ++j) a(i, j) = a(j, i);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) v(i, j) = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        off += fabs(a(0, 1)); off += fabs(a(0, 2)); off += fabs(a(1, 2));
        if (off == 0.0) break;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2, r = 3 - p - q;
            double apq = a(p, q);
            if (apq == 0.0) continue;
            double app = a(p, p), aqq = a(q, q);
            double g = 100.0 * fabs(apq);
            if (fabs(app) + g == fabs(app) && fabs(aqq) + g == fabs(aqq)) { a(p, q) = 0.0; a(q, p) = 0.0; continue; }
            double theta = (aqq - app) / (2.0 * apq);
            double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
            if (theta < 0.0) t = -t;
            double c = 1.0 / sqrt(t * t + 1.0);
            double s = t * c;
            a(p, p) = app - t * apq; a(q, q) = aqq + t * apq; a(p, q) = 0.0; a(q, p) = 0.0;
            {
                double arp = a(r, p), arq = a(r, q);
                double nrp = c * arp - s * arq, nrq = s * arp + c * arq;
                a(r, p) = nrp; a(p, r) = nrp; a(r, q) = nrq; a(q, r) = nrq;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double vrp = v(k, p), vrq = v(k, q);
                v(k, p) = c * vrp - s * vrq; v(k, q) = s * vrp + c * vrq;
            }
        }
    }
    for (int i = 0; i < 3; ++i) d[i] = a(i, i);
    for (int i = 0; i < 2; ++i) {
        int k = i; double ek = d[i];
        for (int j = i + 1; j < 3; ++j) if (d[j] < ek) { k = j; ek = d[j]; }
        if (k != i) {
            double t = d[i]; d[i] = d[k]; d[k] = t;
            for (int r = 0; r < 3; ++r) { double u = v(r, i); v(r, i) = v(r, k); v(r, k) = u; }
        }
    }
}

} // namespace wf
