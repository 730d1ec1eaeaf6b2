// Host-side mirror of the small-vector surface of the reference's External/matvec.h (Blitz++ TinyVector/TinyMatrix
// typedefs Vec3_t / Mat3_t) that the example drivers and GeneralAfter callbacks use: constructors, (i) / [i] access,
// scalar assignment, element-wise arithmetic, norm / dot / cross, stream output. Plain fp64 PODs -- the numerics of the
// hot path live on the device, these types only carry values across the API.
#ifndef WF_HOST_MATVEC_H
#define WF_HOST_MATVEC_H
#include <cmath>
#include <cstdio>
#include <iostream>

#include "array.h"
#include "fatal.h"

struct Vec3_t {
    double d[3];
    Vec3_t() { d[0] = d[1] = d[2] = 0.0; }
    Vec3_t(double s) { d[0] = d[1] = d[2] = s; }
    Vec3_t(double x, double y, double z) { d[0] = x; d[1] = y; d[2] = z; }
    double & operator()(int i) { return d[i]; }
    double operator()(int i) const { return d[i]; }
    double & operator[](int i) { return d[i]; }
    double operator[](int i) const { return d[i]; }
    Vec3_t & operator=(double s) { d[0] = d[1] = d[2] = s; return *this; }
    Vec3_t & operator+=(const Vec3_t & o) { for (int k = 0; k < 3; ++k) d[k] += o.d[k]; return *this; }
    Vec3_t & operator-=(const Vec3_t & o) { for (int k = 0; k < 3; ++k) d[k] -= o.d[k]; return *this; }
    Vec3_t & operator*=(double s) { for (int k = 0; k < 3; ++k) d[k] *= s; return *this; }
    Vec3_t & operator/=(double s) { for (int k = 0; k < 3; ++k) d[k] /= s; return *this; }
    const double * data() const { return d; }
};
inline Vec3_t operator+(const Vec3_t & a, const Vec3_t & b) { return Vec3_t(a.d[0] + b.d[0], a.d[1] + b.d[1], a.d[2] + b.d[2]); }
inline Vec3_t operator-(const Vec3_t & a, const Vec3_t & b) { return Vec3_t(a.d[0] - b.d[0], a.d[1] - b.d[1], a.d[2] - b.d[2]); }
inline Vec3_t operator-(const Vec3_t & a) { return Vec3_t(-a.d[0], -a.d[1], -a.d[2]); }
inline Vec3_t operator*(double s, const Vec3_t & a) { return Vec3_t(s * a.d[0], s * a.d[1], s * a.d[2]); }
inline Vec3_t operator*(const Vec3_t & a, double s) { return Vec3_t(s * a.d[0], s * a.d[1], s * a.d[2]); }
inline Vec3_t operator*(const Vec3_t & a, const Vec3_t & b) { return Vec3_t(a.d[0] * b.d[0], a.d[1] * b.d[1], a.d[2] * b.d[2]); }
inline Vec3_t operator/(const Vec3_t & a, double s) { return Vec3_t(a.d[0] / s, a.d[1] / s, a.d[2] / s); }
inline double dot(const Vec3_t & a, const Vec3_t & b) { return a.d[0] * b.d[0] + (a.d[1] * b.d[1] + a.d[2] * b.d[2]); }
inline double norm(const Vec3_t & a) { return std::sqrt(dot(a, a)); }
inline Vec3_t cross(const Vec3_t & a, const Vec3_t & b) { return Vec3_t(a.d[1] * b.d[2] - a.d[2] * b.d[1], a.d[2] * b.d[0] - a.d[0] * b.d[2], a.d[0] * b.d[1] - a.d[1] * b.d[0]); }
inline std::ostream & operator<<(std::ostream & os, const Vec3_t & v) { return os << "(" << v.d[0] << "," << v.d[1] << "," << v.d[2] << ")"; }

struct Mat3_t {
    double d[9];
    Mat3_t() { for (int k = 0; k < 9; ++k) d[k] = 0.0; }
    double & operator()(int i, int j) { return d[3 * i + j]; }
    double operator()(int i, int j) const { return d[3 * i + j]; }
    Mat3_t & operator=(double s) { for (int k = 0; k < 9; ++k) d[k] = s; return *this; }
};
inline Mat3_t operator+(const Mat3_t & a, const Mat3_t & b) { Mat3_t r; for (int k = 0; k < 9; ++k) r.d[k] = a.d[k] + b.d[k]; return r; }
inline Mat3_t operator-(const Mat3_t & a, const Mat3_t & b) { Mat3_t r; for (int k = 0; k < 9; ++k) r.d[k] = a.d[k] - b.d[k]; return r; }
inline Mat3_t operator*(double s, const Mat3_t & a) { Mat3_t r; for (int k = 0; k < 9; ++k) r.d[k] = a.d[k] * s; return r; }
inline void set_to_zero(Mat3_t & m) { m = 0.0; }
inline void set_to_zero(Vec3_t & v) { v = 0.0; }
inline std::ostream & operator<<(std::ostream & os, const Mat3_t & m) { for (int i = 0; i < 3; ++i) os << m(i, 0) << " " << m(i, 1) << " " << m(i, 2) << "\n"; return os; }

#endif
