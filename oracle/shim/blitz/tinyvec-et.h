// TEST INFRASTRUCTURE (oracle build only) -- never part of the product path.
// Minimal stand-in for Blitz++ 0.9 <blitz/tinyvec-et.h>, written from scratch for this repo so that
// the untouched reference sources under /root/reference compile here (Blitz++ itself is not installed).
// Only the surface the reference touches (External/matvec.h:542-547, Source/*.cpp) is provided:
// value-initialised TinyVector<T,N>, comma initialiser, element-wise arithmetic, dot/norm, stream output.
// Arithmetic order notes (they matter for bit-level parity):
//   * every binary operator is evaluated eagerly, component by component, exactly as written by the caller;
//   * dot(a,b) follows Blitz's recursive meta-program  a0*b0 + (a1*b1 + (a2*b2)) (right-associated);
//   * norm(a) = sqrt(dot(a,a)).
#ifndef WF_SHIM_BLITZ_TINYVEC_ET_H
#define WF_SHIM_BLITZ_TINYVEC_ET_H
#include <cmath>
#include <cstddef>
#include <iostream>

namespace blitz {

template <typename T, int N> class TinyVector;

// helper implementing  V = a, b, c;
template <typename T> class ListInit_ {
    T * p_; T * end_;
public:
    ListInit_(T * p, T * end) : p_(p), end_(end) {}
    ListInit_ operator, (T v) { if (p_ < end_) *p_ = v; return ListInit_(p_ + 1, end_); }
};

template <typename T, int N>
class TinyVector {
    T d_[N];
public:
    TinyVector() { for (int i = 0; i < N; ++i) d_[i] = T(); }
    TinyVector(T v) { for (int i = 0; i < N; ++i) d_[i] = v; }
    TinyVector(T a, T b) { for (int i = 0; i < N; ++i) d_[i] = T(); d_[0] = a; if (N > 1) d_[1] = b; }
    TinyVector(T a, T b, T c) { for (int i = 0; i < N; ++i) d_[i] = T(); d_[0] = a; if (N > 1) d_[1] = b; if (N > 2) d_[2] = c; }
    TinyVector(const TinyVector & o) { for (int i = 0; i < N; ++i) d_[i] = o.d_[i]; }
    TinyVector & operator=(const TinyVector & o) { for (int i = 0; i < N; ++i) d_[i] = o.d_[i]; return *this; }
    ListInit_<T> operator=(T v) { for (int i = 0; i < N; ++i) d_[i] = v; return ListInit_<T>(d_ + 1, d_ + N); }

    T & operator()(int i) { return d_[i]; }
    const T & operator()(int i) const { return d_[i]; }
    T & operator[](int i) { return d_[i]; }
    const T & operator[](int i) const { return d_[i]; }
    T * data() { return d_; }
    const T * data() const { return d_; }
    int length() const { return N; }

    TinyVector & operator+=(const TinyVector & o) { for (int i = 0; i < N; ++i) d_[i] += o.d_[i]; return *this; }
    TinyVector & operator-=(const TinyVector & o) { for (int i = 0; i < N; ++i) d_[i] -= o.d_[i]; return *this; }
    TinyVector & operator*=(const TinyVector & o) { for (int i = 0; i < N; ++i) d_[i] *= o.d_[i]; return *this; }
    TinyVector & operator/=(const TinyVector & o) { for (int i = 0; i < N; ++i) d_[i] /= o.d_[i]; return *this; }
    TinyVector & operator+=(T s) { for (int i = 0; i < N; ++i) d_[i] += s; return *this; }
    TinyVector & operator-=(T s) { for (int i = 0; i < N; ++i) d_[i] -= s; return *this; }
    TinyVector & operator*=(T s) { for (int i = 0; i < N; ++i) d_[i] *= s; return *this; }
    TinyVector & operator/=(T s) { for (int i = 0; i < N; ++i) d_[i] /= s; return *this; }
};

#define WF_SHIM_VV(op) \
template <typename T, int N> inline TinyVector<T,N> operator op (const TinyVector<T,N> & a, const TinyVector<T,N> & b) \
{ TinyVector<T,N> r; for (int i = 0; i < N; ++i) r[i] = a[i] op b[i]; return r; }
WF_SHIM_VV(+) WF_SHIM_VV(-) WF_SHIM_VV(*) WF_SHIM_VV(/)
#undef WF_SHIM_VV

// scalar on either side; the scalar may be int/float/double in the reference, promote through double
#define WF_SHIM_VS(op, S) \
template <int N> inline TinyVector<double,N> operator op (const TinyVector<double,N> & a, S s) \
{ TinyVector<double,N> r; for (int i = 0; i < N; ++i) r[i] = a[i] op s; return r; } \
template <int N> inline TinyVector<double,N> operator op (S s, const TinyVector<double,N> & a) \
{ TinyVector<double,N> r; for (int i = 0; i < N; ++i) r[i] = s op a[i]; return r; }
#define WF_SHIM_VS_ALL(S) WF_SHIM_VS(+, S) WF_SHIM_VS(-, S) WF_SHIM_VS(*, S) WF_SHIM_VS(/, S)
WF_SHIM_VS_ALL(double) WF_SHIM_VS_ALL(float) WF_SHIM_VS_ALL(int)
#undef WF_SHIM_VS_ALL
#undef WF_SHIM_VS

template <typename T, int N> inline TinyVector<T,N> operator-(const TinyVector<T,N> & a)
{ TinyVector<T,N> r; for (int i = 0; i < N; ++i) r[i] = -a[i]; return r; }

template <typename T, int N> struct DotRec_ {
    static inline T f(const TinyVector<T,N> & a, const TinyVector<T,N> & b, int i)
    { return (i == N - 1) ? a[i] * b[i] : a[i] * b[i] + f(a, b, i + 1); }
};
template <typename T, int N> inline T dot(const TinyVector<T,N> & a, const TinyVector<T,N> & b)
{ return DotRec_<T,N>::f(a, b, 0); }
template <typename T, int N> inline T norm(const TinyVector<T,N> & a) { return std::sqrt(dot(a, a)); }

template <typename T, int N> inline std::ostream & operator<<(std::ostream & os, const TinyVector<T,N> & v)
{ os << "("; for (int i = 0; i < N; ++i) { os << v[i]; if (i + 1 < N) os << ","; } os << ")"; return os; }

} // namespace blitz
#endif
