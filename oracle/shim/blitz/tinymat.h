// TEST INFRASTRUCTURE (oracle build only). Stand-in for Blitz++ <blitz/tinymat.h>; see tinyvec-et.h.
// TinyMatrix<T,R,C>: value-initialised, row-major, (i,j) access, comma initialiser, data().
// All matrix arithmetic the reference needs is defined by the reference itself (External/matvec.h:536-575).
#ifndef WF_SHIM_BLITZ_TINYMAT_H
#define WF_SHIM_BLITZ_TINYMAT_H
#include "tinyvec-et.h"
namespace blitz {
template <typename T, int R, int C>
class TinyMatrix {
    T d_[R * C];
public:
    TinyMatrix() { for (int i = 0; i < R * C; ++i) d_[i] = T(); }
    TinyMatrix(const TinyMatrix & o) { for (int i = 0; i < R * C; ++i) d_[i] = o.d_[i]; }
    TinyMatrix & operator=(const TinyMatrix & o) { for (int i = 0; i < R * C; ++i) d_[i] = o.d_[i]; return *this; }
    ListInit_<T> operator=(T v) { for (int i = 0; i < R * C; ++i) d_[i] = v; return ListInit_<T>(d_ + 1, d_ + R * C); }
    T & operator()(int i, int j) { return d_[i * C + j]; }
    const T & operator()(int i, int j) const { return d_[i * C + j]; }
    T * data() { return d_; }
    const T * data() const { return d_; }
    int rows() const { return R; }
    int cols() const { return C; }
};
template <typename T, int R, int C> inline std::ostream & operator<<(std::ostream & os, const TinyMatrix<T,R,C> & m)
{ for (int i = 0; i < R; ++i) { for (int j = 0; j < C; ++j) os << m(i, j) << " "; os << "\n"; } return os; }
} // namespace blitz
#endif
