// TEST INFRASTRUCTURE (oracle build only) -- never part of the product path.
// Stand-in for the handful of GSL entry points External/matvec.h:494-533,759-804 uses (GSL is not installed
// here). The symmetric eigen-solver is a cyclic Jacobi iteration written for this repo; the same sequence of
// operations is restated in oracle/wf_oracle.cpp (sym_eig3) and weldform_b200/csrc (wf_sym_eig3) so that all three
// agree bit for bit when FMA contraction is off. Real GSL uses tridiagonal QR: results agree to round-off only,
// and enter the reference solely through V*diag*V^T (Source/Domain.cpp:1559-1569), which is invariant to
// eigenvector sign and order. PARITY UNPINNED at this third-party boundary (GSL version unpinned, README.md:46).
#ifndef WF_SHIM_GSL_EIGEN_H
#define WF_SHIM_GSL_EIGEN_H
#include <cmath>
#include <cstddef>
#include <cstdlib>

struct gsl_block { size_t size; double * data; };
struct gsl_vector { size_t size; size_t stride; double * data; gsl_block * block; int owner; };
struct gsl_matrix { size_t size1; size_t size2; size_t tda; double * data; gsl_block * block; int owner; };
struct gsl_vector_view { gsl_vector vector; };
struct gsl_matrix_view { gsl_matrix matrix; };
struct gsl_eigen_symm_workspace  { size_t n; };
struct gsl_eigen_symmv_workspace { size_t n; };
enum gsl_eigen_sort_t { GSL_EIGEN_SORT_VAL_ASC, GSL_EIGEN_SORT_VAL_DESC, GSL_EIGEN_SORT_ABS_ASC, GSL_EIGEN_SORT_ABS_DESC };

inline gsl_vector * gsl_vector_alloc(size_t n)
{ gsl_vector * v = (gsl_vector*)std::malloc(sizeof(gsl_vector)); v->size = n; v->stride = 1; v->data = (double*)std::calloc(n, sizeof(double)); v->block = 0; v->owner = 1; return v; }
inline void gsl_vector_free(gsl_vector * v) { if (v) { if (v->owner) std::free(v->data); std::free(v); } }
inline double gsl_vector_get(const gsl_vector * v, size_t i) { return v->data[i * v->stride]; }
inline void gsl_vector_set(gsl_vector * v, size_t i, double x) { v->data[i * v->stride] = x; }
inline gsl_matrix * gsl_matrix_alloc(size_t r, size_t c)
{ gsl_matrix * m = (gsl_matrix*)std::malloc(sizeof(gsl_matrix)); m->size1 = r; m->size2 = c; m->tda = c; m->data = (double*)std::calloc(r * c, sizeof(double)); m->block = 0; m->owner = 1; return m; }
inline void gsl_matrix_free(gsl_matrix * m) { if (m) { if (m->owner) std::free(m->data); std::free(m); } }
inline double gsl_matrix_get(const gsl_matrix * m, size_t i, size_t j) { return m->data[i * m->tda + j]; }
inline void gsl_matrix_set(gsl_matrix * m, size_t i, size_t j, double x) { m->data[i * m->tda + j] = x; }
inline gsl_matrix_view gsl_matrix_view_array(double * base, size_t r, size_t c)
{ gsl_matrix_view v; v.matrix.size1 = r; v.matrix.size2 = c; v.matrix.tda = c; v.matrix.data = base; v.matrix.block = 0; v.matrix.owner = 0; return v; }
inline gsl_vector_view gsl_matrix_column(gsl_matrix * m, size_t j)
{ gsl_vector_view v; v.vector.size = m->size1; v.vector.stride = m->tda; v.vector.data = m->data + j; v.vector.block = 0; v.vector.owner = 0; return v; }
inline gsl_eigen_symm_workspace  * gsl_eigen_symm_alloc (size_t n) { gsl_eigen_symm_workspace  * w = (gsl_eigen_symm_workspace *)std::malloc(sizeof(*w)); w->n = n; return w; }
inline gsl_eigen_symmv_workspace * gsl_eigen_symmv_alloc(size_t n) { gsl_eigen_symmv_workspace * w = (gsl_eigen_symmv_workspace*)std::malloc(sizeof(*w)); w->n = n; return w; }
inline void gsl_eigen_symm_free (gsl_eigen_symm_workspace  * w) { std::free(w); }
inline void gsl_eigen_symmv_free(gsl_eigen_symmv_workspace * w) { std::free(w); }

// Cyclic Jacobi for a symmetric n x n matrix (n <= 8). a: row-major n*n (destroyed), d: eigenvalues, v: eigenvectors
// in columns (row-major n*n) or NULL.
inline void wf_shim_jacobi(size_t n, double * a, size_t tda, double * d, double * v, size_t vtda)
{
    if (v) for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) v[i * vtda + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (size_t p = 0; p < n; ++p) for (size_t q = p + 1; q < n; ++q) off += std::fabs(a[p * tda + q]);
        if (off == 0.0) break;
        for (size_t p = 0; p < n; ++p) for (size_t q = p + 1; q < n; ++q) {
            double apq = a[p * tda + q];
            if (apq == 0.0) continue;
            double app = a[p * tda + p], aqq = a[q * tda + q];
            double g = 100.0 * std::fabs(apq);
            if (std::fabs(app) + g == std::fabs(app) && std::fabs(aqq) + g == std::fabs(aqq)) {
                a[p * tda + q] = 0.0; a[q * tda + p] = 0.0; continue;
            }
            double theta = (aqq - app) / (2.0 * apq);
            double t = 1.0 / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
            if (theta < 0.0) t = -t;
            double c = 1.0 / std::sqrt(t * t + 1.0);
            double s = t * c;
            a[p * tda + p] = app - t * apq;
            a[q * tda + q] = aqq + t * apq;
            a[p * tda + q] = 0.0; a[q * tda + p] = 0.0;
            for (size_t r = 0; r < n; ++r) {
                if (r == p || r == q) continue;
                double arp = a[r * tda + p], arq = a[r * tda + q];
                double nrp = c * arp - s * arq;
                double nrq = s * arp + c * arq;
                a[r * tda + p] = nrp; a[p * tda + r] = nrp;
                a[r * tda + q] = nrq; a[q * tda + r] = nrq;
            }
            if (v) for (size_t r = 0; r < n; ++r) {
                double vrp = v[r * vtda + p], vrq = v[r * vtda + q];
                v[r * vtda + p] = c * vrp - s * vrq;
                v[r * vtda + q] = s * vrp + c * vrq;
            }
        }
    }
    for (size_t i = 0; i < n; ++i) d[i] = a[i * tda + i];
}
inline int gsl_eigen_symm(gsl_matrix * A, gsl_vector * eval, gsl_eigen_symm_workspace *)
{
    // symmetrise from the lower triangle like GSL does
    size_t n = A->size1;
    for (size_t i = 0; i < n; ++i) for (size_t j = i + 1; j < n; ++j) A->data[i * A->tda + j] = A->data[j * A->tda + i];
    double d[8]; wf_shim_jacobi(n, A->data, A->tda, d, 0, 0);
    for (size_t i = 0; i < n; ++i) gsl_vector_set(eval, i, d[i]);
    return 0;
}
inline int gsl_eigen_symmv(gsl_matrix * A, gsl_vector * eval, gsl_matrix * evec, gsl_eigen_symmv_workspace *)
{
    size_t n = A->size1;
    for (size_t i = 0; i < n; ++i) for (size_t j = i + 1; j < n; ++j) A->data[i * A->tda + j] = A->data[j * A->tda + i];
    double d[8]; wf_shim_jacobi(n, A->data, A->tda, d, evec->data, evec->tda);
    for (size_t i = 0; i < n; ++i) gsl_vector_set(eval, i, d[i]);
    return 0;
}
// selection sort of eigenpairs, same scheme GSL documents (swap columns of evec)
inline int gsl_eigen_symmv_sort(gsl_vector * eval, gsl_matrix * evec, gsl_eigen_sort_t type)
{
    size_t n = eval->size;
    for (size_t i = 0; i + 1 < n; ++i) {
        size_t k = i; double ek = gsl_vector_get(eval, i);
        for (size_t j = i + 1; j < n; ++j) {
            double ej = gsl_vector_get(eval, j); bool better = false;
            switch (type) {
                case GSL_EIGEN_SORT_VAL_ASC:  better = (ej < ek); break;
                case GSL_EIGEN_SORT_VAL_DESC: better = (ej > ek); break;
                case GSL_EIGEN_SORT_ABS_ASC:  better = (std::fabs(ej) < std::fabs(ek)); break;
                case GSL_EIGEN_SORT_ABS_DESC: better = (std::fabs(ej) > std::fabs(ek)); break;
            }
            if (better) { k = j; ek = ej; }
        }
        if (k != i) {
            double t = gsl_vector_get(eval, i); gsl_vector_set(eval, i, gsl_vector_get(eval, k)); gsl_vector_set(eval, k, t);
            for (size_t r = 0; r < evec->size1; ++r) {
                double u = gsl_matrix_get(evec, r, i); gsl_matrix_set(evec, r, i, gsl_matrix_get(evec, r, k)); gsl_matrix_set(evec, r, k, u);
            }
        }
    }
    return 0;
}
#endif
