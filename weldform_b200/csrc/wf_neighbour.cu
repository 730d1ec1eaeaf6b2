// Cell-linked-list neighbour build on the device. Replaces Domain::CellInitiate / ListGenerate / MainNeighbourSearch /
// AllocateNbPair / CalcPairPosList / SaveNeighbourData (Source/Neighbour.cpp:10-694), Domain::CalculateSurface
// (Source/Domain.cpp:1415-1491) and Domain::ContactNbSearch (Source/Contact.cpp:57-132).
//
// Design: cell key on the reference's own grid -> 64-bit radix sort (cell, descending creation index) -> all particle
// arrays are permuted into cell order (so the pair kernels gather from nearly contiguous memory) -> cell start/end
// table -> one thread per particle walks its 27 cells in the order in which the reference's serial sweep would have
// DISCOVERED each pair and writes a per-particle neighbour list (ELL layout, transposed for coalescing):
//   rows [0, nb_hi)          neighbours with larger creation index  (the reference's Anei/Aref "front", j > i)
//   rows [K-1, K-nb_lo] down neighbours with smaller creation index (the Nishiura table "back", j < i)
// each in discovery order, which is exactly the order AccelReduction/DensReduction/RateTensorsReduction sum in
// (InteractionAlt.cpp:237-260,450-479,613-640) when Nproc = 1. There is no fixed MAX_NB_PER_PART cap: K grows.
#include <cub/cub.cuh>
#include <algorithm>
#include <vector>
#include <cmath>

#include "wf_internal.cuh"

using namespace wf;

namespace {

struct Grid {
    double BLPF[3], CellSize[3], Cellfac, hmax; int CellNo[3]; int ax[3];   // ax[0] fastest .. ax[2] slowest storage axis
    __host__ __device__ int lin(const int c[3]) const { return (c[ax[2]] * CellNo[ax[1]] + c[ax[1]]) * CellNo[ax[0]] + c[ax[0]]; }
    __host__ __device__ void unlin(int l, int c[3]) const { c[ax[0]] = l % CellNo[ax[0]]; l /= CellNo[ax[0]]; c[ax[1]] = l % CellNo[ax[1]]; c[ax[2]] = l / CellNo[ax[1]]; }
};

// Neighbour.cpp:185-225 ListGenerate: cell index with clamping
// mig != 0 (multi-GPU): particles whose slow-axis cell layer left [layer_lo, layer_hi) get a sentinel key (they sort to
// the tail and are dropped) and are listed for migration to the adjacent rank (emig[dir*emig_cap + k], cnt[dir])
// mesh_tail != 0: rigid-mesh element particles (creation index >= nsph) are keyed in a cell space of their own (bit 30 set):
// they sort behind every SPH particle -- a contiguous tail of the local region, by cell, descending creation index inside a
// cell -- stay out of the SPH cell table, the halo slices and the migration, and get a cell table of their own, which
// k_cont_list walks. Every rank of a multi-GPU run keeps a replica of all of them.
#define WF_MESH_CELL_BIT 0x40000000u
struct MigArgs { int on, layer_lo, layer_hi; long long plane; int *emig; int *cnt; long long emig_cap; int mesh_tail, nsph; };
__global__ void k_cell_keys(Grid g, const double4 *Q0, const int *orig, int64_t n, int64_t s_off, unsigned long long *keys, int *vals, MigArgs mg)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    s += s_off;
    double4 q = Q0[s]; double x[3] = { q.x, q.y, q.z }; int c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int ci = (int)(floor((x[k] - g.BLPF[k]) / g.CellSize[k]));
        if (ci < 0) ci = 0;
        if (ci >= g.CellNo[k]) ci = g.CellNo[k] - 1;
        c[k] = ci;
    }
    // descending creation index inside a cell == the order the reference's push-front linked list is walked in
    const int lin = g.lin(c);
    if (mg.mesh_tail && orig[s] >= mg.nsph) {
        keys[s] = ((unsigned long long)((unsigned)lin | WF_MESH_CELL_BIT) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)orig[s]);
        vals[s] = (int)s;
        return;
    }
    if (mg.on) {
        const int layer = (int)(lin / mg.plane);
        if (layer < mg.layer_lo || layer >= mg.layer_hi) {
            const int dir = layer < mg.layer_lo ? 0 : 1;
            const int k = atomicAdd(&mg.cnt[dir], 1);
            if (k < mg.emig_cap) mg.emig[(long long)dir * mg.emig_cap + k] = (int)s;
            keys[s] = 0xFFFFFFFF00000000ull | (unsigned long long)(0xFFFFFFFFu - (unsigned)orig[s]);
            vals[s] = (int)s;
            return;
        }
    }
    keys[s] = ((unsigned long long)(unsigned)lin << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)orig[s]);
    vals[s] = (int)s;
}

__global__ void k_permute_q(int64_t n, const int *perm, const double4 *a0, const double4 *a1, const double4 *a2, const double4 *a3, const double4 *a4, const double4 *a5,
                            double4 *b0, double4 *b1, double4 *b2, double4 *b3, double4 *b4, double4 *b5)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int p = perm[s];
    b0[s] = a0[p]; b1[s] = a1[p]; b2[s] = a2[p]; b3[s] = a3[p]; b4[s] = a4[p]; b5[s] = a5[p];
}
// WF_PERM_FIELDS own fields per thread: one read of perm[s] and that many independent 8-byte gathers in flight
#define WF_PERM_FIELDS 8
__global__ void k_permute_f(int64_t n, int64_t cap, int nf, const int *perm, const double *a, double *b)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int p = perm[s];
    const int f0 = blockIdx.y * WF_PERM_FIELDS;
    double v[WF_PERM_FIELDS];
#pragma unroll
    for (int k = 0; k < WF_PERM_FIELDS; ++k) if (f0 + k < nf) v[k] = a[(int64_t)(f0 + k) * cap + p];
#pragma unroll
    for (int k = 0; k < WF_PERM_FIELDS; ++k) if (f0 + k < nf) b[(int64_t)(f0 + k) * cap + s] = v[k];
}
__global__ void k_permute_i(int64_t n, int64_t cap, const int *perm, const int *a, int *b, int *slot_of)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int64_t off = (int64_t)blockIdx.y * cap;
    int v = a[off + perm[s]];
    b[off + s] = v;
    if (blockIdx.y == I_LID && slot_of) slot_of[v] = (int)s;     // one GPU: local index == creation index
}
__global__ void k_cell_bounds(int64_t n, const unsigned long long *keys, int *cell_start, int *cell_end, int *mcell_start, int *mcell_end)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int cell = (int)(keys[s] >> 32);
    if (cell < 0) return;                      // migration sentinel
    int *cs = cell_start, *ce = cell_end; int idx = cell;
    if ((unsigned)cell & WF_MESH_CELL_BIT) { cs = mcell_start; ce = mcell_end; idx = (int)((unsigned)cell & ~WF_MESH_CELL_BIT); }
    if (s == 0 || (int)(keys[s - 1] >> 32) != cell) cs[idx] = (int)s;
    if (s == n - 1 || (int)(keys[s + 1] >> 32) != cell) ce[idx] = (int)s + 1;
}

// The 27 cells around (q1,q2,q3) in the order in which the reference's sweep (q1 outer, q3, q2 inner;
// Neighbour.cpp:346-359,418-477) emits pairs that involve a particle of the centre cell. role = 1: the centre-cell
// particle is temp1 ("first") of the pair, 0: it is temp2. Entry 9 is the centre cell itself.
__constant__ signed char c_walk[27][4] = {
    {-1,-1,-1,0}, {-1, 0,-1,0}, {-1, 1,-1,0}, {-1,-1, 0,0}, {-1, 0, 0,0},      // owner cells with smaller q1
    { 0,-1,-1,0}, { 0, 0,-1,0}, { 0, 1,-1,0}, { 0,-1, 0,0},                    // same q1, processed earlier
    { 0, 0, 0,2},                                                              // own cell (split by creation index)
    { 1, 0, 0,1},                                                              // (q1+1,q2,q3)
    {-1, 1, 0,1}, { 0, 1, 0,1}, { 1, 1, 0,1},                                  // q2+1 row
    {-1,-1, 1,1}, { 0,-1, 1,1}, { 1,-1, 1,1}, {-1, 0, 1,1}, { 0, 0, 1,1}, { 1, 0, 1,1}, {-1, 1, 1,1}, { 0, 1, 1,1}, { 1, 1, 1,1},   // q3+1 plane
    { 1,-1,-1,0}, { 1, 0,-1,0}, { 1, 1,-1,0}, { 1,-1, 0,0}                     // owner cells with larger q1
};

__device__ inline bool is_mesh_id(int id, int4 mids) { return id == mids.x || id == mids.y || id == mids.z || id == mids.w; }

// One thread per particle. Neighbour.cpp:374-417 AllocateNbPair + :361-372 CheckRadius, gather form.
__global__ void __launch_bounds__(128)
k_nb_list(Grid g, int64_t n, int64_t cap, int nsph, int K, const unsigned long long *keys, const int *cell_start, const int *cell_end,
          const double4 *__restrict__ Q0, const double4 *__restrict__ Q2, const int *__restrict__ orig, const int *__restrict__ ID,
          const int *__restrict__ cls_of, const int *__restrict__ is_free, const ClassProps *__restrict__ classes, int contact, int4 mids, int check_material, int uniform_h, int walk_order,
          uint32_t *nbr, int *nb_cnt, int *nb_hi, int *nbI, double *normal_x, double *normal_y, double *normal_z, int *d_maxcnt)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int total = 0;
    if (s < n) {
        const int oi = orig[s];
        int nhi = 0, nlo = 0;
        double nx = 0.0, ny = 0.0, nz = 0.0;
        bool active = oi < nsph;
        if (active && contact && is_mesh_id(ID[s], mids)) active = false;
        // fixed particles (IsFree = false) only ever enter FSMPairs (Neighbour.cpp:400-411), which no pair loop of the new
        // solvers reads: they keep an empty SM list and are invisible to their neighbours
        if (active && is_free && !is_free[s]) active = false;
        if (active) {
            const double4 qi = Q0[s];
            const V3 xi = v3(qi.x, qi.y, qi.z);
            const int mat_i = check_material ? classes[cls_of[s]].Material : 0;
            const double lim_u = g.Cellfac * ((qi.w + qi.w) / 2), thr_u = (lim_u * lim_u) * (1.0 + 8e-15);     // uniform h: (h+h)/2 == h
            int c[3]; g.unlin((int)(keys[s] >> 32), c);
            // distance from x_i to the lower / upper face of its own cell along each axis: a neighbour cell none of whose points can
            // lie within the largest possible cutoff Cellfac*(h_i + hmax)/2 is skipped (about a quarter of the 27 for a lattice).
            // Conservative: faces nudged by 1e-9 cell sizes (floor((x - BLPF)/CellSize) rounds), clamped particles count as on the face.
            const double xi3[3] = { qi.x, qi.y, qi.z };
            double glo[3], ghi[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double lo = g.BLPF[k] + c[k] * g.CellSize[k], eps = 1e-9 * g.CellSize[k];
                glo[k] = fmax(xi3[k] - lo - eps, 0.0); ghi[k] = fmax(lo + g.CellSize[k] - xi3[k] - eps, 0.0);
            }
            const double lim_c = g.Cellfac * ((qi.w + g.hmax) / 2), thr_c = (lim_c * lim_c) * (1.0 + 1e-9);
            for (int e = 0; e < 27; ++e) {
                int d[3] = { c[0] + c_walk[e][0], c[1] + c_walk[e][1], c[2] + c_walk[e][2] };
                if (d[0] < 0 || d[0] >= g.CellNo[0] || d[1] < 0 || d[1] >= g.CellNo[1] || d[2] < 0 || d[2] >= g.CellNo[2]) continue;
                double gap2 = 0.0;
#pragma unroll
                for (int k = 0; k < 3; ++k) { const double gk = c_walk[e][k] < 0 ? glo[k] : (c_walk[e][k] > 0 ? ghi[k] : 0.0); gap2 += gk * gk; }
                if (gap2 > thr_c) continue;
                const int l = g.lin(d);
                const int js = cell_start[l], je = cell_end[l];
                const int role = c_walk[e][3];
                // two phases per cell: (1) a store-free squared-distance sweep over the cell's candidates, 32 at a time, whose
                // loads the compiler can batch (lanes of one cell read the same record: one L1 wavefront); (2) the reference's
                // exact test and the list update, in candidate order, for the few that passed.
                for (int base = js; base < je; base += 32) {
                    const int m = min(32, je - base);
                    unsigned mask = 0u;
#pragma unroll 4
                    for (int k = 0; k < m; ++k) {
                        const double4 qj = Q0[base + k];        // two LDG.128: measured faster here than one LDG.256 (lanes of a cell read the same record)
                        const double ex = xi.x - qj.x, ey = xi.y - qj.y, ez = xi.z - qj.z;
                        const double d2 = ex * ex + (ey * ey + ez * ez);          // == dot(xij, xij) for either sign of xij
                        double thr = thr_u;
                        if (!uniform_h) { const double hh = (qi.w + qj.w) / 2, lim = g.Cellfac * hh; thr = (lim * lim) * (1.0 + 8e-15); }
                        if (!(d2 > thr)) mask |= 1u << k;
                    }
                    while (mask) {
                        const int j = base + (__ffs(mask) - 1); mask &= mask - 1u;
                        if (j == (int)s) continue;
                        const int oj = orig[j];
                        if (oj >= nsph) continue;                               // rigid-mesh particle: never an SM pair
                        if (contact && is_mesh_id(ID[j], mids)) continue;
                        if (is_free && !is_free[j]) continue;
                        const double4 qj = Q0[j];
                        // own cell: particles stored before s have a larger creation index and act as temp1
                        const int ifirst = (role == 2) ? (j > (int)s ? 1 : 0) : role;
                        const double hh = (qi.w + qj.w) / 2;
                        const V3 xj = v3(qj.x, qj.y, qj.z);
                        const V3 xij = ifirst ? (xi - xj) : (xj - xi);
                        // CheckRadius (Neighbour.cpp:361-372): sqrt(|xij|^2)/hh <= Cellfac. The squared form decides every candidate
                        // that is not within a few ulp of the cutoff (sqrt and the division each round within 2^-53 relative);
                        // only those borderline ones take the exact expression.
                        const double d2 = dot(xij, xij), lim = g.Cellfac * hh, lim2 = lim * lim;
                        if (d2 > lim2 * (1.0 + 8e-15)) continue;
                        if (!(d2 < lim2 * (1.0 - 8e-15))) { const double rij = sqrt(d2); if (!((rij / hh) <= g.Cellfac)) continue; }
                        if (check_material && classes[cls_of[j]].Material != mat_i) continue;
                        uint32_t entry = (uint32_t)j | (ifirst ? 0x80000000u : 0u);
                        // walk_order (gradKernelCorr: the reference then sums per pair with locks, SolverFraser.cpp:112-115): one list in
                        // plain discovery order, which is the order the pair loop adds into this particle
                        if (oj > oi || walk_order) { if (nhi + nlo < K) nbr[(int64_t)nhi * cap + s] = entry; ++nhi; }
                        else         { if (nhi + nlo < K) nbr[(int64_t)(K - 1 - nlo) * cap + s] = entry; ++nlo; }
                        if (contact) {
                            // Domain.cpp:1446-1463: P1->normal += mj*xij ; P2->normal -= mi*xij, in discovery order
                            const double mj = Q2[j].x;
                            if (ifirst) { nx += mj * xij.x; ny += mj * xij.y; nz += mj * xij.z; }
                            else        { nx -= mj * xij.x; ny -= mj * xij.y; nz -= mj * xij.z; }
                        }
                    }
                }
            }
        }
        total = nhi + nlo;
        nb_cnt[s] = total; nb_hi[s] = nhi; nbI[s] = total;
        if (contact && oi < nsph) { normal_x[s] = nx; normal_y[s] = ny; normal_z[s] = nz; }
    }
    // block max of the neighbour count
    typedef cub::BlockReduce<int, 128> BR; __shared__ typename BR::TempStorage tmp;
    int m = BR(tmp).Reduce(total, cub::Max());
    if (threadIdx.x == 0 && m > 0) atomicMax(d_maxcnt, m);
}

// Domain.cpp:1426-1429,1470-1487 CalculateSurface tail: restore IDs, scale the normal, flag free-surface particles
__global__ void k_surface(int64_t n, int nsph_first_fem, double inv_totmass, int surf_id, const double4 *Q0, const int *orig, const int *nb,
                          const int *nowrite, const int *id_orig, int *ID, int *oldsurf, double *nx, double *ny, double *nz, int *d_surfcount)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    if (orig[s] >= nsph_first_fem) return;
    oldsurf[s] = (ID[s] == surf_id) ? 1 : 0;
    int id = id_orig[s];
    V3 nr = v3(nx[s], ny[s], nz[s]);
    nr.x *= inv_totmass; nr.y *= inv_totmass; nr.z *= inv_totmass;
    nx[s] = nr.x; ny[s] = nr.y; nz[s] = nr.z;
    if (norm(nr) >= 0.25 * Q0[s].w && nb[s] <= 46) {
        if (!nowrite[s]) id = surf_id;
        atomicAdd(d_surfcount, 1);
    }
    ID[s] = id;
}

// RIGPairs -> ContPairs (Neighbour.cpp:375-397 with the stale surface IDs, then Contact.cpp:57-132 with the fresh ones):
// per free-surface SPH particle, the rigid-mesh element particles within h_i + h_j, in discovery order.
__global__ void __launch_bounds__(128)
k_cont_list(Grid g, int64_t n, int64_t cap, int nsph, int K, const unsigned long long *keys, const int *cell_start, const int *cell_end,
            const double4 *__restrict__ Q0, const int *__restrict__ orig, const int *__restrict__ ID, const int *__restrict__ oldsurf,
            int surf_id, int require_old, int *cont, int *cont_cnt, int *d_maxcnt)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int cnt = 0;
    if (s < n) {
        const int oi = orig[s];
        if (oi < nsph && ID[s] == surf_id && (!require_old || oldsurf[s])) {
            const double4 qi = Q0[s]; const V3 xi = v3(qi.x, qi.y, qi.z);
            int c[3]; g.unlin((int)(keys[s] >> 32), c);
            for (int e = 0; e < 27; ++e) {
                int d[3] = { c[0] + c_walk[e][0], c[1] + c_walk[e][1], c[2] + c_walk[e][2] };
                if (d[0] < 0 || d[0] >= g.CellNo[0] || d[1] < 0 || d[1] >= g.CellNo[1] || d[2] < 0 || d[2] >= g.CellNo[2]) continue;
                const int l = g.lin(d);
                const int js = cell_start[l], je = cell_end[l];
                const int role = c_walk[e][3];
                for (int j = js; j < je; ++j) {
                    const int oj = orig[j];
                    if (oj < nsph) continue;
                    const double4 qj = Q0[j]; const V3 xj = v3(qj.x, qj.y, qj.z);
                    const int ifirst = (role == 2) ? (j > (int)s ? 1 : 0) : role;
                    const V3 xij = ifirst ? (xi - xj) : (xj - xi);
                    if (norm(xij) < (ifirst ? (qi.w + qj.w) : (qj.w + qi.w))) { if (cnt < K) cont[(int64_t)cnt * cap + s] = oj - nsph; ++cnt; }
                }
            }
        }
        cont_cnt[s] = cnt;
    }
    typedef cub::BlockReduce<int, 128> BR; __shared__ typename BR::TempStorage tmp;
    int m = BR(tmp).Reduce(cnt, cub::Max());
    if (threadIdx.x == 0 && m > 0) atomicMax(d_maxcnt, m);
}

// pairs (i, j > i) of slot s: rows [0, nb_hi) hold them unless the list is in plain discovery order (walk_order), where
// every front row is tested
// CalcGradCorrMatrix, second loop (Domain.cpp:1870-1898): the FSM pairs -- one particle fixed -- add the same term as the SM pairs
// to BOTH members. The neighbour lists hold SM pairs only, so every free particle walks its 27 cells for fixed partners (same
// filters as AllocateNbPair, Neighbour.cpp:374-417) in the reference's discovery order, after its SM sums (k_gradcorr).
__global__ void __launch_bounds__(128)
k_gradcorr_fsm(Grid g, int64_t n, int64_t cap, int nsph, int dim, int kt, const unsigned long long *keys, const int *cell_start, const int *cell_end,
               const double4 *__restrict__ Q0, const double4 *__restrict__ Q2, const int *__restrict__ orig, const int *__restrict__ ID, const int *__restrict__ cls_of,
               const int *__restrict__ is_free, const ClassProps *__restrict__ classes, int contact, int4 mids, double *F)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    if (orig[s] >= nsph || !is_free[s]) return;
    if (contact && is_mesh_id(ID[s], mids)) return;
    const double4 qi = Q0[s]; const V3 xi = v3(qi.x, qi.y, qi.z);
    const int mat_i = classes[cls_of[s]].Material;
    double T[9];
    for (int k = 0; k < 9; ++k) T[k] = F[(int64_t)(F_GC + k) * cap + s];
    int c[3]; g.unlin((int)(keys[s] >> 32), c);
    for (int e = 0; e < 27; ++e) {
        int d[3] = { c[0] + c_walk[e][0], c[1] + c_walk[e][1], c[2] + c_walk[e][2] };
        if (d[0] < 0 || d[0] >= g.CellNo[0] || d[1] < 0 || d[1] >= g.CellNo[1] || d[2] < 0 || d[2] >= g.CellNo[2]) continue;
        const int l = g.lin(d);
        const int js = cell_start[l], je = cell_end[l], role = c_walk[e][3];
        for (int j = js; j < je; ++j) {
            if (j == (int)s || orig[j] >= nsph || is_free[j]) continue;
            if (contact && is_mesh_id(ID[j], mids)) continue;
            const double4 qj = Q0[j]; const V3 xj = v3(qj.x, qj.y, qj.z);
            const int ifirst = (role == 2) ? (j > (int)s ? 1 : 0) : role;
            const V3 xij = ifirst ? (xi - xj) : (xj - xi);
            const double hh = (qi.w + qj.w) / 2;
            const double rij = norm(xij);
            if (!((rij / hh) <= g.Cellfac)) continue;                         // CheckRadius
            if (classes[cls_of[j]].Material != mat_i) continue;
            const double hg = (qi.w + qj.w) / 2.0, q = rij / hg;
            const double GK = (kt == 0 && dim == 3) ? cubicGW_c(1.0 / (hg * hg * hg * hg * WF_PI), q, hg) : gradW(dim, kt, q, hg);
            const double4 q2j = Q2[j];
            const double f = q2j.x / q2j.y;
            const V3 gv = GK * xij;
            const double gx[3] = { gv.x, gv.y, gv.z }, xx[3] = { xij.x, xij.y, xij.z };
            for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) T[3 * r + cc] = T[3 * r + cc] - (gx[r] * xx[cc]) * f;
        }
    }
    for (int k = 0; k < 9; ++k) F[(int64_t)(F_GC + k) * cap + s] = T[k];
}

__global__ void k_export_pairs(int64_t n, int64_t cap, int K, const uint32_t *nbr, const int *nb_hi, const int *orig, const int64_t *offs, int64_t *out, int64_t *count)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int64_t o = out ? offs[s] : 0; int oi = orig[s]; int m = 0;
    for (int k = 0; k < nb_hi[s]; ++k) {
        int j = (int)(nbr[(int64_t)k * cap + s] & 0x7fffffffu);
        if (orig[j] > oi) { if (out) { out[2 * (o + m)] = oi; out[2 * (o + m) + 1] = orig[j]; } ++m; }
    }
    if (count) count[s] = m;
}
__global__ void k_export_cont(int64_t n, int64_t cap, int nsph, const int *cont, const int *cont_cnt, const int *orig, const int64_t *offs, int64_t *out)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int64_t o = offs[s]; int oi = orig[s];
    for (int k = 0; k < cont_cnt[s]; ++k) { out[2 * (o + k)] = oi; out[2 * (o + k) + 1] = (int64_t)nsph + cont[(int64_t)k * cap + s]; }
}
__global__ void k_cnt_to_i64(int64_t n, const int *a, int64_t *b) { int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (s < n) b[s] = a[s]; }

Grid make_grid(const wf_ctx *c)
{
    Grid g; for (int k = 0; k < 3; ++k) { g.BLPF[k] = c->BLPF[k]; g.CellSize[k] = c->CellSize[k]; g.CellNo[k] = c->CellNo[k]; }
    g.Cellfac = c->Cellfac; g.hmax = c->hmax; g.ax[0] = c->ax_fast; g.ax[1] = c->ax_mid; g.ax[2] = c->ax_slow; return g;
}

} // namespace

namespace wf {

// Neighbour.cpp:10-139 CellInitiate (3-D, non-periodic). Done once per run, as in the reference; a one-off host pass
// over a downloaded copy of the positions (setup, not hot path). Multi-GPU: bounding box and hmax are all-reduced, so
// every rank builds the same global grid; the slab (cell layers of the slowest axis) each rank owns is balanced by count.
int nb_cell_initiate(wf_ctx *c)
{
    const int64_t n = c->n_loc;
    std::vector<double4> q(n);
    WF_CUDA(c, cudaMemcpy(q.data(), c->Q0[c->cur], n * sizeof(double4), cudaMemcpyDeviceToHost));
    double BL[3] = { q[0].x, q[0].y, q[0].z }, TR[3] = { q[0].x, q[0].y, q[0].z }; double hmax = q[0].w;
    for (int64_t s = 0; s < n; ++s) {
        double x[3] = { q[s].x, q[s].y, q[s].z };
        for (int k = 0; k < 3; ++k) { if (x[k] > TR[k]) TR[k] = x[k]; if (x[k] < BL[k]) BL[k] = x[k]; }
        if (q[s].w > hmax) hmax = q[s].w;
    }
    WF_CHECK(comm_global_grid(c, BL, TR, &hmax));
    const double pBL[3] = { BL[0], BL[1], BL[2] }, pTR[3] = { TR[0], TR[1], TR[2] };      // extent the particles occupy
    for (int k = 0; k < 3; ++k) { if (c->prm.dom_max[k] > TR[k]) TR[k] = c->prm.dom_max[k]; if (c->prm.dom_min[k] < BL[k]) BL[k] = c->prm.dom_min[k]; }
    for (int k = 0; k < 3; ++k) { TR[k] += hmax / 2; BL[k] -= hmax / 2; }
    int64_t nc = 1;
    for (int k = 0; k < 3; ++k) {
        double qq = (TR[k] - BL[k]) / (c->Cellfac * hmax);
        if ((double)(std::ceil(qq) - qq) < (hmax / 10.0)) c->CellNo[k] = (int)std::ceil(qq); else c->CellNo[k] = (int)std::floor(qq);
        if (c->CellNo[k] <= 0) return fail(c, WF_ERR_ARG, "CellInitiate: empty cell grid along axis %d (Neighbour.cpp:108)", k);
        c->CellSize[k] = (TR[k] - BL[k]) / c->CellNo[k];
        c->BLPF[k] = BL[k]; c->TRPR[k] = TR[k];
        nc *= c->CellNo[k];
    }
    if (nc >= ((int64_t)1 << 31) - 2) return fail(c, WF_ERR_UNSUPPORTED, "cell grid too large (%lld cells)", (long long)nc);
    c->hmax = hmax; c->ncells = nc;
    // storage order of cells: the axis along which the particles occupy most cell layers is the slowest (the slab
    // decomposition splits along it), the one with the fewest the fastest
    int ax[3] = { 0, 1, 2 }; double occ[3];
    for (int k = 0; k < 3; ++k) occ[k] = (pTR[k] - pBL[k]) / c->CellSize[k];
    std::sort(ax, ax + 3, [&](int a, int b) { return occ[a] < occ[b] || (occ[a] == occ[b] && a < b); });
    c->ax_fast = ax[0]; c->ax_mid = ax[1]; c->ax_slow = ax[2];
    cudaFree(c->cell_start); cudaFree(c->cell_end);
    WF_CUDA(c, cudaMalloc(&c->cell_start, nc * sizeof(int))); WF_CUDA(c, cudaMalloc(&c->cell_end, nc * sizeof(int)));
    c->cell_cap = nc;
    if (c->mesh_tail) {
        if (nc >= (int64_t)WF_MESH_CELL_BIT) return fail(c, WF_ERR_UNSUPPORTED, "cell grid too large for the mesh cell table (%lld cells)", (long long)nc);
        cudaFree(c->mcell_start); cudaFree(c->mcell_end);
        WF_CUDA(c, cudaMalloc(&c->mcell_start, nc * sizeof(int))); WF_CUDA(c, cudaMalloc(&c->mcell_end, nc * sizeof(int)));
    }
    if (c->comm.nranks > 1) {
        const int slow = c->ax_slow, L = c->CellNo[slow];
        std::vector<long long> hist(L, 0);
        for (int64_t s = 0; s < n - c->nmeshp; ++s) {          // SPH particles only: element particles are replicated, not owned
            const double xs = slow == 0 ? q[s].x : (slow == 1 ? q[s].y : q[s].z);
            int ci = (int)std::floor((xs - c->BLPF[slow]) / c->CellSize[slow]);
            if (ci < 0) ci = 0; if (ci >= L) ci = L - 1;
            hist[ci]++;
        }
        WF_CHECK(comm_plan_layers(c, hist));
        c->comm.emig_cap = std::max<int64_t>(4096, n / 16);
        cudaFree(c->comm.d_emig); WF_CUDA(c, cudaMalloc(&c->comm.d_emig, 2 * c->comm.emig_cap * sizeof(int)));
    }
    return 0;
}

int nb_gradcorr_fsm(wf_ctx *c)
{
    if (!c->any_fixed) return 0;
    Grid g = make_grid(c);
    int4 mids = make_int4(-2147483647, -2147483647, -2147483647, -2147483647);
    if (c->mesh_ids.size() > 0) mids.x = c->mesh_ids[0]; if (c->mesh_ids.size() > 1) mids.y = c->mesh_ids[1];
    if (c->mesh_ids.size() > 2) mids.z = c->mesh_ids[2]; if (c->mesh_ids.size() > 3) mids.w = c->mesh_ids[3];
    k_gradcorr_fsm<<<grid_for(c->n_loc, 128), 128, 0, c->stream>>>(g, c->n_loc, c->cap, (int)c->nsph, c->prm.dimension, c->prm.kernel_type, c->keys[1], c->cell_start, c->cell_end,
        c->Q0[c->cur], c->Q2[c->cur], c->Ip(I_ORIG), c->Ip(I_ID), c->Ip(I_CLS), c->Ip(I_FREE), c->d_classes, c->prm.contact ? 1 : 0, mids, c->F[c->cur]);
    WF_LAUNCH_CHECK(c); c->launches++;
    return 0;
}

int nb_build_impl(wf_ctx *c)
{
    const int64_t cap = c->cap; const int B = 256;
    cudaStream_t st = c->stream;
    Grid g = make_grid(c);
    const bool multi = c->comm.nranks > 1;
    int64_t n = c->n_loc, n_out = 0;                  // slots to sort; emigrants among them
    const int a = c->cur, b = 1 - c->cur;
    if (multi && c->solver_ready) WF_CHECK(comm_rebalance(c));       // may move this rank's layer bounds by one layer (the cell-key pass below migrates)
    MigArgs mg; mg.on = multi ? 1 : 0; mg.layer_lo = c->comm.layer_lo; mg.layer_hi = c->comm.layer_hi;
    mg.plane = (long long)c->CellNo[c->ax_mid] * c->CellNo[c->ax_fast]; mg.emig = c->comm.d_emig; mg.cnt = c->comm.d_cnt; mg.emig_cap = c->comm.emig_cap;
    mg.mesh_tail = c->mesh_tail ? 1 : 0; mg.nsph = (int)c->nsph;
    {   ProfScope ps(c, PK_NB_KEYS);
        if (multi) WF_CUDA(c, cudaMemsetAsync(c->comm.d_cnt, 0, 4 * sizeof(int), st));
        k_cell_keys<<<grid_for(n, B), B, 0, st>>>(g, c->Q0[a], c->Ip(I_ORIG), n, 0, c->keys[0], c->vals[0], mg); WF_LAUNCH_CHECK(c); c->launches++;
        if (multi) {
            // migration: emigrants out, immigrants appended at [n_loc, n); their keys computed here (they are inside the slab by construction)
            int64_t n_sort = n;
            WF_CHECK(comm_migrate(c, &n_sort, &n_out));
            if (n_sort > n) { MigArgs off = mg; off.on = 0; off.mesh_tail = 0;
                k_cell_keys<<<grid_for(n_sort - n, B), B, 0, st>>>(g, c->Q0[a], c->Ip(I_ORIG), n_sort - n, n, c->keys[0], c->vals[0], off); WF_LAUNCH_CHECK(c); c->launches++; }
            n = n_sort;
        }
    }
    {   ProfScope ps(c, PK_NB_SORT);
        int cellbits = 1; while (((int64_t)1 << cellbits) < c->ncells) ++cellbits;
        const int end_bit = (multi || c->mesh_tail) ? 64 : 32 + cellbits;
        size_t need = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, need, c->keys[0], c->keys[1], c->vals[0], c->vals[1], (int)n, 0, end_bit, st);
        if (need > c->cub_tmp_bytes) { cudaFree(c->cub_tmp); c->cub_tmp = nullptr; WF_CUDA(c, cudaMalloc(&c->cub_tmp, need)); c->cub_tmp_bytes = need; }
        WF_CUDA(c, cub::DeviceRadixSort::SortPairs(c->cub_tmp, need, c->keys[0], c->keys[1], c->vals[0], c->vals[1], (int)n, 0, end_bit, st));
        c->launches += 4; }
    {   ProfScope ps(c, PK_NB_PERMUTE);
        const int *perm = c->vals[1];
        const int64_t keep = n - n_out;               // emigrants carry the sentinel key: they are the tail and are dropped
        k_permute_q<<<grid_for(keep, B), B, 0, st>>>(keep, perm, c->Q0[a], c->Q1[a], c->Q2[a], c->S0[a], c->S1[a], c->S2[a], c->Q0[b], c->Q1[b], c->Q2[b], c->S0[b], c->S1[b], c->S2[b]); WF_LAUNCH_CHECK(c);
        // gradCorrM (the last 9 own fields) lives only with gradKernelCorr
        const int nf = c->prm.grad_kernel_corr ? F_COUNT : F_GC;
        dim3 gf(grid_for(keep, B), (nf + WF_PERM_FIELDS - 1) / WF_PERM_FIELDS); k_permute_f<<<gf, B, 0, st>>>(keep, cap, nf, perm, c->F[a], c->F[b]); WF_LAUNCH_CHECK(c);
        dim3 gi(grid_for(keep, B), I_COUNT); k_permute_i<<<gi, B, 0, st>>>(keep, cap, perm, c->I[a], c->I[b], multi ? nullptr : c->slot_of); WF_LAUNCH_CHECK(c);
        c->cur = b; c->launches += 3;
        c->n_loc = keep; c->n = keep; n = keep; }
    if (multi) {
        ProfScope ps(c, PK_HALO);
        WF_CHECK(comm_halo_setup(c, c->keys[1]));     // boundary slices, halo regions filled: c->n = n_loc + halos
        if (c->n > c->n_loc) { MigArgs off = mg; off.on = 0; off.mesh_tail = 0;
            k_cell_keys<<<grid_for(c->n - c->n_loc, B), B, 0, st>>>(g, c->Q0[c->cur], c->Ip(I_ORIG), c->n - c->n_loc, c->n_loc, c->keys[1], c->vals[0], off); WF_LAUNCH_CHECK(c); c->launches++; }
        WF_CHECK(comm_local_index(c));
        WF_CHECK(ghost_localize(c));                  // ghost pairs as local indices of the particles this rank holds now
    }
    {   ProfScope ps(c, PK_NB_CELLS);
        WF_CUDA(c, cudaMemsetAsync(c->cell_start, 0, c->ncells * sizeof(int), st)); WF_CUDA(c, cudaMemsetAsync(c->cell_end, 0, c->ncells * sizeof(int), st));
        if (c->mesh_tail) { WF_CUDA(c, cudaMemsetAsync(c->mcell_start, 0, c->ncells * sizeof(int), st)); WF_CUDA(c, cudaMemsetAsync(c->mcell_end, 0, c->ncells * sizeof(int), st)); }
        k_cell_bounds<<<grid_for(c->n, B), B, 0, st>>>(c->n, c->keys[1], c->cell_start, c->cell_end, c->mcell_start, c->mcell_end); WF_LAUNCH_CHECK(c); c->launches++; }
    n = c->n_loc;                                     // lists are built for this rank's own particles; candidates include the halos
    int4 mids = make_int4(-2147483647, -2147483647, -2147483647, -2147483647);
    if (c->mesh_ids.size() > 0) mids.x = c->mesh_ids[0]; if (c->mesh_ids.size() > 1) mids.y = c->mesh_ids[1];
    if (c->mesh_ids.size() > 2) mids.z = c->mesh_ids[2]; if (c->mesh_ids.size() > 3) mids.w = c->mesh_ids[3];
    bool check_material = false;
    for (int k = 0; k < c->n_classes; ++k) if (!c->h_classes[k].IsMesh && c->h_classes[k].Material != c->h_classes[0].Material) check_material = true;
    if (c->nbr_k == 0) c->nbr_k = 64;
    for (int attempt = 0; attempt < 4; ++attempt) {
        ProfScope ps(c, PK_NB_LIST);
        const int K = c->nbr_k;
        if ((int64_t)K * cap > c->nbr_alloc) { cudaFree(c->nbr); c->nbr = nullptr; WF_CUDA(c, cudaMalloc(&c->nbr, (size_t)K * cap * sizeof(uint32_t))); c->nbr_alloc = (int64_t)K * cap; }
        WF_CUDA(c, cudaMemsetAsync(c->d_maxcnt, 0, 4 * sizeof(int), st));
        k_nb_list<<<grid_for(n, 128), 128, 0, st>>>(g, n, cap, (int)c->nsph, K, c->keys[1], c->cell_start, c->cell_end, c->Q0[c->cur], c->Q2[c->cur], c->Ip(I_ORIG), c->Ip(I_ID),
            c->Ip(I_CLS), c->any_fixed ? c->Ip(I_FREE) : nullptr, c->d_classes, c->prm.contact ? 1 : 0, mids, check_material ? 1 : 0, c->uniform_h ? 1 : 0, c->prm.grad_kernel_corr ? 1 : 0, c->nbr, c->nb_cnt, c->nb_hi, c->Ip(I_NB), c->Fp(F_NX), c->Fp(F_NY), c->Fp(F_NZ), c->d_maxcnt);
        WF_LAUNCH_CHECK(c); c->launches++;
        int mx = 0; WF_CUDA(c, cudaMemcpyAsync(&mx, c->d_maxcnt, sizeof(int), cudaMemcpyDeviceToHost, st)); WF_CUDA(c, cudaStreamSynchronize(st));
        if (mx <= K) break;
        c->nbr_k = ((mx + 7) / 8) * 8;       // no silent MAX_NB_PER_PART cap (Domain.h:49): grow and redo
        if (attempt == 3) return fail(c, WF_ERR_NOMEM, "neighbour list did not fit after regrowing (max count %d)", mx);
    }
    if (c->prm.contact) {
        {   ProfScope ps(c, PK_NB_SURFACE);
            WF_CUDA(c, cudaMemsetAsync(c->d_maxcnt + 1, 0, sizeof(int), st));
            const int first_fem = c->meshes.empty() ? (int)c->n : (int)c->meshes[0].first_particle;
            k_surface<<<grid_for(n, B), B, 0, st>>>(n, first_fem, 1. / c->totmass_mean, 1, c->Q0[c->cur], c->Ip(I_ORIG), c->Ip(I_NB), c->Ip(I_NOWRITE), c->Ip(I_IDORIG), c->Ip(I_ID),
                c->Ip(I_OLDSURF), c->Fp(F_NX), c->Fp(F_NY), c->Fp(F_NZ), c->d_maxcnt + 1);
            WF_LAUNCH_CHECK(c); c->launches++; }
        if (c->cont_k == 0) c->cont_k = 16;
        for (int attempt = 0; attempt < 4; ++attempt) {
            ProfScope ps(c, PK_NB_CONTACT);
            const int K = c->cont_k;
            if ((int64_t)K * cap > c->cont_alloc) { cudaFree(c->cont); c->cont = nullptr; WF_CUDA(c, cudaMalloc(&c->cont, (size_t)K * cap * sizeof(int))); c->cont_alloc = (int64_t)K * cap; }
            WF_CUDA(c, cudaMemsetAsync(c->d_maxcnt + 2, 0, sizeof(int), st));
            // mesh_tail layout: the element particles have a cell table of their own (same walk, same order inside a cell)
            k_cont_list<<<grid_for(n, 128), 128, 0, st>>>(g, n, cap, (int)c->nsph, K, c->keys[1], c->mesh_tail ? c->mcell_start : c->cell_start, c->mesh_tail ? c->mcell_end : c->cell_end, c->Q0[c->cur], c->Ip(I_ORIG), c->Ip(I_ID), c->Ip(I_OLDSURF),
                1, c->solver_ready ? 1 : 0, c->cont, c->cont_cnt, c->d_maxcnt + 2);
            WF_LAUNCH_CHECK(c); c->launches++;
            int mx[2] = { 0, 0 }; WF_CUDA(c, cudaMemcpyAsync(mx, c->d_maxcnt + 1, 2 * sizeof(int), cudaMemcpyDeviceToHost, st)); WF_CUDA(c, cudaStreamSynchronize(st));
            if (attempt == 0) {   // the reference counts the surface particles of the whole domain (Domain.cpp:1489); a slab may well own none. Every
                // rank takes the same decision (once per build: the retries below are rank-local), so that none is left waiting in the next collective.
                double lo = (double)mx[0], hi = (double)mx[0];
                WF_CHECK(comm_global_minmax(c, &lo, &hi));
                if (hi == 0.0) return fail(c, WF_ERR_ARG, "ERROR: No external particles found. Please check particle masses (Domain.cpp:1489)");
            }
            if (mx[1] <= K) break;
            c->cont_k = ((mx[1] + 7) / 8) * 8;
            if (attempt == 3) return fail(c, WF_ERR_NOMEM, "contact candidate list did not fit after regrowing (max count %d)", mx[1]);
        }
    }
    c->nb_cleared = false;
    c->dens_valid = false;           // a density rate summed over the old list is not what the reference would compute now
    c->nb_rebuilds++;
    return 0;
}

} // namespace wf

static int export_common(wf_ctx *c, bool contact_pairs, int64_t *ij, int64_t capn, int64_t *n_pairs)
{
    if (!c || !n_pairs) return WF_ERR_ARG;
    if (!c->nbr) return fail(c, WF_ERR_ARG, "export pairs: no neighbour list built yet");
    if (contact_pairs && !c->cont) { *n_pairs = 0; return 0; }
    // ClearNbData (Neighbour.cpp:479-497) empties SMPairs/ContPairs at the end of a cadence window; mirror that state
    if (c->nb_cleared) { *n_pairs = 0; return 0; }
    cudaSetDevice(c->device);
    const int64_t n = c->n_loc; cudaStream_t st = c->stream;
    int64_t *d_cnt = nullptr, *d_off = nullptr;
    WF_CUDA(c, cudaMalloc(&d_cnt, (n + 1) * sizeof(int64_t))); WF_CUDA(c, cudaMalloc(&d_off, (n + 1) * sizeof(int64_t)));
    WF_CUDA(c, cudaMemsetAsync(d_cnt, 0, (n + 1) * sizeof(int64_t), st));
    if (contact_pairs) k_cnt_to_i64<<<grid_for(n, 256), 256, 0, st>>>(n, c->cont_cnt, d_cnt);
    else k_export_pairs<<<grid_for(n, 256), 256, 0, st>>>(n, c->cap, c->nbr_k, c->nbr, c->nb_hi, c->Ip(I_ORIG), nullptr, nullptr, d_cnt);
    size_t need = 0; cub::DeviceScan::ExclusiveSum(nullptr, need, d_cnt, d_off, (int)(n + 1), st);
    void *tmp = nullptr; WF_CUDA(c, cudaMalloc(&tmp, need));
    cub::DeviceScan::ExclusiveSum(tmp, need, d_cnt, d_off, (int)(n + 1), st);
    int64_t total = 0; WF_CUDA(c, cudaMemcpyAsync(&total, d_off + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st)); WF_CUDA(c, cudaStreamSynchronize(st));
    *n_pairs = total;
    int rc = 0;
    if (ij && total > 0) {
        if (total > capn) rc = fail(c, WF_ERR_ARG, "export pairs: buffer holds %lld pairs, need %lld", (long long)capn, (long long)total);
        else {
            int64_t *d_out = nullptr; WF_CUDA(c, cudaMalloc(&d_out, 2 * total * sizeof(int64_t)));
            if (contact_pairs) k_export_cont<<<grid_for(n, 256), 256, 0, st>>>(n, c->cap, (int)c->nsph, c->cont, c->cont_cnt, c->Ip(I_ORIG), d_off, d_out);
            else k_export_pairs<<<grid_for(n, 256), 256, 0, st>>>(n, c->cap, c->nbr_k, c->nbr, c->nb_hi, c->Ip(I_ORIG), d_off, d_out, nullptr);
            std::vector<int64_t> h(2 * total);
            WF_CUDA(c, cudaMemcpyAsync(h.data(), d_out, 2 * total * sizeof(int64_t), cudaMemcpyDeviceToHost, st)); WF_CUDA(c, cudaStreamSynchronize(st));
            cudaFree(d_out);
            std::vector<std::pair<int64_t, int64_t> > pr(total);
            for (int64_t k = 0; k < total; ++k) pr[k] = std::make_pair(std::min(h[2 * k], h[2 * k + 1]), std::max(h[2 * k], h[2 * k + 1]));
            std::sort(pr.begin(), pr.end());
            for (int64_t k = 0; k < total; ++k) { ij[2 * k] = pr[k].first; ij[2 * k + 1] = pr[k].second; }
        }
    }
    cudaFree(tmp); cudaFree(d_cnt); cudaFree(d_off);
    return rc;
}
extern "C" int wf_nb_export_pairs(wf_ctx *c, int64_t *ij, int64_t cap, int64_t *n_pairs) { return export_common(c, false, ij, cap, n_pairs); }
extern "C" int wf_cont_export_pairs(wf_ctx *c, int64_t *ij, int64_t cap, int64_t *n_pairs) { return export_common(c, true, ij, cap, n_pairs); }
