// Multi-GPU support: one process per GPU, 1-D slab decomposition of the reference's cell grid along its longest axis,
// NCCL point-to-point halos over NVLink. The reference is single-process ("distributed
// communication backend: none"); this is the B200-native scaling path of BASELINE.json's north_star.
//
//   * ownership: rank r owns the cell layers [layer_lo, layer_hi) of the slowest axis, balanced by particle count;
//   * at every neighbour rebuild: particles that left the slab migrate to the adjacent rank (whole state, packed), the
//     local region is re-sorted by cell, and the one-cell-thick boundary layers -- CONTIGUOUS slices of the cell-sorted
//     arrays, because the slab axis is the most significant sort digit -- are sent to the neighbours' halo regions;
//   * per step: three refreshes of those same slices, dictated by the data dependencies of the Fraser step
//       m/rho            (after the density update, before the rate-tensor pass)
//       m,rho,P/rho^2, sigma/rho^2, TIR   (after the constitutive update, before the momentum pass)
//       x, v, T          (after the move / temperature update, before the conduction + energy + next-density pass)
//     zero-copy: ncclSend/ncclRecv straight from/into the particle arrays, no pack/unpack kernels;
//   * global scalars (max pl_strain, max_disp, energies ...): one 32-double all-gather per step, reduced in rank order
//     on every rank (deterministic, identical on all ranks).
// Neighbour lists keep the reference's discovery order (it depends only on global cell coordinates and creation
// indices), so every per-particle result is bit-identical to the single-GPU run.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- inside a torch process that is torch's own copy), so the library
// has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>
#include <cub/cub.cuh>
#include <algorithm>
#include <cstring>
#include <vector>

#include "wf_internal.cuh"

using namespace wf;

namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false; std::string err;
};

NcclApi *nccl_api()
{
    static NcclApi a; static bool tried = false;
    if (tried) return &a;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { a.err = std::string("cannot load libnccl.so.2: ") + dlerror(); return &a; }
#define WF_SYM(name) do { *(void **)(&a.name) = dlsym(h, "nccl" #name); if (!a.name) { a.err = "libnccl.so.2 lacks nccl" #name; return &a; } } while (0)
    WF_SYM(GetUniqueId); WF_SYM(CommInitRank); WF_SYM(CommDestroy); WF_SYM(Send); WF_SYM(Recv); WF_SYM(AllGather); WF_SYM(AllReduce);
    WF_SYM(GroupStart); WF_SYM(GroupEnd); WF_SYM(GetErrorString);
#undef WF_SYM
    a.ok = true;
    return &a;
}

#define WF_NCCL(c, expr) do { ncclResult_t r_ = (expr); if (r_ != ncclSuccess) return wf::fail((c), WF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, nccl_api()->GetErrorString(r_), __FILE__, __LINE__); } while (0)

struct QPtrs { double4 *q[6]; };

// whole-state pack / unpack of migrating particles: doubles [field][k]; fields = 24 gathered components, F_COUNT own
// doubles, I_COUNT own ints (stored as doubles, exact)
__global__ void k_pack(QPtrs Q, const double *F, const int *I, int64_t cap, const int *slots, int n, double *out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x; if (k >= n) return;
    const int f = blockIdx.y; const int64_t s = slots[k];
    double v;
    if (f < 24) v = ((const double *)&Q.q[f >> 2][s])[f & 3];
    else if (f < 24 + F_COUNT) v = F[(int64_t)(f - 24) * cap + s];
    else v = (double)I[(int64_t)(f - 24 - F_COUNT) * cap + s];
    out[(int64_t)f * n + k] = v;
}
__global__ void k_unpack(QPtrs Q, double *F, int *I, int64_t cap, int64_t base, int n, const double *in)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x; if (k >= n) return;
    const int f = blockIdx.y; const int64_t s = base + k;
    const double v = in[(int64_t)f * n + k];
    if (f < 24) ((double *)&Q.q[f >> 2][s])[f & 3] = v;
    else if (f < 24 + F_COUNT) F[(int64_t)(f - 24) * cap + s] = v;
    else I[(int64_t)(f - 24 - F_COUNT) * cap + s] = (int)v;
}
// first slot whose sorted key has a cell index >= target (binary search, one thread per target)
__global__ void k_lower_bound(const unsigned long long *keys, int64_t n, const long long *targets, int nt, long long *out)
{
    const int t = threadIdx.x; if (t >= nt) return;
    const unsigned long long tk = (unsigned long long)targets[t] << 32;
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (keys[mid] < tk) lo = mid + 1; else hi = mid; }
    out[t] = lo;
}
__global__ void k_iota_orig(int64_t n, const int *orig, int *keys, int *vals)
{ int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (s < n) { keys[s] = orig[s]; vals[s] = (int)s; } }
__global__ void k_set_lid(int64_t n, const int *sorted_slots, int *slot_of, int *lid)
{ int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (k < n) { const int s = sorted_slots[k]; slot_of[k] = s; lid[s] = (int)k; } }

QPtrs qptrs(wf_ctx *c, int buf) { QPtrs q; q.q[0] = c->Q0[buf]; q.q[1] = c->Q1[buf]; q.q[2] = c->Q2[buf]; q.q[3] = c->S0[buf]; q.q[4] = c->S1[buf]; q.q[5] = c->S2[buf]; return q; }

} // namespace

namespace wf {

// balanced contiguous split of the slow-axis layers by particle count (same code for the library and the host planner)
static void balance_layers(const std::vector<long long> &hist, int nranks, std::vector<int> &bounds)
{
    const int L = (int)hist.size();
    std::vector<long long> cum(L + 1, 0);
    for (int l = 0; l < L; ++l) cum[l + 1] = cum[l] + hist[l];
    const long long total = cum[L];
    bounds.assign(nranks + 1, 0); bounds[nranks] = L;
    for (int r = 1; r < nranks; ++r) {
        // the layer boundary whose cumulative count is NEAREST to r/nranks of the particles (the step time is set by the fullest
        // rank), keeping at least one layer for every rank before and after
        const int lo = bounds[r - 1] + 1, hi = L - (nranks - r);
        int best = lo; long long bestd = -1;
        for (int b = lo; b <= hi; ++b) {
            long long d = cum[b] * nranks - (long long)r * total; if (d < 0) d = -d;
            if (bestd < 0 || d < bestd) { bestd = d; best = b; }
        }
        bounds[r] = best;
    }
}

int comm_global_grid(wf_ctx *c, double BL[3], double TR[3], double *hmax)
{
    if (c->comm.nranks <= 1) return 0;
    NcclApi *A = nccl_api();
    double h[8] = { -BL[0], -BL[1], -BL[2], TR[0], TR[1], TR[2], *hmax, 0.0 };
    WF_CHECK(ensure_stage(c, 256));
    WF_CUDA(c, cudaMemcpyAsync(c->stage, h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
    WF_NCCL(c, A->AllReduce(c->stage, c->stage, 8, ncclDouble, ncclMax, (ncclComm_t)c->comm.nccl, c->stream));
    WF_CUDA(c, cudaMemcpyAsync(h, c->stage, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    WF_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 3; ++k) { BL[k] = -h[k]; TR[k] = h[3 + k]; }
    *hmax = h[6];
    return 0;
}

// Property classes are deduplicated per rank from that rank's particles, so the same tuple can get different indices on different ranks --
// and the class index of a HALO particle is read by its neighbours (k_T, Alpha, TI ... of the other particle of a pair). Every rank therefore
// adopts one table: the union of all ranks' classes in rank order; remap[old local index] = index in the union.
int comm_unify_classes(wf_ctx *c, std::vector<ClassProps> &cls, std::vector<int> &remap)
{
    wf_ctx::Comm &m = c->comm;
    remap.resize(cls.size()); for (size_t k = 0; k < cls.size(); ++k) remap[k] = (int)k;
    if (m.nranks <= 1) return 0;
    static_assert(sizeof(ClassProps) % sizeof(double) == 0, "ClassProps travels as doubles");
    const int MAXC = 256, per = (int)(sizeof(ClassProps) / sizeof(double)), rec = 1 + MAXC * per;
    if ((int)cls.size() > MAXC) return fail(c, WF_ERR_UNSUPPORTED, "more than %d distinct per-particle property tuples on one rank of a multi-GPU run", MAXC);
    NcclApi *A = nccl_api();
    std::vector<double> snd((size_t)rec, 0.0), all((size_t)rec * m.nranks);
    snd[0] = (double)cls.size();
    if (!cls.empty()) std::memcpy(&snd[1], cls.data(), cls.size() * sizeof(ClassProps));
    double *d_s = nullptr, *d_r = nullptr;
    WF_CUDA(c, cudaMalloc(&d_s, snd.size() * sizeof(double))); WF_CUDA(c, cudaMalloc(&d_r, all.size() * sizeof(double)));
    WF_CUDA(c, cudaMemcpyAsync(d_s, snd.data(), snd.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    WF_NCCL(c, A->AllGather(d_s, d_r, (size_t)rec, ncclDouble, (ncclComm_t)m.nccl, c->stream));
    WF_CUDA(c, cudaMemcpyAsync(all.data(), d_r, all.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    WF_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(d_s); cudaFree(d_r);
    std::vector<ClassProps> uni;
    for (int r = 0; r < m.nranks; ++r) {
        const double *b = &all[(size_t)r * rec]; const int nr = (int)b[0];
        for (int k = 0; k < nr; ++k) {
            ClassProps cp; std::memcpy(&cp, b + 1 + (size_t)k * per, sizeof(cp));
            int found = -1;
            for (size_t u = 0; u < uni.size() && found < 0; ++u) if (std::memcmp(&uni[u], &cp, sizeof(cp)) == 0) found = (int)u;
            if (found < 0) { found = (int)uni.size(); uni.push_back(cp); }
            if (r == m.rank) remap[(size_t)k] = found;
        }
    }
    cls.swap(uni);
    return 0;
}

// min / max of a per-rank value over all ranks (e.g. "is the particle mass the same on every rank?")
int comm_global_minmax(wf_ctx *c, double *mn, double *mx)
{
    if (c->comm.nranks <= 1) return 0;
    NcclApi *A = nccl_api();
    double h[2] = { -*mn, *mx };
    WF_CHECK(ensure_stage(c, 256));
    WF_CUDA(c, cudaMemcpyAsync(c->stage, h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
    WF_NCCL(c, A->AllReduce(c->stage, c->stage, 2, ncclDouble, ncclMax, (ncclComm_t)c->comm.nccl, c->stream));
    WF_CUDA(c, cudaMemcpyAsync(h, c->stage, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    WF_CUDA(c, cudaStreamSynchronize(c->stream));
    *mn = -h[0]; *mx = h[1];
    return 0;
}

// A rank-local failure inside a collective section (capacity checks of the migration and the halo set-up) must become every rank's failure: a rank
// that simply returned would leave its peers blocked in the next send / receive. All ranks call this with their local status (0 or a negative
// wf_status, message already recorded by fail()); all return the same: 0, or the failing rank's status on that rank and WF_ERR_ARG elsewhere.
int comm_agree(wf_ctx *c, int local_status, const char *what)
{
    if (c->comm.nranks <= 1) return local_status;
    double lo = (double)local_status, hi = (double)local_status;
    WF_CHECK(comm_global_minmax(c, &lo, &hi));
    if (lo == 0.0) return 0;
    if (local_status != 0) return local_status;
    return fail(c, WF_ERR_ARG, "%s failed on another GPU rank (status %d); this rank stops with it", what, (int)lo);
}

// CalculateSurface (Domain.cpp:1421-1425) sums the particle masses serially in creation order; with the particles spread over
// ranks every rank scatters its masses into a zeroed creation-ordered array, the arrays are summed (exact: one non-zero
// contributor per entry) and each rank repeats the reference's serial sum over the result.
static __global__ void k_scatter_mass(const double *m, const long long *gidx, long long n, double *all)
{ long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) all[gidx[i]] = m[i]; }
int comm_serial_mass_sum(wf_ctx *c, const double *mass, const int64_t *gidx, int64_t n_local, int64_t n_global, double *tot)
{
    NcclApi *A = nccl_api();
    double *d_all = nullptr, *d_m = nullptr; long long *d_g = nullptr;
    WF_CUDA(c, cudaMalloc(&d_all, n_global * sizeof(double))); WF_CUDA(c, cudaMalloc(&d_m, n_local * sizeof(double))); WF_CUDA(c, cudaMalloc(&d_g, n_local * sizeof(long long)));
    WF_CUDA(c, cudaMemsetAsync(d_all, 0, n_global * sizeof(double), c->stream));
    WF_CUDA(c, cudaMemcpyAsync(d_m, mass, n_local * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    WF_CUDA(c, cudaMemcpyAsync(d_g, gidx, n_local * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    k_scatter_mass<<<grid_for(n_local, 256), 256, 0, c->stream>>>(d_m, d_g, n_local, d_all); WF_LAUNCH_CHECK(c);
    WF_NCCL(c, A->AllReduce(d_all, d_all, (size_t)n_global, ncclDouble, ncclSum, (ncclComm_t)c->comm.nccl, c->stream));
    std::vector<double> h(n_global);
    WF_CUDA(c, cudaMemcpyAsync(h.data(), d_all, n_global * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    WF_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(d_all); cudaFree(d_m); cudaFree(d_g);
    double t = 0.; for (int64_t i = 0; i < n_global; ++i) t += h[i];
    *tot = t;
    return 0;
}

int comm_plan_layers(wf_ctx *c, const std::vector<long long> &hist_local)
{
    if (c->comm.nranks <= 1) return 0;
    NcclApi *A = nccl_api();
    const size_t L = hist_local.size();
    if ((int)L < c->comm.nranks) return fail(c, WF_ERR_ARG, "slab decomposition: %d cell layers along the longest axis for %d ranks", (int)L, c->comm.nranks);
    WF_CHECK(ensure_stage(c, L * sizeof(long long)));
    std::vector<long long> hist(hist_local);
    WF_CUDA(c, cudaMemcpyAsync(c->stage, hist.data(), L * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    WF_NCCL(c, A->AllReduce(c->stage, c->stage, L, ncclInt64, ncclSum, (ncclComm_t)c->comm.nccl, c->stream));
    WF_CUDA(c, cudaMemcpyAsync(hist.data(), c->stage, L * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    WF_CUDA(c, cudaStreamSynchronize(c->stream));
    std::vector<int> b; balance_layers(hist, c->comm.nranks, b);
    c->comm.layer_lo = b[c->comm.rank]; c->comm.layer_hi = b[c->comm.rank + 1];
    return 0;
}

// sends/receives the boundary slices of the selected gathered arrays (zero-copy); halo regions were sized at the rebuild
static int halo_exchange(wf_ctx *c, int sets, cudaStream_t st)
{
    wf_ctx::Comm &m = c->comm;
    NcclApi *A = nccl_api();
    double4 *arr[6]; int na = 0; const int b = c->cur;
    if (sets & HALO_Q0) arr[na++] = c->Q0[b];
    if (sets & HALO_Q1) arr[na++] = c->Q1[b];
    if (sets & HALO_Q2) arr[na++] = c->Q2[b];
    if (sets & HALO_S) { arr[na++] = c->S0[b]; arr[na++] = c->S1[b]; arr[na++] = c->S2[b]; }
    ncclComm_t comm = (ncclComm_t)m.nccl;
    WF_NCCL(c, A->GroupStart());
    for (int k = 0; k < na; ++k) {
        if (m.rank > 0) {
            if (m.send_lo_n) WF_NCCL(c, A->Send(arr[k] + m.send_lo_start, 4 * m.send_lo_n, ncclDouble, m.rank - 1, comm, st));
            if (m.halo_lo_n) WF_NCCL(c, A->Recv(arr[k] + m.halo_lo_start, 4 * m.halo_lo_n, ncclDouble, m.rank - 1, comm, st));
        }
        if (m.rank < m.nranks - 1) {
            if (m.send_hi_n) WF_NCCL(c, A->Send(arr[k] + m.send_hi_start, 4 * m.send_hi_n, ncclDouble, m.rank + 1, comm, st));
            if (m.halo_hi_n) WF_NCCL(c, A->Recv(arr[k] + m.halo_hi_start, 4 * m.halo_hi_n, ncclDouble, m.rank + 1, comm, st));
        }
    }
    WF_NCCL(c, A->GroupEnd());
    c->launches += 1;
    return 0;
}
int comm_halo_refresh(wf_ctx *c, int sets)
{
    if (c->comm.nranks <= 1) return 0;
    ProfScope ps(c, PK_HALO);
    return halo_exchange(c, sets, c->stream);
}
// Overlap of the halo exchange with interior pair work: the exchange is issued on the communication stream once everything
// queued on the compute stream so far (the producers of the slices) has finished; the caller evaluates the interior range
// meanwhile and calls comm_halo_end before touching the boundary layers. Only the one-cell-thick boundary layers -- the send
// slices themselves -- can hold a particle with a halo neighbour, and they are contiguous in the cell-sorted order.
int comm_halo_begin(wf_ctx *c, int sets, int64_t *i0, int64_t *i1)
{
    wf_ctx::Comm &m = c->comm;
    *i0 = 0; *i1 = c->n_loc;
    if (m.nranks <= 1) return 0;
    if (c->prm.tuning & 512) {                        // bit 9 of wf_params::tuning: blocking refresh (A/B measurement)
        WF_CHECK(comm_halo_refresh(c, sets));
        return 0;
    }
    if (m.rank > 0) *i0 = m.send_lo_start + m.send_lo_n;
    if (m.rank < m.nranks - 1) *i1 = m.send_hi_start;
    if (*i1 < *i0) *i1 = *i0;
    WF_CUDA(c, cudaEventRecord(m.ev_ready, c->stream));
    WF_CUDA(c, cudaStreamWaitEvent(m.stream, m.ev_ready, 0));
    WF_CHECK(halo_exchange(c, sets, m.stream));
    WF_CUDA(c, cudaEventRecord(m.ev_done, m.stream));
    m.pending = true;
    return 0;
}
int comm_halo_end(wf_ctx *c)
{
    wf_ctx::Comm &m = c->comm;
    if (!m.pending) return 0;
    ProfScope ps(c, PK_HALO);
    WF_CUDA(c, cudaStreamWaitEvent(c->stream, m.ev_done, 0));
    m.pending = false;
    return 0;
}

// ---- time-based slab rebalancing -----------------------------------------------------------------------------------------
// Equal particle counts do not give equal step times: ranks differ in neighbours per particle (end caps, free surfaces) and
// the GPUs of a box do not clock identically under load, and every halo exchange makes the fast ranks wait for the slowest.
// Each rank times its interior momentum launch (the largest kernel) over the steps between two rebuilds; at the rebuild the
// ranks exchange (time, particles, particles in their first / last layer) and every boundary moves by at most ONE layer if
// that lowers the larger of the two adjacent estimated step times by more than 1 %. The decision is a deterministic function
// of the gathered numbers, so all ranks take it alike; the particles of the moved layer travel through the ordinary migration
// (they are simply outside their old owner's [layer_lo, layer_hi) at the next cell-key pass). Results do not depend on the
// decomposition (per-particle sums follow the reference's order), so this never changes a bit of the state.
// stats: 8 doubles per rank = time per particle, particles, particles in the first layer, in the last layer, emigrant-list capacity, -, -, -
static void rebalance_decide(int nranks, const double *all, bool forced, std::vector<int> &lo, std::vector<int> &hi)
{
    std::vector<double> cost(nranks), load(nranks);
    bool any_zero = false;
    for (int r = 0; r < nranks; ++r) { cost[r] = all[8 * r]; load[r] = cost[r] * all[8 * r + 1]; if (!(cost[r] > 0.0)) any_zero = true; }
    for (int b = 0; b + 1 < nranks; ++b) {               // boundary between rank b and b + 1, left to right with updated loads
        const int r = b, q = b + 1;
        const double n_up = all[8 * r + 3], n_dn = all[8 * q + 2];            // rank r's last layer / rank q's first layer
        int move = 0;                                                          // +1: r gives its last layer to q; -1: q gives its first layer to r
        if (forced) move = (b == 0) ? +1 : 0;
        else if (!any_zero) {
            const double cur = std::max(load[r], load[q]);
            const double a_ = std::max(load[r] - cost[r] * n_up, load[q] + cost[q] * n_up);
            const double b_ = std::max(load[r] + cost[r] * n_dn, load[q] - cost[q] * n_dn);
            if (a_ < cur * 0.99 && a_ <= b_) move = +1; else if (b_ < cur * 0.99) move = -1;
        }
        if (move > 0 && (hi[r] - lo[r] < 4 || n_up > all[8 * r + 4] * 0.9)) move = 0;      // the giving rank's emigrant list must hold the layer
        if (move < 0 && (hi[q] - lo[q] < 4 || n_dn > all[8 * q + 4] * 0.9)) move = 0;
        if (move > 0) { hi[r] -= 1; lo[q] -= 1; load[r] -= cost[r] * n_up; load[q] += cost[q] * n_up; }
        if (move < 0) { hi[r] += 1; lo[q] += 1; load[r] += cost[r] * n_dn; load[q] -= cost[q] * n_dn; }
        if (move != 0) ++b;                              // a rank takes part in at most one move per rebuild: skip its other boundary
    }
}

int comm_time_begin(wf_ctx *c)
{
    wf_ctx::Comm &m = c->comm;
    if (m.nranks <= 1 || (c->prm.tuning & 16384)) return 0;       // bit 14 of wf_params::tuning: rebalancing off (A/B measurement)
    if (m.t_used >= 64) return 0;
    if (m.t_used >= (int)m.t_events.size()) { cudaEvent_t a, b; WF_CUDA(c, cudaEventCreate(&a)); WF_CUDA(c, cudaEventCreate(&b)); m.t_events.push_back(std::make_pair(a, b)); }
    WF_CUDA(c, cudaEventRecord(m.t_events[m.t_used].first, c->stream));
    return 0;
}
int comm_time_end(wf_ctx *c)
{
    wf_ctx::Comm &m = c->comm;
    if (m.nranks <= 1 || (c->prm.tuning & 16384)) return 0;
    if (m.t_used >= (int)m.t_events.size()) return 0;
    WF_CUDA(c, cudaEventRecord(m.t_events[m.t_used].second, c->stream));
    m.t_used++;
    return 0;
}
int comm_rebalance(wf_ctx *c)
{
    wf_ctx::Comm &m = c->comm;
    if (m.nranks <= 1 || (c->prm.tuning & 16384)) return 0;
    m.rebuilds_seen++;
    const bool forced = (c->prm.tuning & 8192) && m.rebuilds_seen == 3;     // bit 13: force one move of boundary 0 at the third rebuild (tests)
    if (m.t_used == 0 && !forced) return 0;
    double t = 0.0;
    if (m.t_used > 0) {
        WF_CUDA(c, cudaEventSynchronize(m.t_events[m.t_used - 1].second));
        for (int k = 0; k < m.t_used; ++k) { float ms = 0.f; cudaEventElapsedTime(&ms, m.t_events[k].first, m.t_events[k].second); t += ms; }
        t /= m.t_used;
    }
    m.t_used = 0;
    // the interior launch covers n_loc minus the boundary layers: scale to the whole rank
    const double n_int = (double)(c->n_loc - c->nmeshp) - (double)(m.rank > 0 ? m.send_lo_n : 0) - (double)(m.rank < m.nranks - 1 ? m.send_hi_n : 0);
    double mine[8] = { n_int > 0 ? t / n_int : 0.0, (double)(c->n_loc - c->nmeshp), (double)m.send_lo_n, (double)m.send_hi_n, (double)m.emig_cap, 0., 0., 0. };
    WF_CUDA(c, cudaMemcpyAsync(m.d_rank, mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream));
    WF_CHECK(comm_gather_scalars(c, 8));
    std::vector<double> all(8 * (size_t)m.nranks);
    WF_CUDA(c, cudaMemcpyAsync(all.data(), m.d_all, all.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    WF_CUDA(c, cudaStreamSynchronize(c->stream));
    // gather all layer bounds too (they are needed to keep >= 3 layers per rank): derive from a second gather
    double lb[2] = { (double)m.layer_lo, (double)m.layer_hi };
    WF_CUDA(c, cudaMemcpyAsync(m.d_rank, lb, sizeof(lb), cudaMemcpyHostToDevice, c->stream));
    WF_CHECK(comm_gather_scalars(c, 2));
    std::vector<double> bounds(2 * (size_t)m.nranks);
    WF_CUDA(c, cudaMemcpyAsync(bounds.data(), m.d_all, bounds.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    WF_CUDA(c, cudaStreamSynchronize(c->stream));
    std::vector<int> lo(m.nranks), hi(m.nranks);
    for (int r = 0; r < m.nranks; ++r) { lo[r] = (int)bounds[2 * r]; hi[r] = (int)bounds[2 * r + 1]; }
    const int lo_before = lo[m.rank], hi_before = hi[m.rank];
    rebalance_decide(m.nranks, all.data(), forced, lo, hi);
    if (lo[m.rank] != lo_before || hi[m.rank] != hi_before) m.rebalance_moves++;
    m.layer_lo = lo[m.rank]; m.layer_hi = hi[m.rank];
    return 0;
}

// wf_destroy: release the NCCL communicator (collective-free: ncclCommDestroy only tears down this rank's resources)
void comm_destroy(wf_ctx *c)
{
    if (!c->comm.nccl) return;
    NcclApi *A = nccl_api();
    if (A && A->CommDestroy) A->CommDestroy((ncclComm_t)c->comm.nccl);
    c->comm.nccl = nullptr;
}

int comm_gather_scalars_on(wf_ctx *c, const double *src, double *dst, int nvals, cudaStream_t st)
{
    wf_ctx::Comm &m = c->comm;
    if (m.nranks <= 1) return 0;
    NcclApi *A = nccl_api();
    WF_NCCL(c, A->AllGather(src, dst, (size_t)nvals, ncclDouble, (ncclComm_t)m.nccl, st));
    return 0;
}
int comm_red_wait(wf_ctx *c)
{
    wf_ctx::Comm &m = c->comm;
    for (int p = 0; p < 2; ++p) if (m.red_pending[p]) { WF_CUDA(c, cudaStreamWaitEvent(c->stream, m.ev_red_done[p], 0)); m.red_pending[p] = false; }
    return 0;
}

int comm_gather_scalars(wf_ctx *c, int nvals)
{
    wf_ctx::Comm &m = c->comm;
    if (m.nranks <= 1) return 0;
    NcclApi *A = nccl_api();
    WF_NCCL(c, A->AllGather(m.d_rank, m.d_all, (size_t)nvals, ncclDouble, (ncclComm_t)m.nccl, c->stream));
    return 0;
}

static int exchange_counts(wf_ctx *c, const int send[2], int recv[2])
{
    // send[0] -> rank-1, send[1] -> rank+1; recv[0] <- rank-1, recv[1] <- rank+1
    wf_ctx::Comm &m = c->comm; NcclApi *A = nccl_api(); ncclComm_t comm = (ncclComm_t)m.nccl;
    int h[4] = { send[0], send[1], 0, 0 };
    WF_CUDA(c, cudaMemcpyAsync(m.d_cnt + 4, h, 4 * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    WF_NCCL(c, A->GroupStart());
    if (m.rank > 0) { WF_NCCL(c, A->Send(m.d_cnt + 4, 1, ncclInt32, m.rank - 1, comm, c->stream)); WF_NCCL(c, A->Recv(m.d_cnt + 6, 1, ncclInt32, m.rank - 1, comm, c->stream)); }
    if (m.rank < m.nranks - 1) { WF_NCCL(c, A->Send(m.d_cnt + 5, 1, ncclInt32, m.rank + 1, comm, c->stream)); WF_NCCL(c, A->Recv(m.d_cnt + 7, 1, ncclInt32, m.rank + 1, comm, c->stream)); }
    WF_NCCL(c, A->GroupEnd());
    WF_CUDA(c, cudaMemcpyAsync(h, m.d_cnt + 4, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    WF_CUDA(c, cudaStreamSynchronize(c->stream));
    recv[0] = m.rank > 0 ? h[2] : 0; recv[1] = m.rank < m.nranks - 1 ? h[3] : 0;
    return 0;
}

// Phase 1 of the multi-GPU neighbour build (before the sort): emigrants were listed by k_cell_keys_mig in
// d_emig[dir*emig_cap ..] with counts in d_cnt[0..1]; ship them, append the immigrants at [n_loc, n_loc + n_in).
// Returns the number of slots to sort in *n_sort and the number of emigrants in *n_out.
int comm_migrate(wf_ctx *c, int64_t *n_sort, int64_t *n_out)
{
    wf_ctx::Comm &m = c->comm; NcclApi *A = nccl_api(); ncclComm_t comm = (ncclComm_t)m.nccl;
    int out[2] = { 0, 0 }, in[2] = { 0, 0 };
    WF_CUDA(c, cudaMemcpyAsync(out, m.d_cnt, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    WF_CUDA(c, cudaStreamSynchronize(c->stream));
    int st = 0;
    if (out[0] > m.emig_cap || out[1] > m.emig_cap) st = fail(c, WF_ERR_NOMEM, "migration: %d/%d particles leave the slab at once (capacity %lld)", out[0], out[1], (long long)m.emig_cap);
    else if (m.rank == 0 && out[0]) st = fail(c, WF_ERR_ARG, "migration: %d particles left the cell grid below rank 0", out[0]);
    else if (m.rank == m.nranks - 1 && out[1]) st = fail(c, WF_ERR_ARG, "migration: %d particles left the cell grid above the last rank", out[1]);
    WF_CHECK(comm_agree(c, st, "particle migration"));
    WF_CHECK(exchange_counts(c, out, in));
    const int NF = 24 + F_COUNT + I_COUNT;
    const int64_t total_in = (int64_t)in[0] + in[1];
    st = 0;
    if (c->n_loc + total_in > c->cap) st = fail(c, WF_ERR_NOMEM, "migration: %lld local + %lld incoming particles exceed the %lld allocated slots", (long long)c->n_loc, (long long)total_in, (long long)c->cap);
    WF_CHECK(comm_agree(c, st, "particle migration"));
    const size_t need = (size_t)NF * std::max(std::max(out[0], out[1]), std::max(in[0], in[1])) * sizeof(double);
    if (need > m.stage_cap) {
        for (int d = 0; d < 2; ++d) { cudaFree(m.stage_send[d]); cudaFree(m.stage_recv[d]); m.stage_send[d] = m.stage_recv[d] = nullptr; }
        const size_t cap2 = need * 2 + 4096;
        for (int d = 0; d < 2; ++d) { WF_CUDA(c, cudaMalloc(&m.stage_send[d], cap2)); WF_CUDA(c, cudaMalloc(&m.stage_recv[d], cap2)); }
        m.stage_cap = cap2;
    }
    const int b = c->cur;
    for (int d = 0; d < 2; ++d) if (out[d]) {
        dim3 g((out[d] + 127) / 128, NF);
        k_pack<<<g, 128, 0, c->stream>>>(qptrs(c, b), c->F[b], c->I[b], c->cap, m.d_emig + (int64_t)d * m.emig_cap, out[d], m.stage_send[d]);
        WF_LAUNCH_CHECK(c); c->launches++;
    }
    if (out[0] || out[1] || in[0] || in[1]) {
        WF_NCCL(c, A->GroupStart());
        if (out[0]) WF_NCCL(c, A->Send(m.stage_send[0], (size_t)NF * out[0], ncclDouble, m.rank - 1, comm, c->stream));
        if (in[0]) WF_NCCL(c, A->Recv(m.stage_recv[0], (size_t)NF * in[0], ncclDouble, m.rank - 1, comm, c->stream));
        if (out[1]) WF_NCCL(c, A->Send(m.stage_send[1], (size_t)NF * out[1], ncclDouble, m.rank + 1, comm, c->stream));
        if (in[1]) WF_NCCL(c, A->Recv(m.stage_recv[1], (size_t)NF * in[1], ncclDouble, m.rank + 1, comm, c->stream));
        WF_NCCL(c, A->GroupEnd());
    }
    int64_t base = c->n_loc;
    for (int d = 0; d < 2; ++d) if (in[d]) {
        dim3 g((in[d] + 127) / 128, NF);
        k_unpack<<<g, 128, 0, c->stream>>>(qptrs(c, b), c->F[b], c->I[b], c->cap, base, in[d], m.stage_recv[d]);
        WF_LAUNCH_CHECK(c); c->launches++;
        base += in[d];
    }
    m.migrated_out += out[0] + out[1]; m.migrated_in += total_in;
    *n_sort = c->n_loc + total_in; *n_out = (int64_t)out[0] + out[1];
    return 0;
}

// Phase 2 (after sort + permutation of the local region, n_loc final): boundary slices, halo sizes, first halo fill
// of all gathered arrays + creation index / ID / class of the halo particles.
int comm_halo_setup(wf_ctx *c, const unsigned long long *sorted_keys)
{
    wf_ctx::Comm &m = c->comm; NcclApi *A = nccl_api(); ncclComm_t comm = (ncclComm_t)m.nccl;
    const long long plane = (long long)c->CellNo[c->ax_mid] * c->CellNo[c->ax_fast];
    long long tg[4] = { (long long)m.layer_lo * plane, (long long)(m.layer_lo + 1) * plane, (long long)(m.layer_hi - 1) * plane, (long long)m.layer_hi * plane };
    long long *d_t = (long long *)(m.d_cnt + 8), res[4];
    WF_CUDA(c, cudaMemcpyAsync(d_t, tg, sizeof(tg), cudaMemcpyHostToDevice, c->stream));
    k_lower_bound<<<1, 32, 0, c->stream>>>(sorted_keys, c->n_loc, d_t, 4, d_t + 4); WF_LAUNCH_CHECK(c);
    WF_CUDA(c, cudaMemcpyAsync(res, d_t + 4, sizeof(res), cudaMemcpyDeviceToHost, c->stream));
    WF_CUDA(c, cudaStreamSynchronize(c->stream));
    m.send_lo_start = res[0]; m.send_lo_n = (m.rank > 0) ? res[1] - res[0] : 0;
    m.send_hi_start = res[2]; m.send_hi_n = (m.rank < m.nranks - 1) ? res[3] - res[2] : 0;
    int snd[2] = { (int)m.send_lo_n, (int)m.send_hi_n }, rcv[2];
    WF_CHECK(exchange_counts(c, snd, rcv));
    m.halo_lo_start = c->n_loc; m.halo_lo_n = rcv[0];
    m.halo_hi_start = c->n_loc + rcv[0]; m.halo_hi_n = rcv[1];
    c->n = c->n_loc + rcv[0] + rcv[1];
    { int st = 0;
      if (c->n > c->cap) st = fail(c, WF_ERR_NOMEM, "halo: %lld local + %d + %d halo particles exceed the %lld allocated slots", (long long)c->n_loc, rcv[0], rcv[1], (long long)c->cap);
      WF_CHECK(comm_agree(c, st, "halo set-up")); }
    m.halo_bytes_step = 8LL * 32 * (m.send_lo_n + m.send_hi_n);
    WF_CHECK(comm_halo_refresh(c, HALO_Q0 | HALO_Q1 | HALO_Q2 | HALO_S));
    const int ifields[4] = { I_ORIG, I_ID, I_CLS, I_FREE };
    WF_NCCL(c, A->GroupStart());
    for (int k = 0; k < 4; ++k) {
        int *arr = c->Ip(ifields[k]);
        if (m.rank > 0) { if (m.send_lo_n) WF_NCCL(c, A->Send(arr + m.send_lo_start, m.send_lo_n, ncclInt32, m.rank - 1, comm, c->stream));
                          if (m.halo_lo_n) WF_NCCL(c, A->Recv(arr + m.halo_lo_start, m.halo_lo_n, ncclInt32, m.rank - 1, comm, c->stream)); }
        if (m.rank < m.nranks - 1) { if (m.send_hi_n) WF_NCCL(c, A->Send(arr + m.send_hi_start, m.send_hi_n, ncclInt32, m.rank + 1, comm, c->stream));
                                     if (m.halo_hi_n) WF_NCCL(c, A->Recv(arr + m.halo_hi_start, m.halo_hi_n, ncclInt32, m.rank + 1, comm, c->stream)); }
    }
    WF_NCCL(c, A->GroupEnd());
    return 0;
}

// local index (position in ascending creation-index order of this rank's particles) <-> slot
int comm_local_index(wf_ctx *c)
{
    const int64_t n = c->n_loc; cudaStream_t st = c->stream;
    int *k_in = c->vals[0], *v_in = c->vals[0] + 0;   // reuse the sort buffers: keys need int buffers of their own
    (void)k_in; (void)v_in;
    int *buf = nullptr; const size_t bytes = 4 * (size_t)n * sizeof(int);
    WF_CHECK(ensure_stage(c, bytes + 256));
    buf = (int *)c->stage;
    int *ko = buf, *vo = buf + n, *ks = buf + 2 * n, *vs = buf + 3 * n;
    k_iota_orig<<<grid_for(n, 256), 256, 0, st>>>(n, c->Ip(I_ORIG), ko, vo); WF_LAUNCH_CHECK(c);
    size_t need = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, need, ko, ks, vo, vs, (int)n, 0, 32, st);
    if (need > c->cub_tmp_bytes) { cudaFree(c->cub_tmp); c->cub_tmp = nullptr; WF_CUDA(c, cudaMalloc(&c->cub_tmp, need)); c->cub_tmp_bytes = need; }
    WF_CUDA(c, cub::DeviceRadixSort::SortPairs(c->cub_tmp, need, ko, ks, vo, vs, (int)n, 0, 32, st));
    k_set_lid<<<grid_for(n, 256), 256, 0, st>>>(n, vs, c->slot_of, c->Ip(I_LID)); WF_LAUNCH_CHECK(c);
    c->launches += 3;
    return 0;
}

} // namespace wf

// ================================================================================================================
extern "C" int wf_comm_unique_id(void *out128)
{
    if (!out128) return WF_ERR_ARG;
    NcclApi *A = nccl_api();
    if (!A->ok) return wf::fail(nullptr, WF_ERR_UNSUPPORTED, "NCCL unavailable: %s", A->err.c_str());
    ncclUniqueId id; ncclResult_t r = A->GetUniqueId(&id);
    if (r != ncclSuccess) return wf::fail(nullptr, WF_ERR_CUDA, "ncclGetUniqueId: %s", A->GetErrorString(r));
    std::memcpy(out128, id.internal, NCCL_UNIQUE_ID_BYTES);
    return 0;
}

extern "C" int wf_comm_init(wf_ctx *c, int rank, int nranks, const void *id128)
{
    if (!c || nranks < 1 || rank < 0 || rank >= nranks) return WF_ERR_ARG;
    if (c->have_particles) return wf::fail(c, WF_ERR_ARG, "wf_comm_init: call before wf_set_particles");
    c->comm.rank = rank; c->comm.nranks = nranks;
    if (nranks == 1) return 0;
    if (!id128) return WF_ERR_ARG;
    NcclApi *A = nccl_api();
    if (!A->ok) return wf::fail(c, WF_ERR_UNSUPPORTED, "NCCL unavailable: %s", A->err.c_str());
    cudaSetDevice(c->device);
    ncclUniqueId id; std::memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
    ncclComm_t comm;
    WF_NCCL(c, A->CommInitRank(&comm, nranks, id, rank));
    c->comm.nccl = comm;
    {   // highest priority: the exchange kernels must get SM slots while the interior pair kernel still has blocks queued
        int prio_lo = 0, prio_hi = 0; cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        WF_CUDA(c, cudaStreamCreateWithPriority(&c->comm.stream, cudaStreamNonBlocking, prio_hi)); }
    WF_CUDA(c, cudaEventCreateWithFlags(&c->comm.ev_ready, cudaEventDisableTiming)); WF_CUDA(c, cudaEventCreateWithFlags(&c->comm.ev_done, cudaEventDisableTiming));
    WF_CUDA(c, cudaMalloc(&c->comm.d_cnt, 64 * sizeof(int) + 16 * sizeof(long long)));
    WF_CUDA(c, cudaMemset(c->comm.d_cnt, 0, 64 * sizeof(int) + 16 * sizeof(long long)));
    WF_CUDA(c, cudaMalloc(&c->comm.d_all, (size_t)nranks * 32 * sizeof(double)));
    WF_CUDA(c, cudaEventCreateWithFlags(&c->comm.ev_red_ready, cudaEventDisableTiming));
    for (int p = 0; p < 2; ++p) {
        WF_CUDA(c, cudaMalloc(&c->comm.d_red_rank[p], 32 * sizeof(double))); WF_CUDA(c, cudaMalloc(&c->comm.d_red_all[p], (size_t)nranks * 32 * sizeof(double)));
        WF_CUDA(c, cudaEventCreateWithFlags(&c->comm.ev_red_done[p], cudaEventDisableTiming));
    }
    return 0;
}

extern "C" int wf_comm_info(wf_ctx *c, int64_t out[8])
{
    if (!c || !out) return WF_ERR_ARG;
    out[0] = c->n_loc; out[1] = c->n - c->n_loc; out[2] = c->comm.migrated_out; out[3] = c->comm.migrated_in;
    out[4] = c->comm.layer_lo; out[5] = c->comm.layer_hi; out[6] = c->comm.halo_bytes_step; out[7] = c->CellNo[c->ax_slow];
    return 0;
}

// Host-only: the boundary moves comm_rebalance would decide from the gathered per-rank statistics (8 doubles per rank: time per
// particle, particles, particles in the first / last layer, emigrant-list capacity, 3 unused); bounds = nranks + 1 layer indices,
// updated in place. Same code as the library's collective; exposed for CPU tests and for drivers that want to log the decision.
extern "C" int wf_plan_rebalance(const double *stats, int nranks, int32_t *bounds)
{
    if (!stats || !bounds || nranks < 1) return WF_ERR_ARG;
    std::vector<int> lo(nranks), hi(nranks);
    for (int r = 0; r < nranks; ++r) { lo[r] = bounds[r]; hi[r] = bounds[r + 1]; }
    wf::rebalance_decide(nranks, stats, false, lo, hi);
    for (int r = 0; r < nranks; ++r) bounds[r] = lo[r];
    bounds[nranks] = hi[nranks - 1];
    return 0;
}

// Host-side planner (no GPU, no NCCL): the slab a particle set would be split into by `nranks` ranks -- the same grid
// (Neighbour.cpp:10-139) and the same balancing code the library uses. rank_of[i] in [0, nranks), bounds[nranks+1] layers.
extern "C" int wf_plan_slabs(const double *x, const double *h, int64_t n, int kernel_type, const double dom_min[3], const double dom_max[3],
                             int nranks, int32_t *rank_of, int32_t *bounds)
{ return wf_plan_slabs_ex(x, h, n, n, kernel_type, dom_min, dom_max, nranks, rank_of, bounds); }

// With rigid meshes: x, h hold the n_sph SPH particles followed by the element particles (centroid, hfac*radius). All of
// them shape the cell grid (bounding box, hmax -- Neighbour.cpp:10-139 runs over every particle), only the SPH particles are
// distributed (rank_of = -1 for the element particles: every rank keeps a replica of the meshes).
extern "C" int wf_plan_slabs_ex(const double *x, const double *h, int64_t n, int64_t n_sph, int kernel_type, const double dom_min[3], const double dom_max[3],
                                int nranks, int32_t *rank_of, int32_t *bounds)
{
    if (!x || !h || n <= 0 || n_sph <= 0 || n_sph > n || nranks < 1) return WF_ERR_ARG;
    const double Cellfac = (kernel_type == 2) ? 3.0 : 2.0;
    double BL[3] = { x[0], x[1], x[2] }, TR[3] = { x[0], x[1], x[2] }, hmax = h[0];
    for (int64_t i = 0; i < n; ++i) { for (int k = 0; k < 3; ++k) { if (x[3 * i + k] > TR[k]) TR[k] = x[3 * i + k]; if (x[3 * i + k] < BL[k]) BL[k] = x[3 * i + k]; } if (h[i] > hmax) hmax = h[i]; }
    const double pBL[3] = { BL[0], BL[1], BL[2] }, pTR[3] = { TR[0], TR[1], TR[2] };
    if (dom_max) for (int k = 0; k < 3; ++k) if (dom_max[k] > TR[k]) TR[k] = dom_max[k];
    if (dom_min) for (int k = 0; k < 3; ++k) if (dom_min[k] < BL[k]) BL[k] = dom_min[k];
    int CellNo[3]; double CellSize[3];
    for (int k = 0; k < 3; ++k) { TR[k] += hmax / 2; BL[k] -= hmax / 2; }
    for (int k = 0; k < 3; ++k) {
        double qq = (TR[k] - BL[k]) / (Cellfac * hmax);
        if ((double)(std::ceil(qq) - qq) < (hmax / 10.0)) CellNo[k] = (int)std::ceil(qq); else CellNo[k] = (int)std::floor(qq);
        if (CellNo[k] <= 0) return WF_ERR_ARG;
        CellSize[k] = (TR[k] - BL[k]) / CellNo[k];
    }
    int ax[3] = { 0, 1, 2 }; double occ[3];
    for (int k = 0; k < 3; ++k) occ[k] = (pTR[k] - pBL[k]) / CellSize[k];
    std::sort(ax, ax + 3, [&](int a, int b) { return occ[a] < occ[b] || (occ[a] == occ[b] && a < b); });
    const int slow = ax[2], L = CellNo[slow];
    if (L < nranks) return WF_ERR_ARG;
    std::vector<long long> hist(L, 0); std::vector<int> layer(n);
    for (int64_t i = 0; i < n; ++i) {
        int ci = (int)std::floor((x[3 * i + slow] - BL[slow]) / CellSize[slow]);
        if (ci < 0) ci = 0; if (ci >= L) ci = L - 1;
        layer[i] = ci; if (i < n_sph) hist[ci]++;
    }
    std::vector<int> b; wf::balance_layers(hist, nranks, b);
    if (bounds) for (int r = 0; r <= nranks; ++r) bounds[r] = b[r];
    if (rank_of) for (int64_t i = 0; i < n; ++i) { int r = (int)(std::upper_bound(b.begin(), b.end(), layer[i]) - b.begin()) - 1; rank_of[i] = i < n_sph ? std::min(std::max(r, 0), nranks - 1) : -1; }
    return 0;
}
