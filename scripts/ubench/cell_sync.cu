// Micro-benchmark for the NEXT design of the pair walk: per-lane list walk (what the library does) against a CELL-SYNCHRONOUS walk.
//
// The contracted pair kernels are bound by L1 wavefronts: at a given list position the 32 lanes of a warp read 32 records in ~25 different
// 128-byte lines (profiles/README_r2.md), because the particles of one cell have 0..5 hits in each neighbour cell and so drift apart in their
// lists. In the cell-synchronous form the warp visits the 27 neighbour cells together: at cell e every lane reads its own hits IN THAT CELL
// (rows [e][t] of a per-cell padded list, one byte per entry: the index inside the cell), so a request touches the ~14 records of 2-3 adjacent
// cells -- ~8 lines -- at the price of idle lanes (a lane with 2 hits waits for the lane with 5).
//
// Same cell-sorted lattice (cell = 2.4 dx like the library's Cellfac*h with h = 1.2 dx), same 56 neighbours, same five gathered 32-byte records
// and the same ~80 dependent-free FP64 operations per pair in both kernels; the difference is the order of the visits alone.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 cell_sync.cu -o cell_sync ; run on a B200 (setup ~20 s on the host).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ double4 ld256(const double4 *p)
{
    double4 r; asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p)); return r;
}
// ~80 FP64 operations on the 20 gathered doubles (independent chains, like a pair term evaluated with FMAs)
__device__ __forceinline__ void pair_math(const double4 &a, const double4 &b, const double4 &c, const double4 &d, const double4 &e, double xi, double &s0, double &s1, double &s2)
{
    const double dx = xi - a.x, dy = xi - a.y, dz = xi - a.z;
    const double r2 = fma(dx, dx, fma(dy, dy, dz * dz)) + 1.0;
    const double ux = b.x - xi, uy = b.y - xi, uz = b.z - xi;
    const double vx = fma(ux, dx, fma(uy, dy, uz * dz));
    double w = fma(r2, 0.25, -3.0); w = fma(w, r2, a.w); w = fma(w, w, b.w); w = fma(w, vx, 1.0);
    double t = fma(vx, 0.5, r2); t = fma(t, t, w); t = fma(t, 0.125, vx);
    const double mxx = fma(c.x, t, d.x) + w, myy = fma(c.y, t, d.y) + w, mzz = fma(c.z, t, d.z) + w;
    const double mxy = fma(c.w, t, d.w), myz = fma(e.x, t, e.y), mxz = fma(e.z, t, e.w);
    const double gx = w * dx, gy = w * dy, gz = w * dz;
    s0 = fma(gx, mxx, fma(gy, mxy, fma(gz, mxz, s0)));
    s1 = fma(gx, mxy, fma(gy, myy, fma(gz, myz, s1)));
    s2 = fma(gx, mxz, fma(gy, myz, fma(gz, mzz, s2)));
    double p = fma(t, t, 1.0); p = fma(p, w, t); p = fma(p, p, r2); p = fma(p, 0.5, w); p = fma(p, vx, t);
    s0 = fma(p, 1e-9, s0);
}

// (a) per-lane list walk: ELL transposed, list in cell-walk order
__global__ void __launch_bounds__(128, 3) k_list(int n, const uint32_t *nbr, const int *cnt, const double4 *A, const double4 *B, const double4 *C, const double4 *D, const double4 *E, double *out)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x; if (s >= n) return;
    const double xi = A[s].x; double s0 = 0., s1 = 0., s2 = 0.;
    const int c = cnt[s];
#pragma unroll 2
    for (int k = 0; k < c; ++k) {
        const int j = (int)nbr[(size_t)k * n + s];
        const double4 a = ld256(A + j), b = ld256(B + j), cc = ld256(C + j), d = ld256(D + j), e = ld256(E + j);
        pair_math(a, b, cc, d, e, xi, s0, s1, s2);
    }
    out[s] = s0 + s1 + s2;
}
// (b) cell-synchronous walk: for each of the 27 neighbour cells, rows [e*P + t] hold the index (inside that cell) of this particle's t-th hit, 255 = none
__global__ void __launch_bounds__(128, 3) k_cell(int n, int P, const uint8_t *hit, const int *cell_of, const int *cell_start, const int *cell_off, int ncell,
                                                 const double4 *A, const double4 *B, const double4 *C, const double4 *D, const double4 *E, double *out)
{
    const int s0i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = s0i < n; const int s = live ? s0i : n - 1;
    const double xi = A[s].x; double s0 = 0., s1 = 0., s2 = 0.;
    const int mycell = cell_of[s];
    for (int e = 0; e < 27; ++e) {
        const int nc = mycell + cell_off[e];
        const int cs = (nc >= 0 && nc < ncell) ? cell_start[nc] : -1;
        for (int t = 0; t < P; ++t) {
            const unsigned idx = (cs >= 0 && live) ? hit[(size_t)(e * P + t) * n + s] : 255u;
            if (!__any_sync(0xffffffffu, idx != 255u)) break;
            if (idx != 255u) {
                const int j = cs + (int)idx;
                const double4 a = ld256(A + j), b = ld256(B + j), cc = ld256(C + j), d = ld256(D + j), ee = ld256(E + j);
                pair_math(a, b, cc, d, ee, xi, s0, s1, s2);
            }
        }
    }
    if (live) out[s] = s0 + s1 + s2;
}

int main()
{
    const int L = 120; const double cellsz = 2.4, cut2 = 2.4 * 2.4;
    const int NC = (int)(L / cellsz) + 1;                 // cells per axis
    const int n = L * L * L;
    // particles on the lattice, sorted by cell (x fastest), raster order inside a cell
    std::vector<int> cell(n), order(n);
    auto cid = [&](int x, int y, int z) { return ((int)(z / cellsz) * NC + (int)(y / cellsz)) * NC + (int)(x / cellsz); };
    for (int z = 0; z < L; ++z) for (int y = 0; y < L; ++y) for (int x = 0; x < L; ++x) { const int p = (z * L + y) * L + x; cell[p] = cid(x, y, z); order[p] = p; }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cell[a] < cell[b]; });
    std::vector<int> slot_of(n), px(n), py(n), pz(n), cell_of(n);
    const int ncell = NC * NC * NC; std::vector<int> cstart(ncell + 1, 0);
    for (int s = 0; s < n; ++s) { const int p = order[s]; slot_of[p] = s; px[s] = p % L; py[s] = (p / L) % L; pz[s] = p / (L * L); cell_of[s] = cell[p]; cstart[cell[p] + 1]++; }
    for (int c = 0; c < ncell; ++c) cstart[c + 1] += cstart[c];
    int off[27], k27 = 0;
    for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) off[k27++] = (dz * NC + dy) * NC + dx;
    // neighbour lists both ways
    const int K = 64; int P = 0;
    std::vector<uint32_t> nbr((size_t)K * n, 0); std::vector<int> cnt(n, 0);
    std::vector<std::vector<uint8_t> > hits(27, std::vector<uint8_t>());
    std::vector<int> percell((size_t)27 * n, 0);
    for (int pass = 0; pass < 2; ++pass) {
        std::vector<uint8_t> hit; if (pass == 1) hit.assign((size_t)27 * P * n, 255);
        for (int s = 0; s < n; ++s) {
            const int cx = (int)(px[s] / cellsz), cy = (int)(py[s] / cellsz), cz = (int)(pz[s] / cellsz); int c = 0;
            for (int e = 0; e < 27; ++e) {
                const int dx = e % 3 - 1, dy = (e / 3) % 3 - 1, dz = e / 9 - 1;
                const int X = cx + dx, Y = cy + dy, Z = cz + dz; int h = 0;
                if (X >= 0 && Y >= 0 && Z >= 0 && X < NC && Y < NC && Z < NC) {
                    const int nc = (Z * NC + Y) * NC + X;
                    for (int j = cstart[nc]; j < cstart[nc + 1]; ++j) {
                        if (j == s) continue;
                        const double ddx = px[s] - px[j], ddy = py[s] - py[j], ddz = pz[s] - pz[j];
                        if (ddx * ddx + ddy * ddy + ddz * ddz <= cut2) {
                            if (pass == 0) { if (c < K) nbr[(size_t)c * n + s] = (uint32_t)j; ++c; }
                            else hit[(size_t)(e * P + h) * n + s] = (uint8_t)(j - cstart[nc]);
                            ++h;
                        }
                    }
                }
                if (pass == 0) { if (h > P) P = h; }
            }
            if (pass == 0) cnt[s] = c < K ? c : K;
        }
        if (pass == 1) hits[0].swap(hit);
    }
    long long pairs = 0; for (int s = 0; s < n; ++s) pairs += cnt[s];
    std::printf("n = %d, cells %d^3 (%.1f particles per cell), pair visits %lld (%.1f per particle), max hits per (particle, cell) P = %d\n", n, NC, (double)n / ncell, pairs, (double)pairs / n, P);
    // device
    uint32_t *d_nbr; int *d_cnt, *d_cell_of, *d_cstart, *d_off; uint8_t *d_hit; double4 *arr[5]; double *out;
    cudaMalloc(&d_nbr, nbr.size() * 4); cudaMemcpy(d_nbr, nbr.data(), nbr.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&d_cnt, n * 4); cudaMemcpy(d_cnt, cnt.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&d_cell_of, n * 4); cudaMemcpy(d_cell_of, cell_of.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&d_cstart, (ncell + 1) * 4); cudaMemcpy(d_cstart, cstart.data(), (ncell + 1) * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&d_off, 27 * 4); cudaMemcpy(d_off, off, 27 * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&d_hit, hits[0].size()); cudaMemcpy(d_hit, hits[0].data(), hits[0].size(), cudaMemcpyHostToDevice);
    std::vector<double4> h((size_t)n);
    for (int k = 0; k < 5; ++k) {
        for (int s = 0; s < n; ++s) h[s] = make_double4(px[s] * 0.01 + k, py[s] * 0.01, pz[s] * 0.01, 1.0 + 1e-3 * ((s * 7 + k) % 13));
        cudaMalloc(&arr[k], (size_t)n * sizeof(double4)); cudaMemcpy(arr[k], h.data(), (size_t)n * sizeof(double4), cudaMemcpyHostToDevice);
    }
    cudaMalloc(&out, n * sizeof(double));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int g = (n + 127) / 128;
    std::vector<double> r0(n), r1(n);
    for (int mode = 0; mode < 2; ++mode) {
        float best = 1e9f;
        for (int t = 0; t < 5; ++t) {
            cudaEventRecord(e0);
            if (mode == 0) k_list<<<g, 128>>>(n, d_nbr, d_cnt, arr[0], arr[1], arr[2], arr[3], arr[4], out);
            else k_cell<<<g, 128>>>(n, P, d_hit, d_cell_of, d_cstart, d_off, ncell, arr[0], arr[1], arr[2], arr[3], arr[4], out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (t > 0 && ms < best) best = ms;
        }
        cudaMemcpy(mode == 0 ? r0.data() : r1.data(), out, n * sizeof(double), cudaMemcpyDeviceToHost);
        const cudaError_t e = cudaGetLastError();
        std::printf("%-44s %8.3f ms   %.2f ns per particle   %s\n", mode == 0 ? "per-lane list walk (library's form)" : "cell-synchronous walk", best, 1e6 * best / n, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    double md = 0.0, mx = 0.0; for (int s = 0; s < n; ++s) { md = std::max(md, std::abs(r0[s] - r1[s])); mx = std::max(mx, std::abs(r0[s])); }
    std::printf("same sums up to order: max |diff| / max = %.2e\n", md / mx);
    return 0;
}
