// Micro-benchmark: L1 data-pipe cost of per-lane 32-byte record gathers on sm_100a (records resident in L1).
// Patterns: how many lanes of a warp fall into the same 128-byte line, 2 x LDG.128 vs 1 x LDG.256, and LDS from shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 l1_gather.cu -o l1_gather ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <bool W256> __device__ __forceinline__ double4 ldrec(const double4 *p)
{
    if (W256) { double4 r; asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p)); return r; }
    return *p;
}
// each lane walks its own index list (idx[it*32 + lane]); records live in a small table (L1 resident)
template <bool W256> __global__ void k_gather(const double4 *tab, const int *idx, int iters, int reps, double *out)
{
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int r = 0; r < reps; ++r)
        for (int it = 0; it < iters; ++it) {
            const int j = idx[it * 32 + lane];
            const double4 v = ldrec<W256>(tab + j);
            acc += v.x + v.w;
        }
    if (acc == 1.2345) out[0] = acc;
}
__global__ void k_gather_smem(const double4 *tab, int ntab, const int *idx, int iters, int reps, double *out)
{
    extern __shared__ double4 sm[];
    for (int i = threadIdx.x; i < ntab; i += blockDim.x) sm[i] = tab[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int r = 0; r < reps; ++r)
        for (int it = 0; it < iters; ++it) {
            const int j = idx[it * 32 + lane];
            const double4 v = sm[j];
            acc += v.x + v.w;
        }
    if (acc == 1.2345) out[0] = acc;
}

int main()
{
    const int ntab = 1024;            // 32 KB of records
    const int iters = 256, reps = 200;
    double4 *tab; cudaMalloc(&tab, ntab * sizeof(double4)); cudaMemset(tab, 0, ntab * sizeof(double4));
    double *out; cudaMalloc(&out, 8);
    int *d_idx; cudaMalloc(&d_idx, iters * 32 * sizeof(int));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *names[] = { "coalesced (32 consecutive records)", "groups of 4 lanes share a 128B line, lines random", "pairs of lanes share a 64B half line, random",
                            "fully random records", "all lanes same record (broadcast)", "lane l -> record 4*l (one sector per line)", "random, ~22 distinct lines with jitter (lattice-like)" };
    for (int pat = 0; pat < 7; ++pat) {
        std::vector<int> h(iters * 32);
        srand(1234);
        for (int it = 0; it < iters; ++it) {
            int base = rand() % (ntab - 64);
            for (int l = 0; l < 32; ++l) {
                int j;
                switch (pat) {
                case 0: j = (base & ~3) + l; break;
                case 1: { static int g[8]; if (l % 4 == 0) g[l / 4] = (rand() % (ntab / 4)) * 4; j = g[l / 4] + (l % 4); } break;
                case 2: { static int g[16]; if (l % 2 == 0) g[l / 2] = (rand() % (ntab / 2)) * 2; j = g[l / 2] + (l % 2); } break;
                case 3: j = rand() % ntab; break;
                case 4: j = base; break;
                case 5: j = (4 * l + (base & ~3)) % ntab; break;
                default: j = (base + l + (rand() % 5) - 2 + ntab) % ntab; break;
                }
                h[it * 32 + l] = j;
            }
        }
        cudaMemcpy(d_idx, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice);
        for (int mode = 0; mode < 3; ++mode) {
            const int blocks = 148 * 3, threads = 128;
            float best = 1e9f;
            for (int t = 0; t < 3; ++t) {
                cudaEventRecord(e0);
                if (mode == 0) k_gather<false><<<blocks, threads>>>(tab, d_idx, iters, reps, out);
                else if (mode == 1) k_gather<true><<<blocks, threads>>>(tab, d_idx, iters, reps, out);
                else k_gather_smem<<<blocks, threads, ntab * sizeof(double4)>>>(tab, ntab, d_idx, iters, reps, out);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            const double warp_reqs = (double)blocks * (threads / 32) * iters * reps;
            const double cyc_per_req = best * 1e-3 * 1.965e9 * 148 / warp_reqs;     // SM-cycles per warp-level 1 KB record request
            printf("%-60s %-12s %8.3f ms  %6.2f SM-cycles per warp request (1024 B)  %7.1f GB/s/SM-equivalent\n", names[pat],
                   mode == 0 ? "2xLDG.128" : (mode == 1 ? "LDG.256" : "LDS (smem)"), best, cyc_per_req, 1024.0 / cyc_per_req * 1.965);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
