// Micro-benchmark for the next kernel experiment: what does the momentum pass pay for gathering SIX separate 32-byte
// records per neighbour (Q0, Q1, Q2, S0, S1, S2: six sectors in six arrays) compared with ONE packed record of 20 doubles (x, v, m, rho,
// sigma/rho^2, TIR: five consecutive sectors)? Same lattice-like neighbour lists (56 neighbours within 2.4 dx on a cubic lattice, ELL
// transposed like the library's), same trivial arithmetic, so the difference is the L1 / L2 data path alone.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 rec_layout.cu -o rec_layout ; run on a B200 (a few seconds).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ double4 ld256(const double4 *p)
{
    double4 r; asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p)); return r;
}
// six arrays, one sector each
__global__ void __launch_bounds__(128, 3) k_six(int n, int K, const int *nbr, const int *cnt, const double4 *a0, const double4 *a1, const double4 *a2,
                                                 const double4 *a3, const double4 *a4, const double4 *a5, double *out)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x; if (s >= n) return;
    double acc = 0.0; const int c = cnt[s];
    for (int k = 0; k < c; ++k) {
        const int j = nbr[(size_t)k * n + s];
        const double4 q0 = ld256(a0 + j), q1 = ld256(a1 + j), q2 = ld256(a2 + j), q3 = ld256(a3 + j), q4 = ld256(a4 + j), q5 = ld256(a5 + j);
        acc += (q0.x + q1.y) + (q2.z + q3.w) + (q4.x + q5.y) + (q0.w + q1.w) + (q2.x + q3.x) + (q4.w + q5.w);
    }
    out[s] = acc;
}
// one packed record: 5 consecutive sectors (stride 160 bytes)
__global__ void __launch_bounds__(128, 3) k_packed(int n, int K, const int *nbr, const int *cnt, const double4 *rec, double *out)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x; if (s >= n) return;
    double acc = 0.0; const int c = cnt[s];
    for (int k = 0; k < c; ++k) {
        const int j = nbr[(size_t)k * n + s];
        const double4 *p = rec + (size_t)5 * j;
        const double4 q0 = ld256(p), q1 = ld256(p + 1), q2 = ld256(p + 2), q3 = ld256(p + 3), q4 = ld256(p + 4);
        acc += (q0.x + q1.y) + (q2.z + q3.w) + (q4.x + q0.y) + (q0.w + q1.w) + (q2.x + q3.x) + (q4.w + q1.x);
    }
    out[s] = acc;
}

int main()
{
    const int L = 160, n = L * L * L;                 // 4.1 M particles, x fastest: neighbours of consecutive slots overlap like in cell order
    std::vector<int> off;
    for (int dz = -2; dz <= 2; ++dz) for (int dy = -2; dy <= 2; ++dy) for (int dx = -2; dx <= 2; ++dx) {
        const int d2 = dx * dx + dy * dy + dz * dz;
        if (d2 > 0 && d2 * 1.0 <= 2.4 * 2.4) off.push_back((dz * L + dy) * L + dx);
    }
    const int K = (int)off.size();
    std::printf("n = %d, neighbours per interior particle = %d\n", n, K);
    std::vector<int> nbr((size_t)K * n), cnt(n);
    for (int z = 0; z < L; ++z) for (int y = 0; y < L; ++y) for (int x = 0; x < L; ++x) {
        const int s = (z * L + y) * L + x; int c = 0;
        for (int dz = -2; dz <= 2; ++dz) for (int dy = -2; dy <= 2; ++dy) for (int dx = -2; dx <= 2; ++dx) {
            const int d2 = dx * dx + dy * dy + dz * dz;
            if (d2 == 0 || d2 * 1.0 > 2.4 * 2.4) continue;
            const int X = x + dx, Y = y + dy, Z = z + dz;
            if (X < 0 || Y < 0 || Z < 0 || X >= L || Y >= L || Z >= L) continue;
            nbr[(size_t)c * n + s] = (Z * L + Y) * L + X; ++c;
        }
        cnt[s] = c;
    }
    int *d_nbr, *d_cnt; double4 *a[6], *rec; double *out;
    cudaMalloc(&d_nbr, nbr.size() * sizeof(int)); cudaMalloc(&d_cnt, n * sizeof(int)); cudaMalloc(&out, n * sizeof(double));
    cudaMemcpy(d_nbr, nbr.data(), nbr.size() * sizeof(int), cudaMemcpyHostToDevice); cudaMemcpy(d_cnt, cnt.data(), n * sizeof(int), cudaMemcpyHostToDevice);
    for (int k = 0; k < 6; ++k) { cudaMalloc(&a[k], (size_t)n * sizeof(double4)); cudaMemset(a[k], 0, (size_t)n * sizeof(double4)); }
    cudaMalloc(&rec, (size_t)5 * n * sizeof(double4)); cudaMemset(rec, 0, (size_t)5 * n * sizeof(double4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int g = (n + 127) / 128;
    for (int mode = 0; mode < 2; ++mode) {
        float best = 1e9f;
        for (int t = 0; t < 5; ++t) {
            cudaEventRecord(e0);
            if (mode == 0) k_six<<<g, 128>>>(n, K, d_nbr, d_cnt, a[0], a[1], a[2], a[3], a[4], a[5], out);
            else k_packed<<<g, 128>>>(n, K, d_nbr, d_cnt, rec, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (t > 0 && ms < best) best = ms;
        }
        const cudaError_t e = cudaGetLastError();
        std::printf("%-34s %8.3f ms   %.2f ns per particle   %s\n", mode == 0 ? "six arrays, six sectors" : "one packed record, five sectors", best, 1e6 * best / n,
                    e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
