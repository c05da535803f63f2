// Microbenchmark: throughput of warp-uniform ("broadcast") loads from an L1-resident table on sm_100a, as the taumol
// kernels issue them: every lane of a warp reads the same 16 bytes (or 8) of a k-table row and multiplies them into
// its own accumulators.  Compares LDG.128 / LDG.64 / LDS.128 / LDS.64 and per-lane-distinct rows.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256, 2) k(const double *__restrict__ tab, double *out, int iters, int nrows)
{
    extern __shared__ double s[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (MODE >= 2 && MODE <= 3) { for (int i = threadIdx.x; i < nrows * 16; i += blockDim.x) s[i] = tab[i]; __syncthreads(); }
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int row = (blockIdx.x * 8 + wid) % nrows;
    for (int it = 0; it < iters; ++it) {
        int r = MODE == 4 ? (row + lane * 7) % nrows : row;      // MODE 4: every lane its own row
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0 || MODE == 4) { const double2 v = __ldg(reinterpret_cast<const double2 *>(tab + r * 16) + j); a0 = fma(v.x, 1.0001, a0); a1 = fma(v.y, 1.0001, a1); }
            if (MODE == 1) { const double v = __ldg(tab + r * 16 + 2 * j), u = __ldg(tab + r * 16 + 2 * j + 1); a0 = fma(v, 1.0001, a0); a1 = fma(u, 1.0001, a1); }
            if (MODE == 2) { const double2 v = reinterpret_cast<const double2 *>(s + r * 16)[j]; a2 = fma(v.x, 1.0001, a2); a3 = fma(v.y, 1.0001, a3); }
            if (MODE == 3) { const double v = s[r * 16 + 2 * j], u = s[r * 16 + 2 * j + 1]; a2 = fma(v, 1.0001, a2); a3 = fma(u, 1.0001, a3); }
        }
        row = (row + 3) % nrows;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}

int main()
{
    const int nrows = 512, iters = 2000, blocks = 148 * 2;
    double *tab, *out;
    cudaMalloc(&tab, nrows * 16 * 8); cudaMalloc(&out, blocks * 256 * 8);
    cudaMemset(tab, 0, nrows * 16 * 8);
    const char *names[5] = {"LDG.128 uniform", "LDG.64 x2 uniform", "LDS.128 uniform", "LDS.64 x2 uniform", "LDG.128 per-lane rows"};
    for (int m = 0; m < 5; ++m) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            const size_t sm = nrows * 16 * 8;
            if (m == 0) k<0><<<blocks, 256, 0>>>(tab, out, iters, nrows);
            if (m == 1) k<1><<<blocks, 256, 0>>>(tab, out, iters, nrows);
            if (m == 2) { cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); k<2><<<blocks, 256, sm>>>(tab, out, iters, nrows); }
            if (m == 3) { cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); k<3><<<blocks, 256, sm>>>(tab, out, iters, nrows); }
            if (m == 4) k<4><<<blocks, 256, 0>>>(tab, out, iters, nrows);
            cudaEventRecord(b); cudaEventSynchronize(b);
        }
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double bytes_per_sm = (double)2 * 8 * iters * 8 * 32 * 16;       // 2 blocks x 8 warps x iters x 8 x (32 lanes x 16 B)
        printf("%-24s %8.3f ms  %7.1f B/clk/SM returned to registers (at 1.965 GHz)  err=%s\n", names[m], ms,
               bytes_per_sm / (ms * 1e-3 * 1.965e9), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
