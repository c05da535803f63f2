// Microbenchmark 2: the same broadcast-heavy table reads when the lanes of a warp spread over D distinct rows
// (D = 1, 2, 4), from L1 (LDG.128), from shared memory with explicit LDS.128 (row stride 128 B, and padded to 144 B),
// and from shared memory through a generic pointer (LD.E.128).  Decides whether staging k-table windows in shared
// memory can pay in taumol.
#include <cstdio>
#include <cuda_runtime.h>

// SRC: 0 global (__ldg), 1 shared (explicit), 2 shared through a generic pointer the compiler cannot classify
template <int SRC, int D, int STRIDE>
__global__ void __launch_bounds__(256, 2) k(const double *__restrict__ tab, double *out, int iters, int nrows, int flag)
{
    extern __shared__ __align__(16) double s[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (SRC) { for (int i = threadIdx.x; i < nrows * 16; i += blockDim.x) s[(i / 16) * STRIDE + (i & 15)] = tab[i]; __syncthreads(); }
    const double *gen = flag ? s : tab;          // flag = 1 at run time: generic pointer into shared memory
    double a0 = 0, a1 = 0;
    int row = (blockIdx.x * 8 + wid) % (nrows - 8);
    for (int it = 0; it < iters; ++it) {
        const int r = row + (lane % D);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            double2 v;
            if (SRC == 0) v = __ldg(reinterpret_cast<const double2 *>(tab + r * 16) + j);
            if (SRC == 1) v = reinterpret_cast<const double2 *>(s + r * STRIDE)[j];
            if (SRC == 2) v = reinterpret_cast<const double2 *>(gen + r * STRIDE)[j];
            a0 = fma(v.x, 1.0001, a0); a1 = fma(v.y, 1.0001, a1);
        }
        row = (row + 3) % (nrows - 8);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1;
}

template <int SRC, int D, int STRIDE>
void run(const char *name, const double *tab, double *out)
{
    const int nrows = 256, iters = 2000, blocks = 148 * 2;
    const size_t sm = SRC ? (size_t)nrows * STRIDE * 8 : 0;
    cudaFuncSetAttribute(k<SRC, D, STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a);
        k<SRC, D, STRIDE><<<blocks, 256, sm>>>(tab, out, iters, nrows, SRC == 2);
        cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
    }
    const double bytes_per_sm = (double)2 * 8 * iters * 8 * 32 * 16;
    printf("%-44s %8.3f ms  %7.1f B/clk/SM  (%4.2f clk per warp load)  err=%s\n", name, ms, bytes_per_sm / (ms * 1e-3 * 1.965e9),
           (ms * 1e-3 * 1.965e9) / (2.0 * 8 * iters * 8), cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    double *tab, *out;
    cudaMalloc(&tab, 256 * 16 * 8); cudaMalloc(&out, 148 * 2 * 256 * 8);
    cudaMemset(tab, 0, 256 * 16 * 8);
    run<0, 1, 16>("LDG.128, 1 row per warp", tab, out);
    run<0, 2, 16>("LDG.128, 2 rows per warp", tab, out);
    run<0, 4, 16>("LDG.128, 4 rows per warp", tab, out);
    run<1, 1, 16>("LDS.128, 1 row per warp", tab, out);
    run<1, 2, 16>("LDS.128, 2 rows per warp, stride 128 B", tab, out);
    run<1, 4, 16>("LDS.128, 4 rows per warp, stride 128 B", tab, out);
    run<1, 2, 18>("LDS.128, 2 rows per warp, stride 144 B", tab, out);
    run<1, 4, 18>("LDS.128, 4 rows per warp, stride 144 B", tab, out);
    run<2, 1, 16>("generic LD.128 -> shared, 1 row per warp", tab, out);
    run<2, 2, 16>("generic LD.128 -> shared, 2 rows per warp", tab, out);
    return 0;
}
