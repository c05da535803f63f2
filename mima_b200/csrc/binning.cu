// binning.cu -- orders the columns of every layer by the k-table row key of their cells (sm_100a).
//
// Why: taumol evaluates tau(g) = sum_k w_k * T[row_k][g] with thread <-> (column, layer) cell, and a 16-byte table
// load costs one L1 data-pipe pass per DISTINCT row among the 32 lanes of the warp.  With a warp = 32 adjacent columns
// of one layer the lanes of the bench columns spread over 2.1 (single-species bands) to 4.2 (binary-species bands)
// rows per load; ordered by (lower, jp, jt, jt1, js) they share one (1.0 / 1.02, measured on the T170L60 columns).
// The reference has no counterpart: its column loop (rrtmg_lw_rad.nomcica.f90:453, rrtmg_sw_rad.nomcica.f90:500) treats
// cells one at a time, and the order in which independent cells are evaluated does not change any result.
//
// One block per layer: histogram of the layer's keys in shared memory (16384 bins, 64 KB), exclusive scan, then the
// scatter perm[lay][offset[key]++] = col.  Lanes with equal keys are aggregated with match.any so that a bin sees one
// atomic per warp.  The order inside a bin depends on the schedule; every cell's result is independent of it.
#include "rrtmg_dev.cuh"

namespace rrtmg {

constexpr int BIN_THREADS = 1024;

__global__ void __launch_bounds__(BIN_THREADS) bin_cells_kernel(const uint16_t *__restrict__ skey, int *__restrict__ perm, int nc)
{
    extern __shared__ unsigned s_bin[];                 // BIN_COUNT counters, then offsets
    __shared__ unsigned s_warp[BIN_THREADS / 32];
    const int lay = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const uint16_t *key = skey + (size_t)lay * nc;
    int *out = perm + (size_t)lay * nc;
    for (int i = tid; i < BIN_COUNT; i += BIN_THREADS) s_bin[i] = 0u;
    __syncthreads();
    const int nround = (nc + BIN_THREADS - 1) / BIN_THREADS;
    for (int r = 0; r < nround; ++r) {
        const int c = r * BIN_THREADS + tid;
        const bool live = c < nc;
        const unsigned k = live ? key[c] : 0xffffffffu;
        const unsigned m = __match_any_sync(0xffffffffu, k);
        if (live && lane == __ffs(m) - 1) atomicAdd(&s_bin[k], (unsigned)__popc(m));
    }
    __syncthreads();
    // exclusive scan of the 16384 counters: 16 per thread, warp scan, scan of the 32 warp totals
    constexpr int PER = BIN_COUNT / BIN_THREADS;
    unsigned v[PER], sum = 0u;
#pragma unroll
    for (int j = 0; j < PER; ++j) { v[j] = s_bin[tid * PER + j]; sum += v[j]; }
    unsigned inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp[tid >> 5] = inc;
    __syncthreads();
    if (tid < 32) {
        unsigned w = s_warp[tid], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        s_warp[tid] = wi - w;
    }
    __syncthreads();
    unsigned run = s_warp[tid >> 5] + inc - sum;
#pragma unroll
    for (int j = 0; j < PER; ++j) { s_bin[tid * PER + j] = run; run += v[j]; }
    __syncthreads();
    for (int r = 0; r < nround; ++r) {
        const int c = r * BIN_THREADS + tid;
        const bool live = c < nc;
        const unsigned k = live ? key[c] : 0xffffffffu;
        const unsigned m = __match_any_sync(0xffffffffu, k);
        const int leader = __ffs(m) - 1;
        unsigned base = 0u;
        if (live && lane == leader) base = atomicAdd(&s_bin[k], (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (live) out[base + __popc(m & ((1u << lane) - 1u))] = c;
    }
}

int bin_cells(const uint16_t *skey, int *perm, int nc, int nlay, cudaStream_t s)
{
    const size_t smem = (size_t)BIN_COUNT * sizeof(unsigned);
    cudaFuncSetAttribute(bin_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bin_cells_kernel<<<nlay, BIN_THREADS, smem, s>>>(skey, perm, nc);
    return 1;
}

} // namespace rrtmg
