// lw_solver.cu -- RRTMG longwave radiative transfer on sm_100a: the STAGED solvers (they read the [col][lay][g] staging that
// lw_taumol_kernel writes).  They serve idrv = 1, cloudy skies and stage capture; clear-sky calls without derivatives run
// the fused column kernel of lw_column.cu instead (option lw_fused = 0 sends those here as well).
//
// rtrn: LW/src/rrtmg_lw_rtrnmr.f90:481-777 ("Clear layer" branches; identical in rtrnmc.f90:407-432,481-503).
//       taut = taug + tauaer (rad.nomcica:514-519, iaer = 10 forced).
//
// Block <-> column, thread <-> g-point (the staging fields are [col][lay][g], so one level step of a warp
// reads 256 contiguous bytes per field).  Levels are processed four at a time: the eight staging loads of
// a group are issued before any of its arithmetic, which is what keeps enough bytes in flight to stream
// from HBM.  The up sweep recomputes the layer transmittance and source from taug/fracs instead of
// storing them (a second read of 16 B per cell instead of a write plus a read).  The sum over g-points
// goes through shared memory in batches of 16 levels (tile_reduce16, fixed summation order).
//
// This translation unit is compiled with FMA contraction on (build.py): the flux arithmetic has no
// index/branch decisions that depend on the last bit, unlike setcoef in lw_kernels.cu.
#include "rrtmg_dev.cuh"

namespace rrtmg {

struct LwSolverConst {
    double delwave[NBNDLW];
    double heatfac, fluxfac, bpade;
    unsigned char ngb[NGPTLW];
};
__constant__ LwSolverConst c_ls;

int lw_solver_upload_const(const LwConst &c, const unsigned char *ngb)
{
    LwSolverConst h;
    for (int b = 0; b < NBNDLW; ++b) h.delwave[b] = c.delwave[b];
    h.heatfac = c.heatfac; h.fluxfac = c.fluxfac; h.bpade = c.bpade;
    for (int g = 0; g < NGPTLW; ++g) h.ngb[g] = ngb[g];
    return cudaMemcpyToSymbol(c_ls, &h, sizeof h) == cudaSuccess ? 0 : -1;
}

constexpr int RT_THREADS = 160;   // 140 g-points -> 5 warps
constexpr int RT_S = 141;         // tile row stride (odd)
constexpr int RT_U = 4;           // levels per load group
constexpr int RT_WARPS = RT_THREADS / 32;
constexpr int RT_WS = 34;         // row stride of the warp-local tile (upper half-warp shifted by one: conflict-free)

// sum over the 32 lanes of each of the 8 rows of a warp-private tile; lanes 4r..4r+3 return the sum of row r
__device__ __forceinline__ double warp_rows8(const double *tile, int lane)
{
    __syncwarp();
    const int q = lane & 3;
    const double *src = tile + (lane >> 2) * RT_WS + 17 * (q >> 1) + 8 * (q & 1);
    double acc = src[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) acc += src[j];
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    __syncwarp();
    return acc;
}

// layer transmittance and Planck-weighted sources of one (g, layer) cell (:589-607)
template <bool DOWN>
__device__ __forceinline__ void lw_layer(const double2 *__restrict__ et, double bpade, double secd, double taut,
                                         double plfrac, double blay, double dplankup, double dplankdn,
                                         double &atrans, double &bbd, double &bbugas)
{
    const double rec_6 = 0.166667;
    double odepth = secd * taut;
    if (odepth < 0.0) odepth = 0.0;
    if (odepth <= 0.06) {
        atrans = odepth - 0.5 * odepth * odepth;
        odepth = rec_6 * odepth;
        if (DOWN) bbd = plfrac * (blay + dplankdn * odepth);
        else bbugas = plfrac * (blay + dplankup * odepth);
    } else {
        const double tblind = odepth * rcp_fast(bpade + odepth);
        const int itr = (int)(10000.0 * tblind + 0.5);
        const double2 e = ld_tbl(et + itr);
        atrans = 1. - e.x;
        if (DOWN) bbd = plfrac * (blay + e.y * dplankdn);
        else bbugas = plfrac * (blay + e.y * dplankup);
    }
}

#ifdef RRTMG_B200_DEV_VARIANTS       // direct-load form of the clear-sky kernel (option lw_rtrn_variant = 0, 1): development builds only
// The Planck sources of the column ([lay][16] and [lev][16], 16 KB at 60 layers) are staged in shared memory
// once per block: the 140 g-threads need them 2-3 times per level and they are shared by all g-points of a band.
// WR: warp-local g-sums -- every warp adds up its own 32 lanes per level (batches of 8 levels through a
// 2 KB warp-private tile, no block barrier inside the sweeps) and the five warp partials of a level are
// combined in a fixed order once at the end.
template <bool AER, bool WR>
__global__ void __launch_bounds__(RT_THREADS) lw_rtrn_kernel(LwTables T, LwIn in, LwOut out, LwWork w)
{
    __shared__ double s_tile[WR ? RT_WARPS * 8 * RT_WS : 16 * RT_S];
    __shared__ double s_part[WR ? RT_WARPS * 2 * (MAXLAY + 1) : 16 * (RT_THREADS / 16 + 1)];
    __shared__ double s_dn[MAXLAY + 1], s_up[MAXLAY + 1];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double *wt = s_tile + (WR ? wid * 8 * RT_WS + lane + (lane >> 4) : 0);
    extern __shared__ __align__(16) double s_planck[];          // pl[nlay][16] then pv[nlay+1][16]
    const int col = blockIdx.x;
    const int nlay = w.nlay;
    const int g = threadIdx.x;
    const bool active = g < NGPTLW;
    const int band = active ? c_ls.ngb[g] : 0;
    {
        const double2 *src = reinterpret_cast<const double2 *>(w.planklay + (size_t)col * nlay * 16);
        double2 *dst = reinterpret_cast<double2 *>(s_planck);
        for (int i = threadIdx.x; i < nlay * 8; i += RT_THREADS) dst[i] = src[i];
        src = reinterpret_cast<const double2 *>(w.planklev + (size_t)col * (nlay + 1) * 16);
        dst = reinterpret_cast<double2 *>(s_planck + nlay * 16);
        for (int i = threadIdx.x; i < (nlay + 1) * 8; i += RT_THREADS) dst[i] = src[i];
    }
    const double secd = w.secdiff[(size_t)col * 16 + band];
    const double wgt = active ? 0.5 * c_ls.delwave[band] : 0.0;     // wtdiff * delwave
    const double bpade = c_ls.bpade;
    const double2 *__restrict__ et = reinterpret_cast<const double2 *>(T.exptfn);
    const double *__restrict__ taug = w.taug + (size_t)col * nlay * NGPTLW + (active ? g : 0);
    const double *__restrict__ fracs = w.fracs + (size_t)col * nlay * NGPTLW + (active ? g : 0);
    const double *pl = s_planck + band;
    const double *pv = s_planck + nlay * 16 + band;
    const double *taer = AER ? in.tauaer + col + (size_t)band * nlay * in.ld : nullptr;
    __syncthreads();

    // ---- downward sweep (:505-618), k counts layers from the top
    double radld = 0.0;
    double plfrac1 = 0.0;
    double pup = pv[nlay * 16];                       // Planck at the upper interface of the current layer
    double tgn[RT_U], frn[RT_U];          // loads of the next group, issued before the current group's arithmetic
#pragma unroll
    for (int j = 0; j < RT_U; ++j) {
        const int lay = max(nlay - 1 - j, 0);
        tgn[j] = taug[(size_t)lay * NGPTLW];
        frn[j] = fracs[(size_t)lay * NGPTLW];
        if (AER) tgn[j] = tgn[j] + taer[(size_t)lay * in.ld];
    }
    for (int k0 = 0; k0 < nlay; k0 += RT_U) {
        double tg[RT_U], fr[RT_U];
        const bool full = k0 + RT_U <= nlay;
#pragma unroll
        for (int j = 0; j < RT_U; ++j) { tg[j] = tgn[j]; fr[j] = frn[j]; }
        if (k0 + RT_U < nlay) {
#pragma unroll
            for (int j = 0; j < RT_U; ++j) {
                const int lay = max(nlay - 1 - (k0 + RT_U + j), 0);      // clamped in the ragged tail
                tgn[j] = taug[(size_t)lay * NGPTLW];
                frn[j] = fracs[(size_t)lay * NGPTLW];
                if (AER) tgn[j] = tgn[j] + taer[(size_t)lay * in.ld];
            }
        }
#pragma unroll
        for (int j = 0; j < RT_U; ++j) {
            const int k = k0 + j;
            if (full || k < nlay) {
                const int lay = nlay - 1 - k;
                const double blay = pl[lay * 16], pdn = pv[lay * 16];
                double atrans, bbd, bbugas;
                lw_layer<true>(et, bpade, secd, tg[j], fr[j], blay, pup - blay, pdn - blay, atrans, bbd, bbugas);
                pup = pdn;
                radld = radld + (bbd - radld) * atrans;
                if (WR) wt[(k & 7) * RT_WS] = radld * wgt;
                else if (active) s_tile[(k & 15) * RT_S + g] = radld * wgt;
                if (k == nlay - 1) plfrac1 = fr[j];
            }
        }
        const int klast = min(k0 + RT_U, nlay) - 1;
        if (WR) {
            if ((klast & 7) == 7 || klast == nlay - 1) {
                const double sum = warp_rows8(s_tile + wid * 8 * RT_WS, lane);
                const int kk = (klast & ~7) + (lane >> 2);
                if ((lane & 3) == 0 && kk <= klast) s_part[(wid * 2) * (MAXLAY + 1) + nlay - 1 - kk] = sum;
            }
        } else if ((klast & 15) == 15 || klast == nlay - 1) {
            const double sum = tile_reduce16<RT_THREADS, NGPTLW, RT_S>(s_tile, s_part);
            const int kk = (klast & ~15) + threadIdx.x;
            if (threadIdx.x < 16 && kk <= klast) s_dn[nlay - 1 - kk] = sum * c_ls.fluxfac;
        }
    }
    if (threadIdx.x == 0) s_dn[nlay] = 0.0;   // no downward flux enters at the top (drad(nlayers) = 0)
    if (WR && lane == 0) s_part[(wid * 2) * (MAXLAY + 1) + nlay] = 0.0;

    // ---- surface (:628-636) and upward sweep (:649-711); level k = 0 is the surface, level k > 0 the top
    //      of layer k (1-based).  Groups are aligned to the 16-level batches of the reduction.
    double radlu = 0.0;
#pragma unroll
    for (int j = 0; j < RT_U; ++j) {
        const int lay = min(max(j, 1), nlay) - 1;
        tgn[j] = taug[(size_t)lay * NGPTLW];
        frn[j] = fracs[(size_t)lay * NGPTLW];
        if (AER) tgn[j] = tgn[j] + taer[(size_t)lay * in.ld];
    }
    for (int k0 = 0; k0 <= nlay; k0 += RT_U) {
        double tg[RT_U], fr[RT_U];
        const bool full = k0 > 0 && k0 + RT_U - 1 <= nlay;
#pragma unroll
        for (int j = 0; j < RT_U; ++j) { tg[j] = tgn[j]; fr[j] = frn[j]; }
        if (k0 + RT_U <= nlay) {
#pragma unroll
            for (int j = 0; j < RT_U; ++j) {
                const int lay = min(k0 + RT_U + j, nlay) - 1;
                tgn[j] = taug[(size_t)lay * NGPTLW];
                frn[j] = fracs[(size_t)lay * NGPTLW];
                if (AER) tgn[j] = tgn[j] + taer[(size_t)lay * in.ld];
            }
        }
#pragma unroll
        for (int j = 0; j < RT_U; ++j) {
            const int k = k0 + j;
            if (!full && k == 0) {
                const double semiss = in.emis ? in.emis[col + (size_t)band * in.ld] : 1.0;
                const double rad0 = plfrac1 * w.plankbnd[(size_t)col * 16 + band];
                const double reflect = 1. - semiss;
                radlu = rad0 + reflect * radld;
                if (WR) wt[0] = radlu * wgt;
                else if (active) s_tile[g] = radlu * wgt;
            } else if (full || k <= nlay) {
                const int lay = k - 1;
                const double blay = pl[lay * 16];
                double atrans, bbd, bbugas;
                lw_layer<false>(et, bpade, secd, tg[j], fr[j], blay, pv[(lay + 1) * 16] - blay, 0.0, atrans, bbd, bbugas);
                radlu = radlu + (bbugas - radlu) * atrans;
                if (WR) wt[(k & 7) * RT_WS] = radlu * wgt;
                else if (active) s_tile[(k & 15) * RT_S + g] = radlu * wgt;
            }
        }
        const int klast = min(k0 + RT_U - 1, nlay);
        if (WR) {
            if ((klast & 7) == 7 || klast == nlay) {
                const double sum = warp_rows8(s_tile + wid * 8 * RT_WS, lane);
                const int kk = (klast & ~7) + (lane >> 2);
                if ((lane & 3) == 0 && kk <= klast) s_part[(wid * 2 + 1) * (MAXLAY + 1) + kk] = sum;
            }
        } else if ((klast & 15) == 15 || klast == nlay) {
            const double sum = tile_reduce16<RT_THREADS, NGPTLW, RT_S>(s_tile, s_part);
            const int kk = (klast & ~15) + threadIdx.x;
            if (threadIdx.x < 16 && kk <= klast) s_up[kk] = sum * c_ls.fluxfac;
        }
    }
    __syncthreads();
    if (WR) {
        for (int lev = threadIdx.x; lev <= nlay; lev += RT_THREADS) {
            double d = 0.0, u = 0.0;
#pragma unroll
            for (int i = 0; i < RT_WARPS; ++i) {
                d += s_part[(i * 2) * (MAXLAY + 1) + lev];
                u += s_part[(i * 2 + 1) * (MAXLAY + 1) + lev];
            }
            s_dn[lev] = d * c_ls.fluxfac;
            s_up[lev] = u * c_ls.fluxfac;
        }
        __syncthreads();
    }

    // ---- fluxes and heating rates (:751-777), copy-out (rad.nomcica:546-555)
    for (int lev = threadIdx.x; lev <= nlay; lev += RT_THREADS) {
        const size_t o = col + (size_t)lev * out.ld;
        const double u = s_up[lev], d = s_dn[lev];
        out.uflx[o] = u; out.dflx[o] = d;
        out.uflxc[o] = u; out.dflxc[o] = d;
        if (lev < nlay) {
            const double fnet0 = u - d, fnet1 = s_up[lev + 1] - s_dn[lev + 1];
            const double pz0 = in.plev[col + (size_t)lev * in.ld], pz1 = in.plev[col + (size_t)(lev + 1) * in.ld];
            const double h = c_ls.heatfac * (fnet0 - fnet1) / (pz0 - pz1);
            out.hr[o] = h;
            out.hrc[o] = h;
        }
    }
}


#endif  // RRTMG_B200_DEV_VARIANTS

// =====================================================================================================
// Variant 2: the staging rows arrive by TMA.  The column streams through a ring of NST shared-memory stages
// filled by cp.async.bulk (1-D bulk copies completing on an mbarrier): per stage the taug and fracs rows of
// CH layers (2 x CH x 1120 B, contiguous in the [col][lay][g] staging layout) and the matching Planck rows
// (planklay CH x 128 B, planklev (CH+1) x 128 B).  The five g-point warps wait on the stage's "full"
// barrier, read everything with shared-memory loads and release the stage through an "empty" barrier (one
// arrival per warp); a sixth warp (one lane) re-arms the barrier and issues the copies of the chunk NST
// positions ahead (having the last g-point warp issue them instead measured 2 % slower).  No block barrier.  The chunk sequence of a column is fixed: down sweep top to bottom, then up sweep
// bottom to top (second read, mostly L2).  What this buys: the streaming loads no longer occupy the L1
// tag/miss path, which the 16-byte exp/tfn table gathers need (profiles/r01_summary.md), the Planck rows
// need no per-block prologue, and 64 registers suffice.  Arithmetic, warp-local g-sums and epilogue are
// those of lw_rtrn_kernel<., true>.
// =====================================================================================================
template <int CH> struct TmaGeom {
    static constexpr int STAGE = CH * 2 * NGPTLW + CH * 16 + (CH + 1) * 16;   // doubles per stage
};

// chunk q of the column's sequence: layers [lo, lo + n)
template <int CH>
__device__ __forceinline__ void tma_chunk(int q, int nd, int nlay, int &lo, int &n)
{
    if (q < nd) { const int hi = nlay - 1 - q * CH; lo = max(hi - CH + 1, 0); n = hi - lo + 1; }
    else { lo = (q - nd) * CH; n = min(CH, nlay - lo); }
}
template <int CH>
__device__ __forceinline__ void tma_issue(const LwWork &w, int col, int nlay, int lo, int n, double *d, uint64_t *bar)
{
    const double *gt = w.taug + ((size_t)col * nlay + lo) * NGPTLW, *gf = w.fracs + ((size_t)col * nlay + lo) * NGPTLW;
    const double *gl = w.planklay + ((size_t)col * nlay + lo) * 16, *gv = w.planklev + ((size_t)col * (nlay + 1) + lo) * 16;
    mbar_expect_tx(bar, (uint32_t)(n * (2 * NGPTLW + 16) * 8 + (n + 1) * 128));
    tma_load_1d(d, gt, (uint32_t)(n * NGPTLW * 8), bar);
    tma_load_1d(d + CH * NGPTLW, gf, (uint32_t)(n * NGPTLW * 8), bar);
    tma_load_1d(d + 2 * CH * NGPTLW, gl, (uint32_t)(n * 128), bar);
    tma_load_1d(d + 2 * CH * NGPTLW + CH * 16, gv, (uint32_t)((n + 1) * 128), bar);
}

// DRV (idrv = 1): also the derivative of the upward flux with respect to the surface temperature
// (rtrnmr.f90:629-646, 686-689, 736-746): d_radlu_dt starts as fracs(1) * dplankbnd_dt and is attenuated by
// (1 - atrans) per layer; its g-sums go through a second warp tile.
template <bool AER, int LMAX, int NST, int CH, bool DRV>
__global__ void __launch_bounds__(RT_THREADS + 32) lw_rtrn_tma_kernel(LwTables T, LwIn in, LwOut out, LwWork w)
{
    constexpr int STAGE = TmaGeom<CH>::STAGE;
    extern __shared__ __align__(128) double s_stage[];           // NST stages of STAGE doubles
    __shared__ __align__(8) uint64_t s_full[NST], s_empty[NST];
    __shared__ double s_tile[RT_WARPS * 8 * RT_WS];
    __shared__ double s_part[RT_WARPS * 2 * (LMAX + 1)];
    static_assert(2 * (LMAX + 1) <= RT_WARPS * 8 * RT_WS, "the level fluxes reuse the tile storage");
    __shared__ double s_dtile[DRV ? RT_WARPS * 8 * RT_WS : 1];
    __shared__ double s_dpart[DRV ? RT_WARPS * (LMAX + 1) : 1];
    double *s_dn = s_tile, *s_up = s_tile + LMAX + 1;           // only used after the sweeps
    const int col = blockIdx.x;
    const int nlay = w.nlay;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nd = (nlay + CH - 1) / CH;                         // chunks per sweep
    if (threadIdx.x == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], RT_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (wid == RT_WARPS) {
        // producer warp (one lane): chunk q goes to stage q % NST as soon as the five g-point warps have left it
        if (lane == 0) {
            for (int q = 0; q < 2 * nd; ++q) {
                const int st = q % NST, it = q / NST;
                int lo, n;
                tma_chunk<CH>(q, nd, nlay, lo, n);
                if (it > 0) mbar_wait(&s_empty[st], (uint32_t)((it - 1) & 1));
                tma_issue<CH>(w, col, nlay, lo, n, s_stage + (size_t)st * STAGE, &s_full[st]);
            }
        }
    } else {

    const int g = threadIdx.x;
    const bool active = g < NGPTLW;
    const int gc = active ? g : 0;
    const int band = c_ls.ngb[gc];
    const double secd = w.secdiff[(size_t)col * 16 + band];
    const double wgt = active ? 0.5 * c_ls.delwave[band] : 0.0;
    const double bpade = c_ls.bpade;
    const double2 *__restrict__ et = reinterpret_cast<const double2 *>(T.exptfn);
    const double *taer = AER ? in.tauaer + col + (size_t)band * nlay * in.ld : nullptr;
    double *wt = s_tile + wid * 8 * RT_WS + lane + (lane >> 4);
    // This warp has read everything it needs from stage `st`.  The stage is then overwritten through the ASYNC proxy
    // (TMA) while it was read through the generic proxy: the release needs fence.proxy.async before the arrival.
    // Without it 1-3 columns per 16384 saw a stage refilled under their loads in about one call out of thirty
    // (tools/stress_lw.py; tests/test_gpu_parity.py::test_repeated_calls_are_bitwise_reproducible).  The caller also
    // pins the values computed from the stage with keep() (an empty asm that consumes the register) so that the
    // shared-memory loads have returned before the arrival is issued.
    auto keep = [](double v) { asm volatile("" ::"d"(v) : "memory"); };
    auto release = [&](int st) {
        // generic-proxy reads of the stage -> async-proxy (TMA) overwrite: cross-proxy ordering needs the proxy fence
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[st]);
    };
    double radld = 0.0, plfrac1 = 0.0;
    int q = 0;
    // downward sweep (:505-618)
    for (; q < nd; ++q) {
        const int st = q % NST, it = q / NST;
        const int hi = nlay - 1 - q * CH, lo = max(hi - CH + 1, 0), n = hi - lo + 1;
        mbar_wait(&s_full[st], (uint32_t)(it & 1));
        const double *d = s_stage + (size_t)st * STAGE;
        const double *stg = d + gc, *sfr = d + CH * NGPTLW + gc;
        const double *spl = d + 2 * CH * NGPTLW + band, *spv = spl + CH * 16;
        double at[CH], bd[CH], frv[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int r = max(n - 1 - j, 0);                  // row inside the stage, top layer first
            double tg = stg[r * NGPTLW];
            if (AER) tg = tg + taer[(size_t)(lo + r) * in.ld];
            frv[j] = sfr[r * NGPTLW];
            const double blay = spl[r * 16];
            double bbu;
            lw_layer<true>(et, bpade, secd, tg, frv[j], blay, spv[(r + 1) * 16] - blay, spv[r * 16] - blay, at[j], bd[j], bbu);
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) { keep(at[j]); keep(bd[j]); }      // every load of the stage feeds these
        release(st);
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            if (j < n) {
                const int k = nlay - 1 - (hi - j);
                radld = radld + (bd[j] - radld) * at[j];
                wt[(k & 7) * RT_WS] = radld * wgt;
                if (k == nlay - 1) plfrac1 = frv[j];
                if ((k & 7) == 7 || k == nlay - 1) {
                    const double sum = warp_rows8(s_tile + wid * 8 * RT_WS, lane);
                    const int kk = (k & ~7) + (lane >> 2);
                    if ((lane & 3) == 0 && kk <= k) s_part[(wid * 2) * (LMAX + 1) + nlay - 1 - kk] = sum;
                }
            }
        }
    }
    if (lane == 0) s_part[(wid * 2) * (LMAX + 1) + nlay] = 0.0;
    // surface (:628-636)
    double radlu;
    {
        const double semiss = in.emis ? in.emis[col + (size_t)band * in.ld] : 1.0;
        const double rad0 = plfrac1 * w.plankbnd[(size_t)col * 16 + band];
        radlu = rad0 + (1. - semiss) * radld;
        wt[0] = radlu * wgt;
    }
    double drad = 0.0;
    double *dwt = s_dtile + (DRV ? wid * 8 * RT_WS + lane + (lane >> 4) : 0);
    if (DRV) {
        drad = plfrac1 * w.dplankbnd[(size_t)col * 16 + band];
        dwt[0] = drad * wgt;
    }
    // upward sweep (:649-711)
    for (; q < 2 * nd; ++q) {
        const int st = q % NST, it = q / NST;
        const int lo = (q - nd) * CH, n = min(CH, nlay - lo);
        mbar_wait(&s_full[st], (uint32_t)(it & 1));
        const double *d = s_stage + (size_t)st * STAGE;
        const double *stg = d + gc, *sfr = d + CH * NGPTLW + gc;
        const double *spl = d + 2 * CH * NGPTLW + band, *spv = spl + CH * 16;
        double at[CH], bu[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int r = min(j, n - 1);
            double tg = stg[r * NGPTLW];
            if (AER) tg = tg + taer[(size_t)(lo + r) * in.ld];
            const double blay = spl[r * 16];
            double bbd;
            lw_layer<false>(et, bpade, secd, tg, sfr[r * NGPTLW], blay, spv[(r + 1) * 16] - blay, 0.0, at[j], bbd, bu[j]);
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) { keep(at[j]); keep(bu[j]); }      // every load of the stage feeds these
        release(st);
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            if (j < n) {
                const int k = lo + j + 1;
                radlu = radlu + (bu[j] - radlu) * at[j];
                wt[(k & 7) * RT_WS] = radlu * wgt;
                if (DRV) {
                    drad = drad * (1.0 - at[j]);
                    dwt[(k & 7) * RT_WS] = drad * wgt;
                }
                if ((k & 7) == 7 || k == nlay) {
                    const double sum = warp_rows8(s_tile + wid * 8 * RT_WS, lane);
                    const int kk = (k & ~7) + (lane >> 2);
                    if ((lane & 3) == 0 && kk <= k) s_part[(wid * 2 + 1) * (LMAX + 1) + kk] = sum;
                    if (DRV) {
                        const double dsum = warp_rows8(s_dtile + wid * 8 * RT_WS, lane);
                        if ((lane & 3) == 0 && kk <= k) s_dpart[wid * (LMAX + 1) + kk] = dsum;
                    }
                }
            }
        }
    }
    }
    __syncthreads();
    for (int lev = threadIdx.x; lev <= nlay; lev += RT_THREADS + 32) {
        double d = 0.0, u = 0.0;
#pragma unroll
        for (int i = 0; i < RT_WARPS; ++i) {
            d += s_part[(i * 2) * (LMAX + 1) + lev];
            u += s_part[(i * 2 + 1) * (LMAX + 1) + lev];
        }
        s_dn[lev] = d * c_ls.fluxfac;
        s_up[lev] = u * c_ls.fluxfac;
    }
    __syncthreads();
    // fluxes and heating rates (:751-777), copy-out (rad.nomcica:546-555)
    for (int lev = threadIdx.x; lev <= nlay; lev += RT_THREADS + 32) {
        const size_t o = col + (size_t)lev * out.ld;
        const double u = s_up[lev], d = s_dn[lev];
        out.uflx[o] = u; out.dflx[o] = d;
        out.uflxc[o] = u; out.dflxc[o] = d;
        if (lev < nlay) {
            const double fnet0 = u - d, fnet1 = s_up[lev + 1] - s_dn[lev + 1];
            const double pz0 = in.plev[col + (size_t)lev * in.ld], pz1 = in.plev[col + (size_t)(lev + 1) * in.ld];
            const double h = c_ls.heatfac * (fnet0 - fnet1) / (pz0 - pz1);
            out.hr[o] = h;
            out.hrc[o] = h;
        }
        if (DRV) {
            double dd = 0.0;
#pragma unroll
            for (int i = 0; i < RT_WARPS; ++i) dd += s_dpart[i * (LMAX + 1) + lev];
            dd = dd * c_ls.fluxfac;
            out.duflx_dt[o] = dd;
            out.duflxc_dt[o] = dd;
        }
    }
}

template <bool AER, int LMAX, int NST, int CH>
static void launch_tma(const LwTables &t, const LwIn &in, const LwOut &out, LwWork &w, cudaStream_t s)
{
    const size_t smem = (size_t)NST * TmaGeom<CH>::STAGE * sizeof(double);
    if (w.idrv) {
        cudaFuncSetAttribute(lw_rtrn_tma_kernel<AER, LMAX, NST, CH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        lw_rtrn_tma_kernel<AER, LMAX, NST, CH, true><<<w.nc, RT_THREADS + 32, smem, s>>>(t, in, out, w);
        return;
    }
    cudaFuncSetAttribute(lw_rtrn_tma_kernel<AER, LMAX, NST, CH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    lw_rtrn_tma_kernel<AER, LMAX, NST, CH, false><<<w.nc, RT_THREADS + 32, smem, s>>>(t, in, out, w);
}
template <bool AER, int LMAX>
static void launch_tma_pick(const LwTables &t, const LwIn &in, const LwOut &out, LwWork &w, cudaStream_t s, int v)
{
    // measured at T170L60 (stages x layers per stage): 2x4 7.60 ms, 2x3 7.68, 3x3 7.77, 2x5 8.02, 4x2 8.06, 3x2 8.07, 2x6 8.31, 3x4 8.39,
    // 5x2 8.65, 6x1 9.74, 4x4 10.1; a suspend-time hint on the consumers' try_wait (200 ns .. 20 us): no change
    // (a 56-register build that fits six blocks per SM measured 8.4 ms)
#ifdef RRTMG_B200_DEV_VARIANTS
    if (v == 3) { launch_tma<AER, LMAX, 3, 4>(t, in, out, w, s); return; }
#endif
    (void)v;
    launch_tma<AER, LMAX, 2, 4>(t, in, out, w, s);
}

// =====================================================================================================
// Cloudy sky (icld >= 1, cloud optical depth given per band: cldprop's inflag = 0, rrtmg_lw_cldprop.f90:154-176).
// MR = false: random overlap, rrtmg_lw_rtrn.f90:302-316, 371-375, 480-485.  MR = true: maximum/random overlap,
// rrtmg_lw_rtrnmr.f90:316-479 (overlap factors, one thread per column), :569-588 and :653-674 (the cloudy and clear
// parts of the radiance carried separately and re-partitioned at every cloudy level).  The layer optics are the three
// branches of :516-567.  Not MiMA's configuration and not tuned: the layout of lw_rtrn_kernel (block <-> column,
// thread <-> g-point, the up sweep recomputes the layer quantities), direct loads, block-level g-sums over four
// values per level (total, clear, and with DRV the two surface-temperature derivatives).
// =====================================================================================================
struct LwCloudCell { double atrans, atot, bbd, bbdtot, gassrc, bbugas, bbutot; };

__device__ __forceinline__ void lw_cloud_cell(const double2 *__restrict__ et, double bpade, double secd, double taut, double plfrac,
                                              double blay, double dplankup, double dplankdn, bool cloudy, double odcld, LwCloudCell &c)
{
    const double rec_6 = 0.166667, tblint = 10000.0;
    double odepth = secd * taut;
    if (odepth < 0.0) odepth = 0.0;
    c.atot = 0.0; c.bbdtot = 0.0; c.gassrc = 0.0; c.bbutot = 0.0;
    if (cloudy) {
        double odtot = odepth + odcld;
        if (odtot < 0.06) {
            c.atrans = odepth - 0.5 * odepth * odepth;
            const double odepth_rec = rec_6 * odepth;
            c.gassrc = plfrac * (blay + dplankdn * odepth_rec) * c.atrans;
            c.atot = odtot - 0.5 * odtot * odtot;
            const double odtot_rec = rec_6 * odtot;
            c.bbdtot = plfrac * (blay + dplankdn * odtot_rec);
            c.bbd = plfrac * (blay + dplankdn * odepth_rec);
            c.bbugas = plfrac * (blay + dplankup * odepth_rec);
            c.bbutot = plfrac * (blay + dplankup * odtot_rec);
        } else if (odepth <= 0.06) {
            c.atrans = odepth - 0.5 * odepth * odepth;
            const double odepth_rec = rec_6 * odepth;
            c.gassrc = plfrac * (blay + dplankdn * odepth_rec) * c.atrans;
            const double tblind = odtot / (bpade + odtot);
            const int ittot = (int)(tblint * tblind + 0.5);
            const double2 e = __ldg(et + ittot);
            c.bbdtot = plfrac * (blay + e.y * dplankdn);
            c.bbd = plfrac * (blay + dplankdn * odepth_rec);
            c.atot = 1. - e.x;
            c.bbugas = plfrac * (blay + dplankup * odepth_rec);
            c.bbutot = plfrac * (blay + e.y * dplankup);
        } else {
            double tblind = odepth / (bpade + odepth);
            const int itgas = (int)(tblint * tblind + 0.5);
            // tau_tbl(itgas) (rrtmg_lw_init.f90:106-123), evaluated instead of stored
            const double tfn = (double)itgas / tblint;
            odepth = itgas >= 10000 ? 1.e10 : (itgas <= 0 ? 0.0 : bpade * tfn / (1.0 - tfn));
            const double2 eg = __ldg(et + itgas);
            c.atrans = 1. - eg.x;
            c.gassrc = c.atrans * plfrac * (blay + eg.y * dplankdn);
            odtot = odepth + odcld;
            tblind = odtot / (bpade + odtot);
            const int ittot = (int)(tblint * tblind + 0.5);
            const double2 e = __ldg(et + ittot);
            c.bbdtot = plfrac * (blay + e.y * dplankdn);
            c.bbd = plfrac * (blay + eg.y * dplankdn);
            c.atot = 1. - e.x;
            c.bbugas = plfrac * (blay + eg.y * dplankup);
            c.bbutot = plfrac * (blay + e.y * dplankup);
        }
    } else {
        if (odepth <= 0.06) {
            c.atrans = odepth - 0.5 * odepth * odepth;
            odepth = rec_6 * odepth;
            c.bbd = plfrac * (blay + dplankdn * odepth);
            c.bbugas = plfrac * (blay + dplankup * odepth);
        } else {
            const double tblind = odepth / (bpade + odepth);
            const int itr = (int)(tblint * tblind + 0.5);
            const double2 e = __ldg(et + itr);
            c.atrans = 1. - e.x;
            c.bbd = plfrac * (blay + e.y * dplankdn);
            c.bbugas = plfrac * (blay + e.y * dplankup);
        }
    }
}

enum { CF_CLD1, CF_CLD2, CF_CLR1, CF_CLR2, CF_CMB1, CF_CMB2, CF_CLD1D, CF_CLD2D, CF_CLR1D, CF_CLR2D, CF_CMB1D, CF_CMB2D, CF_COUNT };

// maximum/random overlap factors of one column (rtrnmr.f90:326-479); arrays carry the Fortran indices 0..nlayers+1
__device__ void lw_overlap_factors(int nlayers, const double *cldfrac, const unsigned char *icldlyr, unsigned char *istcld,
                                   unsigned char *istcldd, double (*f)[MAXLAY + 2])
{
    double rat1 = 0., rat2 = 0.;
    double *faccld1 = f[CF_CLD1], *faccld2 = f[CF_CLD2], *facclr1 = f[CF_CLR1], *facclr2 = f[CF_CLR2], *faccmb1 = f[CF_CMB1], *faccmb2 = f[CF_CMB2];
    double *faccld1d = f[CF_CLD1D], *faccld2d = f[CF_CLD2D], *facclr1d = f[CF_CLR1D], *facclr2d = f[CF_CLR2D], *faccmb1d = f[CF_CMB1D], *faccmb2d = f[CF_CMB2D];
    istcld[1] = 1;
    istcldd[nlayers] = 1;
    for (int lev = 1; lev <= nlayers; ++lev) {
        if (icldlyr[lev]) {
            istcld[lev + 1] = 0;
            if (lev == nlayers) {
                faccld1[lev + 1] = 0.; faccld2[lev + 1] = 0.; facclr1[lev + 1] = 0.;
                facclr2[lev + 1] = 0.; faccmb1[lev + 1] = 0.; faccmb2[lev + 1] = 0.;
            } else if (cldfrac[lev + 1] >= cldfrac[lev]) {
                faccld1[lev + 1] = 0.;
                faccld2[lev + 1] = 0.;
                if (istcld[lev] == 1) {
                    facclr1[lev + 1] = 0.;
                    facclr2[lev + 1] = 0.;
                    if (cldfrac[lev] < 1.) facclr2[lev + 1] = (cldfrac[lev + 1] - cldfrac[lev]) / (1. - cldfrac[lev]);
                    facclr2[lev] = 0.;
                    faccld2[lev] = 0.;
                } else {
                    const double fmx = fmax(cldfrac[lev], cldfrac[lev - 1]);
                    if (cldfrac[lev + 1] > fmx) {
                        facclr1[lev + 1] = rat2;
                        facclr2[lev + 1] = (cldfrac[lev + 1] - fmx) / (1. - fmx);
                    } else if (cldfrac[lev + 1] < fmx) {
                        facclr1[lev + 1] = (cldfrac[lev + 1] - cldfrac[lev]) / (cldfrac[lev - 1] - cldfrac[lev]);
                        facclr2[lev + 1] = 0.;
                    } else {
                        facclr1[lev + 1] = rat2;
                        facclr2[lev + 1] = 0.;
                    }
                }
                if (facclr1[lev + 1] > 0. || facclr2[lev + 1] > 0.) { rat1 = 1.; rat2 = 0.; }
                else { rat1 = 0.; rat2 = 0.; }
            } else {
                facclr1[lev + 1] = 0.;
                facclr2[lev + 1] = 0.;
                if (istcld[lev] == 1) {
                    faccld1[lev + 1] = 0.;
                    faccld2[lev + 1] = (cldfrac[lev] - cldfrac[lev + 1]) / cldfrac[lev];
                    facclr2[lev] = 0.;
                    faccld2[lev] = 0.;
                } else {
                    const double fmn = fmin(cldfrac[lev], cldfrac[lev - 1]);
                    if (cldfrac[lev + 1] <= fmn) {
                        faccld1[lev + 1] = rat1;
                        faccld2[lev + 1] = (fmn - cldfrac[lev + 1]) / fmn;
                    } else {
                        faccld1[lev + 1] = (cldfrac[lev] - cldfrac[lev + 1]) / (cldfrac[lev] - fmn);
                        faccld2[lev + 1] = 0.;
                    }
                }
                if (faccld1[lev + 1] > 0. || faccld2[lev + 1] > 0.) { rat1 = 0.; rat2 = 1.; }
                else { rat1 = 0.; rat2 = 0.; }
            }
            faccmb1[lev + 1] = facclr1[lev + 1] * faccld2[lev] * cldfrac[lev - 1];
            faccmb2[lev + 1] = faccld1[lev + 1] * facclr2[lev] * (1. - cldfrac[lev - 1]);
        } else {
            istcld[lev + 1] = 1;
        }
    }
    for (int lev = nlayers; lev >= 1; --lev) {
        if (icldlyr[lev]) {
            istcldd[lev - 1] = 0;
            if (lev == 1) {
                faccld1d[lev - 1] = 0.; faccld2d[lev - 1] = 0.; facclr1d[lev - 1] = 0.;
                facclr2d[lev - 1] = 0.; faccmb1d[lev - 1] = 0.; faccmb2d[lev - 1] = 0.;
            } else if (cldfrac[lev - 1] >= cldfrac[lev]) {
                faccld1d[lev - 1] = 0.;
                faccld2d[lev - 1] = 0.;
                if (istcldd[lev] == 1) {
                    facclr1d[lev - 1] = 0.;
                    facclr2d[lev - 1] = 0.;
                    if (cldfrac[lev] < 1.) facclr2d[lev - 1] = (cldfrac[lev - 1] - cldfrac[lev]) / (1. - cldfrac[lev]);
                    facclr2d[lev] = 0.;
                    faccld2d[lev] = 0.;
                } else {
                    const double fmx = fmax(cldfrac[lev], cldfrac[lev + 1]);
                    if (cldfrac[lev - 1] > fmx) {
                        facclr1d[lev - 1] = rat2;
                        facclr2d[lev - 1] = (cldfrac[lev - 1] - fmx) / (1. - fmx);
                    } else if (cldfrac[lev - 1] < fmx) {
                        facclr1d[lev - 1] = (cldfrac[lev - 1] - cldfrac[lev]) / (cldfrac[lev + 1] - cldfrac[lev]);
                        facclr2d[lev - 1] = 0.;
                    } else {
                        facclr1d[lev - 1] = rat2;
                        facclr2d[lev - 1] = 0.;
                    }
                }
                if (facclr1d[lev - 1] > 0. || facclr2d[lev - 1] > 0.) { rat1 = 1.; rat2 = 0.; }
                else { rat1 = 0.; rat2 = 0.; }
            } else {
                facclr1d[lev - 1] = 0.;
                facclr2d[lev - 1] = 0.;
                if (istcldd[lev] == 1) {
                    faccld1d[lev - 1] = 0.;
                    faccld2d[lev - 1] = (cldfrac[lev] - cldfrac[lev - 1]) / cldfrac[lev];
                    facclr2d[lev] = 0.;
                    faccld2d[lev] = 0.;
                } else {
                    const double fmn = fmin(cldfrac[lev], cldfrac[lev + 1]);
                    if (cldfrac[lev - 1] <= fmn) {
                        faccld1d[lev - 1] = rat1;
                        faccld2d[lev - 1] = (fmn - cldfrac[lev - 1]) / fmn;
                    } else {
                        faccld1d[lev - 1] = (cldfrac[lev] - cldfrac[lev - 1]) / (cldfrac[lev] - fmn);
                        faccld2d[lev - 1] = 0.;
                    }
                }
                if (faccld1d[lev - 1] > 0. || faccld2d[lev - 1] > 0.) { rat1 = 0.; rat2 = 1.; }
                else { rat1 = 0.; rat2 = 0.; }
            }
            faccmb1d[lev - 1] = facclr1d[lev - 1] * faccld2d[lev] * cldfrac[lev + 1];
            faccmb2d[lev - 1] = faccld1d[lev - 1] * facclr2d[lev] * (1. - cldfrac[lev + 1]);
        } else {
            istcldd[lev - 1] = 1;
        }
    }
}

template <bool MR, bool DRV>
__global__ void __launch_bounds__(RT_THREADS) lw_rtrn_cloud_kernel(LwTables T, LwIn in, LwOut out, LwWork w)
{
    __shared__ double s_tile[16 * RT_S];
    __shared__ double s_part[16 * (RT_THREADS / 16 + 1)];
    __shared__ double s_flux[6][MAXLAY + 1];                 // down, down clear, up, up clear, d up/dT, d up clear/dT
    __shared__ double s_cf[MAXLAY + 2];                      // cldfrac(0:nlayers+1), zero at both ends
    __shared__ double s_fac[MR ? CF_COUNT : 1][MAXLAY + 2];
    __shared__ unsigned char s_cld[MAXLAY + 2], s_st[MAXLAY + 2], s_std[MAXLAY + 2];
    const int col = blockIdx.x;
    const int nlay = w.nlay;
    const int g = threadIdx.x;
    const bool active = g < NGPTLW;
    const int band = active ? c_ls.ngb[g] : 0;
    const size_t ld = (size_t)in.ld;
    // cloud fraction, cloudy-layer flag (:316-324) and cldprop's test for a layer with cloud optical depth (cldprop.f90:158-168)
    for (int i = threadIdx.x; i <= nlay + 1; i += RT_THREADS) {
        double cf = 0.0;
        if (i >= 1 && i <= nlay) cf = in.cldfr[col + (size_t)(i - 1) * ld];
        s_cf[i] = cf;
        s_cld[i] = cf >= 1.e-6;
        s_st[i] = 0; s_std[i] = 0;
        if (MR)
            for (int q = 0; q < CF_COUNT; ++q) s_fac[q][i] = 0.0;
    }
    __syncthreads();
    if (MR && threadIdx.x == 0) lw_overlap_factors(nlay, s_cf, s_cld, s_st, s_std, s_fac);
    __syncthreads();

    const double secd = w.secdiff[(size_t)col * 16 + band];
    const double wgt = active ? 0.5 * c_ls.delwave[band] : 0.0;
    const double bpade = c_ls.bpade;
    const double2 *__restrict__ et = reinterpret_cast<const double2 *>(T.exptfn);
    const double *__restrict__ taug = w.taug + (size_t)col * nlay * NGPTLW + (active ? g : 0);
    const double *__restrict__ fracs = w.fracs + (size_t)col * nlay * NGPTLW + (active ? g : 0);
    const double *__restrict__ pl = w.planklay + (size_t)col * nlay * 16 + band;
    const double *__restrict__ pv = w.planklev + (size_t)col * (nlay + 1) * 16 + band;
    const double *taer = in.tauaer ? in.tauaer + col + (size_t)band * nlay * ld : nullptr;
    // cloud band of this g-point's spectral band (ipat, rtrnmr.f90:243-245) for the ncbands cldprop left behind
    const int ncb = w.ncbands[col];
    const int ibc = ncb == 16 ? band : (ncb == 5 ? (band <= 1 ? band : (band <= 4 ? 2 : (band <= 7 ? 3 : 4))) : 0);
    const double *tcld = w.taucloud + (size_t)col * nlay * 16 + ibc;
    const double secdc = w.secdiff[(size_t)col * 16 + ibc];          // odcld(lay, ib) = secdiff(ib) * taucloud(lay, ib), ib the CLOUD band

    auto cell = [&](int lay, LwCloudCell &c) {          // lay 0-based; Fortran lev = lay + 1
        double taut = taug[(size_t)lay * NGPTLW];
        if (taer) taut = taut + taer[(size_t)lay * ld];
        const double blay = __ldg(pl + lay * 16);
        const bool cloudy = s_cld[lay + 1];
        const double odcld = cloudy ? secdc * tcld[16 * (size_t)lay] : 0.0;
        lw_cloud_cell(et, bpade, secd, taut, fracs[(size_t)lay * NGPTLW], blay, __ldg(pv + (lay + 1) * 16) - blay,
                      __ldg(pv + lay * 16) - blay, cloudy, odcld, c);
        return odcld;
    };

    // ---- downward sweep
    double radld = 0.0, radclrd = 0.0, cldradd = 0.0, clrradd = 0.0, rad = 0.0;
    bool iclddn = false;
    for (int k = 0; k < nlay; ++k) {
        const int lay = nlay - 1 - k, lev = lay + 1;
        LwCloudCell c;
        const double odcld = cell(lay, c);
        if (s_cld[lev]) {
            iclddn = true;
            const double cf = s_cf[lev];
            if (MR) {
                if (s_std[lev]) {
                    cldradd = cf * radld;
                    clrradd = radld - cldradd;
                    rad = 0.;
                }
                const double ttot = 1. - c.atot;
                const double cldsrc = c.bbdtot * c.atot;
                cldradd = cldradd * ttot + cf * cldsrc;
                clrradd = clrradd * (1. - c.atrans) + (1. - cf) * c.gassrc;
                radld = cldradd + clrradd;
                const double radmod = rad * (s_fac[MR ? CF_CLR1D : 0][lev - 1] * (1. - c.atrans) + s_fac[MR ? CF_CLD1D : 0][lev - 1] * ttot) -
                                      s_fac[MR ? CF_CMB1D : 0][lev - 1] * c.gassrc + s_fac[MR ? CF_CMB2D : 0][lev - 1] * cldsrc;
                const double oldcld = cldradd - radmod;
                const double oldclr = clrradd + radmod;
                rad = -radmod + s_fac[MR ? CF_CLR2D : 0][lev - 1] * oldclr - s_fac[MR ? CF_CLD2D : 0][lev - 1] * oldcld;
                cldradd = cldradd + rad;
                clrradd = clrradd - rad;
            } else {
                const double efclfrac = (1. - exp(-odcld)) * cf;
                radld = radld - radld * (c.atrans + efclfrac * (1. - c.atrans)) + c.gassrc + cf * (c.bbdtot * c.atot - c.gassrc);
            }
        } else {
            radld = radld + (c.bbd - radld) * c.atrans;
        }
        if (iclddn) radclrd = radclrd + (c.bbd - radclrd) * c.atrans;
        else radclrd = radld;
        if (active) {
            s_tile[((k & 3) * 4 + 0) * RT_S + g] = radld * wgt;
            s_tile[((k & 3) * 4 + 1) * RT_S + g] = radclrd * wgt;
            s_tile[((k & 3) * 4 + 2) * RT_S + g] = 0.0;
            s_tile[((k & 3) * 4 + 3) * RT_S + g] = 0.0;
        }
        if ((k & 3) == 3 || k == nlay - 1) {
            const double sum = tile_reduce16<RT_THREADS, NGPTLW, RT_S>(s_tile, s_part);
            const int kk = (k & ~3) + (threadIdx.x >> 2), q = threadIdx.x & 3;
            if (threadIdx.x < 16 && kk <= k && q < 2) s_flux[q][nlay - 1 - kk] = sum * c_ls.fluxfac;
        }
    }
    if (threadIdx.x < 2) s_flux[threadIdx.x][nlay] = 0.0;

    // ---- surface (:628-647) and upward sweep
    double radlu = 0.0, radclru = 0.0, cldradu = 0.0, clrradu = 0.0, d_radlu_dt = 0.0, d_radclru_dt = 0.0;
    rad = 0.0;
    for (int k = 0; k <= nlay; ++k) {
        if (k == 0) {
            const double semiss = in.emis ? in.emis[col + (size_t)band * ld] : 1.0;
            const double plfrac1 = fracs[0];
            const double rad0 = plfrac1 * w.plankbnd[(size_t)col * 16 + band];
            const double reflect = 1. - semiss;
            radlu = rad0 + reflect * radld;
            radclru = rad0 + reflect * radclrd;
            if (DRV) { d_radlu_dt = plfrac1 * w.dplankbnd[(size_t)col * 16 + band]; d_radclru_dt = d_radlu_dt; }
        } else {
            const int lay = k - 1, lev = k;
            LwCloudCell c;
            const double odcld = cell(lay, c);
            if (s_cld[lev]) {
                const double cf = s_cf[lev];
                const double gassrc = c.bbugas * c.atrans;
                if (MR) {
                    if (s_st[lev]) {
                        cldradu = cf * radlu;
                        clrradu = radlu - cldradu;
                        rad = 0.;
                    }
                    const double ttot = 1. - c.atot;
                    const double cldsrc = c.bbutot * c.atot;
                    cldradu = cldradu * ttot + cf * cldsrc;
                    clrradu = clrradu * (1.0 - c.atrans) + (1. - cf) * gassrc;
                    radlu = cldradu + clrradu;
                    const double radmod = rad * (s_fac[MR ? CF_CLR1 : 0][lev + 1] * (1.0 - c.atrans) + s_fac[MR ? CF_CLD1 : 0][lev + 1] * ttot) -
                                          s_fac[MR ? CF_CMB1 : 0][lev + 1] * gassrc + s_fac[MR ? CF_CMB2 : 0][lev + 1] * cldsrc;
                    const double oldcld = cldradu - radmod;
                    const double oldclr = clrradu + radmod;
                    rad = -radmod + s_fac[MR ? CF_CLR2 : 0][lev + 1] * oldclr - s_fac[MR ? CF_CLD2 : 0][lev + 1] * oldcld;
                    cldradu = cldradu + rad;
                    clrradu = clrradu - rad;
                } else {
                    const double efclfrac = (1. - exp(-odcld)) * cf;
                    radlu = radlu - radlu * (c.atrans + efclfrac * (1. - c.atrans)) + gassrc + cf * (c.bbutot * c.atot - gassrc);
                }
                if (DRV) d_radlu_dt = d_radlu_dt * cf * (1.0 - c.atot) + d_radlu_dt * (1.0 - cf) * (1.0 - c.atrans);
            } else {
                radlu = radlu + (c.bbugas - radlu) * c.atrans;
                if (DRV) d_radlu_dt = d_radlu_dt * (1.0 - c.atrans);
            }
            if (iclddn) {
                radclru = radclru + (c.bbugas - radclru) * c.atrans;
                if (DRV) d_radclru_dt = d_radclru_dt * (1.0 - c.atrans);
            } else {
                radclru = radlu;
                if (DRV) d_radclru_dt = d_radlu_dt;
            }
        }
        if (active) {
            s_tile[((k & 3) * 4 + 0) * RT_S + g] = radlu * wgt;
            s_tile[((k & 3) * 4 + 1) * RT_S + g] = radclru * wgt;
            s_tile[((k & 3) * 4 + 2) * RT_S + g] = d_radlu_dt * wgt;
            s_tile[((k & 3) * 4 + 3) * RT_S + g] = d_radclru_dt * wgt;
        }
        if ((k & 3) == 3 || k == nlay) {
            const double sum = tile_reduce16<RT_THREADS, NGPTLW, RT_S>(s_tile, s_part);
            const int kk = (k & ~3) + (threadIdx.x >> 2), q = threadIdx.x & 3;
            if (threadIdx.x < 16 && kk <= k) s_flux[2 + q][kk] = sum * c_ls.fluxfac;
        }
    }
    __syncthreads();

    // ---- fluxes and heating rates (:751-777), copy-out (rad.nomcica:546-564)
    for (int lev = threadIdx.x; lev <= nlay; lev += RT_THREADS) {
        const size_t o = col + (size_t)lev * out.ld;
        const double u = s_flux[2][lev], d = s_flux[0][lev], uc = s_flux[3][lev], dc = s_flux[1][lev];
        out.uflx[o] = u; out.dflx[o] = d;
        out.uflxc[o] = uc; out.dflxc[o] = dc;
        if (DRV) { out.duflx_dt[o] = s_flux[4][lev]; out.duflxc_dt[o] = s_flux[5][lev]; }
        if (lev < nlay) {
            const double pz0 = in.plev[col + (size_t)lev * ld], pz1 = in.plev[col + (size_t)(lev + 1) * ld];
            out.hr[o] = c_ls.heatfac * ((u - d) - (s_flux[2][lev + 1] - s_flux[0][lev + 1])) / (pz0 - pz1);
            out.hrc[o] = c_ls.heatfac * ((uc - dc) - (s_flux[3][lev + 1] - s_flux[1][lev + 1])) / (pz0 - pz1);
        }
    }
}

int lw_launch_rtrn(const LwTables &t, const LwIn &in, const LwOut &out, LwWork &w, cudaStream_t s)
{
    // variant 2 (default): TMA-fed ring of 2 stages x 4 layers, warp-local g-sums; 3: 3 stages; 1: direct loads, warp-local
    // g-sums; 0: direct loads, block-level g-sums (two barriers per 16 levels)
    if (in.icld >= 1) {       // cloudy sky: rtrn (icld = 1) or rtrnmr (icld = 2, 3), rad.nomcica:527-541
        const bool mr = in.icld != 1;
        if (mr) { if (w.idrv) lw_rtrn_cloud_kernel<true, true><<<w.nc, RT_THREADS, 0, s>>>(t, in, out, w);
                  else lw_rtrn_cloud_kernel<true, false><<<w.nc, RT_THREADS, 0, s>>>(t, in, out, w); }
        else { if (w.idrv) lw_rtrn_cloud_kernel<false, true><<<w.nc, RT_THREADS, 0, s>>>(t, in, out, w);
               else lw_rtrn_cloud_kernel<false, false><<<w.nc, RT_THREADS, 0, s>>>(t, in, out, w); }
        return 1;
    }
#ifdef RRTMG_B200_DEV_VARIANTS
    if (g_tune.lw_rtrn_variant >= 2 || w.idrv)          // the derivative outputs are built in the TMA kernel only
#endif
    {
        const int v = g_tune.lw_rtrn_variant;
        if (w.nlay <= 64) { if (in.tauaer) launch_tma_pick<true, 64>(t, in, out, w, s, v); else launch_tma_pick<false, 64>(t, in, out, w, s, v); }
        else { if (in.tauaer) launch_tma_pick<true, MAXLAY>(t, in, out, w, s, v); else launch_tma_pick<false, MAXLAY>(t, in, out, w, s, v); }
        return 1;
    }
#ifdef RRTMG_B200_DEV_VARIANTS
    const bool wr = g_tune.lw_rtrn_variant != 0;
    const size_t smem = (size_t)(2 * w.nlay + 1) * 16 * sizeof(double) + (size_t)g_tune.lw_rtrn_pad_kb * 1024;
#define RT_LAUNCH(A, W) do { \
        cudaFuncSetAttribute(lw_rtrn_kernel<A, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        lw_rtrn_kernel<A, W><<<w.nc, RT_THREADS, smem, s>>>(t, in, out, w); } while (0)
    if (in.tauaer) { if (wr) RT_LAUNCH(true, true); else RT_LAUNCH(true, false); }
    else { if (wr) RT_LAUNCH(false, true); else RT_LAUNCH(false, false); }
#undef RT_LAUNCH
    return 1;
#endif
}

} // namespace rrtmg
