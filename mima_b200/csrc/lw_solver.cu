// lw_solver.cu -- RRTMG longwave clear-sky radiative transfer on sm_100a.
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
        const double2 e = __ldg(et + itr);
        atrans = 1. - e.x;
        if (DOWN) bbd = plfrac * (blay + e.y * dplankdn);
        else bbugas = plfrac * (blay + e.y * dplankup);
    }
}

// The Planck sources of the column ([lay][16] and [lev][16], 16 KB at 60 layers) are staged in shared memory
// once per block: the 140 g-threads need them 2-3 times per level and they are shared by all g-points of a band.
template <bool AER>
__global__ void __launch_bounds__(RT_THREADS) lw_rtrn_kernel(LwTables T, LwIn in, LwOut out, LwWork w)
{
    __shared__ double s_tile[16 * RT_S];
    __shared__ double s_part[16 * (RT_THREADS / 16 + 1)];
    __shared__ double s_dn[MAXLAY + 1], s_up[MAXLAY + 1];
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
                if (active) s_tile[(k & 15) * RT_S + g] = radld * wgt;
                if (k == nlay - 1) plfrac1 = fr[j];
            }
        }
        const int klast = min(k0 + RT_U, nlay) - 1;
        if ((klast & 15) == 15 || klast == nlay - 1) {
            const double sum = tile_reduce16<RT_THREADS, NGPTLW, RT_S>(s_tile, s_part);
            const int kk = (klast & ~15) + threadIdx.x;
            if (threadIdx.x < 16 && kk <= klast) s_dn[nlay - 1 - kk] = sum * c_ls.fluxfac;
        }
    }
    if (threadIdx.x == 0) s_dn[nlay] = 0.0;   // no downward flux enters at the top (drad(nlayers) = 0)

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
                if (active) s_tile[g] = radlu * wgt;
            } else if (full || k <= nlay) {
                const int lay = k - 1;
                const double blay = pl[lay * 16];
                double atrans, bbd, bbugas;
                lw_layer<false>(et, bpade, secd, tg[j], fr[j], blay, pv[(lay + 1) * 16] - blay, 0.0, atrans, bbd, bbugas);
                radlu = radlu + (bbugas - radlu) * atrans;
                if (active) s_tile[(k & 15) * RT_S + g] = radlu * wgt;
            }
        }
        const int klast = min(k0 + RT_U - 1, nlay);
        if ((klast & 15) == 15 || klast == nlay) {
            const double sum = tile_reduce16<RT_THREADS, NGPTLW, RT_S>(s_tile, s_part);
            const int kk = (klast & ~15) + threadIdx.x;
            if (threadIdx.x < 16 && kk <= klast) s_up[kk] = sum * c_ls.fluxfac;
        }
    }
    __syncthreads();

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


// =====================================================================================================
// Variant 2 ("slice"): block = one warp <-> (column, slice of 28 g-points); 5 slices cover the 140 g-points.
// The down sweep evaluates every cell once and leaves {atrans, bbugas} in shared memory ([lay][28] pairs,
// 448 B per layer: 26.9 KB at 60 layers -> 7 warps resident per SM); the up sweep is then two FP64
// instructions per cell fed by one 16-byte shared-memory load, and taug/fracs are streamed from HBM exactly
// once (evict-first loads).  There is no block barrier anywhere: the g-sum of a slice goes through a small
// tile (8 levels x 28) that four lanes per level add up, and the five slice partials of a column are
// combined in a fixed order by lw_flux_finish_kernel, which also transposes through shared memory so that
// the six (ncol, nlay+1) outputs are written in 256-byte runs.
// =====================================================================================================
constexpr int RS_W = 28;                  // g-points per slice
constexpr int RS_NSL = NGPTLW / RS_W;     // 5
constexpr int RS_TS = 29;                 // tile row stride
static_assert(RS_W * RS_NSL == NGPTLW, "slices must tile the g-points");

// sum of the 28 values of each of the 8 tile rows; lanes 4r..4r+3 return the sum of row r
__device__ __forceinline__ double slice_reduce8(const double *tile, int lane)
{
    __syncwarp();
    const double *src = tile + (lane >> 2) * RS_TS + (lane & 3);
    double acc = src[0];
#pragma unroll
    for (int j = 1; j < 7; ++j) acc += src[4 * j];
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    __syncwarp();
    return acc;
}

template <bool AER>
__global__ void __launch_bounds__(32) lw_rtrn_slice_kernel(LwTables T, LwIn in, LwWork w, double *__restrict__ part)
{
    extern __shared__ __align__(16) double2 s_res[];          // [nlay][RS_W] {atrans, bbugas}
    __shared__ double s_tile[8 * RS_TS];
    const int lane = threadIdx.x;
    const int col = blockIdx.x / RS_NSL, sl = blockIdx.x - col * RS_NSL;
    const int nlay = w.nlay;
    const bool active = lane < RS_W;
    const int gl = active ? lane : RS_W - 1;                  // idle lanes shadow the last g-point (no stores)
    const int g = sl * RS_W + gl;
    const int band = c_ls.ngb[g];
    const double secd = w.secdiff[(size_t)col * 16 + band];
    const double wgt = 0.5 * c_ls.delwave[band];              // wtdiff * delwave
    const double bpade = c_ls.bpade;
    const double2 *__restrict__ et = reinterpret_cast<const double2 *>(T.exptfn);
    const double *__restrict__ taug = w.taug + (size_t)col * nlay * NGPTLW + g;
    const double *__restrict__ fracs = w.fracs + (size_t)col * nlay * NGPTLW + g;
    const double *__restrict__ pl = w.planklay + (size_t)col * nlay * 16 + band;
    const double *__restrict__ pv = w.planklev + (size_t)col * (nlay + 1) * 16 + band;
    const double *taer = AER ? in.tauaer + col + (size_t)band * nlay * in.ld : nullptr;
    double *pdn = part + ((size_t)(col * RS_NSL + sl) * 2) * (nlay + 1);   // downward partials [lev]
    double *pup = pdn + (nlay + 1);                                         // upward partials [lev]
    const double rec_6 = 0.166667;

    // ---- downward sweep (:505-618), k counts layers from the top; loads of the next group of four
    //      layers are issued before the arithmetic of the current one
    double radld = 0.0, plfrac1 = 0.0;
    double pupper = __ldg(pv + (size_t)nlay * 16);            // Planck at the upper interface of the layer
    double tgn[RT_U], frn[RT_U], bln[RT_U], pvn[RT_U];
#pragma unroll
    for (int j = 0; j < RT_U; ++j) {
        const int lay = max(nlay - 1 - j, 0);
        tgn[j] = __ldcs(taug + (size_t)lay * NGPTLW);
        frn[j] = __ldcs(fracs + (size_t)lay * NGPTLW);
        if (AER) tgn[j] = tgn[j] + taer[(size_t)lay * in.ld];
        bln[j] = __ldg(pl + (size_t)lay * 16);
        pvn[j] = __ldg(pv + (size_t)lay * 16);
    }
    for (int k0 = 0; k0 < nlay; k0 += RT_U) {
        double tg[RT_U], fr[RT_U], bl[RT_U], pw[RT_U];
#pragma unroll
        for (int j = 0; j < RT_U; ++j) { tg[j] = tgn[j]; fr[j] = frn[j]; bl[j] = bln[j]; pw[j] = pvn[j]; }
        if (k0 + RT_U < nlay) {
#pragma unroll
            for (int j = 0; j < RT_U; ++j) {
                const int lay = max(nlay - 1 - (k0 + RT_U + j), 0);      // clamped in the ragged tail
                tgn[j] = __ldcs(taug + (size_t)lay * NGPTLW);
                frn[j] = __ldcs(fracs + (size_t)lay * NGPTLW);
                if (AER) tgn[j] = tgn[j] + taer[(size_t)lay * in.ld];
                bln[j] = __ldg(pl + (size_t)lay * 16);
                pvn[j] = __ldg(pv + (size_t)lay * 16);
            }
        }
        // non-recurrent part of the four cells first (independent -> the table gathers overlap) ...
        double at[RT_U], bd[RT_U], bu[RT_U];
#pragma unroll
        for (int j = 0; j < RT_U; ++j) {
            const double blay = bl[j];
            const double dup = (j == 0 ? pupper : pw[j - 1]) - blay, ddn = pw[j] - blay;
            double odepth = secd * tg[j];
            if (odepth < 0.0) odepth = 0.0;
            double tfac;
            if (odepth <= 0.06) {
                at[j] = odepth - 0.5 * odepth * odepth;
                tfac = rec_6 * odepth;
            } else {
                const double tblind = odepth * rcp_fast(bpade + odepth);
                const int itr = (int)(10000.0 * tblind + 0.5);
                const double2 e = __ldg(et + itr);
                at[j] = 1. - e.x;
                tfac = e.y;
            }
            bd[j] = fr[j] * (blay + tfac * ddn);
            bu[j] = fr[j] * (blay + tfac * dup);
        }
        pupper = pw[RT_U - 1];
        // ... then the recurrence
#pragma unroll
        for (int j = 0; j < RT_U; ++j) {
            const int k = k0 + j;
            if (k < nlay) {
                const int lay = nlay - 1 - k;
                radld = radld + (bd[j] - radld) * at[j];
                if (active) {
                    s_res[lay * RS_W + lane] = make_double2(at[j], bu[j]);
                    s_tile[(k & 7) * RS_TS + lane] = radld * wgt;
                }
                if (k == nlay - 1) plfrac1 = fr[j];
            }
        }
        const int klast = min(k0 + RT_U, nlay) - 1;
        if ((klast & 7) == 7 || klast == nlay - 1) {
            const double sum = slice_reduce8(s_tile, lane);
            const int kk = (klast & ~7) + (lane >> 2);
            if ((lane & 3) == 0 && kk <= klast) pdn[nlay - 1 - kk] = sum;
        }
    }
    if (lane == 0) pdn[nlay] = 0.0;            // no downward flux enters at the top

    // ---- surface (:628-636) and upward sweep (:649-711); level k = 0 is the surface
    double radlu;
    {
        const double semiss = in.emis ? in.emis[col + (size_t)band * in.ld] : 1.0;
        const double rad0 = plfrac1 * w.plankbnd[(size_t)col * 16 + band];
        radlu = rad0 + (1. - semiss) * radld;
        if (active) s_tile[lane] = radlu * wgt;
    }
    for (int k = 1; k <= nlay; ++k) {
        const double2 r = s_res[(k - 1) * RS_W + gl];
        radlu = radlu + (r.y - radlu) * r.x;
        if (active) s_tile[(k & 7) * RS_TS + lane] = radlu * wgt;
        if ((k & 7) == 7 || k == nlay) {
            const double sum = slice_reduce8(s_tile, lane);
            const int kk = (k & ~7) + (lane >> 2);
            if ((lane & 3) == 0 && kk <= k) pup[kk] = sum;
        }
    }
}

// Sum of the slice partials, fluxes, heating rates (:751-777) and copy-out (rad.nomcica:546-555).
// Block <-> 32 adjacent columns; phase 1 (lanes = levels, coalesced partial reads) -> shared memory ->
// phase 2 (lanes = columns, coalesced interface writes).
constexpr int FF_COLS = 32;
__global__ void __launch_bounds__(256) lw_flux_finish_kernel(LwIn in, LwOut out, LwWork w, const double *__restrict__ part)
{
    extern __shared__ double s_flux[];                         // [2][nlay+1][FF_COLS+1]
    const int nlev = w.nlay + 1;
    const int c0 = blockIdx.x * FF_COLS;
    const int ncb = min(FF_COLS, w.nc - c0);
    double *s_up = s_flux, *s_dn = s_flux + (size_t)nlev * (FF_COLS + 1);
    for (int i = threadIdx.x; i < ncb * nlev; i += blockDim.x) {
        const int c = i / nlev, lev = i - c * nlev;
        const double *p = part + (size_t)(c0 + c) * RS_NSL * 2 * nlev + lev;
        double dn = 0.0, up = 0.0;
#pragma unroll
        for (int s = 0; s < RS_NSL; ++s) {
            dn += p[(size_t)(2 * s) * nlev];
            up += p[(size_t)(2 * s + 1) * nlev];
        }
        s_up[lev * (FF_COLS + 1) + c] = up * c_ls.fluxfac;
        s_dn[lev * (FF_COLS + 1) + c] = dn * c_ls.fluxfac;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nlev * FF_COLS; i += blockDim.x) {
        const int lev = i / FF_COLS, c = i - lev * FF_COLS;
        if (c >= ncb) continue;
        const int col = c0 + c;
        const size_t o = col + (size_t)lev * out.ld;
        const double u = s_up[lev * (FF_COLS + 1) + c], d = s_dn[lev * (FF_COLS + 1) + c];
        out.uflx[o] = u; out.dflx[o] = d;
        out.uflxc[o] = u; out.dflxc[o] = d;
        if (lev < w.nlay) {
            const double fnet0 = u - d;
            const double fnet1 = s_up[(lev + 1) * (FF_COLS + 1) + c] - s_dn[(lev + 1) * (FF_COLS + 1) + c];
            const double pz0 = in.plev[col + (size_t)lev * in.ld], pz1 = in.plev[col + (size_t)(lev + 1) * in.ld];
            const double h = c_ls.heatfac * (fnet0 - fnet1) / (pz0 - pz1);
            out.hr[o] = h;
            out.hrc[o] = h;
        }
    }
}

static void launch_rtrn_slice(const LwTables &t, const LwIn &in, const LwOut &out, LwWork &w, cudaStream_t s)
{
    const size_t smem = (size_t)w.nlay * RS_W * sizeof(double2);
    const unsigned nblk = (unsigned)w.nc * RS_NSL;
    if (in.tauaer) {
        cudaFuncSetAttribute(lw_rtrn_slice_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        lw_rtrn_slice_kernel<true><<<nblk, 32, smem, s>>>(t, in, w, w.part);
    } else {
        cudaFuncSetAttribute(lw_rtrn_slice_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        lw_rtrn_slice_kernel<false><<<nblk, 32, smem, s>>>(t, in, w, w.part);
    }
    const size_t fsmem = (size_t)2 * (w.nlay + 1) * (FF_COLS + 1) * sizeof(double);
    cudaFuncSetAttribute(lw_flux_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
    lw_flux_finish_kernel<<<(w.nc + FF_COLS - 1) / FF_COLS, 256, fsmem, s>>>(in, out, w, w.part);
}

int lw_launch_rtrn(const LwTables &t, const LwIn &in, const LwOut &out, LwWork &w, cudaStream_t s)
{
    if (g_tune.lw_rtrn_variant == 1) { launch_rtrn_slice(t, in, out, w, s); return 2; }
    const size_t smem = (size_t)(2 * w.nlay + 1) * 16 * sizeof(double) + (size_t)g_tune.lw_rtrn_pad_kb * 1024;
    if (in.tauaer) {
        cudaFuncSetAttribute(lw_rtrn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        lw_rtrn_kernel<true><<<w.nc, RT_THREADS, smem, s>>>(t, in, out, w);
    } else {
        cudaFuncSetAttribute(lw_rtrn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        lw_rtrn_kernel<false><<<w.nc, RT_THREADS, smem, s>>>(t, in, out, w);
    }
    return 1;
}

} // namespace rrtmg
