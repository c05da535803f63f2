// lw_kernels.cu -- RRTMG longwave on sm_100a: prep (inatm + setcoef), the staged taumol kernel, cldprop.  The fused clear-sky
// kernel is in lw_column.cu, the staged solvers in lw_solver.cu; the per-cell arithmetic they share in lw_bands.cuh.
//
// What is computed follows the reference routines (cited per kernel); how it is computed is a GPU-first design:
//   lw_prep_cell_kernel  thread <-> (column, layer): unit conversion, column amounts, interpolation indices and factors
//                        (lw_cell), the per-cell terms of the column sums; for the column kernel the state goes to the
//                        tile-major field, for the staged kernels the Planck sources of the layer and its upper interface;
//   lw_prep_kernel       thread <-> column: the column sums in the reference's order (laytrop, precipitable water ->
//                        diffusivity secant), surface Planck terms;
//   lw_taumol_kernel     (staged path: clouds, idrv = 1, stage capture) thread <-> (column, layer) cell, lanes = 32 adjacent
//                        columns of one layer; the cell's setcoef state is evaluated in place (lw_cell), every term
//                        w * T[row][.] of a band formula is consumed at once into ng register accumulators, and the warps
//                        of a block walk the 16 bands together (one block barrier per four bands) so that the ~300 KB of
//                        straight-line band code is fetched once per block;
//   lw_cldprop_kernel    thread <-> column (cloudy sky only).
// Compiled with -fmad=false: fused multiply-adds appear only where written as fma().
#include "lw_bands.cuh"

namespace rrtmg {

__constant__ unsigned char c_lw_ngb[NGPTLW];   // band (0-based) of each g-point

__device__ LwCldConst d_lwcld;          // 19 KB: global memory, read through the read-only path
int lw_upload_cld(const LwCldConst &c) { return cudaMemcpyToSymbol(d_lwcld, &c, sizeof c) == cudaSuccess ? 0 : -1; }

int lw_upload_const(const LwConst &c)
{
    unsigned char ngb[NGPTLW];
    for (int b = 0; b < NBNDLW; ++b)
        for (int i = 0; i < c.band[b].ng; ++i) ngb[c.band[b].g0 + i] = (unsigned char)b;
    if (cudaMemcpyToSymbol(c_lw, &c, sizeof(LwConst)) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(c_lw_ngb, ngb, sizeof ngb) != cudaSuccess) return -1;
    if (lw_column_upload_const(c)) return -1;
    return lw_solver_upload_const(c, ngb);
}

// =====================================================================================================
// prep: column-integrated quantities -- laytrop (setcoef.f90:293-294), precipitable water and the
//       diffusivity secant per band (rtrnmr.f90:259-280) -- and the Planck sources (setcoef.f90:154-249).
//       With w.f != nullptr (stage capture, test hook) the per-cell setcoef state is also written out.
// =====================================================================================================
// Stage 1, thread <-> (column, layer) cell (lanes = adjacent columns): everything that is local to the cell -- the
// Planck sources of the layer and of its upper interface, and the three numbers the column sums need (dry column,
// H2O column, "below 100 hPa" flag).  60x more threads than a thread-per-column sweep: at T42 (8192 columns) the
// old kernel occupied 64 of the 148 SMs with one latency-bound warp each.
__global__ void __launch_bounds__(128) lw_prep_cell_kernel(LwTables T, LwIn in, LwWork w)
{
    const int nc = w.nc, nlay = w.nlay;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nc * nlay) return;
    const int l = (int)(i / nc);
    const int col = (int)(i - (size_t)l * nc);
    const size_t ld = (size_t)in.ld;
    const size_t o = col + (size_t)l * ld;
    LwPair p;
    double wkl1;
    const bool lower = lw_cell(in, col, l, p, wkl1);
    w.cs_coldry[i] = p.coldry;
    w.cs_wkl1[i] = wkl1;
    w.cs_lower[i] = lower ? 1 : 0;
    if (w.fused) {             // tile-major state for the fused column kernel, which also interpolates the Planck sources itself
        double *f = w.f + w.tfld(l, col);
#define TF(k) f[(k) * 32]
        TF(LF_COUNT) = __hiloint2double(0, (int)lw_pack(p.jp, p.jt, p.jt1, p.inds, p.indf, p.indm));
        TF(LF_FAC00) = p.fac00; TF(LF_FAC01) = p.fac01; TF(LF_FAC10) = p.fac10; TF(LF_FAC11) = p.fac11;
        TF(LF_COLH2O) = p.colh2o; TF(LF_COLCO2) = p.colco2; TF(LF_COLO3) = p.colo3;
        TF(LF_COLN2O) = p.coln2o; TF(LF_COLCO) = p.colco; TF(LF_COLCH4) = p.colch4;
        TF(LF_COLO2) = p.colo2; TF(LF_COLBRD) = p.colbrd;
        TF(LF_SELFFAC) = p.selffac; TF(LF_SELFFRAC) = p.selffrac;
        TF(LF_FORFAC) = p.forfac; TF(LF_FORFRAC) = p.forfrac;
        TF(LF_MINORFRAC) = p.minorfrac; TF(LF_SCALEMINOR) = p.scaleminor;
        TF(LF_SCALEMINORN2) = p.scaleminorn2; TF(LF_COLDRY) = p.coldry;
        TF(LF_PAVEL) = p.pavel;
        if (in.ccl4) TF(LF_WX1) = p.wx1;
        if (in.cfc11) TF(LF_WX2) = p.wx2;
        if (in.cfc12) TF(LF_WX3) = p.wx3;
        if (in.cfc22) TF(LF_WX4) = p.wx4;
#undef TF
        return;
    }
    {
        const size_t wo = i;
        w.idx[wo] = lw_pack(p.jp, p.jt, p.jt1, p.inds, p.indf, p.indm);
        w.fld(LF_FAC00)[wo] = p.fac00; w.fld(LF_FAC01)[wo] = p.fac01;
        w.fld(LF_FAC10)[wo] = p.fac10; w.fld(LF_FAC11)[wo] = p.fac11;
        w.fld(LF_COLH2O)[wo] = p.colh2o; w.fld(LF_COLCO2)[wo] = p.colco2; w.fld(LF_COLO3)[wo] = p.colo3;
        w.fld(LF_COLN2O)[wo] = p.coln2o; w.fld(LF_COLCO)[wo] = p.colco; w.fld(LF_COLCH4)[wo] = p.colch4;
        w.fld(LF_COLO2)[wo] = p.colo2; w.fld(LF_COLBRD)[wo] = p.colbrd;
        w.fld(LF_SELFFAC)[wo] = p.selffac; w.fld(LF_SELFFRAC)[wo] = p.selffrac;
        w.fld(LF_FORFAC)[wo] = p.forfac; w.fld(LF_FORFRAC)[wo] = p.forfrac;
        w.fld(LF_MINORFRAC)[wo] = p.minorfrac; w.fld(LF_SCALEMINOR)[wo] = p.scaleminor;
        w.fld(LF_SCALEMINORN2)[wo] = p.scaleminorn2; w.fld(LF_COLDRY)[wo] = p.coldry;
        w.fld(LF_PAVEL)[wo] = p.pavel;
        w.fld(LF_WX1)[wo] = p.wx1; w.fld(LF_WX2)[wo] = p.wx2; w.fld(LF_WX3)[wo] = p.wx3; w.fld(LF_WX4)[wo] = p.wx4;
    }
    // ---- setcoef: Planck sources (setcoef.f90:154-249)
    const double tavel = in.tlay[o], tz = in.tlev[o + ld];
    int indlay = (int)(tavel - 159.);
    indlay = indlay < 1 ? 1 : (indlay > 180 ? 180 : indlay);
    const double tlayfrac = tavel - 159. - (double)indlay;
    int indlev = (int)(tz - 159.);
    indlev = indlev < 1 ? 1 : (indlev > 180 ? 180 : indlev);
    const double tlevfrac = tz - 159. - (double)indlev;
    double2 *play2 = reinterpret_cast<double2 *>(w.planklay + ((size_t)col * nlay + l) * 16);
    double2 *plev2 = reinterpret_cast<double2 *>(w.planklev + ((size_t)col * (nlay + 1) + l + 1) * 16);
#pragma unroll 2
    for (int ib = 0; ib < 16; ib += 2) {
        double2 a, b;
        const double *tp = T.totplnk + ib * 181;
        double d = __ldg(tp + indlay) - __ldg(tp + indlay - 1);
        a.x = __ldg(tp + indlay - 1) + tlayfrac * d;
        d = __ldg(tp + indlev) - __ldg(tp + indlev - 1);
        b.x = __ldg(tp + indlev - 1) + tlevfrac * d;
        tp += 181;
        d = __ldg(tp + indlay) - __ldg(tp + indlay - 1);
        a.y = __ldg(tp + indlay - 1) + tlayfrac * d;
        d = __ldg(tp + indlev) - __ldg(tp + indlev - 1);
        b.y = __ldg(tp + indlev - 1) + tlevfrac * d;
        play2[ib >> 1] = a;
        plev2[ib >> 1] = b;
    }
}

// Stage 2, thread <-> column: the column-integrated quantities in the reference's summation order -- laytrop
// (setcoef.f90:293-294), precipitable water and the diffusivity secant per band (rtrnmr.f90:259-280) -- and the
// surface / level-0 Planck terms.
__global__ void __launch_bounds__(128) lw_prep_kernel(LwTables T, LwIn in, LwWork w)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= w.nc) return;
    const int nlay = w.nlay, nc = w.nc;
    const size_t ld = (size_t)in.ld;
    const double amd = 28.9660, amw = 18.0160, grav = 9.8066;

    double amttl = 0.0, wvttl = 0.0;
    const double tbound = in.tsfc[col];
    const double pz0 = in.plev[col];
    const double tz0 = in.tlev[col];
    int laytrop = 0;

    // surface / level-0 Planck terms
    int indbound = (int)(tbound - 159.);
    indbound = indbound < 1 ? 1 : (indbound > 180 ? 180 : indbound);
    const double tbndfrac = tbound - 159. - (double)indbound;
    int indlev0 = (int)(tz0 - 159.);
    indlev0 = indlev0 < 1 ? 1 : (indlev0 > 180 ? 180 : indlev0);
    const double t0frac = tz0 - 159. - (double)indlev0;
    {
        double *pb = w.plankbnd + (size_t)col * 16;
        double *pl0 = w.planklev + (size_t)col * (nlay + 1) * 16;
#pragma unroll 4
        for (int ib = 0; ib < 16; ++ib) {
            const double *tp = T.totplnk + ib * 181;
            const double semiss = in.emis ? in.emis[col + ib * ld] : 1.0;
            double dbdtlev = __ldg(tp + indbound) - __ldg(tp + indbound - 1);
            pb[ib] = semiss * (__ldg(tp + indbound - 1) + tbndfrac * dbdtlev);
            dbdtlev = __ldg(tp + indlev0) - __ldg(tp + indlev0 - 1);
            if (!w.fused) pl0[ib] = __ldg(tp + indlev0 - 1) + t0frac * dbdtlev;
            if (w.idrv) {       // setcoef.f90:197-201
                const double *td = T.totplnkderiv + ib * 181;
                dbdtlev = __ldg(td + indbound) - __ldg(td + indbound - 1);
                w.dplankbnd[(size_t)col * 16 + ib] = semiss * (__ldg(td + indbound - 1) + tbndfrac * dbdtlev);
            }
        }
    }
    for (int l = 0; l < nlay; ++l) {
        const size_t i = (size_t)l * nc + col;
        const double wkl1 = w.cs_wkl1[i];
        if (w.cs_lower[i]) laytrop = laytrop + 1;
        amttl = amttl + w.cs_coldry[i] + wkl1;
        wvttl = wvttl + wkl1;
    }
    w.laytrop[col] = laytrop;
    // precipitable water and diffusivity secant per band
    const double wvsh = (amw * wvttl) / (amd * amttl);
    const double pwvcm = wvsh * (1.e3 * pz0) / (1.e2 * grav);
    double *sd = w.secdiff + (size_t)col * 16;
    for (int ib = 0; ib < 16; ++ib) {
        double s;
        if (ib == 0 || ib == 3 || ib >= 9) {
            s = 1.66;
        } else {
            s = c_lw.a0[ib] + c_lw.a1[ib] * exp(c_lw.a2[ib] * pwvcm);
            if (s > 1.80) s = 1.80;
            if (s < 1.50) s = 1.50;
        }
        sd[ib] = s;
    }
}

// =====================================================================================================
// taumol: LW/src/rrtmg_lw_taumol.f90:260-3147 (taugb1..16)
//
// Every gas optical depth of RRTMG is a weighted sum of k-table rows, tau(g) = sum_k w_k * T[row_k][g],
// with (w_k, row_k) depending only on the (column, layer) cell and the band.  Thread <-> cell (lanes =
// 32 adjacent columns of one layer): the thread walks the band formula of taugbN, and every term is
// consumed at once into NG register accumulators (NG = g-points of the band, compile-time), the table
// row being read as NG/2 16-byte loads through the read-only path (rows of neighbouring columns mostly
// coincide -> L1 broadcast).  No plan is ever stored.  The finished NG values per cell are transposed
// through a per-warp shared-memory slab so that the staging fields are written [col][lay][g] (g
// fastest, what the solver's g-lanes read) in 16-byte pieces.
// =====================================================================================================
constexpr int TM_STRIDE = 18;      // slab row stride in doubles (36 words: conflict-free 16-byte accesses)

template <int NG>
struct BandAcc {
    double t[NG];                    // taug accumulators of this cell
    const double *__restrict__ tab;  // band table, [row][NG]
    double *sf;                      // this lane's fracs row in the slab
    __device__ __forceinline__ void clear()
    {
#pragma unroll
        for (int g = 0; g < NG; ++g) t[g] = 0.0;
    }
    __device__ __forceinline__ void add(int off, double wgt)
    {
        const double2 *__restrict__ q = reinterpret_cast<const double2 *>(tab + off);
#pragma unroll
        for (int j = 0; j < NG / 2; ++j) {
            const double2 v = __ldg(q + j);
            t[2 * j] = fma(wgt, v.x, t[2 * j]);
            t[2 * j + 1] = fma(wgt, v.y, t[2 * j + 1]);
        }
    }
    // terms whose weight is exactly zero for every column of MiMA's default configuration (the CFC cross-sections when
    // no CFC array is passed, the O2 continuum when O2 is absent): t + 0 * T = t bit for bit, so the row is not read
    __device__ __forceinline__ void add_nz(int off, double wgt)
    {
        if (wgt != 0.0) add(off, wgt);
    }
    __device__ __forceinline__ void scale(int off)      // taug(g) *= T[off + g]
    {
        const double2 *__restrict__ q = reinterpret_cast<const double2 *>(tab + off);
#pragma unroll
        for (int j = 0; j < NG / 2; ++j) {
            const double2 v = __ldg(q + j);
            t[2 * j] = t[2 * j] * v.x;
            t[2 * j + 1] = t[2 * j + 1] * v.y;
        }
    }
    __device__ __forceinline__ void frac1(int off)      // fracs(g) = T[off + g]
    {
        const double2 *__restrict__ q = reinterpret_cast<const double2 *>(tab + off);
#pragma unroll
        for (int j = 0; j < NG / 2; ++j) reinterpret_cast<double2 *>(sf)[j] = __ldg(q + j);
    }
    __device__ __forceinline__ void frac2(int o0, double w0, int o1, double w1)   // w0*T[o0+g] + w1*T[o1+g]
    {
        const double2 *__restrict__ q0 = reinterpret_cast<const double2 *>(tab + o0);
        const double2 *__restrict__ q1 = reinterpret_cast<const double2 *>(tab + o1);
#pragma unroll
        for (int j = 0; j < NG / 2; ++j) {
            const double2 a = __ldg(q0 + j), b = __ldg(q1 + j);
            double2 r;
            r.x = fma(w1, b.x, w0 * a.x);
            r.y = fma(w1, b.y, w0 * a.y);
            reinterpret_cast<double2 *>(sf)[j] = r;
        }
    }
    __device__ __forceinline__ void fzero()
    {
#pragma unroll
        for (int j = 0; j < NG / 2; ++j) reinterpret_cast<double2 *>(sf)[j] = make_double2(0.0, 0.0);
    }
};

// One band of one warp's 32 cells: accumulate in registers, transpose through the warp's slab, write the
// 32 x NG block of taug and fracs with 16-byte stores (g fastest).
template <int BAND>
__device__ __forceinline__ void lw_band(const LwTables &T, const LwPair &p, bool valid, bool lower, double *slab,
                                        double *__restrict__ taug, double *__restrict__ fracs,
                                        size_t cell0, size_t colstride, int nvalid)
{
    constexpr int NG = lw_ng(BAND);
    const int lane = threadIdx.x & 31;
    const LwBand &B = c_lw.band[BAND];
    double *st = slab + lane * TM_STRIDE;                       // taug row of this lane
    double *sf = slab + (32 + lane) * TM_STRIDE;                // fracs row
    if (valid) {
        BandAcc<NG> pw;
        pw.tab = T.tab + B.base;
        pw.sf = sf;
        pw.clear();
        lw_band_terms<BAND>(p, lower, pw);
#pragma unroll
        for (int j = 0; j < NG / 2; ++j) reinterpret_cast<double2 *>(st)[j] = make_double2(pw.t[2 * j], pw.t[2 * j + 1]);
    }
    __syncwarp();
    constexpr int HP = NG / 2;                                  // 16-byte pieces per cell
    const int g0 = B.g0;
#pragma unroll
    for (int i = lane; i < 32 * HP; i += 32) {
        const int c = i / HP, j = i - c * HP;
        if (c < nvalid) {
            const double2 a = reinterpret_cast<const double2 *>(slab + c * TM_STRIDE)[j];
            const double2 b = reinterpret_cast<const double2 *>(slab + (32 + c) * TM_STRIDE)[j];
            const size_t o = cell0 + (size_t)c * colstride + g0 + 2 * j;
            *reinterpret_cast<double2 *>(taug + o) = a;
            *reinterpret_cast<double2 *>(fracs + o) = b;
        }
    }
    __syncwarp();
}

// The 16 warps of a block walk through the bands together (one barrier per band): the straight-line band
// code (~300 KB for all bands) is then fetched once per block instead of once per warp -- with
// independent warps the kernel was bound by instruction-cache misses.  Work items are (32-column tile,
// layer) pairs, linearised so that no warp idles when nlay is not a multiple of the block's warp count.
constexpr int TM_BLOCK_WARPS = 8;
__global__ void __launch_bounds__(32 * TM_BLOCK_WARPS, 2) lw_taumol_kernel(LwTables T, LwIn in, LwWork w, int g_tm_sync)
{
    extern __shared__ __align__(16) double s_dyn[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nlay = w.nlay, nc = w.nc;
    const int ntile = (nc + 31) / 32;
    const long item = (long)blockIdx.x * TM_BLOCK_WARPS + wid;
    const bool live = item < (long)ntile * nlay;
    const int tile = live ? (int)(item / nlay) : 0;
    const int lay = live ? (int)(item - (long)tile * nlay) : 0;     // 0-based layer
    const int c0 = tile * 32;
    const int col = c0 + lane;
    const bool valid = live && col < nc;
    const int nvalid = live ? min(32, nc - c0) : 0;

    LwPair p;
    bool lower = false;
    if (valid) {
        double wkl1;
        lw_cell(in, col, lay, p, wkl1);
        lower = (lay + 1) <= w.laytrop[col];
    }
    double *slab = s_dyn + (size_t)wid * (64 * TM_STRIDE);
    const size_t colstride = (size_t)nlay * NGPTLW;
    const size_t cell0 = ((size_t)c0 * nlay + lay) * NGPTLW;
#define LW_BAND(b) lw_band<b>(T, p, valid, lower, slab, w.taug, w.fracs, cell0, colstride, nvalid); if (((b) & (g_tm_sync - 1)) == g_tm_sync - 1) __syncthreads()
    LW_BAND(0); LW_BAND(1); LW_BAND(2); LW_BAND(3); LW_BAND(4); LW_BAND(5); LW_BAND(6); LW_BAND(7);
    LW_BAND(8); LW_BAND(9); LW_BAND(10); LW_BAND(11); LW_BAND(12); LW_BAND(13); LW_BAND(14); LW_BAND(15);
#undef LW_BAND
}

// =====================================================================================================
// cldprop (rrtmg_lw_cldprop.f90:31-276), thread <-> column: the routine carries state from layer to layer (ncbands, the
// abscoice / abscoliq vectors), so the layers of a column are walked in order.  inflag = 0: optical depth as given;
// 1: abscld1 * water path; 2: ice (iceflag 0-3) and liquid (liqflag 0-1) parameterisations in the effective radii.
// Out: taucloud [col][lay][16] (zero where the layer is not cloudy), ncbands per column (selects ipat in rtrn/rtrnmr),
// and the number of the Fortran `stop` a column ran into (w.err, atomicMax): 1 ICE RADIUS TOO SMALL, 2 ICE RADIUS OUT
// OF BOUNDS, 3 ICE GENERALIZED EFFECTIVE SIZE OUT OF BOUNDS, 4 LIQUID EFFECTIVE RADIUS OUT OF BOUNDS.
// =====================================================================================================
__global__ void __launch_bounds__(64) lw_cldprop_kernel(LwIn in, LwWork w)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= w.nc) return;
    const int nlay = w.nlay;
    const size_t ld = (size_t)in.ld;
    const double cldmin = 1.e-20;
    const LwCldConst &K = d_lwcld;
    double abscoice[17], abscoliq[17];
    for (int ib = 0; ib <= 16; ++ib) { abscoice[ib] = 0.; abscoliq[ib] = 0.; }
    int ncbands = 1, iceind = 0, liqind = 0, stop = 0;
    auto icb = [](int ib, int ind) {           // :147-149, ib 1-based
        if (ind == 0) return 1;
        if (ind == 2) return ib;
        return ib <= 2 ? ib : (ib <= 5 ? 3 : (ib <= 8 ? 4 : 5));
    };
    for (int lay = 0; lay < nlay && !stop; ++lay) {
        const size_t o = col + (size_t)lay * ld;
        double *tc = w.taucloud + ((size_t)col * nlay + lay) * 16;
        double tauctot = 0.;
        for (int ib = 0; ib < 16; ++ib) {
            tc[ib] = 0.0;
            tauctot = tauctot + in.taucld[ib + 16 * o];
        }
        const double ciwp = in.cicewp ? in.cicewp[o] : 0., clwp = in.cliqwp ? in.cliqwp[o] : 0.;
        const double cwp = ciwp + clwp;
        if (!(in.cldfr[o] >= cldmin && (cwp >= cldmin || tauctot >= cldmin))) continue;
        if (in.inflg == 0) {
            ncbands = 16;
            for (int ib = 0; ib < 16; ++ib) tc[ib] = in.taucld[ib + 16 * o];
        } else if (in.inflg == 1) {
            ncbands = 16;
            for (int ib = 0; ib < 16; ++ib) tc[ib] = K.abscld1 * cwp;
        } else {
            const double radice = in.reice ? in.reice[o] : 0.;
            if (ciwp == 0.0) {
                abscoice[1] = 0.0;
                iceind = 0;
            } else if (in.iceflg == 0) {
                if (radice < 10.0) { stop = 1; break; }
                abscoice[1] = K.absice0[0] + K.absice0[1] / radice;
                iceind = 0;
            } else if (in.iceflg == 1) {
                if (radice < 13.0 || radice > 130.) { stop = 2; break; }
                ncbands = 5;
                for (int ib = 1; ib <= 5; ++ib) abscoice[ib] = K.absice1[0 + 2 * (ib - 1)] + K.absice1[1 + 2 * (ib - 1)] / radice;
                iceind = 1;
            } else if (in.iceflg == 2) {
                if (radice < 5.0 || radice > 131.0) { stop = 2; break; }
                ncbands = 16;
                const double factor = (radice - 2.) / 3.;
                int index = (int)factor;
                if (index == 43) index = 42;
                const double fint = factor - (double)index;
                for (int ib = 1; ib <= 16; ++ib) {
                    const double a0 = K.absice2[(index - 1) + 43 * (ib - 1)], a1 = K.absice2[index + 43 * (ib - 1)];
                    abscoice[ib] = a0 + fint * (a1 - (a0));
                }
                iceind = 2;
            } else if (in.iceflg == 3) {
                if (radice < 5.0 || radice > 140.0) { stop = 3; break; }
                ncbands = 16;
                const double factor = (radice - 2.) / 3.;
                int index = (int)factor;
                if (index == 46) index = 45;
                const double fint = factor - (double)index;
                for (int ib = 1; ib <= 16; ++ib) {
                    const double a0 = K.absice3[(index - 1) + 46 * (ib - 1)], a1 = K.absice3[index + 46 * (ib - 1)];
                    abscoice[ib] = a0 + fint * (a1 - (a0));
                }
                iceind = 2;
            }
            if (clwp == 0.0) {
                abscoliq[1] = 0.0;
                liqind = 0;
                if (iceind == 1) iceind = 2;
            } else if (in.liqflg == 0) {
                abscoliq[1] = K.absliq0;
                liqind = 0;
                if (iceind == 1) iceind = 2;
            } else if (in.liqflg == 1) {
                const double radliq = in.reliq ? in.reliq[o] : 0.;
                if (radliq < 2.5 || radliq > 60.) { stop = 4; break; }
                int index = (int)(radliq - 1.5);
                if (index == 0) index = 1;
                if (index == 58) index = 57;
                const double fint = radliq - 1.5 - (double)index;
                ncbands = 16;
                for (int ib = 1; ib <= 16; ++ib) {
                    const double a0 = K.absliq1[(index - 1) + 58 * (ib - 1)], a1 = K.absliq1[index + 58 * (ib - 1)];
                    abscoliq[ib] = a0 + fint * (a1 - (a0));
                }
                liqind = 2;
            }
            for (int ib = 1; ib <= ncbands; ++ib) tc[ib - 1] = ciwp * abscoice[icb(ib, iceind)] + clwp * abscoliq[icb(ib, liqind)];
        }
    }
    w.ncbands[col] = ncbands;
    if (stop) atomicMax(w.err, stop);
}

int lw_run_pass(const LwTables &t, const LwIn &in, const LwOut &out, LwWork &w, cudaStream_t s, double *cap)
{
    // clear sky without derivatives and without stage capture: taumol and rtrn fused per column (lw_column.cu)
    w.fused = g_tune.lw_fused && in.icld == 0 && !w.idrv && !cap;
    ktimer_begin(K_LW_PREP, s);
    lw_prep_cell_kernel<<<(unsigned)(((size_t)w.nc * w.nlay + 127) / 128), 128, 0, s>>>(t, in, w);
    lw_prep_kernel<<<(w.nc + 127) / 128, 128, 0, s>>>(t, in, w);
    ktimer_end(s);
    if (w.fused) {
        ktimer_begin(K_LW_COLUMN, s);
        const int n = lw_launch_column(t, in, out, w, s);
        ktimer_end(s);
        return 2 + n;
    }
    {
        const long items = (long)((w.nc + 31) / 32) * w.nlay;
        const size_t smem = (size_t)TM_BLOCK_WARPS * 64 * TM_STRIDE * sizeof(double);
        cudaFuncSetAttribute(lw_taumol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ktimer_begin(K_LW_TAUMOL, s);
        lw_taumol_kernel<<<(unsigned)((items + TM_BLOCK_WARPS - 1) / TM_BLOCK_WARPS), 32 * TM_BLOCK_WARPS, smem, s>>>(t, in, w, g_tune.taumol_sync > 0 ? g_tune.taumol_sync : 1);
        ktimer_end(s);
    }
    if (cap) {
        const size_t n = (size_t)w.nc * w.nlay * NGPTLW;
        cudaMemcpyAsync(cap, w.taug, n * 8, cudaMemcpyDeviceToDevice, s);
        cudaMemcpyAsync(cap + n, w.fracs, n * 8, cudaMemcpyDeviceToDevice, s);
    }
    int ncld = 0;
    if (in.icld >= 1) { lw_cldprop_kernel<<<(w.nc + 63) / 64, 64, 0, s>>>(in, w); ncld = 1; }
    ktimer_begin(K_LW_RTRN, s);
    const int nrt = lw_launch_rtrn(t, in, out, w, s);
    ktimer_end(s);
    return 3 + nrt + ncld;
}

} // namespace rrtmg
