// lw_column.cu -- RRTMG longwave, clear sky: taumol and rtrn fused per (column, g-point slice) on sm_100a.
//
// What is computed: LW/src/rrtmg_lw_taumol.f90:260-3147 (taugb1..16, through lw_band_terms in lw_bands.cuh),
// the Planck sources of LW/src/rrtmg_lw_setcoef.f90:154-249, and the clear branches of LW/src/rrtmg_lw_rtrnmr.f90:481-777
// (identical in rtrnmc.f90:407-432, 481-503); taut = taug + tauaer (rad.nomcica:514-519, iaer = 10 forced).
//
// How: lanes = 32 adjacent columns, a warp = (32-column tile, task), a task = up to eight g-points of one band, a block = 16
// tiles on ONE task.  The block first stages the task's slice of its band's k-distribution table -- every row of the band table
// restricted to the task's g-points, at most 164 KB -- into shared memory with one TMA bulk copy (cp.async.bulk + mbarrier,
// stage_to_shared); the slices are laid out at init (ColSlices, api.cu) with a row stride of an odd number of 16-byte units, so
// the two to four distinct rows that the 32 columns of a warp read with one LDS.128 fall into different banks.  Per-lane reads
// of warp-uniform or nearly uniform table rows cost the L1 data pipe one pass per quarter-warp AND distinct row
// (tools/micro/l1_rows.cu: 4.1 clk per row from L1, 2.2 clk from shared memory); with the rows read through L1 that pipe was
// the kernel's bound (75 % busy; rows made loop-invariant: -1.9 ms, profiles/r02_summary.md experiment 6).
// The thread then walks its column from the top layer down: it reads the cell's setcoef state (written once per cell by
// lw_prep_cell, read back by the 23 tasks of the tile while it is still in L2), evaluates the band formula for its g-points into
// registers, interpolates the three Planck values of the layer, and feeds the N downward recurrences at once -- the optical depths
// never leave the registers, and the N independent chains hide the latency of the one-level-at-a-time recurrence.  What the
// upward sweep needs again goes to a scratch field in whole 512-byte rows per warp: the layer's optical depths (8 bytes per
// g-point, two per 16-byte store) and, per task, how the layer's Planck fractions were formed (a weight and two row offsets);
// the upward sweep streams that back and forms absorptivity, fractions and upward source again with the operations of the
// downward sweep -- bitwise the same values -- instead of reading {absorptivity, source} pairs of 16 bytes per g-point.  So the
// staging traffic is one write and one read of 8 bytes per (cell, g-point) plus 16 per (cell, task), every access a whole
// number of 128-byte lines, where the staged pipeline (lw_taumol -> [col][lay][g] -> lw_rtrn) wrote 16 bytes per (cell,
// g-point) in scattered 16-128 byte pieces and read them twice.  Each
// warp leaves its g-sums per level in a partial field [task][level][column]; lw_finish adds the 23 partials of a level in
// task order (fixed summation order: results are reproducible bit for bit) and writes fluxes and heating rates.
//
// Used for icld = 0, idrv = 0 without stage capture; everything else keeps the staged kernels (lw_kernels.cu, lw_solver.cu).
// Compiled with -fmad=false like the other setcoef/taumol code; the recurrences spell their fma() out.
#include "lw_bands.cuh"

namespace rrtmg {

// This unit's copy of the constant block carries the row stride of the task slices (col_slice_rs) instead of the stride of the
// band tables in global memory: the band formulas of lw_bands.cuh form every row offset as row * B.rs.
int lw_column_upload_const(const LwConst &c)
{
    static LwConst k;
    k = c;
    for (int b = 0; b < NBNDLW; ++b) k.band[b].rs = lw_slice_rs(b);
    return cudaMemcpyToSymbol(c_lw, &k, sizeof(LwConst)) == cudaSuccess ? 0 : -1;
}

// launch order of the tasks of a tile group: the long ones (binary-species bands, eight g-points) first
__constant__ unsigned char c_task_order[LW_NTASK] = {4, 5, 8, 9, 6, 7, 11, 12, 14, 15, 18, 19, 22, 21, 0, 13, 10, 17, 2, 3, 1, 16, 20};

// 16 bytes of a table row from the block's shared-memory copy of the task slice (LDS.128 with an immediate offset)
template <int IMM>
__device__ __forceinline__ double2 lds2(uint32_t a)
{
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(a), "n"(IMM));
    return v;
}
template <int N, int J = 0, class F>
__device__ __forceinline__ void row_pairs(uint32_t a, F f)
{
    if constexpr (J < N / 2) {
        const double2 v = lds2<J * 16>(a);
        f(2 * J, v.x);
        f(2 * J + 1, v.y);
        row_pairs<N, J + 1>(a, f);
    }
}
// accumulator policy of lw_band_terms for the g-point slice of a task: taug and fracs stay in registers, the table rows come
// from the task's slice in shared memory (row offsets `off` in doubles, stride col_slice_rs)
template <int N>
struct SliceAcc {
    double t[N], f[N];
    uint32_t tab;                    // shared-memory address of the slice
    // how the Planck fractions of the layer were formed, for the upward sweep to form them again: none (0), one row (1),
    // (1 - fw1) * row(fo0) + fw1 * row(fo1) (2)
    int fkind, fo0, fo1;
    double fw1;
    __device__ __forceinline__ void clear()
    {
#pragma unroll
        for (int g = 0; g < N; ++g) { t[g] = 0.0; f[g] = 0.0; }
    }
    __device__ __forceinline__ void add(int off, double wgt)
    {
        row_pairs<N>(tab + off * 8, [&](int g, double v) { t[g] = fma(wgt, v, t[g]); });
    }
    __device__ __forceinline__ void add_nz(int off, double wgt)
    {
        if (wgt != 0.0) add(off, wgt);
    }
    __device__ __forceinline__ void scale(int off)
    {
        row_pairs<N>(tab + off * 8, [&](int g, double v) { t[g] = t[g] * v; });
    }
    __device__ __forceinline__ void frac1(int off)
    {
        fkind = 1; fo0 = off; fo1 = 0; fw1 = 0.0;
        row_pairs<N>(tab + off * 8, [&](int g, double v) { f[g] = v; });
    }
    // every caller passes w0 = 1 - w1 (frac_eta in lw_bands.cuh): the upward sweep forms it the same way
    __device__ __forceinline__ void frac2(int o0, double w0, int o1, double w1)
    {
        fkind = 2; fo0 = o0; fo1 = o1; fw1 = w1;
        row_pairs<N>(tab + o0 * 8, [&](int g, double v) { f[g] = w0 * v; });
        row_pairs<N>(tab + o1 * 8, [&](int g, double v) { f[g] = fma(w1, v, f[g]); });
    }
    __device__ __forceinline__ void fzero()
    {
        fkind = 0; fo0 = 0; fo1 = 0; fw1 = 0.0;
#pragma unroll
        for (int g = 0; g < N; ++g) f[g] = 0.0;
    }
    // the fractions again from what frac1 / frac2 / fzero recorded (bitwise the same values)
    __device__ __forceinline__ void refrac(int kind, int o0, int o1, double w1)
    {
        if (kind == 2) {
            const double w0 = 1. - w1;
            row_pairs<N>(tab + o0 * 8, [&](int g, double v) { f[g] = w0 * v; });
            row_pairs<N>(tab + o1 * 8, [&](int g, double v) { f[g] = fma(w1, v, f[g]); });
        } else if (kind == 1) {
            row_pairs<N>(tab + o0 * 8, [&](int g, double v) { f[g] = v; });
        } else {
#pragma unroll
            for (int g = 0; g < N; ++g) f[g] = 0.0;
        }
    }
};

// the cell's setcoef state as lw_prep_cell left it (tile-major: f points at [lay][tile][0][lane], field k at f[k * 32]); a band
// uses a few of these, the rest of the loads are dropped by the compiler
__device__ __forceinline__ void lw_load_pair(const double *__restrict__ f, const LwIn &in, LwPair &p)
{
#define LWF(k) __ldg(f + (k) * 32)
    const LwIdx ix = lw_unpack((uint32_t)__double2loint(LWF(LF_COUNT)));
    p.jp = ix.jp; p.jt = ix.jt; p.jt1 = ix.jt1; p.inds = ix.inds; p.indf = ix.indf; p.indm = ix.indm;
    p.fac00 = LWF(LF_FAC00); p.fac01 = LWF(LF_FAC01); p.fac10 = LWF(LF_FAC10); p.fac11 = LWF(LF_FAC11);
    p.colh2o = LWF(LF_COLH2O); p.colco2 = LWF(LF_COLCO2); p.colo3 = LWF(LF_COLO3); p.coln2o = LWF(LF_COLN2O);
    p.colco = LWF(LF_COLCO); p.colch4 = LWF(LF_COLCH4); p.colo2 = LWF(LF_COLO2); p.colbrd = LWF(LF_COLBRD);
    p.selffac = LWF(LF_SELFFAC); p.selffrac = LWF(LF_SELFFRAC); p.forfac = LWF(LF_FORFAC); p.forfrac = LWF(LF_FORFRAC);
    p.minorfrac = LWF(LF_MINORFRAC); p.scaleminor = LWF(LF_SCALEMINOR); p.scaleminorn2 = LWF(LF_SCALEMINORN2);
    p.coldry = LWF(LF_COLDRY); p.pavel = LWF(LF_PAVEL);
    // the cross-section amounts are exactly zero when the caller passes no CFC array (MiMA's configuration)
    p.wx1 = in.ccl4 ? LWF(LF_WX1) : 0.0; p.wx2 = in.cfc11 ? LWF(LF_WX2) : 0.0;
    p.wx3 = in.cfc12 ? LWF(LF_WX3) : 0.0; p.wx4 = in.cfc22 ? LWF(LF_WX4) : 0.0;
#undef LWF
}

__device__ __forceinline__ void pf_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// the scratch field is written once and read once, 60 layers apart: keep it out of L1, which the k-tables, the exp/tfn
// table and the prefetched setcoef state need
__device__ __forceinline__ double2 ld_scratch(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_scratch(double2 *p, double a, double b)
{
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
// absorptivity and the Pade "tau function" of one (cell, g-point) from its optical depth (rtrnmr.f90:536-560): series below 0.06,
// the reference's look-up table above; called by both sweeps, so that the upward one reproduces the downward one's values
__device__ __forceinline__ void lw_trans(const double2 *__restrict__ et, double bpade, double odepth, double &at, double &tf)
{
    if (odepth <= 0.06) {
        at = odepth - 0.5 * odepth * odepth;
        tf = 0.166667 * odepth;
    } else {
        const double tblind = odepth * rcp_fast(bpade + odepth);
        const int itr = (int)(10000.0 * tblind + 0.5);
        const double2 e = ld_tbl(et + itr);
        at = 1. - e.x;
        tf = e.y;
    }
}
// integrated Planck function of one band at temperature t (setcoef.f90:154-249: linear in the 1 K table)
__device__ __forceinline__ double lw_planck(const double *__restrict__ tp, double t)
{
    int ind = (int)(t - 159.);
    ind = ind < 1 ? 1 : (ind > 180 ? 180 : ind);
    const double frac = t - 159. - (double)ind;
    const double lo = __ldg(tp + ind - 1);
    const double d = __ldg(tp + ind) - lo;
    return lo + frac * d;
}

template <int BAND, int G0, int N, bool AER>
__device__ __forceinline__ void lw_column_task(const LwTables &T, const LwIn &in, const LwWork &w, int task, int tile, int lane,
                                               uint32_t slice)
{
    const int nc = w.nc, nlay = w.nlay;
    const int col = tile * 32 + lane;
    const bool valid = col < nc;
    const int cc = valid ? col : nc - 1;                       // idle lanes of the last tile repeat its last column
    const size_t ld = (size_t)in.ld;
    const size_t ncp = (size_t)w.ncp;
    const LwBand &B = c_lw.band[BAND];
    const int gfirst = B.g0 + G0;                              // first g-point of the task in the 140-vector
    const double secd = w.secdiff[(size_t)cc * 16 + BAND];
    const double wgt = 0.5 * c_lw.delwave[BAND];
    const double bpade = c_lw.bpade;
    const int laytrop = w.laytrop[cc];
    const double2 *__restrict__ et = reinterpret_cast<const double2 *>(T.exptfn);
    const double *__restrict__ tp = T.totplnk + BAND * 181;
    const double *taer = AER ? in.tauaer + cc + (size_t)BAND * nlay * ld : nullptr;
    // scratch: per (tile, layer) LW_CSLOT 16-byte slots of 32 lanes; a task owns N / 2 slots of optical-depth pairs and one slot
    // {weight, packed row offsets} that says how the layer's Planck fractions were formed
    constexpr int NP = N / 2;
    double2 *__restrict__ sc = reinterpret_cast<double2 *>(w.colst) + ((size_t)tile * nlay * LW_CSLOT + (gfirst >> 1) + task) * 32 + lane;
    double *__restrict__ pdn = w.part + ((size_t)task * 2 * (nlay + 1)) * ncp + col;
    double *__restrict__ pup = pdn + (size_t)(nlay + 1) * ncp;

    SliceAcc<N> pw;
    pw.tab = slice;
    double rad[N];
#pragma unroll
    for (int k = 0; k < N; ++k) rad[k] = 0.0;
    if (valid) pdn[(size_t)nlay * ncp] = 0.0;                  // no downward flux at the top
    double plev_up = lw_planck(tp, in.tlev[cc + (size_t)nlay * ld]);

    // downward sweep (:505-618), top layer first
    LwPair p;
    const size_t fstep = (size_t)(w.ncp >> 5) * (LF_SLOTS * 32);           // one layer of the tile-major state
    const double *__restrict__ fp = w.f + w.tfld(nlay - 1, cc);
    const double *__restrict__ tlp = in.tlay + cc + (size_t)(nlay - 1) * ld, *__restrict__ tvp = in.tlev + cc + (size_t)(nlay - 1) * ld;
    for (int lay = nlay - 1; lay >= 0; --lay) {
        lw_load_pair(fp, in, p);
        const double tl = *tlp, tv = *tvp;
        fp -= fstep; tlp -= ld; tvp -= ld;
        const bool lower = (lay + 1) <= laytrop;
        pw.clear();
        lw_band_terms<BAND>(p, lower, pw);
        const double blay = lw_planck(tp, tl);
        const double plev_dn = lw_planck(tp, tv);
        const double dplankup = plev_up - blay, dplankdn = plev_dn - blay;
        plev_up = plev_dn;
        double ta = 0.0;
        if (AER) ta = taer[(size_t)lay * ld];
        double2 *__restrict__ s = sc + (size_t)lay * (LW_CSLOT * 32);
        double sum = 0.0, od0 = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            double tg = pw.t[k];
            if (AER) tg = tg + ta;
            const double plfrac = pw.f[k];
            double odepth = secd * tg;
            if (odepth < 0.0) odepth = 0.0;
            double at, tf;
            lw_trans(et, bpade, odepth, at, tf);
            const double bbd = plfrac * fma(tf, dplankdn, blay);
            rad[k] = fma(bbd - rad[k], at, rad[k]);
            sum = fma(rad[k], wgt, sum);
            if (k & 1) { if (valid) st_scratch(s + (k >> 1) * 32, od0, odepth); }
            else od0 = odepth;
        }
        if (valid) {
            st_scratch(s + NP * 32, pw.fw1, __hiloint2double(pw.fo1 | (pw.fkind << 28), pw.fo0));
            pdn[(size_t)lay * ncp] = sum;
        }
    }
    // surface (:628-636): after the last iteration pw.f holds the Planck fractions of layer 1
    {
        const double semiss = in.emis ? in.emis[cc + (size_t)BAND * ld] : 1.0;
        const double pb = w.plankbnd[(size_t)cc * 16 + BAND];
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const double rad0 = pw.f[k] * pb;
            rad[k] = rad0 + (1. - semiss) * rad[k];
            sum = fma(rad[k], wgt, sum);
        }
        if (valid) pup[0] = sum;
    }
    // upward sweep (:649-711): the optical depths and the fraction recipe come back from the scratch field, one layer ahead in
    // registers and LC_AHEAD layers ahead on their way into L2; absorptivity, Planck fractions and the upward source are formed
    // again by the operations of the downward sweep (bitwise the same values).  Half the scratch bytes of storing {absorptivity,
    // source} per g-point, for arithmetic the kernel has room for
    {
        constexpr int LC_AHEAD = 4, NS = NP + 1;
        const double2 zero2 = make_double2(0.0, 0.0);
        const double *__restrict__ tlp = in.tlay + cc, *__restrict__ tvp = in.tlev + cc + ld;   // layer 1, its upper interface
        double2 v[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) v[j] = valid ? ld_scratch(sc + j * 32) : zero2;
        double tl = *tlp, tv = *tvp;
#pragma unroll 2
        for (int lay = 0; lay < nlay; ++lay) {
            double2 nx[NS];
            double tln = 0.0, tvn = 0.0;
            if (lay + 1 < nlay) {
                const double2 *__restrict__ s = sc + (size_t)(lay + 1) * (LW_CSLOT * 32);
#pragma unroll
                for (int j = 0; j < NS; ++j) nx[j] = valid ? ld_scratch(s + j * 32) : zero2;
                tlp += ld; tvp += ld;
                tln = *tlp; tvn = *tvp;
            }
            if (lay + LC_AHEAD < nlay) {
                const double2 *__restrict__ s = sc + (size_t)(lay + LC_AHEAD) * (LW_CSLOT * 32);
#pragma unroll
                for (int j = 0; j < NS; ++j) pf_l2(s + j * 32);
            }
            const double blay = lw_planck(tp, tl);
            const double dplankup = lw_planck(tp, tv) - blay;
            const int hi = __double2hiint(v[NP].y), lo = __double2loint(v[NP].y);
            pw.refrac(valid ? (hi >> 28) : 0, lo, hi & 0x0fffffff, v[NP].x);
            double sum = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) {
                const double odepth = (k & 1) ? v[k >> 1].y : v[k >> 1].x;
                double at, tf;
                lw_trans(et, bpade, odepth, at, tf);
                const double bbu = pw.f[k] * fma(tf, dplankup, blay);
                rad[k] = fma(bbu - rad[k], at, rad[k]);
                sum = fma(rad[k], wgt, sum);
            }
            if (valid) pup[(size_t)(lay + 1) * ncp] = sum;
#pragma unroll
            for (int j = 0; j < NS; ++j) v[j] = nx[j];
            tl = tln; tv = tvn;
        }
    }
}

// One 16-warp block per SM, all of its warps on the same task: the loop body of a task is 10-20 KB of straight-line code and the
// instruction cache behind the 6 KB L0 holds 32 KB -- with four 4-warp blocks of different tasks per SM the kernel spent 22 of 23
// issue slots waiting for instructions (profiles/r02_summary.md, experiment 5) -- and the task's table slice is staged once per
// block.  The dynamic shared memory of a launch is the largest slice (a launch has one size), which leaves room for one block per
// SM; option "col_warps" = 8 runs 8-warp blocks (a test knob: half the warps per SM).
template <bool AER, int WARPS, int BLOCKS>
__global__ void __launch_bounds__(32 * WARPS, BLOCKS) lw_column_kernel(LwTables T, LwIn in, LwWork w)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // Block order: super-groups of COL_SUPER_COLS columns; inside one, all tile groups of the longest task first, then all
    // of the next, ... (c_task_order).  The 23 tasks of a tile then still run close enough in time to share the tile's
    // setcoef state through L2 (a super-group's state is ~50 MB), and what is left for the end of the grid are the short
    // tasks of the last super-group: with ~5 blocks per SM (a rank's 16384 columns at 8 GPUs) the kernel's tail is a short
    // block instead of a long one.
    constexpr int SG = COL_SUPER_COLS / (32 * WARPS);
    const int ngrp = ((w.nc + 31) / 32 + WARPS - 1) / WARPS;
    const int sg = blockIdx.x / (SG * LW_NTASK);
    const int gcount = min(SG, ngrp - sg * SG);
    const int r = blockIdx.x - sg * (SG * LW_NTASK);
    const int rank = r / gcount;
    const int grp = sg * SG + (r - rank * gcount);
    const int task = c_task_order[rank];
    const int tile = grp * WARPS + wid;
    // the task's slice of the band table: one bulk copy (TMA) into shared memory per block
    extern __shared__ __align__(128) unsigned char s_slice[];
    __shared__ uint64_t s_bar;
    const uint32_t slice = stage_to_shared(s_slice, &s_bar, T.sl.data + T.sl.off[task], (uint32_t)T.sl.bytes[task]);
    if (tile * 32 >= w.nc) return;
#define LC_TASK(t) case t: lw_column_task<lw_task(t).band, lw_task(t).g0, lw_task(t).n, AER>(T, in, w, t, tile, lane, slice); break
    switch (task) {
        LC_TASK(0); LC_TASK(1); LC_TASK(2); LC_TASK(3); LC_TASK(4); LC_TASK(5); LC_TASK(6); LC_TASK(7);
        LC_TASK(8); LC_TASK(9); LC_TASK(10); LC_TASK(11); LC_TASK(12); LC_TASK(13); LC_TASK(14); LC_TASK(15);
        LC_TASK(16); LC_TASK(17); LC_TASK(18); LC_TASK(19); LC_TASK(20); LC_TASK(21); LC_TASK(22);
    }
#undef LC_TASK
}

// The partial g-sums of a level, added in task order; fluxes, heating rates (:751-777) and the copy-out (rad.nomcica:546-555).
// Block = 32 columns x 8 level lanes; the level fluxes of the tile pass through shared memory for the flux differences.
constexpr int LF_ROWS = 8;
__global__ void __launch_bounds__(32 * LF_ROWS) lw_finish_kernel(LwIn in, LwOut out, LwWork w)
{
    extern __shared__ double s_fx[];                            // [2][nlay + 1][32]
    const int nlay = w.nlay, nlev = nlay + 1;
    const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + lane;
    const bool valid = col < w.nc;
    const size_t ncp = (size_t)w.ncp;
    double *s_dn = s_fx, *s_up = s_fx + (size_t)nlev * 32;
    for (int lev = row; lev < nlev; lev += LF_ROWS) {
        double d = 0.0, u = 0.0;
        if (valid) {
            const double *pd = w.part + (size_t)lev * ncp + col;
#pragma unroll
            for (int t = 0; t < LW_NTASK; ++t) {
                d += pd[(size_t)t * 2 * nlev * ncp];
                u += pd[((size_t)t * 2 + 1) * nlev * ncp];
            }
        }
        s_dn[lev * 32 + lane] = d * c_lw.fluxfac;
        s_up[lev * 32 + lane] = u * c_lw.fluxfac;
    }
    __syncthreads();
    if (!valid) return;
    for (int lev = row; lev < nlev; lev += LF_ROWS) {
        const size_t o = col + (size_t)lev * out.ld;
        const double u = s_up[lev * 32 + lane], d = s_dn[lev * 32 + lane];
        out.uflx[o] = u; out.dflx[o] = d;
        out.uflxc[o] = u; out.dflxc[o] = d;
        if (lev < nlay) {
            const double fnet0 = u - d, fnet1 = s_up[(lev + 1) * 32 + lane] - s_dn[(lev + 1) * 32 + lane];
            const double pz0 = in.plev[col + (size_t)lev * in.ld], pz1 = in.plev[col + (size_t)(lev + 1) * in.ld];
            const double h = c_lw.heatfac * (fnet0 - fnet1) / (pz0 - pz1);
            out.hr[o] = h;
            out.hrc[o] = h;
        }
    }
}

template <bool AER, int WARPS, int BLOCKS>
static void lw_launch_column_geom(const LwTables &t, const LwIn &in, LwWork &w, cudaStream_t s)
{
    const int ntile = (w.nc + 31) / 32;
    const unsigned grid = (unsigned)((ntile + WARPS - 1) / WARPS) * LW_NTASK;
    const size_t smem = (size_t)t.sl.max_bytes;
    cudaFuncSetAttribute(lw_column_kernel<AER, WARPS, BLOCKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    lw_column_kernel<AER, WARPS, BLOCKS><<<grid, 32 * WARPS, smem, s>>>(t, in, w);
}

// returns the number of launches
int lw_launch_column(const LwTables &t, const LwIn &in, const LwOut &out, LwWork &w, cudaStream_t s)
{
    const int ntile = (w.nc + 31) / 32;
    const bool wide = g_tune.col_warps != 8;
    if (in.tauaer) { if (wide) lw_launch_column_geom<true, 16, 1>(t, in, w, s); else lw_launch_column_geom<true, 8, 2>(t, in, w, s); }
    else { if (wide) lw_launch_column_geom<false, 16, 1>(t, in, w, s); else lw_launch_column_geom<false, 8, 2>(t, in, w, s); }
    const size_t smem = (size_t)2 * (w.nlay + 1) * 32 * sizeof(double);
    if (smem > 48 * 1024) cudaFuncSetAttribute(lw_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    lw_finish_kernel<<<ntile, 32 * LF_ROWS, smem, s>>>(in, out, w);
    return 2;
}

} // namespace rrtmg
