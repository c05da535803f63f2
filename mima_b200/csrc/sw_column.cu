// sw_column.cu -- RRTMG shortwave, clear sky without aerosols: taumol_sw and the two-stream solver fused per
// (column, g-point slice) on sm_100a.
//
// What is computed: SW/src/rrtmg_sw_taumol.f90:223-1536 (taumol16..29, through sw_band_terms in sw_bands.cuh),
// SW/src/rrtmg_sw_spcvrt.f90:296-619 (clear == total for icld = 0, iaer = 0), rrtmg_sw_reftra.f90:122-303 (sw_reftra in
// sw_twostream.cuh) and rrtmg_sw_vrtqdr.f90:103-150 in the top-down-first form derived in sw_solver.cu: going down, every
// layer leaves the three numbers (zp, zq, rdnd) with which the upward flux follows from the surface value by
// u_above = zp * u_below + zq and the diffuse downward flux as rdnd * u.
//
// How: as lw_column.cu.  Lanes = 32 adjacent columns, a warp = (32-column tile, task), a task = up to six g-points of one
// band, a block = 16 tiles on one task; the block stages the task's slice of the band table into shared memory with one TMA
// bulk copy and reads its rows with LDS.128 (see lw_column.cu).  The thread first evaluates the band formula at the layer that selects the solar source (laysolfr, sw_prep), then
// walks its column from the top layer down: setcoef state of the cell (tile-major fields written by sw_prep_cell), band
// formula into registers, reftra and the downward recurrences for its g-points -- the optical depths never leave the
// registers, neighbouring columns take the same branch of reftra far more often than neighbouring g-points do, and the
// table exponentials of a warp fall into neighbouring entries.  The three numbers per (cell, g-point) and the task's
// direct-plus-diffuse downward sum of the level go to a tile-major scratch field in whole 256-byte rows; the second pass
// streams them back from the surface up.  Per (cell, g-point) that is one write and one read of 24 bytes, against 16 + 16
// of staged optical depths plus 24 + 24 of local-memory stack in the staged pipeline (sw_taumol -> sw_solver_warp).
// sw_cfinish adds the task sums of a level in task order (fixed summation order) and writes fluxes and heating rates.
//
// Used for icld = 0, iaer = 0 without stage capture; everything else keeps the staged kernels.
// Compiled with FMA contraction on, like sw_solver.cu: no table index is taken from contracted arithmetic except the
// binary-species parameter of taumol (specparm), where a last-bit difference moves a weight between two adjacent rows by
// that same bit.
#include "sw_bands.cuh"
#include "sw_twostream.cuh"

namespace rrtmg {

// This unit's copy of the constant block carries the row stride of the task slices (col_slice_rs) instead of the stride of the
// band tables in global memory: the band formulas of sw_bands.cuh form every row offset as row * B.rs.
int sw_column_upload_const(const SwConst &c)
{
    static SwConst k;
    k = c;
    for (int b = 0; b < NBNDSW; ++b) k.band[b].rs = sw_slice_rs(b);
    return cudaMemcpyToSymbol(c_sw, &k, sizeof(SwConst)) == cudaSuccess ? 0 : -1;
}

__host__ __device__ constexpr int sw_band_g0(int band)
{
    constexpr int g0[14] = {0, 6, 18, 26, 34, 44, 54, 56, 66, 74, 80, 86, 94, 100};
    return g0[band];
}
// first scratch slot of a task inside a (tile, layer) row: three slots per g-point and one per task
__host__ __device__ constexpr int sw_task_slot(int t) { return 3 * (sw_band_g0(sw_task(t).band) + sw_task(t).g0) + t; }
// launch order of the tasks inside a super-group, longest first: reftra dominates, so tasks of six g-points come before those of
// four, binary-species bands before single-species ones
__constant__ unsigned char c_sw_task_order[SW_NTASK] = {1, 2, 9, 20, 0, 21, 22, 7, 12, 16, 17, 10, 3, 4, 5, 6, 14, 15, 8, 13, 18, 19, 11};

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
// accumulator policy of sw_band_terms for the g-point slice of a task: everything stays in registers, the table rows come from
// the task's slice in shared memory (row offsets `off` in doubles, stride col_slice_rs)
template <int N>
struct SwSliceAcc {
    double t[N], r[N], sf[N];        // taug; taur of band 24; solar source
    uint32_t tab;                    // shared-memory address of the slice
    __device__ __forceinline__ void clear()
    {
#pragma unroll
        for (int g = 0; g < N; ++g) { t[g] = 0.0; r[g] = 0.0; }
    }
    __device__ __forceinline__ void add(int off, double wgt)
    {
        row_pairs<N>(tab + off * 8, [&](int g, double v) { t[g] = fma(wgt, v, t[g]); });
    }
    __device__ __forceinline__ void addc(double c)
    {
#pragma unroll
        for (int g = 0; g < N; ++g) t[g] = t[g] + c;
    }
    __device__ __forceinline__ void rayl1(int off, double wgt)
    {
        row_pairs<N>(tab + off * 8, [&](int g, double v) { r[g] = wgt * v; });
    }
    __device__ __forceinline__ void rayl2(int o0, double w0, int o1, double w1)
    {
        row_pairs<N>(tab + o0 * 8, [&](int g, double v) { r[g] = w0 * v; });
        row_pairs<N>(tab + o1 * 8, [&](int g, double v) { r[g] = fma(w1, v, r[g]); });
    }
    __device__ __forceinline__ void sflux1(int off, double wgt)
    {
        row_pairs<N>(tab + off * 8, [&](int g, double v) { sf[g] = wgt * v; });
    }
    __device__ __forceinline__ void sflux2(int o0, double w0, int o1, double w1)
    {
        row_pairs<N>(tab + o0 * 8, [&](int g, double v) { sf[g] = w0 * v; });
        row_pairs<N>(tab + o1 * 8, [&](int g, double v) { sf[g] = fma(w1, v, sf[g]); });
    }
};

// the cell's setcoef state as sw_prep_cell left it (tile-major: f points at [lay][tile][0][lane], field k at f[k * 32])
__device__ __forceinline__ void sw_load_pair(const double *__restrict__ f, SwPair &p)
{
#define SWF(k) __ldg(f + (k) * 32)
    const uint32_t v = (uint32_t)__double2loint(SWF(SF_COUNT));
    p.jp = v & 63; p.jt = (v >> 6) & 7; p.jt1 = (v >> 9) & 7; p.inds = (v >> 12) & 15; p.indf = (v >> 16) & 3;
    p.fac00 = SWF(SF_FAC00); p.fac01 = SWF(SF_FAC01); p.fac10 = SWF(SF_FAC10); p.fac11 = SWF(SF_FAC11);
    p.colh2o = SWF(SF_COLH2O); p.colco2 = SWF(SF_COLCO2); p.colo3 = SWF(SF_COLO3); p.colch4 = SWF(SF_COLCH4);
    p.colo2 = SWF(SF_COLO2); p.colmol = SWF(SF_COLMOL); p.coln2o = SWF(SF_COLN2O);
    p.selffac = SWF(SF_SELFFAC); p.selffrac = SWF(SF_SELFFRAC); p.forfac = SWF(SF_FORFAC); p.forfrac = SWF(SF_FORFRAC);
#undef SWF
}

__device__ __forceinline__ double ld_stream(const double *p)
{
    double v;
    asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(double *p, double v)
{
    asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void pf_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int TASK>
__device__ __forceinline__ void sw_column_task(const SwTables &T, const SwIn &in, const SwWork &w, int tile, int lane, uint32_t slice)
{
    constexpr int BAND = sw_task(TASK).band, G0 = sw_task(TASK).g0, N = sw_task(TASK).n;
    constexpr int SLOT = sw_task_slot(TASK);
    constexpr bool R1 = true;
    constexpr bool UVVIS = BAND >= 9 && BAND <= 12;           // bands 25-28 take the UV/visible albedos (rad.nomcica:565-578)
    const int nc = w.nc, klev = w.nlay;
    const int col = tile * 32 + lane;
    const int cc = col < nc ? col : nc - 1;                   // idle lanes of the last tile repeat its last column
    const double prmu0 = in.coszen[cc];
    const bool active = col < nc && !(prmu0 < ZEPZEN);        // night columns: zeros, written by sw_cfinish (rad.nomcica:502-510)
    if (__ballot_sync(0xffffffffu, active) == 0u) return;
    const size_t ncp = (size_t)w.ncp;
    const SwBand &B = c_sw.band[BAND];
    const double bpade = c_sw.bpade;
    const double2 *__restrict__ tb = reinterpret_cast<const double2 *>(T.exptbl);
    const double mu0 = active ? prmu0 : 1.0;
    const double rmu0 = 1. / mu0;
    const int laytrop = max(w.laytrop[cc], 0);
    const size_t fstep = (size_t)(w.ncp >> 5) * (SF_SLOTS * 32);           // one layer of the tile-major state
    const size_t sstep = (size_t)SW_NSLOT * 32;                            // one layer of the scratch field
    double *__restrict__ sc = w.colst + ((size_t)tile * klev * SW_NSLOT + SLOT) * 32 + lane;
    double *__restrict__ pup = w.cpart + ((size_t)TASK * 2 * (klev + 1)) * ncp + col;
    double *__restrict__ pdn = pup + (size_t)(klev + 1) * ncp;

    SwSliceAcc<N> pw;
    pw.tab = slice;
    SwPair p;
    // ---- the solar source: the band formula at the layer the reference leaves sfluxzen from (0 = never written)
    double zinc[N];
    {
#pragma unroll
        for (int k = 0; k < N; ++k) pw.sf[k] = 0.0;
        const int lsf = active ? w.laysolfr[(size_t)cc * 14 + BAND] : 0;
        if (lsf >= 1) {
            sw_load_pair(w.tf + w.tfld(lsf - 1, cc), p);
            pw.clear();
            sw_band_terms<BAND>(p, lsf <= laytrop, true, pw);
        }
#pragma unroll
        for (int k = 0; k < N; ++k) zinc[k] = active ? in.adjflux * pw.sf[k] * prmu0 : 0.0;
    }
    // Rayleigh: taur(g) = colmol * rayl(g) with one table row per band, except band 24 (pw.r)
    double raylg[N];
    if (BAND == 8) {
#pragma unroll
        for (int k = 0; k < N; ++k) raylg[k] = 1.0;
    } else {
        row_pairs<N>(slice + B.sec[SS_RAYL] * B.rs * 8, [&](int g, double v) { raylg[g] = v; });
    }

    // ---- pass 1, top -> surface: reftra + vrtqdr's top-down recurrence
    double tdn[N], rdnd[N], tdbt[N];
#pragma unroll
    for (int k = 0; k < N; ++k) { tdn[k] = 1.; rdnd[k] = 0.; tdbt[k] = 1.; }
    const double *__restrict__ fp = w.tf + w.tfld(klev - 1, cc);
    for (int lay = klev - 1; lay >= 0; --lay) {
        sw_load_pair(fp, p);
        fp -= fstep;
        pw.clear();
        sw_band_terms<BAND>(p, (lay + 1) <= laytrop, false, pw);
        double *__restrict__ s = sc + (size_t)lay * sstep;
        double dn0 = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const double tr = BAND == 8 ? pw.r[k] : p.colmol * raylg[k];
            dn0 = fma(zinc[k], tdn[k], dn0);                  // level lay + 1, above this layer
            const double rd0 = rdnd[k];
            const double dif = tdn[k] - tdbt[k];
            double ref, refd, tra, trad, dbt;
            sw_reftra<R1>(tb, bpade, mu0, rmu0, tr, pw.t[k], ref, refd, tra, trad, dbt);
            const double zreflect = rcp_sel<R1>(1. - refd * rd0);
            const double zp = trad * zreflect;
            const double zq = zinc[k] * ((ref * tdbt[k] + refd * dif) * zreflect);
            tdn[k] = tdbt[k] * tra + (trad * (dif + tdbt[k] * ref * rd0)) * zreflect;
            rdnd[k] = refd + trad * trad * rd0 * zreflect;
            tdbt[k] = dbt * tdbt[k];
            if (active) { st_stream(s + (3 * k) * 32, zp); st_stream(s + (3 * k + 1) * 32, zq); st_stream(s + (3 * k + 2) * 32, rd0); }
        }
        if (active) st_stream(s + (3 * N) * 32, dn0);
    }
    // ---- surface (level 0): upward flux from the albedos (the reference's pfu there)
    double u[N];
    {
        const double sd = UVVIS ? in.asdif[cc] : in.aldif[cc];   // palbd: diffuse
        const double sp = UVVIS ? in.asdir[cc] : in.aldir[cc];   // palbp: direct
        double dn0 = 0.0, pu = 0.0, pd = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            dn0 = fma(zinc[k], tdn[k], dn0);
            u[k] = zinc[k] * ((sp * tdbt[k] + sd * (tdn[k] - tdbt[k])) * rcp_sel<R1>(1. - sd * rdnd[k]));
            pu += u[k];
            pd = fma(rdnd[k], u[k], pd);
        }
        if (active) { pup[0] = pu; pdn[0] = dn0 + pd; }
    }
    // ---- pass 2, surface -> top: one layer ahead in registers, SC_AHEAD layers ahead on their way into L2
    {
        constexpr int SC_AHEAD = 4, NS = 3 * N + 1;
        double v[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) v[j] = active ? ld_stream(sc + j * 32) : 0.0;
#pragma unroll 2
        for (int lay = 0; lay < klev; ++lay) {
            double nx[NS];
            if (lay + 1 < klev) {
                const double *__restrict__ s = sc + (size_t)(lay + 1) * sstep;
#pragma unroll
                for (int j = 0; j < NS; ++j) nx[j] = active ? ld_stream(s + j * 32) : 0.0;
            }
            if (lay + SC_AHEAD < klev) {
                const double *__restrict__ s = sc + (size_t)(lay + SC_AHEAD) * sstep;
#pragma unroll
                for (int j = 0; j < NS; ++j) pf_l2(s + j * 32);
            }
            double pu = 0.0, pd = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) {
                u[k] = fma(v[3 * k], u[k], v[3 * k + 1]);
                pu += u[k];
                pd = fma(v[3 * k + 2], u[k], pd);
            }
            if (active) { pup[(size_t)(lay + 1) * ncp] = pu; pdn[(size_t)(lay + 1) * ncp] = v[3 * N] + pd; }
#pragma unroll
            for (int j = 0; j < NS; ++j) v[j] = nx[j];
        }
    }
}

// One 16-warp block per SM, all of its warps on the same task (instruction cache and table slice: see lw_column.cu).  A task
// body with its six inlined reftra instances is about 25 KB.
template <int WARPS, int BLOCKS>
__global__ void __launch_bounds__(32 * WARPS, BLOCKS) sw_column_kernel(SwTables T, SwIn in, SwWork w)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // block order: super-groups of columns, inside one the tile groups of the longest task first (see lw_column_kernel)
    constexpr int SG = COL_SUPER_COLS / (32 * WARPS);
    const int ngrp = ((w.nc + 31) / 32 + WARPS - 1) / WARPS;
    const int sg = blockIdx.x / (SG * SW_NTASK);
    const int gcount = min(SG, ngrp - sg * SG);
    const int r = blockIdx.x - sg * (SG * SW_NTASK);
    const int rank = r / gcount;
    const int grp = sg * SG + (r - rank * gcount);
    const int task = c_sw_task_order[rank];
    const int tile = grp * WARPS + wid;
    // the task's slice of the band table: one bulk copy (TMA) into shared memory per block
    extern __shared__ __align__(128) unsigned char s_slice[];
    __shared__ uint64_t s_bar;
    const uint32_t slice = stage_to_shared(s_slice, &s_bar, T.sl.data + T.sl.off[task], (uint32_t)T.sl.bytes[task]);
    if (tile * 32 >= w.nc) return;
#define SC_TASK(t) case t: sw_column_task<t>(T, in, w, tile, lane, slice); break
    switch (task) {
        SC_TASK(0); SC_TASK(1); SC_TASK(2); SC_TASK(3); SC_TASK(4); SC_TASK(5); SC_TASK(6); SC_TASK(7);
        SC_TASK(8); SC_TASK(9); SC_TASK(10); SC_TASK(11); SC_TASK(12); SC_TASK(13); SC_TASK(14); SC_TASK(15);
        SC_TASK(16); SC_TASK(17); SC_TASK(18); SC_TASK(19); SC_TASK(20); SC_TASK(21); SC_TASK(22);
    }
#undef SC_TASK
}

// The task sums of a level, added in task order; fluxes and heating rates (rrtmg_sw_rad.nomcica.f90:686-727; clear == total
// for icld = 0).  Block = 32 columns x 8 level lanes; the level fluxes of the tile pass through shared memory.
constexpr int SF_ROWS = 8;
__global__ void __launch_bounds__(32 * SF_ROWS) sw_cfinish_kernel(SwIn in, SwOut out, SwWork w)
{
    extern __shared__ double s_fx[];                            // [2][nlay + 1][32]
    const int klev = w.nlay, nlev = klev + 1;
    const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + lane;
    const bool valid = col < w.nc;
    const bool active = valid && !(in.coszen[col] < ZEPZEN);
    const size_t ncp = (size_t)w.ncp;
    double *s_up = s_fx, *s_dn = s_fx + (size_t)nlev * 32;
    for (int lev = row; lev < nlev; lev += SF_ROWS) {
        double d = 0.0, u = 0.0;
        if (active) {
            const double *pu = w.cpart + (size_t)lev * ncp + col;
#pragma unroll
            for (int t = 0; t < SW_NTASK; ++t) {
                u += pu[(size_t)t * 2 * nlev * ncp];
                d += pu[((size_t)t * 2 + 1) * nlev * ncp];
            }
        }
        s_up[lev * 32 + lane] = u;
        s_dn[lev * 32 + lane] = d;
    }
    __syncthreads();
    if (!valid) return;
    for (int lev = row; lev < nlev; lev += SF_ROWS) {
        const size_t o = col + (size_t)lev * out.ld;
        const double u = s_up[lev * 32 + lane], d = s_dn[lev * 32 + lane];
        out.uflx[o] = u; out.dflx[o] = d; out.uflxc[o] = u; out.dflxc[o] = d;
        if (lev < klev) {
            double h = 0.0;
            if (active && lev < klev - 1) {      // MiMA: no heating in the top layer (rad.nomcica:724-726)
                const double pdp = in.plev[col + (size_t)lev * in.ld] - in.plev[col + (size_t)(lev + 1) * in.ld];
                h = ((s_dn[(lev + 1) * 32 + lane] - s_up[(lev + 1) * 32 + lane]) - (d - u)) * (c_sw.heatfac / pdp);
            }
            out.hr[o] = h;
            out.hrc[o] = h;
        }
    }
}

template <int WARPS, int BLOCKS>
static void sw_launch_column_geom(const SwTables &t, const SwIn &in, SwWork &w, cudaStream_t s)
{
    const int ntile = (w.nc + 31) / 32;
    const unsigned grid = (unsigned)((ntile + WARPS - 1) / WARPS) * SW_NTASK;
    const size_t smem = (size_t)t.sl.max_bytes;
    cudaFuncSetAttribute(sw_column_kernel<WARPS, BLOCKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sw_column_kernel<WARPS, BLOCKS><<<grid, 32 * WARPS, smem, s>>>(t, in, w);
}

// returns the number of launches
int sw_launch_column(const SwTables &t, const SwIn &in, const SwOut &out, SwWork &w, cudaStream_t s)
{
    const int ntile = (w.nc + 31) / 32;
    const bool wide = g_tune.col_warps != 8;
    if (wide) sw_launch_column_geom<16, 1>(t, in, w, s); else sw_launch_column_geom<8, 2>(t, in, w, s);
    const size_t smem = (size_t)2 * (w.nlay + 1) * 32 * sizeof(double);
    if (smem > 48 * 1024) cudaFuncSetAttribute(sw_cfinish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sw_cfinish_kernel<<<ntile, 32 * SF_ROWS, smem, s>>>(in, out, w);
    return 2;
}

} // namespace rrtmg
