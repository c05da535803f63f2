// rrtmg_dev.cuh -- device-side data layout shared by the LW and SW kernels (sm_100a).
//
// HBM layout (one "pass" = a chunk of nc columns, all nlay layers):
//   * interface arrays arrive (ncol, nlay) column-major: the column index is contiguous, so kernels whose lanes are
//     adjacent columns of one layer (prep, taumol, the column kernels) read fully coalesced.
//   * clear sky (lw_column.cu, sw_column.cu): everything between the kernels is tile-major -- per 32-column tile
//     [layer][slot][32 lanes], a slot being one 256-byte row at an immediate offset: the setcoef state written by
//     *_prep_cell (LwWork::f / SwWork::tf), the scratch field between the two sweeps of a column kernel (colst) and the
//     per-task g-sums per level (part / cpart).
//   * staged kernels (clouds, aerosols, idrv = 1, stage capture): the per-cell setcoef state is evaluated in place by
//     taumol (lw_cell / sw_cell; the fields F[field][lay][col] and idx are what the stage-capture hook returns); the
//     taumol -> solver staging fields (LW taug/fracs, SW taug) are stored [col][lay][g] with the g-point index fastest:
//     one warp of a solver (lanes = g-points of a column) reads 256 contiguous bytes per layer step; taumol transposes
//     its per-cell results through a per-warp shared-memory slab and writes whole rows in 16-byte pieces.
//   * k-distribution tables are reduced and transposed at init from the Fortran (row, ig) to [row][ig] with the row
//     stride padded to a power of two, so that the values one interpolation term needs are one <= 128-byte line.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rrtmg {

constexpr int NBNDLW = 16, NGPTLW = 140, NBNDSW = 14, NGPTSW = 112, NTBL = 10000;
constexpr int MAXLAY = 128;
constexpr int SW_STACK_SLOTS = 160 * 32;   // resident one-warp blocks of sw_solver_l2_kernel: SMs x blocks per SM, upper bound

// ------------------------------------------------------------------------------------ LW
// packed per-(lay,col) indices: jp:6 | jt:3 | jt1:3 | indself:4 | indfor:2 | indminor:5
__host__ __device__ inline uint32_t lw_pack(int jp, int jt, int jt1, int inds, int indf, int indm)
{
    return (uint32_t)jp | ((uint32_t)jt << 6) | ((uint32_t)jt1 << 9) | ((uint32_t)inds << 12) |
           ((uint32_t)indf << 16) | ((uint32_t)indm << 18);
}
struct LwIdx { int jp, jt, jt1, inds, indf, indm; };
__host__ __device__ inline LwIdx lw_unpack(uint32_t v)
{
    LwIdx r;
    r.jp = v & 63; r.jt = (v >> 6) & 7; r.jt1 = (v >> 9) & 7; r.inds = (v >> 12) & 15;
    r.indf = (v >> 16) & 3; r.indm = (v >> 18) & 31;
    return r;
}

enum LwField {
    LF_FAC00, LF_FAC01, LF_FAC10, LF_FAC11,
    LF_COLH2O, LF_COLCO2, LF_COLO3, LF_COLN2O, LF_COLCO, LF_COLCH4, LF_COLO2, LF_COLBRD,
    LF_SELFFAC, LF_SELFFRAC, LF_FORFAC, LF_FORFRAC, LF_MINORFRAC, LF_SCALEMINOR, LF_SCALEMINORN2,
    LF_COLDRY, LF_PAVEL, LF_WX1, LF_WX2, LF_WX3, LF_WX4,
    LF_COUNT
};

// sections of one band's table; values are row offsets inside the band table (-1 = absent)
enum LwSec {
    LS_ABSA, LS_ABSB, LS_SELF, LS_FOR, LS_FRACA, LS_FRACB,
    LS_MA1, LS_MA2, LS_MA3, LS_MB1, LS_MB2, LS_X1, LS_X2, LS_GSCALE,
    LS_COUNT
};

// ---- (band, g-point slice) tasks of the fused clear-sky column kernels (lw_column.cu, sw_column.cu) and the per-task copies of
// the k-distribution tables they stage into shared memory.  A task = up to eight (LW) / six (SW) g-points of ONE band; bands of
// more g-points are cut in two so that a task's registers fit the 128-register budget of 16 warps per SM.
constexpr int COL_NTASK = 23;                    // both codes happen to have 23 tasks
struct ColTask { int band, g0, n; };
__host__ __device__ constexpr ColTask lw_task(int t)
{
    constexpr ColTask tk[COL_NTASK] = {
        {0, 0, 6}, {0, 6, 4}, {1, 0, 6}, {1, 6, 6}, {2, 0, 8}, {2, 8, 8}, {3, 0, 8}, {3, 8, 6}, {4, 0, 8}, {4, 8, 8},
        {5, 0, 8}, {6, 0, 6}, {6, 6, 6}, {7, 0, 8}, {8, 0, 6}, {8, 6, 6}, {9, 0, 6}, {10, 0, 8}, {11, 0, 8},
        {12, 0, 4}, {13, 0, 2}, {14, 0, 2}, {15, 0, 2}};
    return tk[t];
}
// SW tasks of at most four g-points (32 tasks) with 16, 20 or 24 warps per block measured within 3 % of this table
// (profiles/r02aa_sweep.txt).
__host__ __device__ constexpr ColTask sw_task(int t)
{
    constexpr ColTask tk[COL_NTASK] = {
        {0, 0, 6}, {1, 0, 6}, {1, 6, 6}, {2, 0, 4}, {2, 4, 4}, {3, 0, 4}, {3, 4, 4}, {4, 0, 6}, {4, 6, 4}, {5, 0, 6}, {5, 6, 4},
        {6, 0, 2}, {7, 0, 6}, {7, 6, 4}, {8, 0, 4}, {8, 4, 4}, {9, 0, 6}, {10, 0, 6}, {11, 0, 4}, {11, 4, 4}, {12, 0, 6},
        {13, 0, 6}, {13, 6, 6}};
    return tk[t];
}
// Row stride (in doubles) of a band's task slices: wide enough for the band's largest task and an ODD number of 16-byte units, so
// that the distinct rows the lanes of a quarter-warp read with one LDS.128 fall into different bank groups unless they are a
// multiple of eight rows apart (microbenchmark tools/micro/l1_rows.cu: two rows 144 B apart cost 2.2 clk, 128 B apart 4.1 clk).
__host__ __device__ constexpr int col_slice_rs(int nmax) { return nmax > 6 ? 10 : (nmax > 2 ? 6 : 2); }
__host__ __device__ constexpr int lw_slice_rs(int band)
{
    int n = 0;
    for (int t = 0; t < COL_NTASK; ++t) if (lw_task(t).band == band && lw_task(t).n > n) n = lw_task(t).n;
    return col_slice_rs(n);
}
__host__ __device__ constexpr int sw_slice_rs(int band)
{
    int n = 0;
    for (int t = 0; t < COL_NTASK; ++t) if (sw_task(t).band == band && sw_task(t).n > n) n = sw_task(t).n;
    return col_slice_rs(n);
}
// The slices: for task t the rows of its band's table restricted to the task's g-points, [row][slice_rs] (zero padded), one
// contiguous piece of `data` per task -- what one bulk copy (TMA) brings into the shared memory of a block working on task t.
struct ColSlices {
    const double *data;
    int off[COL_NTASK];       // first element of the task's slice (multiple of 16 doubles)
    int bytes[COL_NTASK];     // size of the slice (multiple of 16 bytes)
    int max_bytes;
};

struct LwBand {
    int ng;            // reduced g-points in the band
    int rs;            // row stride of the band table in doubles (ng rounded up to a power of two: a row never
                       // straddles a 128-byte line, so one interpolation term costs one L1 wavefront per distinct row)
    int g0;            // first g-point (0-based) in the 140-vector
    int base;          // element offset of the band table in LwTables::tab
    int sec[LS_COUNT]; // row offsets
    double refrat[5];  // planck_a, planck_b, m_a, m_b, m_a3 (band-specific chi_mls ratios)
};

struct LwConst {
    LwBand band[NBNDLW];
    double preflog[59], tref[59];
    double chi_mls[7 * 59];                 // (7,59) column-major
    double rat_h2oco2[59], rat_h2oo3[59], rat_h2on2o[59], rat_h2och4[59], rat_n2oco2[59], rat_o3co2[59];
    double delwave[NBNDLW];
    double a0[NBNDLW], a1[NBNDLW], a2[NBNDLW];
    double heatfac, fluxfac, oneminus, bpade;
};

struct LwTables {             // device pointers
    const double *tab;        // all band tables, [row][rs] (band bases 128-byte aligned)
    const double *totplnk;    // (181,16) column-major as in the Fortran
    const double *totplnkderiv;   // (181,16): d(totplnk)/dT for idrv = 1
    const double *exptfn;     // interleaved {exp_tbl[i], tfn_tbl[i]}, i = 0..NTBL
    ColSlices sl;             // per-task table slices of the column kernel
};

struct LwIn {                 // interface arrays of the current pass (device pointers, may be offset)
    int ld;                   // leading dimension of the interface arrays (= ncol of the call)
    const double *play, *plev, *tlay, *tlev, *tsfc;
    const double *h2o, *o3, *co2, *ch4, *n2o, *o2, *cfc11, *cfc12, *cfc22, *ccl4;
    const double *emis, *tauaer;  // emis (ld,16); tauaer (ld,nlay,16); either may be null
    // clouds (icld >= 1, inflglw = 0): icld = 1 random overlap (rtrn), 2/3 maximum/random overlap (rtrnmr)
    int icld = 0;
    const double *cldfr = nullptr;    // (ld, nlay)
    const double *taucld = nullptr;   // (16, ld, nlay)
    // cloud optics from water paths (cldprop, inflglw = 1, 2): g/m2 and microns, (ld, nlay)
    int inflg = 0, iceflg = 0, liqflg = 0;
    const double *cicewp = nullptr, *cliqwp = nullptr, *reice = nullptr, *reliq = nullptr;
};

// cloud absorption coefficients of cldprop's parameterisations (lwcldpr, rrtmg_lw_init.f90:2018-2656), Fortran order
struct LwCldConst {
    double abscld1, absliq0, absice0[2], absice1[2 * 5], absice2[43 * 16], absice3[46 * 16], absliq1[58 * 16];
    int have;
};
int lw_upload_cld(const LwCldConst &c);

struct LwOut {
    int ld;
    double *uflx, *dflx, *hr, *uflxc, *dflxc, *hrc;
    double *duflx_dt = nullptr, *duflxc_dt = nullptr;   // (ld, nlay+1), written when idrv == 1
};

constexpr int LF_SLOTS = LF_COUNT + 1;
constexpr int LW_NTASK = COL_NTASK;
constexpr int LW_CSLOT = NGPTLW / 2 + LW_NTASK;   // 16-byte scratch slots per (tile, layer) of lw_column: g-point pairs + one per task

struct LwWork {
    // cloudy sky (null otherwise): band optical depths out of cldprop [col][lay][16], ncbands (1, 5, 16) per column,
    // and the number of the Fortran `stop` some column ran into (0 = none)
    double *taucloud = nullptr;
    int *ncbands = nullptr;
    int *err = nullptr;
    int nc, nlay;
    int idrv;                 // 1: also dF_up/dT_surface (rad.nomcica:143-152)
    double *dplankbnd;        // [col][16]: semiss * d(Planck)/dT at the surface temperature (idrv = 1)
    uint32_t *idx;            // [lay][col]
    int *laytrop;             // [col]
    double *cs_coldry, *cs_wkl1;   // [lay][col]: per-cell terms of the column sums (lw_prep_cell -> lw_prep)
    unsigned char *cs_lower;       // [lay][col]: plog > 4.56
    double *f;                // LF_COUNT fields, each [lay][col]
    double *secdiff;          // [col][16]
    double *planklay;         // [col][lay][16]
    double *planklev;         // [col][lay+1][16]  (level 0 = surface)
    double *plankbnd;         // [col][16]
    double *taug, *fracs;     // [col][lay][140]
    // fused clear-sky path (lw_column.cu): ncp = nc rounded up to whole 32-column tiles; colst = the storage of taug and
    // fracs seen as the scratch field between the two sweeps, [tile][lay][LW_CSLOT][32 lanes] of 16 bytes (per task its optical
    // depths in pairs and one {weight, row offsets} slot for the Planck fractions); part = the g-sums of every
    // task per level, [task][down, up][lay+1][ncp]
    int fused = 0, ncp = 0;
    double *colst = nullptr, *part = nullptr;
    // In the fused path the setcoef state is laid out per (layer, 32-column tile): [lay][tile][LF_SLOTS][32 lanes], slot
    // LF_COUNT holding the packed index word -- one base pointer per layer and immediate offsets for the fields.
    __host__ __device__ size_t tfld(int lay, int col) const
    {
        return (((size_t)lay * (ncp >> 5) + (col >> 5)) * LF_SLOTS) * 32 + (col & 31);
    }
    __host__ __device__ const double *fld(int k) const { return f + (size_t)k * nlay * nc; }
    __host__ __device__ double *fld(int k) { return f + (size_t)k * nlay * nc; }
};

// ------------------------------------------------------------------------------------ SW
// packed: jp:6 | jt:3 | jt1:3 | indself:4 | indfor:2
__host__ __device__ inline uint32_t sw_pack(int jp, int jt, int jt1, int inds, int indf)
{
    return (uint32_t)jp | ((uint32_t)jt << 6) | ((uint32_t)jt1 << 9) | ((uint32_t)inds << 12) | ((uint32_t)indf << 16);
}

enum SwField {
    SF_FAC00, SF_FAC01, SF_FAC10, SF_FAC11,
    SF_COLH2O, SF_COLCO2, SF_COLO3, SF_COLCH4, SF_COLO2, SF_COLMOL, SF_COLN2O,
    SF_SELFFAC, SF_SELFFRAC, SF_FORFAC, SF_FORFRAC,
    SF_COUNT
};

enum SwSec {
    SS_ABSA, SS_ABSB, SS_SELF, SS_FOR, SS_SFLUX, SS_RAYL, SS_RAYLB, SS_X1, SS_X2,
    SS_COUNT
};

constexpr int SF_SLOTS = SF_COUNT + 1;
constexpr int SW_NTASK = COL_NTASK;
constexpr int SW_NSLOT = 3 * NGPTSW + SW_NTASK;   // scratch slots per (tile, layer)

struct SwBand {
    int ng, rs, g0, base;   // rs: padded row stride, see LwBand
    int sec[SS_COUNT];
    int nsflux;        // number of sfluxref rows (1, 5 or 9)
    int nrayl;         // number of rayl rows (1 or 9)
};

struct SwConst {
    SwBand band[NBNDSW];
    double preflog[59], tref[59];
    double heatfac, oneminus, bpade;
    double rsrtaua[14][6], rsrpiza[14][6], rsrasya[14][6];   // ECMWF aerosol types (iaer = 6), rrtmg_sw_init.f90:370-470
    int have_aer;                                            // the three tables above were found
};

struct SwTables {
    const double *tab;
    const double *exptbl;     // interleaved {exp_tbl[i], 1/exp_tbl[i]}, i = 0..NTBL
    ColSlices sl;             // per-task table slices of the column kernel
};

struct SwIn {
    int ld;
    const double *play, *plev, *tlay, *tlev, *tsfc;
    const double *h2o, *o3, *co2, *ch4, *n2o, *o2;
    const double *asdir, *asdif, *aldir, *aldif, *coszen;
    double adjflux;           // adjflx * scon / rrsw_scon (same for all bands, rad.nomcica:953-972)
    // optional branches of the interface (null / 0 for MiMA's configuration)
    int icld = 0, iaer = 0;
    const double *cldfr = nullptr;                                                  // (ld, nlay)
    const double *taucld = nullptr, *ssacld = nullptr, *asmcld = nullptr, *fsfcld = nullptr;   // (14, ld, nlay), inflgsw = 0
    const double *tauaer = nullptr, *ssaaer = nullptr, *asmaer = nullptr;           // (ld, nlay, 14), iaer = 10
    const double *ecaer = nullptr;                                                  // (ld, nlay, 6), iaer = 6
    // cloud optics from water paths (cldprop_sw, inflgsw = 2): g/m2 and microns, (ld, nlay)
    int inflg = 0, iceflg = 0, liqflg = 0;
    const double *cicewp = nullptr, *cliqwp = nullptr, *reice = nullptr, *reliq = nullptr;
};

// cloud optical properties of cldprop_sw's parameterisations (swcldpr, rrtmg_sw_init.f90:1519-3341), Fortran order
// (radius index, band 16..29)
struct SwCldConst {
    double extliq1[58 * 14], ssaliq1[58 * 14], asyliq1[58 * 14];
    double extice2[43 * 14], ssaice2[43 * 14], asyice2[43 * 14];
    double extice3[46 * 14], ssaice3[46 * 14], asyice3[46 * 14], fdlice3[46 * 14];
    double abari[5], bbari[5], cbari[5], dbari[5], ebari[5], fbari[5];
    int have;
};
int sw_upload_cld(const SwCldConst &c);

struct SwOut {
    int ld;
    double *uflx, *dflx, *hr, *uflxc, *dflxc, *hrc;
};

struct SwWork {
    int nc, nlay;
    uint32_t *idx;            // [lay][col]
    int *laytrop;             // [col]
    unsigned char *cs_jp;     // [lay][col]: jp | 0x80 * (plog > 4.56)  (sw_prep_cell -> sw_prep)
    int *laysolfr;            // [col][14]: layer (1-based) whose eta selects the solar source; 0 = never written
    double *f;                // SF_COUNT fields, each [lay][col]
    double *taug;             // [col][lay][112]
    double *colmol;           // [col][lay]: taur(g) = colmol * rayl(g) for every band but 24 (evaluated by the solver)
    double *taur24;           // [col][lay][8]: taur of band 24, whose Rayleigh coefficient depends on the cell
    double *taur;             // [col][lay][112], expanded from rdesc only for the stage-capture test hook
    double *sfluxzen;         // [col][112]
    double *part;             // [col][7 half-warps][up, down][lev]: g-point partial sums (sw_solver_warp -> sw_finish)
    double *stack;            // [slot][lev][rdnd, zp, zq][lane]: per-cell stack of sw_solver_l2_kernel, one slot per resident warp
    // general path (icld >= 1 or iaer = 10): per (column, layer, band) {tauc, omgc, asyc (delta-M scaled, cldprop_sw
    // inflag = 0), taua, omga, asya} and the layer cloud fraction; err[0]: bit 0 = partial cloud found, err[1]: number of
    // the cldprop_sw `stop` some cell ran into
    double *opt;              // [col][lay][14][6] or null
    double *clfr;             // [col][lay]
    int *err;
    __host__ __device__ const double *fld(int k) const { return f + (size_t)k * nlay * nc; }
    __host__ __device__ double *fld(int k) { return f + (size_t)k * nlay * nc; }
    // fused clear-sky path (sw_column.cu): ncp = nc rounded up to whole 32-column tiles; tf = the setcoef state per (layer,
    // tile), [lay][tile][SF_SLOTS][32 lanes] with the packed index word in slot SF_COUNT; colst = per (tile, layer) the
    // {zp, zq, rdnd} of every g-point and one downward sum per task, [tile][lay][SW_NSLOT][32 lanes]; cpart = the g-sums of
    // every task per level, [task][up, down][lay+1][ncp]
    int fused = 0, ncp = 0;
    double *tf = nullptr, *colst = nullptr, *cpart = nullptr;
    __host__ __device__ size_t tfld(int lay, int col) const
    {
        return (((size_t)lay * (ncp >> 5) + (col >> 5)) * SF_SLOTS) * 32 + (col & 31);
    }
};


// ------------------------------------------------------------------------------------ radiation driver (driver.cu)
struct RadGeom {
    int si, sj, sk;           // GCM grid of the rank: lon, lat, levels
    int ls;                   // lonstep
    int ni;                   // si / lonstep
    int ncols;                // ni * sj = ncols_rrt
};
struct ZenithArgs {           // scalars of compute_zenith (astro.f90:95-127), evaluated on the host
    double radsec, dt_pi, radpersec, dec_sin, dec_cos, dec_tan;
    int dt;
};
struct PackArgs {
    // GCM state (si, sj, sk) / (si, sj, sk+1) / (si, sj); level 1 = top; Pa, K, kg/kg
    const double *p_full, *p_half, *t, *t_half, *q, *o3f, *coszen, *albedo, *t_surf, *lat;
    const double *qzm;        // zonal-mean q (sj, sk) or null
    const int *top_flag;
    // RRTMG inputs (ncols, sk) / (ncols, sk+1) / (ncols); level 1 = surface; hPa
    double *pfull, *phalf, *tfull, *thalf, *h2o, *o3, *cosz_rr, *albedo_rr, *tsrf;
    double qmin, tmin, tmax, scale_ozone, o3_val;
    int do_fixed_water;
    double fixed_water, fixed_water_pres, fixed_water_lat;
};
struct UnpackArgs {
    const double *swhr, *swuflx, *swdflx, *lwhr, *lwuflx, *lwdflx;   // RRTMG outputs
    double *tdt, *tdt_rad, *tdt_sw, *tdt_lw;                           // (si, sj, sk); any may be null
    double *flux_sw, *flux_lw, *olr, *isr;                             // (si, sj); any may be null
};
int drv_zenith(int n, const double *lat, const double *lon, double *cosz, const ZenithArgs &a, cudaStream_t s);
int drv_interp_temp(int np, int sk, const double *z_full, const double *z_half, const double *t_surf, const double *t,
                    double *t_half, cudaStream_t s);
int drv_pack(const RadGeom &g, const PackArgs &a, double *qzm_buf, int *flag, cudaStream_t s, int top_flag = -1);
int drv_fill(double *p, size_t n, double v, cudaStream_t s);
int drv_unpack(const RadGeom &g, const UnpackArgs &a, double *zm_buf, cudaStream_t s);

// ------------------------------------------------------------------------------------ device math
#ifdef __CUDACC__
// FP64 reciprocal and square root from the 20-bit MUFU seeds plus Newton steps written as fma().  They
// replace the IEEE-rounded `/` and sqrt() sequences (which carry a slow-path call) in the solvers, where
// the divide count bounds the FP64 pipe.  Relative error <= ~2 ulp; arguments here are normal, positive
// or bounded away from zero by the callers, so no special-case handling is needed.
__device__ __forceinline__ double rcp_fast(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    const double t = fma(e, e, e);
    return fma(y, t, y);
}
__device__ __forceinline__ double sqrt_fast(double a)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double g = a * y, h = 0.5 * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    return fma(g, r, g);
}

// One 16-byte gather from the {exp, tfn} / {exp, 1/exp} look-up table of a solver (read-only path).  An L1 evict_last hint
// on these loads (LDG.E.EL.128.CONSTANT) was measured: no change in either solver (profiles/r02_summary.md).
__device__ __forceinline__ double2 ld_tbl(const double2 *__restrict__ p) { return __ldg(p); }

// ---- mbarrier and 1-D bulk-copy (TMA) helpers: the staging ring of lw_rtrn_tma_kernel and the shared-memory look-up tables of
// the column kernels
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Stage `bytes` (a multiple of 16) of global memory into the block's shared memory: thread 0 issues 1-D bulk copies (TMA) that
// complete on one mbarrier, every thread waits for it.  Returns the shared-memory address of the copy, passed through an opaque
// asm behind the wait so that no load from the copy (non-volatile ld.shared asm) can be scheduled in front of it.
__device__ __forceinline__ uint32_t stage_to_shared(void *dst, uint64_t *bar, const void *src, uint32_t bytes)
{
    constexpr uint32_t CH = 32768;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, bytes);
        for (uint32_t o = 0; o < bytes; o += CH)
            tma_load_1d(static_cast<unsigned char *>(dst) + o, static_cast<const unsigned char *>(src) + o, min(CH, bytes - o), bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);
    uint32_t a = smem_u32(dst);
    asm volatile("" : "+r"(a) :: "memory");
    return a;
}

// Sum over g-points of R rows (R = 16 or 32) of per-thread values that the caller has stored in `tile` as
// tile[row * S + g] (S odd, >= NACT).  Stage 1: thread t sums the strided elements of row (t % R) --
// conflict-free because the lanes of a half-warp hold different rows and S is odd; stage 2: threads 0..R-1
// combine the NT/R partials in a fixed order, so the result is bitwise reproducible.  Two barriers; `tile`
// may be refilled right after the call.  Returns the row sum in threads 0..R-1 (row = threadIdx.x).
template <int NT, int R, int NACT, int S>
__device__ __forceinline__ double tile_reduce(const double *tile, double *part)
{
    static_assert(NT % R == 0 && (R & (R - 1)) == 0, "threads must be a multiple of the row count");
    constexpr int P = NT / R;
    const int t = threadIdx.x, row = t & (R - 1), p = t / R;
    __syncthreads();
    double acc = 0.0;
    const double *src = tile + row * S;
#pragma unroll
    for (int j = p; j < NACT; j += P) acc += src[j];
    part[row * (P + 1) + p] = acc;
    __syncthreads();
    double sum = 0.0;
    if (t < R) {
#pragma unroll
        for (int k = 0; k < P; ++k) sum += part[t * (P + 1) + k];
    }
    return sum;
}
template <int NT, int NACT, int S>
__device__ __forceinline__ double tile_reduce16(const double *tile, double *part)
{
    return tile_reduce<NT, 16, NACT, S>(tile, part);
}
#endif

// optional per-kernel CUDA-event timing (api.cu); ids: 0 lw_prep, 1 lw_taumol, 2 lw_rtrn, 3 sw_prep, 4 sw_taumol, 5 sw_solver
enum KernelId { K_LW_PREP, K_LW_TAUMOL, K_LW_RTRN, K_SW_PREP, K_SW_TAUMOL, K_SW_SOLVER, K_LW_COLUMN, K_SW_COLUMN, K_COUNT };
void ktimer_begin(int id, cudaStream_t s);
void ktimer_end(cudaStream_t s);

// launch tuning (api.cu; option keys "lw_rtrn_pad_kb", "sw_solver_pad_kb", "sw_solver_store", "sw_solver_variant"): extra dynamic shared memory per block,
// used to cap the resident blocks per SM so that the sweeps' per-thread state stays L2-resident
struct Tuning {
    int lw_rtrn_pad_kb, sw_solver_pad_kb, sw_solver_store, sw_solver_variant, lw_rtrn_variant, taumol_sync;
    int x[8];                 // experiment knobs ("x0".."x7")
    int lw_fused;             // 1 (default): clear-sky LW without derivatives runs the fused column kernel (lw_column.cu)
    int sw_fused;             // 1 (default): SW without clouds and aerosols runs the fused column kernel (sw_column.cu)
    int col_warps;            // block shape of the fused kernels: 0 = default (LW two 8-warp blocks per SM, SW one 16-warp block), 8 / 16 = forced
};
// columns per super-group of the column kernels' block order (T170L60 step in one pass / in 16384-column passes / T42L40:
// 2048: 23.6 / 26.6 / 1.33 ms, 4096: 23.7 / 25.9 / 1.31, 8192: 24.2 / 26.2 / 1.26; task-fastest order as before: 23.7 / 28.6 / 1.45)
constexpr int COL_SUPER_COLS = 4096;
extern Tuning g_tune;

// solver translation units (lw_solver.cu / sw_solver.cu, compiled with FMA contraction on; see build.py)
int lw_solver_upload_const(const LwConst &c, const unsigned char *ngb);
int sw_solver_upload_const(const SwConst &c, const unsigned char *ngb);
int lw_launch_rtrn(const LwTables &t, const LwIn &in, const LwOut &out, LwWork &w, cudaStream_t s);   // returns the launch count
int lw_column_upload_const(const LwConst &c);
int sw_column_upload_const(const SwConst &c);
int sw_launch_column(const SwTables &t, const SwIn &in, const SwOut &out, SwWork &w, cudaStream_t s);   // fused clear sky; returns the launch count
int lw_launch_column(const LwTables &t, const LwIn &in, const LwOut &out, LwWork &w, cudaStream_t s);   // fused clear sky; returns the launch count
int sw_launch_solver(const SwTables &t, const SwIn &in, const SwOut &out, SwWork &w, cudaStream_t s);   // returns the launch count

// kernel launchers (defined in lw_kernels.cu / sw_kernels.cu); each returns the number of launches
int lw_upload_const(const LwConst &c);
int sw_upload_const(const SwConst &c);
// `cap` (optional): device buffer of 2*nc*nlay*140 doubles receiving a copy of taug|fracs before the solver
// overwrites them in place (test hook).
int lw_run_pass(const LwTables &t, const LwIn &in, const LwOut &out, LwWork &w, cudaStream_t s, double *cap);
int sw_run_pass(const SwTables &t, const SwIn &in, const SwOut &out, SwWork &w, cudaStream_t s);

} // namespace rrtmg
