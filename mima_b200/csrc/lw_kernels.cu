// lw_kernels.cu -- RRTMG longwave on sm_100a: prep (inatm + setcoef), taumol, cldprop.  The solver is in lw_solver.cu.
//
// What is computed follows the reference routines (cited per kernel); how it is computed is a GPU-first design:
//   lw_prep_cell_kernel  thread <-> (column, layer): unit conversion, column amounts, Planck sources of the layer and its
//                        upper interface, the per-cell terms of the column sums;
//   lw_prep_kernel       thread <-> column: the column sums in the reference's order (laytrop, precipitable water ->
//                        diffusivity secant), surface Planck terms;
//   lw_taumol_kernel     thread <-> (column, layer) cell, lanes = 32 adjacent columns of one layer; the cell's setcoef state
//                        is evaluated in place (lw_cell), every term w * T[row][.] of a band formula is consumed at once
//                        into ng register accumulators, and the warps of a block walk the 16 bands together (one block
//                        barrier per four bands) so that the ~300 KB of straight-line band code is fetched once per block;
//   lw_cldprop_kernel    thread <-> column (cloudy sky only).
// Compiled with -fmad=false: fused multiply-adds appear only where written as fma().
#include "rrtmg_dev.cuh"

namespace rrtmg {

__constant__ LwConst c_lw;
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
    return lw_solver_upload_const(c, ngb);
}

#define CHI(m, j) c_lw.chi_mls[((j) - 1) * 7 + ((m) - 1)]

// Interpolation state of one (column, layer) cell: everything setcoef hands to taumol.
struct LwPair {
    int jp, jt, jt1, inds, indf, indm;
    double fac00, fac01, fac10, fac11;
    double colh2o, colco2, colo3, coln2o, colco, colch4, colo2, colbrd;
    double selffac, selffrac, forfac, forfrac, minorfrac, scaleminor, scaleminorn2, coldry, pavel;
    double wx1, wx2, wx3, wx4;
};

// =====================================================================================================
// inatm (LW/src/rrtmg_lw_rad.nomcica.f90:572-901) + setcoef (LW/src/rrtmg_lw_setcoef.f90:251-410) for one
// (column, layer) cell.  Everything here is local to the cell (coldry needs only the two interface
// pressures of the layer), so the taumol kernel evaluates it in place instead of reading it back from HBM;
// the prep kernel calls the same function for the column-integrated quantities (laytrop, pwvcm), which
// keeps the two bit-identical.  Returns true when the layer counts towards laytrop (plog > 4.56).
// `wkl1` (H2O column amount, molecules/cm2) is returned for the precipitable-water integral.
// =====================================================================================================
__device__ __forceinline__ bool lw_cell(const LwIn &in, int col, int l, LwPair &p, double &wkl1_out)
{
    const double amd = 28.9660, amw = 18.0160, amdw = 1.607793, amdo = 0.603428;
    const double grav = 9.8066, avogad = 6.02214199e+23;
    const double stpfac = 296. / 1013.;
    const size_t ld = (size_t)in.ld;
    const size_t o = col + (size_t)l * ld;
    const double pavel = in.play[o], tavel = in.tlay[o];
    const double pzm = in.plev[o], pz = in.plev[o + ld];
    // ---- inatm
    const double q = in.h2o[o];
    double wkl1 = (q / (1.0 - q)) * amdw;
    double wkl2 = in.co2[o];
    double wkl3 = in.o3[o] * amdo;
    double wkl4 = in.n2o ? in.n2o[o] : 0.0;
    double wkl6 = in.ch4 ? in.ch4[o] : 0.0;
    double wkl7 = in.o2 ? in.o2[o] : 0.0;
    const double amm = (1.0 - wkl1) * amd + wkl1 * amw;
    const double coldry = (pzm - pz) * 1.e3 * avogad / (1.e2 * grav * amm * (1.0 + wkl1));
    double summol = 0.0;
    summol = summol + wkl2; summol = summol + wkl3; summol = summol + wkl4;
    summol = summol + 0.0;  summol = summol + wkl6; summol = summol + wkl7;
    const double wbrodl = coldry * (1.0 - summol);
    wkl1 = coldry * wkl1; wkl2 = coldry * wkl2; wkl3 = coldry * wkl3; wkl4 = coldry * wkl4;
    wkl6 = coldry * wkl6; wkl7 = coldry * wkl7;
    const double wkl5 = coldry * 0.0;
    wkl1_out = wkl1;
    p.wx1 = in.ccl4 ? coldry * in.ccl4[o] * 1.e-20 : 0.0;
    p.wx2 = in.cfc11 ? coldry * in.cfc11[o] * 1.e-20 : 0.0;
    p.wx3 = in.cfc12 ? coldry * in.cfc12[o] * 1.e-20 : 0.0;
    p.wx4 = in.cfc22 ? coldry * in.cfc22[o] * 1.e-20 : 0.0;

    // ---- setcoef: interpolation indices and factors
    const double plog = log(pavel);
    int jp = (int)(36. - 5 * (plog + 0.04));
    jp = jp < 1 ? 1 : (jp > 58 ? 58 : jp);
    const double fp = 5. * (c_lw.preflog[jp - 1] - plog);
    const double tr0 = (tavel - c_lw.tref[jp - 1]) / 15.;
    int jt = (int)(3. + tr0);
    jt = jt < 1 ? 1 : (jt > 4 ? 4 : jt);
    const double ft = tr0 - (double)(jt - 3);
    const double tr1 = (tavel - c_lw.tref[jp]) / 15.;
    int jt1 = (int)(3. + tr1);
    jt1 = jt1 < 1 ? 1 : (jt1 > 4 ? 4 : jt1);
    const double ft1 = tr1 - (double)(jt1 - 3);
    const double water = wkl1 / coldry;
    const double scalefac = pavel * stpfac / tavel;
    double forfac, forfrac, selffac, selffrac = 0.0, factor;
    int indfor, indself = 0;
    forfac = scalefac / (1. + water);
    selffac = water * forfac;
    const bool lower = !(plog <= 4.56);
    if (lower) {
        factor = (332.0 - tavel) / 36.0;
        indfor = (int)factor;
        indfor = indfor < 1 ? 1 : (indfor > 2 ? 2 : indfor);
        forfrac = factor - (double)indfor;
        factor = (tavel - 188.0) / 7.2;
        indself = (int)factor - 7;
        indself = indself < 1 ? 1 : (indself > 9 ? 9 : indself);
        selffrac = factor - (double)(indself + 7);
    } else {
        factor = (tavel - 188.0) / 36.0;
        indfor = 3;
        forfrac = factor - 1.0;
    }
    p.scaleminor = pavel / tavel;
    p.scaleminorn2 = (pavel / tavel) * (wbrodl / (coldry + wkl1));
    factor = (tavel - 180.8) / 7.2;
    int indminor = (int)factor;
    indminor = indminor < 1 ? 1 : (indminor > 18 ? 18 : indminor);
    p.minorfrac = factor - (double)indminor;

    p.colh2o = 1.e-20 * wkl1;
    double colco2 = 1.e-20 * wkl2, colo3 = 1.e-20 * wkl3, coln2o = 1.e-20 * wkl4;
    double colco = 1.e-20 * wkl5, colch4 = 1.e-20 * wkl6;
    p.colo2 = 1.e-20 * wkl7;
    if (colco2 == 0.) colco2 = 1.e-32 * coldry;
    if (colo3 == 0.) colo3 = 1.e-32 * coldry;
    if (coln2o == 0.) coln2o = 1.e-32 * coldry;
    if (colco == 0.) colco = 1.e-32 * coldry;
    if (colch4 == 0.) colch4 = 1.e-32 * coldry;
    p.colco2 = colco2; p.colo3 = colo3; p.coln2o = coln2o; p.colco = colco; p.colch4 = colch4;
    p.colbrd = 1.e-20 * wbrodl;
    const double compfp = 1. - fp;
    p.jp = jp; p.jt = jt; p.jt1 = jt1; p.inds = indself; p.indf = indfor; p.indm = indminor;
    p.fac10 = compfp * ft;
    p.fac00 = compfp * (1. - ft);
    p.fac11 = fp * ft1;
    p.fac01 = fp * (1. - ft1);
    p.selffac = p.colh2o * selffac;
    p.selffrac = selffrac;
    p.forfac = p.colh2o * forfac;
    p.forfrac = forfrac;
    p.coldry = coldry;
    p.pavel = pavel;
    return lower;
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
    if (w.f) {
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
            pl0[ib] = __ldg(tp + indlev0 - 1) + t0frac * dbdtlev;
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

struct Eta { double speccomb, specparm, fs; int js; };
// eta = colA/(colA + rat*colB), clamped to oneminus; js = 1+int(mult*eta); fs = mod(mult*eta, 1)
__device__ __forceinline__ Eta binary(double colA, double rat, double colB, double mult)
{
    Eta e;
    e.speccomb = colA + rat * colB;
    e.specparm = colA / e.speccomb;
    if (e.specparm >= c_lw.oneminus) e.specparm = c_lw.oneminus;
    const double specmult = mult * e.specparm;
    const int i = (int)specmult;
    e.js = 1 + i;
    e.fs = specmult - (double)i;
    return e;
}

// rows are Fortran 1-based; `sec` is the section's first row
template <class PW>
__device__ __forceinline__ void key4(PW &pw, const LwBand &B, int sec, int ind0, int ind1, double scale, const LwPair &p)
{
    const int ng = B.rs, o0 = (B.sec[sec] + ind0 - 1) * ng, o1 = (B.sec[sec] + ind1 - 1) * ng;
    pw.add(o0, scale * p.fac00);
    pw.add(o0 + ng, scale * p.fac10);
    pw.add(o1, scale * p.fac01);
    pw.add(o1 + ng, scale * p.fac11);
}
template <class PW>
__device__ __forceinline__ void lerp2(PW &pw, const LwBand &B, int sec, int row, double frac, double scale)
{
    const int ng = B.rs, o = (B.sec[sec] + row - 1) * ng;
    pw.add(o, scale * (1. - frac));
    pw.add(o + ng, scale * frac);
}
// minor gas with eta dimension, Fortran (neta,19,ng): 4-point (eta, T) interpolation
template <class PW>
__device__ __forceinline__ void minor_eta(PW &pw, const LwBand &B, int sec, int neta, int jm, double fm, int indm,
                                          double mf, double scale)
{
    const int ng = B.rs, o = (B.sec[sec] + (indm - 1) * neta + (jm - 1)) * ng;
    pw.add(o, scale * ((1. - mf) * (1. - fm)));
    pw.add(o + ng, scale * ((1. - mf) * fm));
    pw.add(o + neta * ng, scale * (mf * (1. - fm)));
    pw.add(o + (neta + 1) * ng, scale * (mf * fm));
}
// lower-atmosphere binary-species key term: 3-point stencil near eta = 0 / 1, else 2-point
// (template block repeated in taugb3,4,5,7,9,12,13,15,16, e.g. taumol.f90:548-606)
template <class PW>
__device__ __forceinline__ void stencil_lower(PW &pw, const LwBand &B, int ind, const Eta &e, double facA, double facB)
{
    // One instruction stream for the three cases: rows (o, o+1[, o+2]) and (o+9, o+10[, o+11]) with
    //   eta < 0.125 : o = ind,     weights (fk0, fk1, fk2)
    //   eta > 0.875 : o = ind - 1, weights (fk2, fk1, fk0)
    //   otherwise   : o = ind,     weights (1-fs, fs)
    const int ng = B.rs;
    const double sc = e.speccomb;
    const bool lo = e.specparm < 0.125, hi = e.specparm > 0.875;
    const double p = lo ? e.fs - 1 : -e.fs, p4 = (p * p) * (p * p);
    const double fk0 = p4, fk1 = 1 - p - 2.0 * p4, fk2 = p + p4;
    const double w0 = lo ? fk0 : (hi ? fk2 : 1. - e.fs);
    const double w1 = (lo || hi) ? fk1 : e.fs;
    const double w2 = lo ? fk2 : fk0;
    const int o = (B.sec[LS_ABSA] + ind - 1 - (hi ? 1 : 0)) * ng;
    pw.add(o, sc * (w0 * facA));
    pw.add(o + ng, sc * (w1 * facA));
    pw.add(o + 9 * ng, sc * (w0 * facB));
    pw.add(o + 10 * ng, sc * (w1 * facB));
    if (lo || hi) {
        pw.add(o + 2 * ng, sc * (w2 * facA));
        pw.add(o + 11 * ng, sc * (w2 * facB));
    }
}
// upper-atmosphere binary key term (nspb = 5): always 2-point (e.g. taumol.f90:739-750)
template <class PW>
__device__ __forceinline__ void stencil_upper(PW &pw, const LwBand &B, int ind, const Eta &e, double facA, double facB)
{
    const int ng = B.rs, o = (B.sec[LS_ABSB] + ind - 1) * ng;
    const double sc = e.speccomb;
    pw.add(o, sc * ((1. - e.fs) * facA));
    pw.add(o + ng, sc * (e.fs * facA));
    pw.add(o + 5 * ng, sc * ((1. - e.fs) * facB));
    pw.add(o + 6 * ng, sc * (e.fs * facB));
}
template <class PW>
__device__ __forceinline__ void frac_const(PW &pw, const LwBand &B, int sec) { pw.frac1(B.sec[sec] * B.rs); }
template <class PW>
__device__ __forceinline__ void frac_eta(PW &pw, const LwBand &B, int sec, double colA, double refrat, double colB, double mult)
{
    const Eta e = binary(colA, refrat, colB, mult);
    const int o = (B.sec[sec] + e.js - 1) * B.rs;
    pw.frac2(o, 1. - e.fs, o + B.rs, e.fs);
}
// high-CO2 / high-N2O column adjustment (e.g. taumol.f90:529-535)
__device__ __forceinline__ double adjcol(double col, double coldry, double chiref, double thresh, double a, double ex)
{
    const double chi = col / coldry;
    const double rat = 1.e20 * chi / chiref;
    if (rat > thresh) {
        const double adjfac = a + pow(rat - a, ex);
        return adjfac * chiref * coldry * 1.e-20;
    }
    return col;
}

#define IND0A(nsp) (((p.jp - 1) * 5 + (p.jt - 1)) * (nsp))
#define IND1A(nsp) ((p.jp * 5 + (p.jt1 - 1)) * (nsp))
#define IND0B(nsp) (((p.jp - 13) * 5 + (p.jt - 1)) * (nsp))
#define IND1B(nsp) (((p.jp - 12) * 5 + (p.jt1 - 1)) * (nsp))

__host__ __device__ constexpr int lw_ng(int band)
{
    constexpr int ng[16] = {10, 12, 16, 14, 16, 8, 12, 8, 12, 6, 8, 8, 4, 2, 2, 2};
    return ng[band];
}

template <int BAND, class PW>
__device__ __forceinline__ void lw_band_terms(const LwPair &p, bool lower, PW &pw)
{
    const LwBand &B = c_lw.band[BAND];
    if constexpr (BAND == 0) { // band 1: 10-350 cm-1, H2O; N2 continuum minor (:280-373)
        const double scalen2 = p.colbrd * p.scaleminorn2;
        if (lower) {
            double corradj = 1.;
            if (p.pavel < 250.) corradj = 1. - 0.15 * (250. - p.pavel) / 154.4;
            key4(pw, B, LS_ABSA, IND0A(1) + 1, IND1A(1) + 1, corradj * p.colh2o, p);
            lerp2(pw, B, LS_SELF, p.inds, p.selffrac, corradj * p.selffac);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, corradj * p.forfac);
            lerp2(pw, B, LS_MA1, p.indm, p.minorfrac, corradj * scalen2);
            frac_const(pw, B, LS_FRACA);
        } else {
            const double corradj = 1. - 0.15 * (p.pavel / 95.6);
            key4(pw, B, LS_ABSB, IND0B(1) + 1, IND1B(1) + 1, corradj * p.colh2o, p);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, corradj * p.forfac);
            lerp2(pw, B, LS_MB1, p.indm, p.minorfrac, corradj * scalen2);
            frac_const(pw, B, LS_FRACB);
        }
    } else if constexpr (BAND == 1) { // band 2: 350-500, H2O (:376-445)
        if (lower) {
            const double corradj = 1. - .05 * (p.pavel - 100.) / 900.;
            key4(pw, B, LS_ABSA, IND0A(1) + 1, IND1A(1) + 1, corradj * p.colh2o, p);
            lerp2(pw, B, LS_SELF, p.inds, p.selffrac, corradj * p.selffac);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, corradj * p.forfac);
            frac_const(pw, B, LS_FRACA);
        } else {
            key4(pw, B, LS_ABSB, IND0B(1) + 1, IND1B(1) + 1, p.colh2o, p);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            frac_const(pw, B, LS_FRACB);
        }
    } else if constexpr (BAND == 2) { // band 3: 500-630, H2O/CO2 both regions; N2O minor (:448-760)
        const double chin2o = CHI(4, p.jp + 1);
        const double adj = adjcol(p.coln2o, p.coldry, chin2o, 1.5, 0.5, 0.65);
        if (lower) {
            const Eta e0 = binary(p.colh2o, c_lw.rat_h2oco2[p.jp - 1], p.colco2, 8.);
            const Eta e1 = binary(p.colh2o, c_lw.rat_h2oco2[p.jp], p.colco2, 8.);
            const Eta em = binary(p.colh2o, B.refrat[2], p.colco2, 8.);
            stencil_lower(pw, B, IND0A(9) + e0.js, e0, p.fac00, p.fac10);
            stencil_lower(pw, B, IND1A(9) + e1.js, e1, p.fac01, p.fac11);
            lerp2(pw, B, LS_SELF, p.inds, p.selffrac, p.selffac);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            minor_eta(pw, B, LS_MA1, 9, em.js, em.fs, p.indm, p.minorfrac, adj);
            frac_eta(pw, B, LS_FRACA, p.colh2o, B.refrat[0], p.colco2, 8.);
        } else {
            const Eta e0 = binary(p.colh2o, c_lw.rat_h2oco2[p.jp - 1], p.colco2, 4.);
            const Eta e1 = binary(p.colh2o, c_lw.rat_h2oco2[p.jp], p.colco2, 4.);
            const Eta em = binary(p.colh2o, B.refrat[3], p.colco2, 4.);
            stencil_upper(pw, B, IND0B(5) + e0.js, e0, p.fac00, p.fac10);
            stencil_upper(pw, B, IND1B(5) + e1.js, e1, p.fac01, p.fac11);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            minor_eta(pw, B, LS_MB1, 5, em.js, em.fs, p.indm, p.minorfrac, adj);
            frac_eta(pw, B, LS_FRACB, p.colh2o, B.refrat[1], p.colco2, 4.);
        }
    } else if constexpr (BAND == 3) { // band 4: 630-700, H2O/CO2 lower, O3/CO2 upper (:763-1019)
        if (lower) {
            const Eta e0 = binary(p.colh2o, c_lw.rat_h2oco2[p.jp - 1], p.colco2, 8.);
            const Eta e1 = binary(p.colh2o, c_lw.rat_h2oco2[p.jp], p.colco2, 8.);
            stencil_lower(pw, B, IND0A(9) + e0.js, e0, p.fac00, p.fac10);
            stencil_lower(pw, B, IND1A(9) + e1.js, e1, p.fac01, p.fac11);
            lerp2(pw, B, LS_SELF, p.inds, p.selffrac, p.selffac);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            frac_eta(pw, B, LS_FRACA, p.colh2o, B.refrat[0], p.colco2, 8.);
        } else {
            const Eta e0 = binary(p.colo3, c_lw.rat_o3co2[p.jp - 1], p.colco2, 4.);
            const Eta e1 = binary(p.colo3, c_lw.rat_o3co2[p.jp], p.colco2, 4.);
            stencil_upper(pw, B, IND0B(5) + e0.js, e0, p.fac00, p.fac10);
            stencil_upper(pw, B, IND1B(5) + e1.js, e1, p.fac01, p.fac11);
            frac_eta(pw, B, LS_FRACB, p.colo3, B.refrat[1], p.colco2, 4.);
            pw.scale(B.sec[LS_GSCALE] * B.rs);   // stratospheric g-point scaling (:1009-1015)
        }
    } else if constexpr (BAND == 4) { // band 5: 700-820, H2O/CO2 lower, O3/CO2 upper; O3 minor, CCl4 (:1022-1294)
        if (lower) {
            const Eta e0 = binary(p.colh2o, c_lw.rat_h2oco2[p.jp - 1], p.colco2, 8.);
            const Eta e1 = binary(p.colh2o, c_lw.rat_h2oco2[p.jp], p.colco2, 8.);
            const Eta em = binary(p.colh2o, B.refrat[2], p.colco2, 8.);
            stencil_lower(pw, B, IND0A(9) + e0.js, e0, p.fac00, p.fac10);
            stencil_lower(pw, B, IND1A(9) + e1.js, e1, p.fac01, p.fac11);
            lerp2(pw, B, LS_SELF, p.inds, p.selffrac, p.selffac);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            minor_eta(pw, B, LS_MA1, 9, em.js, em.fs, p.indm, p.minorfrac, p.colo3);
            pw.add_nz(B.sec[LS_X1] * B.rs, p.wx1);
            frac_eta(pw, B, LS_FRACA, p.colh2o, B.refrat[0], p.colco2, 8.);
        } else {
            const Eta e0 = binary(p.colo3, c_lw.rat_o3co2[p.jp - 1], p.colco2, 4.);
            const Eta e1 = binary(p.colo3, c_lw.rat_o3co2[p.jp], p.colco2, 4.);
            stencil_upper(pw, B, IND0B(5) + e0.js, e0, p.fac00, p.fac10);
            stencil_upper(pw, B, IND1B(5) + e1.js, e1, p.fac01, p.fac11);
            pw.add_nz(B.sec[LS_X1] * B.rs, p.wx1);
            frac_eta(pw, B, LS_FRACB, p.colo3, B.refrat[1], p.colco2, 4.);
        }
    } else if constexpr (BAND == 5) { // band 6: 820-980, H2O lower; CO2 minor, CFC11, CFC12 (:1297-1380)
        if (lower) {
            const double adj = adjcol(p.colco2, p.coldry, CHI(2, p.jp + 1), 3.0, 2.0, 0.77);
            key4(pw, B, LS_ABSA, IND0A(1) + 1, IND1A(1) + 1, p.colh2o, p);
            lerp2(pw, B, LS_SELF, p.inds, p.selffrac, p.selffac);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            lerp2(pw, B, LS_MA1, p.indm, p.minorfrac, adj);
        }
        pw.add_nz(B.sec[LS_X1] * B.rs, p.wx2);
        pw.add_nz(B.sec[LS_X2] * B.rs, p.wx3);
        frac_const(pw, B, LS_FRACA);
    } else if constexpr (BAND == 6) { // band 7: 980-1080, H2O/O3 lower, O3 upper; CO2 minor (:1383-1654)
        if (lower) {
            const double adj = adjcol(p.colco2, p.coldry, CHI(2, p.jp + 1), 3.0, 3.0, 0.79);
            const Eta e0 = binary(p.colh2o, c_lw.rat_h2oo3[p.jp - 1], p.colo3, 8.);
            const Eta e1 = binary(p.colh2o, c_lw.rat_h2oo3[p.jp], p.colo3, 8.);
            const Eta em = binary(p.colh2o, B.refrat[2], p.colo3, 8.);
            stencil_lower(pw, B, IND0A(9) + e0.js, e0, p.fac00, p.fac10);
            stencil_lower(pw, B, IND1A(9) + e1.js, e1, p.fac01, p.fac11);
            lerp2(pw, B, LS_SELF, p.inds, p.selffrac, p.selffac);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            minor_eta(pw, B, LS_MA1, 9, em.js, em.fs, p.indm, p.minorfrac, adj);
            frac_eta(pw, B, LS_FRACA, p.colh2o, B.refrat[0], p.colo3, 8.);
        } else {
            const double adj = adjcol(p.colco2, p.coldry, CHI(2, p.jp + 1), 3.0, 2.0, 0.79);
            key4(pw, B, LS_ABSB, IND0B(1) + 1, IND1B(1) + 1, p.colo3, p);
            lerp2(pw, B, LS_MB1, p.indm, p.minorfrac, adj);
            frac_const(pw, B, LS_FRACB);
            pw.scale(B.sec[LS_GSCALE] * B.rs);   // (:1645-1650)
        }
    } else if constexpr (BAND == 7) { // band 8: 1080-1180, H2O lower, O3 upper; CO2, O3, N2O minors; CFC12, CFC22 (:1657-1777)
        const double adj = adjcol(p.colco2, p.coldry, CHI(2, p.jp + 1), 3.0, 2.0, 0.65);
        if (lower) {
            key4(pw, B, LS_ABSA, IND0A(1) + 1, IND1A(1) + 1, p.colh2o, p);
            lerp2(pw, B, LS_SELF, p.inds, p.selffrac, p.selffac);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            lerp2(pw, B, LS_MA1, p.indm, p.minorfrac, adj);
            lerp2(pw, B, LS_MA2, p.indm, p.minorfrac, p.colo3);
            lerp2(pw, B, LS_MA3, p.indm, p.minorfrac, p.coln2o);
            frac_const(pw, B, LS_FRACA);
        } else {
            key4(pw, B, LS_ABSB, IND0B(1) + 1, IND1B(1) + 1, p.colo3, p);
            lerp2(pw, B, LS_MB1, p.indm, p.minorfrac, adj);
            lerp2(pw, B, LS_MB2, p.indm, p.minorfrac, p.coln2o);
            frac_const(pw, B, LS_FRACB);
        }
        pw.add_nz(B.sec[LS_X1] * B.rs, p.wx3);
        pw.add_nz(B.sec[LS_X2] * B.rs, p.wx4);
    } else if constexpr (BAND == 8) { // band 9: 1180-1390, H2O/CH4 lower, CH4 upper; N2O minor (:1780-2040)
        const double adj = adjcol(p.coln2o, p.coldry, CHI(4, p.jp + 1), 1.5, 0.5, 0.65);
        if (lower) {
            const Eta e0 = binary(p.colh2o, c_lw.rat_h2och4[p.jp - 1], p.colch4, 8.);
            const Eta e1 = binary(p.colh2o, c_lw.rat_h2och4[p.jp], p.colch4, 8.);
            const Eta em = binary(p.colh2o, B.refrat[2], p.colch4, 8.);
            stencil_lower(pw, B, IND0A(9) + e0.js, e0, p.fac00, p.fac10);
            stencil_lower(pw, B, IND1A(9) + e1.js, e1, p.fac01, p.fac11);
            lerp2(pw, B, LS_SELF, p.inds, p.selffrac, p.selffac);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            minor_eta(pw, B, LS_MA1, 9, em.js, em.fs, p.indm, p.minorfrac, adj);
            frac_eta(pw, B, LS_FRACA, p.colh2o, B.refrat[0], p.colch4, 8.);
        } else {
            key4(pw, B, LS_ABSB, IND0B(1) + 1, IND1B(1) + 1, p.colch4, p);
            lerp2(pw, B, LS_MB1, p.indm, p.minorfrac, adj);
            frac_const(pw, B, LS_FRACB);
        }
    } else if constexpr (BAND == 9) { // band 10: 1390-1480, H2O (:2043-2107)
        if (lower) {
            key4(pw, B, LS_ABSA, IND0A(1) + 1, IND1A(1) + 1, p.colh2o, p);
            lerp2(pw, B, LS_SELF, p.inds, p.selffrac, p.selffac);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            frac_const(pw, B, LS_FRACA);
        } else {
            key4(pw, B, LS_ABSB, IND0B(1) + 1, IND1B(1) + 1, p.colh2o, p);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            frac_const(pw, B, LS_FRACB);
        }
    } else if constexpr (BAND == 10) { // band 11: 1480-1800, H2O; O2 minor (:2110-2187)
        const double scaleo2 = p.colo2 * p.scaleminor;
        if (lower) {
            key4(pw, B, LS_ABSA, IND0A(1) + 1, IND1A(1) + 1, p.colh2o, p);
            lerp2(pw, B, LS_SELF, p.inds, p.selffrac, p.selffac);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            if (scaleo2 != 0.0) lerp2(pw, B, LS_MA1, p.indm, p.minorfrac, scaleo2);
            frac_const(pw, B, LS_FRACA);
        } else {
            key4(pw, B, LS_ABSB, IND0B(1) + 1, IND1B(1) + 1, p.colh2o, p);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            if (scaleo2 != 0.0) lerp2(pw, B, LS_MB1, p.indm, p.minorfrac, scaleo2);
            frac_const(pw, B, LS_FRACB);
        }
    } else if constexpr (BAND == 11) { // band 12: 1800-2080, H2O/CO2 lower; nothing above (:2190-2392)
        if (lower) {
            const Eta e0 = binary(p.colh2o, c_lw.rat_h2oco2[p.jp - 1], p.colco2, 8.);
            const Eta e1 = binary(p.colh2o, c_lw.rat_h2oco2[p.jp], p.colco2, 8.);
            stencil_lower(pw, B, IND0A(9) + e0.js, e0, p.fac00, p.fac10);
            stencil_lower(pw, B, IND1A(9) + e1.js, e1, p.fac01, p.fac11);
            lerp2(pw, B, LS_SELF, p.inds, p.selffrac, p.selffac);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            frac_eta(pw, B, LS_FRACA, p.colh2o, B.refrat[0], p.colco2, 8.);
        } else {
            pw.fzero();
        }
    } else if constexpr (BAND == 12) { // band 13: 2080-2250, H2O/N2O lower; CO2 + CO minors; O3 minor above (:2395-2652)
        if (lower) {
            const Eta e0 = binary(p.colh2o, c_lw.rat_h2on2o[p.jp - 1], p.coln2o, 8.);
            const Eta e1 = binary(p.colh2o, c_lw.rat_h2on2o[p.jp], p.coln2o, 8.);
            const Eta em = binary(p.colh2o, B.refrat[2], p.coln2o, 8.);
            const Eta em3 = binary(p.colh2o, B.refrat[4], p.coln2o, 8.);
            const double adj = adjcol(p.colco2, p.coldry, 3.55e-4, 3.0, 2.0, 0.68);
            stencil_lower(pw, B, IND0A(9) + e0.js, e0, p.fac00, p.fac10);
            stencil_lower(pw, B, IND1A(9) + e1.js, e1, p.fac01, p.fac11);
            lerp2(pw, B, LS_SELF, p.inds, p.selffrac, p.selffac);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            minor_eta(pw, B, LS_MA1, 9, em.js, em.fs, p.indm, p.minorfrac, adj);
            minor_eta(pw, B, LS_MA2, 9, em3.js, em3.fs, p.indm, p.minorfrac, p.colco);
            frac_eta(pw, B, LS_FRACA, p.colh2o, B.refrat[0], p.coln2o, 8.);
        } else {
            lerp2(pw, B, LS_MB1, p.indm, p.minorfrac, p.colo3);
            frac_const(pw, B, LS_FRACB);
        }
    } else if constexpr (BAND == 13) { // band 14: 2250-2380, CO2 (:2655-2713)
        if (lower) {
            key4(pw, B, LS_ABSA, IND0A(1) + 1, IND1A(1) + 1, p.colco2, p);
            lerp2(pw, B, LS_SELF, p.inds, p.selffrac, p.selffac);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            frac_const(pw, B, LS_FRACA);
        } else {
            key4(pw, B, LS_ABSB, IND0B(1) + 1, IND1B(1) + 1, p.colco2, p);
            frac_const(pw, B, LS_FRACB);
        }
    } else if constexpr (BAND == 14) { // band 15: 2380-2600, N2O/CO2 lower; N2 minor; nothing above (:2716-2938)
        if (lower) {
            const Eta e0 = binary(p.coln2o, c_lw.rat_n2oco2[p.jp - 1], p.colco2, 8.);
            const Eta e1 = binary(p.coln2o, c_lw.rat_n2oco2[p.jp], p.colco2, 8.);
            const Eta em = binary(p.coln2o, B.refrat[2], p.colco2, 8.);
            const double scalen2 = p.colbrd * p.scaleminor;
            stencil_lower(pw, B, IND0A(9) + e0.js, e0, p.fac00, p.fac10);
            stencil_lower(pw, B, IND1A(9) + e1.js, e1, p.fac01, p.fac11);
            lerp2(pw, B, LS_SELF, p.inds, p.selffrac, p.selffac);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            minor_eta(pw, B, LS_MA1, 9, em.js, em.fs, p.indm, p.minorfrac, scalen2);
            frac_eta(pw, B, LS_FRACA, p.coln2o, B.refrat[0], p.colco2, 8.);
        } else {
            pw.fzero();
        }
    } else { // band 16: 2600-3250, H2O/CH4 lower, CH4 upper (:2941-3147)
        if (lower) {
            const Eta e0 = binary(p.colh2o, c_lw.rat_h2och4[p.jp - 1], p.colch4, 8.);
            const Eta e1 = binary(p.colh2o, c_lw.rat_h2och4[p.jp], p.colch4, 8.);
            stencil_lower(pw, B, IND0A(9) + e0.js, e0, p.fac00, p.fac10);
            stencil_lower(pw, B, IND1A(9) + e1.js, e1, p.fac01, p.fac11);
            lerp2(pw, B, LS_SELF, p.inds, p.selffrac, p.selffac);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            frac_eta(pw, B, LS_FRACA, p.colh2o, B.refrat[0], p.colch4, 8.);
        } else {
            // The reference sets nspb(16) = 0 (rrtmg_lw_init.f90:209), so taugb16's upper-atmosphere indices
            // ind0 = (...)*nspb(16) + 1 and ind1 collapse to row 1 for every layer (taumol.f90:3135-3136).
            // Reproduced as is: results must match the reference, not the intent.
            key4(pw, B, LS_ABSB, IND0B(0) + 1, IND1B(0) + 1, p.colch4, p);
            frac_const(pw, B, LS_FRACB);
        }
    }
}

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
    ktimer_begin(K_LW_PREP, s);
    lw_prep_cell_kernel<<<(unsigned)(((size_t)w.nc * w.nlay + 127) / 128), 128, 0, s>>>(t, in, w);
    lw_prep_kernel<<<(w.nc + 127) / 128, 128, 0, s>>>(t, in, w);
    ktimer_end(s);
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
