// lw_kernels.cu -- RRTMG longwave on sm_100a: prep (inatm+setcoef), taumol (plan/execute), rtrn (clear sky).
//
// What is computed follows the reference routines (cited per kernel); how it is computed is a
// GPU-first design:
//   lw_prep_kernel    thread <-> column, one sweep over layers: unit conversion, column amounts,
//                     p/T interpolation indices and weights, Planck sources.  Coalesced column-major reads.
//   lw_taumol_kernel  tile = 128 adjacent columns of one layer.  For each of the 16 bands:
//                       plan    (thread <-> column): turn the band formula of taugbN into a short list of
//                               (table row, weight) terms in shared memory -- every gas optical depth of
//                               RRTMG is a weighted sum of k-table rows;
//                       execute (thread <-> (column, g-point), g fastest): tau = sum_k w_k * T[row_k][g],
//                               table rows read as contiguous g-segments through the read-only path.
//   lw_rtrn_kernel    block <-> column, thread <-> g-point: down sweep, surface, up sweep; per-level warp
//                     shuffle reduction over g-points, cross-warp reduction through shared memory.
// Compiled with -fmad=false: fused multiply-adds appear only where written as fma().
#include "rrtmg_dev.cuh"

namespace rrtmg {

__constant__ LwConst c_lw;
__constant__ unsigned char c_lw_ngb[NGPTLW];   // band (0-based) of each g-point

int lw_upload_const(const LwConst &c)
{
    unsigned char ngb[NGPTLW];
    for (int b = 0; b < NBNDLW; ++b)
        for (int i = 0; i < c.band[b].ng; ++i) ngb[c.band[b].g0 + i] = (unsigned char)b;
    if (cudaMemcpyToSymbol(c_lw, &c, sizeof(LwConst)) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(c_lw_ngb, ngb, sizeof ngb) != cudaSuccess) return -1;
    return 0;
}

#define CHI(m, j) c_lw.chi_mls[((j) - 1) * 7 + ((m) - 1)]

// =====================================================================================================
// prep: inatm (LW/src/rrtmg_lw_rad.nomcica.f90:572-901) + setcoef (LW/src/rrtmg_lw_setcoef.f90:31-415)
//       + diffusivity secant (LW/src/rrtmg_lw_rtrnmr.f90:259-280)
// =====================================================================================================
__global__ void __launch_bounds__(128) lw_prep_kernel(LwTables T, LwIn in, LwWork w)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= w.nc) return;
    const int nlay = w.nlay, nc = w.nc;
    const size_t ld = (size_t)in.ld;
    const double amd = 28.9660, amw = 18.0160, amdw = 1.607793, amdo = 0.603428;
    const double grav = 9.8066, avogad = 6.02214199e+23;
    const double stpfac = 296. / 1013.;

    double amttl = 0.0, wvttl = 0.0;
    const double tbound = in.tsfc[col];
    const double pz0 = in.plev[col];
    const double tz0 = in.tlev[col];
    double pzm = pz0;
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
        }
    }

    for (int l = 0; l < nlay; ++l) {
        const size_t o = col + (size_t)l * ld;
        const size_t wo = (size_t)l * nc + col;
        const double pavel = in.play[o], tavel = in.tlay[o];
        const double pz = in.plev[o + ld], tz = in.tlev[o + ld];
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
        pzm = pz;
        double summol = 0.0;
        summol = summol + wkl2; summol = summol + wkl3; summol = summol + wkl4;
        summol = summol + 0.0;  summol = summol + wkl6; summol = summol + wkl7;
        const double wbrodl = coldry * (1.0 - summol);
        wkl1 = coldry * wkl1; wkl2 = coldry * wkl2; wkl3 = coldry * wkl3; wkl4 = coldry * wkl4;
        wkl6 = coldry * wkl6; wkl7 = coldry * wkl7;
        const double wkl5 = coldry * 0.0;
        amttl = amttl + coldry + wkl1;
        wvttl = wvttl + wkl1;
        w.fld(LF_WX1)[wo] = in.ccl4 ? coldry * in.ccl4[o] * 1.e-20 : 0.0;
        w.fld(LF_WX2)[wo] = in.cfc11 ? coldry * in.cfc11[o] * 1.e-20 : 0.0;
        w.fld(LF_WX3)[wo] = in.cfc12 ? coldry * in.cfc12[o] * 1.e-20 : 0.0;
        w.fld(LF_WX4)[wo] = in.cfc22 ? coldry * in.cfc22[o] * 1.e-20 : 0.0;

        // ---- setcoef: Planck sources
        int indlay = (int)(tavel - 159.);
        indlay = indlay < 1 ? 1 : (indlay > 180 ? 180 : indlay);
        const double tlayfrac = tavel - 159. - (double)indlay;
        int indlev = (int)(tz - 159.);
        indlev = indlev < 1 ? 1 : (indlev > 180 ? 180 : indlev);
        const double tlevfrac = tz - 159. - (double)indlev;
        {
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
        if (!(plog <= 4.56)) {
            laytrop = laytrop + 1;
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
        const double scaleminor = pavel / tavel;
        const double scaleminorn2 = (pavel / tavel) * (wbrodl / (coldry + wkl1));
        factor = (tavel - 180.8) / 7.2;
        int indminor = (int)factor;
        indminor = indminor < 1 ? 1 : (indminor > 18 ? 18 : indminor);
        const double minorfrac = factor - (double)indminor;

        const double colh2o = 1.e-20 * wkl1;
        double colco2 = 1.e-20 * wkl2, colo3 = 1.e-20 * wkl3, coln2o = 1.e-20 * wkl4;
        double colco = 1.e-20 * wkl5, colch4 = 1.e-20 * wkl6;
        const double colo2 = 1.e-20 * wkl7;
        if (colco2 == 0.) colco2 = 1.e-32 * coldry;
        if (colo3 == 0.) colo3 = 1.e-32 * coldry;
        if (coln2o == 0.) coln2o = 1.e-32 * coldry;
        if (colco == 0.) colco = 1.e-32 * coldry;
        if (colch4 == 0.) colch4 = 1.e-32 * coldry;
        const double colbrd = 1.e-20 * wbrodl;
        const double compfp = 1. - fp;

        w.idx[wo] = lw_pack(jp, jt, jt1, indself, indfor, indminor);
        w.fld(LF_FAC10)[wo] = compfp * ft;
        w.fld(LF_FAC00)[wo] = compfp * (1. - ft);
        w.fld(LF_FAC11)[wo] = fp * ft1;
        w.fld(LF_FAC01)[wo] = fp * (1. - ft1);
        w.fld(LF_COLH2O)[wo] = colh2o;
        w.fld(LF_COLCO2)[wo] = colco2;
        w.fld(LF_COLO3)[wo] = colo3;
        w.fld(LF_COLN2O)[wo] = coln2o;
        w.fld(LF_COLCO)[wo] = colco;
        w.fld(LF_COLCH4)[wo] = colch4;
        w.fld(LF_COLO2)[wo] = colo2;
        w.fld(LF_COLBRD)[wo] = colbrd;
        w.fld(LF_SELFFAC)[wo] = colh2o * selffac;
        w.fld(LF_SELFFRAC)[wo] = selffrac;
        w.fld(LF_FORFAC)[wo] = colh2o * forfac;
        w.fld(LF_FORFRAC)[wo] = forfrac;
        w.fld(LF_MINORFRAC)[wo] = minorfrac;
        w.fld(LF_SCALEMINOR)[wo] = scaleminor;
        w.fld(LF_SCALEMINORN2)[wo] = scaleminorn2;
        w.fld(LF_COLDRY)[wo] = coldry;
        w.fld(LF_PAVEL)[wo] = pavel;
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
// =====================================================================================================
constexpr int TP = 128;    // columns per tile == threads per block
constexpr int KMAX = 24;   // most terms any band needs (band 13 lower: 6+6+2+2+4+4)

struct PlanSmem {
    double w[KMAX][TP];
    int off[KMAX][TP];
    double wf[2][TP];
    int offf[2][TP];
    int n[TP], nf[TP], gs[TP];
};

struct PW {
    PlanSmem *s;
    int t, n, nf;
    __device__ __forceinline__ void add(int off, double wgt) { s->w[n][t] = wgt; s->off[n][t] = off; ++n; }
    __device__ __forceinline__ void addf(int off, double wgt) { s->wf[nf][t] = wgt; s->offf[nf][t] = off; ++nf; }
};

struct LwPair {
    int jp, jt, jt1, inds, indf, indm;
    double fac00, fac01, fac10, fac11;
    double colh2o, colco2, colo3, coln2o, colco, colch4, colo2, colbrd;
    double selffac, selffrac, forfac, forfrac, minorfrac, scaleminor, scaleminorn2, coldry, pavel;
    double wx1, wx2, wx3, wx4;
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
__device__ __forceinline__ void key4(PW &pw, const LwBand &B, int sec, int ind0, int ind1, double scale, const LwPair &p)
{
    const int ng = B.ng, o0 = (B.sec[sec] + ind0 - 1) * ng, o1 = (B.sec[sec] + ind1 - 1) * ng;
    pw.add(o0, scale * p.fac00);
    pw.add(o0 + ng, scale * p.fac10);
    pw.add(o1, scale * p.fac01);
    pw.add(o1 + ng, scale * p.fac11);
}
__device__ __forceinline__ void lerp2(PW &pw, const LwBand &B, int sec, int row, double frac, double scale)
{
    const int ng = B.ng, o = (B.sec[sec] + row - 1) * ng;
    pw.add(o, scale * (1. - frac));
    pw.add(o + ng, scale * frac);
}
// minor gas with eta dimension, Fortran (neta,19,ng): 4-point (eta, T) interpolation
__device__ __forceinline__ void minor_eta(PW &pw, const LwBand &B, int sec, int neta, int jm, double fm, int indm,
                                          double mf, double scale)
{
    const int ng = B.ng, o = (B.sec[sec] + (indm - 1) * neta + (jm - 1)) * ng;
    pw.add(o, scale * ((1. - mf) * (1. - fm)));
    pw.add(o + ng, scale * ((1. - mf) * fm));
    pw.add(o + neta * ng, scale * (mf * (1. - fm)));
    pw.add(o + (neta + 1) * ng, scale * (mf * fm));
}
// lower-atmosphere binary-species key term: 3-point stencil near eta = 0 / 1, else 2-point
// (template block repeated in taugb3,4,5,7,9,12,13,15,16, e.g. taumol.f90:548-606)
__device__ __forceinline__ void stencil_lower(PW &pw, const LwBand &B, int ind, const Eta &e, double facA, double facB)
{
    const int ng = B.ng, o = (B.sec[LS_ABSA] + ind - 1) * ng;
    const double sc = e.speccomb;
    if (e.specparm < 0.125) {
        const double p = e.fs - 1, p4 = (p * p) * (p * p);
        const double fk0 = p4, fk1 = 1 - p - 2.0 * p4, fk2 = p + p4;
        pw.add(o, sc * (fk0 * facA));
        pw.add(o + ng, sc * (fk1 * facA));
        pw.add(o + 2 * ng, sc * (fk2 * facA));
        pw.add(o + 9 * ng, sc * (fk0 * facB));
        pw.add(o + 10 * ng, sc * (fk1 * facB));
        pw.add(o + 11 * ng, sc * (fk2 * facB));
    } else if (e.specparm > 0.875) {
        const double p = -e.fs, p4 = (p * p) * (p * p);
        const double fk0 = p4, fk1 = 1 - p - 2.0 * p4, fk2 = p + p4;
        pw.add(o - ng, sc * (fk2 * facA));
        pw.add(o, sc * (fk1 * facA));
        pw.add(o + ng, sc * (fk0 * facA));
        pw.add(o + 8 * ng, sc * (fk2 * facB));
        pw.add(o + 9 * ng, sc * (fk1 * facB));
        pw.add(o + 10 * ng, sc * (fk0 * facB));
    } else {
        pw.add(o, sc * ((1. - e.fs) * facA));
        pw.add(o + ng, sc * (e.fs * facA));
        pw.add(o + 9 * ng, sc * ((1. - e.fs) * facB));
        pw.add(o + 10 * ng, sc * (e.fs * facB));
    }
}
// upper-atmosphere binary key term (nspb = 5): always 2-point (e.g. taumol.f90:739-750)
__device__ __forceinline__ void stencil_upper(PW &pw, const LwBand &B, int ind, const Eta &e, double facA, double facB)
{
    const int ng = B.ng, o = (B.sec[LS_ABSB] + ind - 1) * ng;
    const double sc = e.speccomb;
    pw.add(o, sc * ((1. - e.fs) * facA));
    pw.add(o + ng, sc * (e.fs * facA));
    pw.add(o + 5 * ng, sc * ((1. - e.fs) * facB));
    pw.add(o + 6 * ng, sc * (e.fs * facB));
}
__device__ __forceinline__ void frac_const(PW &pw, const LwBand &B, int sec) { pw.addf(B.sec[sec] * B.ng, 1.0); }
__device__ __forceinline__ void frac_eta(PW &pw, const LwBand &B, int sec, double colA, double refrat, double colB, double mult)
{
    const Eta e = binary(colA, refrat, colB, mult);
    const int o = (B.sec[sec] + e.js - 1) * B.ng;
    pw.addf(o, 1. - e.fs);
    pw.addf(o + B.ng, e.fs);
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

__device__ void lw_plan_band(int band, const LwPair &p, bool lower, PW &pw)
{
    const LwBand &B = c_lw.band[band];
    int gs = -1;
    pw.n = 0;
    pw.nf = 0;
    switch (band) {
    case 0: { // band 1: 10-350 cm-1, H2O; N2 continuum minor (:280-373)
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
    } break;
    case 1: { // band 2: 350-500, H2O (:376-445)
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
    } break;
    case 2: { // band 3: 500-630, H2O/CO2 both regions; N2O minor (:448-760)
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
    } break;
    case 3: { // band 4: 630-700, H2O/CO2 lower, O3/CO2 upper (:763-1019)
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
            gs = B.sec[LS_GSCALE] * B.ng;   // stratospheric g-point scaling (:1009-1015)
        }
    } break;
    case 4: { // band 5: 700-820, H2O/CO2 lower, O3/CO2 upper; O3 minor, CCl4 (:1022-1294)
        if (lower) {
            const Eta e0 = binary(p.colh2o, c_lw.rat_h2oco2[p.jp - 1], p.colco2, 8.);
            const Eta e1 = binary(p.colh2o, c_lw.rat_h2oco2[p.jp], p.colco2, 8.);
            const Eta em = binary(p.colh2o, B.refrat[2], p.colco2, 8.);
            stencil_lower(pw, B, IND0A(9) + e0.js, e0, p.fac00, p.fac10);
            stencil_lower(pw, B, IND1A(9) + e1.js, e1, p.fac01, p.fac11);
            lerp2(pw, B, LS_SELF, p.inds, p.selffrac, p.selffac);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            minor_eta(pw, B, LS_MA1, 9, em.js, em.fs, p.indm, p.minorfrac, p.colo3);
            pw.add(B.sec[LS_X1] * B.ng, p.wx1);
            frac_eta(pw, B, LS_FRACA, p.colh2o, B.refrat[0], p.colco2, 8.);
        } else {
            const Eta e0 = binary(p.colo3, c_lw.rat_o3co2[p.jp - 1], p.colco2, 4.);
            const Eta e1 = binary(p.colo3, c_lw.rat_o3co2[p.jp], p.colco2, 4.);
            stencil_upper(pw, B, IND0B(5) + e0.js, e0, p.fac00, p.fac10);
            stencil_upper(pw, B, IND1B(5) + e1.js, e1, p.fac01, p.fac11);
            pw.add(B.sec[LS_X1] * B.ng, p.wx1);
            frac_eta(pw, B, LS_FRACB, p.colo3, B.refrat[1], p.colco2, 4.);
        }
    } break;
    case 5: { // band 6: 820-980, H2O lower; CO2 minor, CFC11, CFC12 (:1297-1380)
        if (lower) {
            const double adj = adjcol(p.colco2, p.coldry, CHI(2, p.jp + 1), 3.0, 2.0, 0.77);
            key4(pw, B, LS_ABSA, IND0A(1) + 1, IND1A(1) + 1, p.colh2o, p);
            lerp2(pw, B, LS_SELF, p.inds, p.selffrac, p.selffac);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            lerp2(pw, B, LS_MA1, p.indm, p.minorfrac, adj);
        }
        pw.add(B.sec[LS_X1] * B.ng, p.wx2);
        pw.add(B.sec[LS_X2] * B.ng, p.wx3);
        frac_const(pw, B, LS_FRACA);
    } break;
    case 6: { // band 7: 980-1080, H2O/O3 lower, O3 upper; CO2 minor (:1383-1654)
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
            gs = B.sec[LS_GSCALE] * B.ng;   // (:1645-1650)
        }
    } break;
    case 7: { // band 8: 1080-1180, H2O lower, O3 upper; CO2, O3, N2O minors; CFC12, CFC22 (:1657-1777)
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
        pw.add(B.sec[LS_X1] * B.ng, p.wx3);
        pw.add(B.sec[LS_X2] * B.ng, p.wx4);
    } break;
    case 8: { // band 9: 1180-1390, H2O/CH4 lower, CH4 upper; N2O minor (:1780-2040)
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
    } break;
    case 9: { // band 10: 1390-1480, H2O (:2043-2107)
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
    } break;
    case 10: { // band 11: 1480-1800, H2O; O2 minor (:2110-2187)
        const double scaleo2 = p.colo2 * p.scaleminor;
        if (lower) {
            key4(pw, B, LS_ABSA, IND0A(1) + 1, IND1A(1) + 1, p.colh2o, p);
            lerp2(pw, B, LS_SELF, p.inds, p.selffrac, p.selffac);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            lerp2(pw, B, LS_MA1, p.indm, p.minorfrac, scaleo2);
            frac_const(pw, B, LS_FRACA);
        } else {
            key4(pw, B, LS_ABSB, IND0B(1) + 1, IND1B(1) + 1, p.colh2o, p);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            lerp2(pw, B, LS_MB1, p.indm, p.minorfrac, scaleo2);
            frac_const(pw, B, LS_FRACB);
        }
    } break;
    case 11: { // band 12: 1800-2080, H2O/CO2 lower; nothing above (:2190-2392)
        if (lower) {
            const Eta e0 = binary(p.colh2o, c_lw.rat_h2oco2[p.jp - 1], p.colco2, 8.);
            const Eta e1 = binary(p.colh2o, c_lw.rat_h2oco2[p.jp], p.colco2, 8.);
            stencil_lower(pw, B, IND0A(9) + e0.js, e0, p.fac00, p.fac10);
            stencil_lower(pw, B, IND1A(9) + e1.js, e1, p.fac01, p.fac11);
            lerp2(pw, B, LS_SELF, p.inds, p.selffrac, p.selffac);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            frac_eta(pw, B, LS_FRACA, p.colh2o, B.refrat[0], p.colco2, 8.);
        }
    } break;
    case 12: { // band 13: 2080-2250, H2O/N2O lower; CO2 + CO minors; O3 minor above (:2395-2652)
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
    } break;
    case 13: { // band 14: 2250-2380, CO2 (:2655-2713)
        if (lower) {
            key4(pw, B, LS_ABSA, IND0A(1) + 1, IND1A(1) + 1, p.colco2, p);
            lerp2(pw, B, LS_SELF, p.inds, p.selffrac, p.selffac);
            lerp2(pw, B, LS_FOR, p.indf, p.forfrac, p.forfac);
            frac_const(pw, B, LS_FRACA);
        } else {
            key4(pw, B, LS_ABSB, IND0B(1) + 1, IND1B(1) + 1, p.colco2, p);
            frac_const(pw, B, LS_FRACB);
        }
    } break;
    case 14: { // band 15: 2380-2600, N2O/CO2 lower; N2 minor; nothing above (:2716-2938)
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
        }
    } break;
    default: { // band 16: 2600-3250, H2O/CH4 lower, CH4 upper (:2941-3147)
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
    } break;
    }
    pw.s->n[pw.t] = pw.n;
    pw.s->nf[pw.t] = pw.nf;
    pw.s->gs[pw.t] = gs;
}

template <int NG>
__device__ __forceinline__ void lw_exec_band(const PlanSmem &s, const double *__restrict__ tab, int g0,
                                             double *__restrict__ taug, double *__restrict__ fracs,
                                             int c0, int nvalid, int lay, int nlay)
{
    for (int cell = threadIdx.x; cell < TP * NG; cell += TP) {
        const int pr = cell / NG, ig = cell - pr * NG;
        if (pr >= nvalid) break;
        const int n = s.n[pr];
        double acc = 0.0;
        for (int k = 0; k < n; ++k) acc = fma(s.w[k][pr], __ldg(tab + s.off[k][pr] + ig), acc);
        const int gs = s.gs[pr];
        if (gs >= 0) acc = acc * __ldg(tab + gs + ig);
        const int nf = s.nf[pr];
        double fr = 0.0;
        if (nf > 0) fr = s.wf[0][pr] * __ldg(tab + s.offf[0][pr] + ig);
        if (nf > 1) fr = fma(s.wf[1][pr], __ldg(tab + s.offf[1][pr] + ig), fr);
        const size_t o = ((size_t)(c0 + pr) * nlay + lay) * NGPTLW + g0 + ig;
        taug[o] = acc;
        fracs[o] = fr;
    }
}

__global__ void __launch_bounds__(TP) lw_taumol_kernel(LwTables T, LwWork w)
{
    __shared__ PlanSmem s;
    const int t = threadIdx.x;
    const int c0 = blockIdx.x * TP;
    const int lay = blockIdx.y;               // 0-based layer
    const int nlay = w.nlay, nc = w.nc;
    const int col = c0 + t;
    const bool valid = col < nc;
    const int nvalid = min(TP, nc - c0);

    LwPair p;
    bool lower = false;
    if (valid) {
        const size_t wo = (size_t)lay * nc + col;
        const LwIdx ix = lw_unpack(w.idx[wo]);
        p.jp = ix.jp; p.jt = ix.jt; p.jt1 = ix.jt1; p.inds = ix.inds; p.indf = ix.indf; p.indm = ix.indm;
        p.fac00 = w.fld(LF_FAC00)[wo]; p.fac01 = w.fld(LF_FAC01)[wo];
        p.fac10 = w.fld(LF_FAC10)[wo]; p.fac11 = w.fld(LF_FAC11)[wo];
        p.colh2o = w.fld(LF_COLH2O)[wo]; p.colco2 = w.fld(LF_COLCO2)[wo]; p.colo3 = w.fld(LF_COLO3)[wo];
        p.coln2o = w.fld(LF_COLN2O)[wo]; p.colco = w.fld(LF_COLCO)[wo]; p.colch4 = w.fld(LF_COLCH4)[wo];
        p.colo2 = w.fld(LF_COLO2)[wo]; p.colbrd = w.fld(LF_COLBRD)[wo];
        p.selffac = w.fld(LF_SELFFAC)[wo]; p.selffrac = w.fld(LF_SELFFRAC)[wo];
        p.forfac = w.fld(LF_FORFAC)[wo]; p.forfrac = w.fld(LF_FORFRAC)[wo];
        p.minorfrac = w.fld(LF_MINORFRAC)[wo]; p.scaleminor = w.fld(LF_SCALEMINOR)[wo];
        p.scaleminorn2 = w.fld(LF_SCALEMINORN2)[wo]; p.coldry = w.fld(LF_COLDRY)[wo];
        p.pavel = w.fld(LF_PAVEL)[wo];
        p.wx1 = w.fld(LF_WX1)[wo]; p.wx2 = w.fld(LF_WX2)[wo]; p.wx3 = w.fld(LF_WX3)[wo]; p.wx4 = w.fld(LF_WX4)[wo];
        lower = (lay + 1) <= w.laytrop[col];
    }
    PW pw;
    pw.s = &s;
    pw.t = t;
    for (int band = 0; band < NBNDLW; ++band) {
        if (valid) lw_plan_band(band, p, lower, pw);
        __syncthreads();
        const LwBand &B = c_lw.band[band];
        const double *tab = T.tab + B.base;
        switch (B.ng) {
        case 16: lw_exec_band<16>(s, tab, B.g0, w.taug, w.fracs, c0, nvalid, lay, nlay); break;
        case 14: lw_exec_band<14>(s, tab, B.g0, w.taug, w.fracs, c0, nvalid, lay, nlay); break;
        case 12: lw_exec_band<12>(s, tab, B.g0, w.taug, w.fracs, c0, nvalid, lay, nlay); break;
        case 10: lw_exec_band<10>(s, tab, B.g0, w.taug, w.fracs, c0, nvalid, lay, nlay); break;
        case 8: lw_exec_band<8>(s, tab, B.g0, w.taug, w.fracs, c0, nvalid, lay, nlay); break;
        case 6: lw_exec_band<6>(s, tab, B.g0, w.taug, w.fracs, c0, nvalid, lay, nlay); break;
        case 4: lw_exec_band<4>(s, tab, B.g0, w.taug, w.fracs, c0, nvalid, lay, nlay); break;
        default: lw_exec_band<2>(s, tab, B.g0, w.taug, w.fracs, c0, nvalid, lay, nlay); break;
        }
        __syncthreads();
    }
}

// =====================================================================================================
// rtrn: clear-sky radiative transfer, LW/src/rrtmg_lw_rtrnmr.f90:481-777 ("Clear layer" branches; identical
// in rtrnmc.f90:407-432,481-503).  taut = taug + tauaer (rad.nomcica:514-519, iaer = 10 forced).
// The staging fields are overwritten in place: taug -> atrans, fracs -> bbugas (needed by the up sweep).
// =====================================================================================================
constexpr int RT_THREADS = 160;   // 140 g-points -> 5 warps
constexpr int RT_S = 141;         // tile row stride (odd)

__global__ void __launch_bounds__(RT_THREADS) lw_rtrn_kernel(LwTables T, LwIn in, LwOut out, LwWork w)
{
    __shared__ double s_tile[16 * RT_S];
    __shared__ double s_part[16 * (RT_THREADS / 16 + 1)];
    __shared__ double s_dn[MAXLAY + 1], s_up[MAXLAY + 1];
    const int col = blockIdx.x;
    const int nlay = w.nlay;
    const int g = threadIdx.x;
    const bool active = g < NGPTLW;
    const int band = active ? c_lw_ngb[g] : 0;
    const double secd = w.secdiff[(size_t)col * 16 + band];
    const double wgt = active ? 0.5 * c_lw.delwave[band] : 0.0;     // wtdiff * delwave
    const double bpade = c_lw.bpade;
    const double rec_6 = 0.166667;
    const double2 *__restrict__ et = reinterpret_cast<const double2 *>(T.exptfn);
    double *taug = w.taug + (size_t)col * nlay * NGPTLW + g;
    double *fracs = w.fracs + (size_t)col * nlay * NGPTLW + g;
    const double *pl = w.planklay + (size_t)col * nlay * 16 + band;
    const double *pv = w.planklev + (size_t)col * (nlay + 1) * 16 + band;
    const double *taer = in.tauaer ? in.tauaer + col + (size_t)band * nlay * in.ld : nullptr;

    // ---- downward sweep (:505-618); batches of 16 levels go through tile_reduce16
    double radld = 0.0;
    double plfrac1 = 0.0;
    for (int k = 0; k < nlay; ++k) {
        const int lev = nlay - k;
        const int slot = k & 15;
        if (active) {
            const size_t o = (size_t)(lev - 1) * NGPTLW;
            const double plfrac = fracs[o];
            const double blay = pl[(lev - 1) * 16];
            const double dplankup = pv[lev * 16] - blay;
            const double dplankdn = pv[(lev - 1) * 16] - blay;
            double taut = taug[o];
            if (taer) taut = taut + taer[(size_t)(lev - 1) * in.ld];
            double odepth = secd * taut;
            if (odepth < 0.0) odepth = 0.0;
            double atrans, bbd, bbugas;
            if (odepth <= 0.06) {
                atrans = odepth - 0.5 * odepth * odepth;
                odepth = rec_6 * odepth;
                bbd = plfrac * (blay + dplankdn * odepth);
                bbugas = plfrac * (blay + dplankup * odepth);
            } else {
                const double tblind = odepth * rcp_fast(bpade + odepth);
                const int itr = (int)(10000.0 * tblind + 0.5);
                const double2 e = __ldg(et + itr);
                atrans = 1. - e.x;
                bbd = plfrac * (blay + e.y * dplankdn);
                bbugas = plfrac * (blay + e.y * dplankup);
            }
            radld = fma(bbd - radld, atrans, radld);
            taug[o] = atrans;
            fracs[o] = bbugas;
            s_tile[slot * RT_S + g] = radld * wgt;
            if (lev == 1) plfrac1 = plfrac;
        }
        if (slot == 15 || k == nlay - 1) {
            const double sum = tile_reduce16<RT_THREADS, NGPTLW, RT_S>(s_tile, s_part);
            const int kk = (k & ~15) + threadIdx.x;
            if (threadIdx.x < 16 && kk <= k) s_dn[nlay - 1 - kk] = sum * c_lw.fluxfac;
        }
    }
    if (threadIdx.x == 0) s_dn[nlay] = 0.0;   // no downward flux enters at the top (drad(nlayers) = 0)

    // ---- surface (:628-636) and upward sweep (:649-711); level k = 0 is the surface
    double radlu = 0.0;
    for (int k = 0; k <= nlay; ++k) {
        const int slot = k & 15;
        if (active) {
            if (k == 0) {
                const double semiss = in.emis ? in.emis[col + (size_t)band * in.ld] : 1.0;
                const double rad0 = plfrac1 * w.plankbnd[(size_t)col * 16 + band];
                const double reflect = 1. - semiss;
                radlu = rad0 + reflect * radld;
            } else {
                const size_t o = (size_t)(k - 1) * NGPTLW;
                const double atrans = taug[o], bbugas = fracs[o];
                radlu = fma(bbugas - radlu, atrans, radlu);
            }
            s_tile[slot * RT_S + g] = radlu * wgt;
        }
        if (slot == 15 || k == nlay) {
            const double sum = tile_reduce16<RT_THREADS, NGPTLW, RT_S>(s_tile, s_part);
            const int kk = (k & ~15) + threadIdx.x;
            if (threadIdx.x < 16 && kk <= k) s_up[kk] = sum * c_lw.fluxfac;
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
            const double h = c_lw.heatfac * (fnet0 - fnet1) / (pz0 - pz1);
            out.hr[o] = h;
            out.hrc[o] = h;
        }
    }
}

int lw_run_pass(const LwTables &t, const LwIn &in, const LwOut &out, LwWork &w, cudaStream_t s, double *cap)
{
    ktimer_begin(K_LW_PREP, s);
    lw_prep_kernel<<<(w.nc + 127) / 128, 128, 0, s>>>(t, in, w);
    ktimer_end(s);
    dim3 grid((w.nc + TP - 1) / TP, w.nlay);
    ktimer_begin(K_LW_TAUMOL, s);
    lw_taumol_kernel<<<grid, TP, 0, s>>>(t, w);
    ktimer_end(s);
    if (cap) {
        const size_t n = (size_t)w.nc * w.nlay * NGPTLW;
        cudaMemcpyAsync(cap, w.taug, n * 8, cudaMemcpyDeviceToDevice, s);
        cudaMemcpyAsync(cap + n, w.fracs, n * 8, cudaMemcpyDeviceToDevice, s);
    }
    ktimer_begin(K_LW_RTRN, s);
    lw_rtrn_kernel<<<w.nc, RT_THREADS, 0, s>>>(t, in, out, w);
    ktimer_end(s);
    return 3;
}

} // namespace rrtmg
