// lw_bands.cuh -- the per-cell arithmetic of RRTMG_LW's setcoef and taumol, shared by the translation units that evaluate
// it (lw_kernels.cu: the staged taumol kernel; lw_column.cu: the fused clear-sky column kernel).  Everything is a
// __forceinline__ device template; the band formulas are written once against an accumulator policy PW (add / add_nz /
// scale / frac1 / frac2 / fzero), so a caller decides where the ng values of a band live.  The constant block c_lw is
// per translation unit (static __constant__): every unit that includes this header uploads its own copy.
// Include only from units compiled with -fmad=false (build.py): index and branch decisions follow IEEE evaluation.
#pragma once
#include "rrtmg_dev.cuh"

namespace rrtmg {

static __constant__ LwConst c_lw;

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

} // namespace rrtmg
