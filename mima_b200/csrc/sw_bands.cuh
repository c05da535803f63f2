// sw_bands.cuh -- the per-cell arithmetic of RRTMG_SW's setcoef_sw and taumol_sw, shared by the translation units that
// evaluate it (sw_kernels.cu: the staged taumol kernel; sw_column.cu: the fused clear-sky column kernel).  The band formulas
// are written once against an accumulator policy PW (add / addc / rayl1 / rayl2 / sflux1 / sflux2), so a caller decides
// where the ng values of a band live.  The constant block c_sw is per translation unit (static __constant__): every unit
// that includes this header uploads its own copy.
#pragma once
#include "rrtmg_dev.cuh"

namespace rrtmg {

static __constant__ SwConst c_sw;

constexpr double ZEPZEN = 1.e-10;

// Interpolation state of one (column, layer) cell: what setcoef_sw hands to taumol_sw.
struct SwPair {
    int jp, jt, jt1, inds, indf;
    double fac00, fac01, fac10, fac11;
    double colh2o, colco2, colo3, colch4, colo2, colmol, coln2o;
    double selffac, selffrac, forfac, forfrac;
};

// =====================================================================================================
// inatm_sw (SW/src/rrtmg_sw_rad.nomcica.f90:761-1101) + setcoef_sw (SW/src/rrtmg_sw_setcoef.f90:30-286)
// for one (column, layer) cell; shared by the prep kernel (column-integrated quantities) and the taumol
// kernel (which evaluates the cell state in place).  Returns true when the layer counts towards laytrop.
// =====================================================================================================
__device__ __forceinline__ bool sw_cell(const SwIn &in, int col, int l, SwPair &p)
{
    const double amd = 28.9660, amw = 18.0160, amdw = 1.607793, amdo = 0.603428;
    const double grav = 9.8066, avogad = 6.02214199e+23;
    const double stpfac = 296. / 1013.;
    const size_t ld = (size_t)in.ld;
    const size_t o = col + (size_t)l * ld;
    const double pavel = in.play[o], tavel = in.tlay[o];
    const double pzm = in.plev[o], pz = in.plev[o + ld];
    const double q = in.h2o[o];
    double wkl1 = (q / (1. - q)) * amdw;
    double wkl2 = in.co2[o];
    double wkl3 = in.o3[o] * amdo;
    double wkl4 = in.n2o ? in.n2o[o] : 0.0;
    double wkl6 = in.ch4 ? in.ch4[o] : 0.0;
    double wkl7 = in.o2 ? in.o2[o] : 0.0;
    const double amm = (1. - wkl1) * amd + wkl1 * amw;
    const double coldry = (pzm - pz) * 1.e3 * avogad / (1.e2 * grav * amm * (1. + wkl1));
    wkl1 = coldry * wkl1; wkl2 = coldry * wkl2; wkl3 = coldry * wkl3; wkl4 = coldry * wkl4;
    wkl6 = coldry * wkl6; wkl7 = coldry * wkl7;

    const double plog = log(pavel);
    int jp = (int)(36. - 5 * (plog + 0.04));
    jp = jp < 1 ? 1 : (jp > 58 ? 58 : jp);
    const double fp = 5. * (c_sw.preflog[jp - 1] - plog);
    const double tr0 = (tavel - c_sw.tref[jp - 1]) / 15.;
    int jt = (int)(3. + tr0);
    jt = jt < 1 ? 1 : (jt > 4 ? 4 : jt);
    const double ft = tr0 - (double)(jt - 3);
    const double tr1 = (tavel - c_sw.tref[jp]) / 15.;
    int jt1 = (int)(3. + tr1);
    jt1 = jt1 < 1 ? 1 : (jt1 > 4 ? 4 : jt1);
    const double ft1 = tr1 - (double)(jt1 - 3);
    const double water = wkl1 / coldry;
    const double scalefac = pavel * stpfac / tavel;
    const double forfac = scalefac / (1. + water);
    double forfrac, selffac = 0.0, selffrac = 0.0, factor;
    int indfor, indself = 0;
    const bool lower = !(plog <= 4.56);
    if (lower) {
        factor = (332.0 - tavel) / 36.0;
        indfor = (int)factor;
        indfor = indfor < 1 ? 1 : (indfor > 2 ? 2 : indfor);
        forfrac = factor - (double)indfor;
        selffac = water * forfac;
        factor = (tavel - 188.0) / 7.2;
        indself = (int)factor - 7;
        indself = indself < 1 ? 1 : (indself > 9 ? 9 : indself);
        selffrac = factor - (double)(indself + 7);
    } else {
        factor = (tavel - 188.0) / 36.0;
        indfor = 3;
        forfrac = factor - 1.0;
    }
    p.colh2o = 1.e-20 * wkl1;
    double colco2 = 1.e-20 * wkl2;
    p.colo3 = 1.e-20 * wkl3;
    double coln2o = 1.e-20 * wkl4, colch4 = 1.e-20 * wkl6, colo2 = 1.e-20 * wkl7;
    p.colmol = 1.e-20 * coldry + p.colh2o;
    if (colco2 == 0.) colco2 = 1.e-32 * coldry;
    if (coln2o == 0.) coln2o = 1.e-32 * coldry;
    if (colch4 == 0.) colch4 = 1.e-32 * coldry;
    if (colo2 == 0.) colo2 = 1.e-32 * coldry;
    p.colco2 = colco2; p.coln2o = coln2o; p.colch4 = colch4; p.colo2 = colo2;
    const double compfp = 1. - fp;
    p.jp = jp; p.jt = jt; p.jt1 = jt1; p.inds = indself; p.indf = indfor;
    p.fac10 = compfp * ft;
    p.fac00 = compfp * (1. - ft);
    p.fac11 = fp * ft1;
    p.fac01 = fp * (1. - ft1);
    p.selffac = selffac; p.selffrac = selffrac; p.forfac = forfac; p.forfrac = forfrac;
    return lower;
}

struct Eta { double speccomb, fs; int js; };
__device__ __forceinline__ Eta binary(double colA, double strrat, double colB, double mult)
{
    Eta e;
    e.speccomb = colA + strrat * colB;
    double specparm = colA / e.speccomb;
    if (specparm >= c_sw.oneminus) specparm = c_sw.oneminus;
    const double specmult = mult * specparm;
    const int i = (int)specmult;
    e.js = 1 + i;
    e.fs = specmult - (double)i;
    return e;
}
template <class PW>
__device__ __forceinline__ void key4(PW &pw, const SwBand &B, int sec, int ind0, int ind1, double scale, const SwPair &p)
{
    const int ng = B.rs, o0 = (B.sec[sec] + ind0 - 1) * ng, o1 = (B.sec[sec] + ind1 - 1) * ng;
    pw.add(o0, scale * p.fac00);
    pw.add(o0 + ng, scale * p.fac10);
    pw.add(o1, scale * p.fac01);
    pw.add(o1 + ng, scale * p.fac11);
}
// 8-point binary key term; one eta for both pressure levels; dT = 9 (lower) / 5 (upper)
template <class PW>
__device__ __forceinline__ void key8(PW &pw, const SwBand &B, int sec, int ind0, int ind1, int dT, const Eta &e, const SwPair &p)
{
    const int ng = B.rs, o0 = (B.sec[sec] + ind0 - 1) * ng, o1 = (B.sec[sec] + ind1 - 1) * ng;
    const double sc = e.speccomb, a = 1. - e.fs, b = e.fs;
    pw.add(o0, sc * (a * p.fac00));
    pw.add(o0 + ng, sc * (b * p.fac00));
    pw.add(o0 + dT * ng, sc * (a * p.fac10));
    pw.add(o0 + (dT + 1) * ng, sc * (b * p.fac10));
    pw.add(o1, sc * (a * p.fac01));
    pw.add(o1 + ng, sc * (b * p.fac01));
    pw.add(o1 + dT * ng, sc * (a * p.fac11));
    pw.add(o1 + (dT + 1) * ng, sc * (b * p.fac11));
}
template <class PW>
__device__ __forceinline__ void lerp2(PW &pw, const SwBand &B, int sec, int row, double frac, double scale)
{
    const int ng = B.rs, o = (B.sec[sec] + row - 1) * ng;
    pw.add(o, scale * (1. - frac));
    pw.add(o + ng, scale * frac);
}
template <class PW>
__device__ __forceinline__ void selffor(PW &pw, const SwBand &B, const SwPair &p, double scale)
{
    lerp2(pw, B, SS_SELF, p.inds, p.selffrac, scale * p.selffac);
    lerp2(pw, B, SS_FOR, p.indf, p.forfrac, scale * p.forfac);
}
template <class PW>
__device__ __forceinline__ void sflux_const(PW &pw, const SwBand &B, double scale) { pw.sflux1(B.sec[SS_SFLUX] * B.rs, scale); }
template <class PW>
__device__ __forceinline__ void sflux_eta(PW &pw, const SwBand &B, const Eta &e)
{
    const int o = (B.sec[SS_SFLUX] + e.js - 1) * B.rs;
    pw.sflux2(o, 1. - e.fs, o + B.rs, e.fs);
}

#define IND0A(nsp) (((p.jp - 1) * 5 + (p.jt - 1)) * (nsp))
#define IND1A(nsp) ((p.jp * 5 + (p.jt1 - 1)) * (nsp))
#define IND0B(nsp) (((p.jp - 13) * 5 + (p.jt - 1)) * (nsp))
#define IND1B(nsp) (((p.jp - 12) * 5 + (p.jt1 - 1)) * (nsp))

__host__ __device__ constexpr int sw_ng(int band)
{
    constexpr int ng[14] = {6, 12, 8, 8, 10, 10, 2, 10, 8, 6, 6, 8, 6, 12};
    return ng[band];
}

// `solar` = this layer is the one whose values the reference leaves in sfluxzen for this band
template <int BAND, class PW>
__device__ __forceinline__ void sw_band_terms(const SwPair &p, bool lower, bool solar, PW &pw)
{
    const SwBand &B = c_sw.band[BAND];
    // Rayleigh: every band but 24 is colmol times one table row (scalar-rayl bands carry a row filled with
    // the scalar) and is evaluated by the solver; band 24 lower is eta-interpolated
    if constexpr (BAND == 0) { // band 16: 2600-3250, H2O/CH4 lower, CH4 upper (:243-339)
        if (lower) {
            const Eta e = binary(p.colh2o, 252.131, p.colch4, 8.);
            key8(pw, B, SS_ABSA, IND0A(9) + e.js, IND1A(9) + e.js, 9, e, p);
            selffor(pw, B, p, p.colh2o);
        } else {
            key4(pw, B, SS_ABSB, IND0B(1) + 1, IND1B(1) + 1, p.colch4, p);
            if (solar) sflux_const(pw, B, 1.0);
        }
    } else if constexpr (BAND == 1) { // band 17: 3250-4000, H2O/CO2 both (:342-462)
        if (lower) {
            const Eta e = binary(p.colh2o, 0.364641, p.colco2, 8.);
            key8(pw, B, SS_ABSA, IND0A(9) + e.js, IND1A(9) + e.js, 9, e, p);
            selffor(pw, B, p, p.colh2o);
        } else {
            const Eta e = binary(p.colh2o, 0.364641, p.colco2, 4.);
            key8(pw, B, SS_ABSB, IND0B(5) + e.js, IND1B(5) + e.js, 5, e, p);
            lerp2(pw, B, SS_FOR, p.indf, p.forfrac, p.colh2o * p.forfac);
            if (solar) sflux_eta(pw, B, e);
        }
    } else if constexpr (BAND == 2 || BAND == 3) { // band 18: 4000-4650 H2O/CH4, CH4 (:465-561); band 19: 4650-5150 H2O/CO2, CO2 (:564-660)
        const double strrat = BAND == 2 ? 38.9589 : 5.49281;
        const double colB = BAND == 2 ? p.colch4 : p.colco2;
        if (lower) {
            const Eta e = binary(p.colh2o, strrat, colB, 8.);
            key8(pw, B, SS_ABSA, IND0A(9) + e.js, IND1A(9) + e.js, 9, e, p);
            selffor(pw, B, p, p.colh2o);
            if (solar) sflux_eta(pw, B, e);
        } else {
            key4(pw, B, SS_ABSB, IND0B(1) + 1, IND1B(1) + 1, colB, p);
        }
    } else if constexpr (BAND == 4) { // band 20: 5150-6150, H2O + CH4 (:663-746)
        if (lower) {
            key4(pw, B, SS_ABSA, IND0A(1) + 1, IND1A(1) + 1, p.colh2o, p);
            selffor(pw, B, p, p.colh2o);
            if (solar) sflux_const(pw, B, 1.0);
        } else {
            key4(pw, B, SS_ABSB, IND0B(1) + 1, IND1B(1) + 1, p.colh2o, p);
            lerp2(pw, B, SS_FOR, p.indf, p.forfrac, p.colh2o * p.forfac);
        }
        pw.add(B.sec[SS_X1] * B.rs, p.colch4);
    } else if constexpr (BAND == 5) { // band 21: 6150-7700, H2O/CO2 both (:749-868)
        if (lower) {
            const Eta e = binary(p.colh2o, 0.0045321, p.colco2, 8.);
            key8(pw, B, SS_ABSA, IND0A(9) + e.js, IND1A(9) + e.js, 9, e, p);
            selffor(pw, B, p, p.colh2o);
            if (solar) sflux_eta(pw, B, e);
        } else {
            const Eta e = binary(p.colh2o, 0.0045321, p.colco2, 4.);
            key8(pw, B, SS_ABSB, IND0B(5) + e.js, IND1B(5) + e.js, 5, e, p);
            lerp2(pw, B, SS_FOR, p.indf, p.forfrac, p.colh2o * p.forfac);
        }
    } else if constexpr (BAND == 6) { // band 22: 7700-8050, H2O/O2 lower, O2 upper, O2 continuum (:871-977)
        const double o2adj = 1.6;
        if (lower) {
            const Eta e = binary(p.colh2o, o2adj * 0.022708, p.colo2, 8.);
            key8(pw, B, SS_ABSA, IND0A(9) + e.js, IND1A(9) + e.js, 9, e, p);
            selffor(pw, B, p, p.colh2o);
            if (solar) sflux_eta(pw, B, e);
        } else {
            key4(pw, B, SS_ABSB, IND0B(1) + 1, IND1B(1) + 1, p.colo2 * o2adj, p);
        }
        pw.addc(4.35e-4 * p.colo2 / (350.0 * 2.0));
    } else if constexpr (BAND == 7) { // band 23: 8050-12850, H2O lower (Giver factor), nothing above (:980-1051)
        if (lower) {
            key4(pw, B, SS_ABSA, IND0A(1) + 1, IND1A(1) + 1, p.colh2o * 1.029, p);
            selffor(pw, B, p, p.colh2o);
            if (solar) sflux_const(pw, B, 1.0);
        }
    } else if constexpr (BAND == 8) { // band 24: 12850-16000, H2O/O2 lower, O2 upper, O3 (:1054-1153)
        if (lower) {
            const Eta e = binary(p.colh2o, 0.124692, p.colo2, 8.);
            key8(pw, B, SS_ABSA, IND0A(9) + e.js, IND1A(9) + e.js, 9, e, p);
            pw.add(B.sec[SS_X1] * B.rs, p.colo3);
            selffor(pw, B, p, p.colh2o);
            if (solar) sflux_eta(pw, B, e);
            const int o = (B.sec[SS_RAYL] + e.js - 1) * B.rs;
            pw.rayl2(o, p.colmol * (1. - e.fs), o + B.rs, p.colmol * e.fs);
        } else {
            key4(pw, B, SS_ABSB, IND0B(1) + 1, IND1B(1) + 1, p.colo2, p);
            pw.add(B.sec[SS_X2] * B.rs, p.colo3);
            pw.rayl1(B.sec[SS_RAYLB] * B.rs, p.colmol);
        }
    } else if constexpr (BAND == 9) { // band 25: 16000-22650, H2O lower, O3 (:1156-1217)
        if (lower) {
            key4(pw, B, SS_ABSA, IND0A(1) + 1, IND1A(1) + 1, p.colh2o, p);
            pw.add(B.sec[SS_X1] * B.rs, p.colo3);
            if (solar) sflux_const(pw, B, 1.0);
        } else {
            pw.add(B.sec[SS_X2] * B.rs, p.colo3);
        }
    } else if constexpr (BAND == 10) { // band 26: 22650-29000, Rayleigh only (:1220-1268)
        if (lower && solar) sflux_const(pw, B, 1.0);
    } else if constexpr (BAND == 11) { // band 27: 29000-38000, O3 (:1271-1347)
        if (lower) {
            key4(pw, B, SS_ABSA, IND0A(1) + 1, IND1A(1) + 1, p.colo3, p);
        } else {
            key4(pw, B, SS_ABSB, IND0B(1) + 1, IND1B(1) + 1, p.colo3, p);
            if (solar) sflux_const(pw, B, 50.15 / 48.37);
        }
    } else if constexpr (BAND == 12) { // band 28: 38000-50000, O3/O2 both (:1350-1455)
        if (lower) {
            const Eta e = binary(p.colo3, 6.67029e-07, p.colo2, 8.);
            key8(pw, B, SS_ABSA, IND0A(9) + e.js, IND1A(9) + e.js, 9, e, p);
        } else {
            const Eta e = binary(p.colo3, 6.67029e-07, p.colo2, 4.);
            key8(pw, B, SS_ABSB, IND0B(5) + e.js, IND1B(5) + e.js, 5, e, p);
            if (solar) sflux_eta(pw, B, e);
        }
    } else { // band 29: 820-2600, H2O lower + CO2, CO2 upper + H2O (:1458-1536)
        if (lower) {
            key4(pw, B, SS_ABSA, IND0A(1) + 1, IND1A(1) + 1, p.colh2o, p);
            selffor(pw, B, p, p.colh2o);
            pw.add(B.sec[SS_X2] * B.rs, p.colco2);
        } else {
            key4(pw, B, SS_ABSB, IND0B(1) + 1, IND1B(1) + 1, p.colco2, p);
            pw.add(B.sec[SS_X1] * B.rs, p.colh2o);
            if (solar) sflux_const(pw, B, 1.0);
        }
    }
}

} // namespace rrtmg
