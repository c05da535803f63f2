// sw_kernels.cu -- RRTMG shortwave on sm_100a: prep (inatm_sw+setcoef_sw), taumol_sw (plan/execute),
// two-stream solver (spcvrt + reftra + vrtqdr, clear sky, no aerosol: the MiMA configuration).
//
//   sw_prep_kernel    thread <-> column: unit conversion, column amounts, p/T interpolation state, and the
//                     per-band layer whose binary-species parameter selects the solar source (laysolfr).
//   sw_taumol_kernel  tile = 128 adjacent columns of one layer; per band plan (thread <-> column) then
//                     execute (thread <-> (column, g)): taug, taur and (at the laysolfr layer) sfluxzen
//                     as weighted sums of table rows.
//   sw_solver_kernel  block <-> column, thread <-> g-point: layer optical properties + PIFM two-stream
//                     R/T (reftra) top-down, adding method bottom-up and top-down (vrtqdr), per-level
//                     shuffle reduction over g-points weighted by the incoming solar flux.
// Night columns (coszen < 1e-10) are skipped and written as zeros (rad.nomcica:502-510).
// Compiled with -fmad=false: fused multiply-adds appear only where written as fma().
#include "rrtmg_dev.cuh"

namespace rrtmg {

__constant__ SwConst c_sw;
__constant__ unsigned char c_sw_ngb[NGPTSW];

int sw_upload_const(const SwConst &c)
{
    unsigned char ngb[NGPTSW];
    for (int b = 0; b < NBNDSW; ++b)
        for (int i = 0; i < c.band[b].ng; ++i) ngb[c.band[b].g0 + i] = (unsigned char)b;
    if (cudaMemcpyToSymbol(c_sw, &c, sizeof(SwConst)) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(c_sw_ngb, ngb, sizeof ngb) != cudaSuccess) return -1;
    return 0;
}

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

// =====================================================================================================
// prep: per column -- night marker, laytrop (setcoef.f90), and the layer whose binary-species parameter
//       selects the solar source of each band (SW/src/rrtmg_sw_taumol.f90, "laysolfr" logic).
//       With w.f != nullptr (stage capture, test hook) the per-cell setcoef state is also written out.
// =====================================================================================================
__global__ void __launch_bounds__(128) sw_prep_kernel(SwIn in, SwWork w)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= w.nc) return;
    const int nlay = w.nlay, nc = w.nc;
    if (in.coszen[col] < ZEPZEN) {
        w.laytrop[col] = -1;    // night column marker
        return;
    }
    unsigned char jpv[MAXLAY + 2];
    int laytrop = 0;
    jpv[0] = 0;
    for (int l = 0; l < nlay; ++l) {
        SwPair p;
        if (sw_cell(in, col, l, p)) laytrop = laytrop + 1;
        jpv[l + 1] = (unsigned char)p.jp;
        if (w.f) {
            const size_t wo = (size_t)l * nc + col;
            w.idx[wo] = sw_pack(p.jp, p.jt, p.jt1, p.inds, p.indf);
            w.fld(SF_FAC00)[wo] = p.fac00; w.fld(SF_FAC01)[wo] = p.fac01;
            w.fld(SF_FAC10)[wo] = p.fac10; w.fld(SF_FAC11)[wo] = p.fac11;
            w.fld(SF_COLH2O)[wo] = p.colh2o; w.fld(SF_COLCO2)[wo] = p.colco2; w.fld(SF_COLO3)[wo] = p.colo3;
            w.fld(SF_COLCH4)[wo] = p.colch4; w.fld(SF_COLO2)[wo] = p.colo2; w.fld(SF_COLMOL)[wo] = p.colmol;
            w.fld(SF_COLN2O)[wo] = p.coln2o;
            w.fld(SF_SELFFAC)[wo] = p.selffac; w.fld(SF_SELFFRAC)[wo] = p.selffrac;
            w.fld(SF_FORFAC)[wo] = p.forfac; w.fld(SF_FORFRAC)[wo] = p.forfrac;
        }
    }
    jpv[nlay + 1] = 0;
    w.laytrop[col] = laytrop;

    // Which layer writes sfluxzen last, per band, replaying the sequential loops of taumolNN:
    //   'u' (upper loop): laysolfr = nlayers; if (jp(lay-1) < layreffr .and. jp(lay) >= layreffr) laysolfr = lay
    //   'l' (lower loop): laysolfr = laytrop; if (jp(lay) < layreffr .and. jp(lay+1) >= layreffr) laysolfr = min(lay+1,laytrop)
    // band:                16   17   18  19  20  21  22  23  24  25  26   27   28   29
    const int layreffr[14] = {18, 30, 6, 3, 3, 8, 2, 6, 1, 2, 0, 32, 58, 49};
    const char kind[14] = {'u', 'u', 'l', 'l', 'l', 'l', 'l', 'l', 'l', 'l', 'c', 'u', 'u', 'u'};
    int *ls = w.laysolfr + (size_t)col * 14;
    for (int b = 0; b < 14; ++b) {
        int last = 0;
        if (kind[b] == 'u') {
            int cur = nlay;
            for (int lay = laytrop + 1; lay <= nlay; ++lay) {
                if (jpv[lay - 1] < layreffr[b] && jpv[lay] >= layreffr[b]) cur = lay;
                if (lay == cur) last = lay;
            }
        } else if (kind[b] == 'l') {
            int cur = laytrop;
            for (int lay = 1; lay <= laytrop; ++lay) {
                if (jpv[lay] < layreffr[b] && jpv[lay + 1] >= layreffr[b]) cur = min(lay + 1, laytrop);
                if (lay == cur) last = lay;
            }
        } else {
            last = laytrop;    // band 26: written at lay == laytrop
        }
        ls[b] = last;
    }
}

// =====================================================================================================
// taumol_sw: SW/src/rrtmg_sw_taumol.f90:223-1536 (taumol16..29).  Same structure as the LW taumol kernel:
// thread <-> (column, layer) cell, register accumulators per band, per-warp slab transpose to
// [col][lay][g].  taur (Rayleigh) is a one- or two-row product and goes straight to the slab; the solar
// source sfluxzen is written by the single cell the reference leaves it from (laysolfr).
// =====================================================================================================
constexpr int TM_WARPS = 4;
constexpr int TM_STRIDE = 18;

template <int NG>
struct BandAcc {
    double t[NG];
    const double *__restrict__ tab;
    double *sr;                      // this lane's taur row in the slab
    double *sflx;                    // this cell's column slot in sfluxzen (+ g0)
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
    __device__ __forceinline__ void addc(double c)
    {
#pragma unroll
        for (int g = 0; g < NG; ++g) t[g] = t[g] + c;
    }
    __device__ __forceinline__ void rayl1(int off, double wgt)
    {
        const double2 *__restrict__ q = reinterpret_cast<const double2 *>(tab + off);
#pragma unroll
        for (int j = 0; j < NG / 2; ++j) {
            const double2 v = __ldg(q + j);
            reinterpret_cast<double2 *>(sr)[j] = make_double2(wgt * v.x, wgt * v.y);
        }
    }
    __device__ __forceinline__ void rayl2(int o0, double w0, int o1, double w1)
    {
        const double2 *__restrict__ q0 = reinterpret_cast<const double2 *>(tab + o0);
        const double2 *__restrict__ q1 = reinterpret_cast<const double2 *>(tab + o1);
#pragma unroll
        for (int j = 0; j < NG / 2; ++j) {
            const double2 a = __ldg(q0 + j), b = __ldg(q1 + j);
            reinterpret_cast<double2 *>(sr)[j] = make_double2(fma(w1, b.x, w0 * a.x), fma(w1, b.y, w0 * a.y));
        }
    }
    __device__ __forceinline__ void sflux1(int off, double wgt)
    {
#pragma unroll
        for (int g = 0; g < NG; ++g) sflx[g] = wgt * __ldg(tab + off + g);
    }
    __device__ __forceinline__ void sflux2(int o0, double w0, int o1, double w1)
    {
#pragma unroll
        for (int g = 0; g < NG; ++g) sflx[g] = fma(w1, __ldg(tab + o1 + g), w0 * __ldg(tab + o0 + g));
    }
};

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
    const int ng = B.ng, o0 = (B.sec[sec] + ind0 - 1) * ng, o1 = (B.sec[sec] + ind1 - 1) * ng;
    pw.add(o0, scale * p.fac00);
    pw.add(o0 + ng, scale * p.fac10);
    pw.add(o1, scale * p.fac01);
    pw.add(o1 + ng, scale * p.fac11);
}
// 8-point binary key term; one eta for both pressure levels; dT = 9 (lower) / 5 (upper)
template <class PW>
__device__ __forceinline__ void key8(PW &pw, const SwBand &B, int sec, int ind0, int ind1, int dT, const Eta &e, const SwPair &p)
{
    const int ng = B.ng, o0 = (B.sec[sec] + ind0 - 1) * ng, o1 = (B.sec[sec] + ind1 - 1) * ng;
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
    const int ng = B.ng, o = (B.sec[sec] + row - 1) * ng;
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
__device__ __forceinline__ void sflux_const(PW &pw, const SwBand &B, double scale) { pw.sflux1(B.sec[SS_SFLUX] * B.ng, scale); }
template <class PW>
__device__ __forceinline__ void sflux_eta(PW &pw, const SwBand &B, const Eta &e)
{
    const int o = (B.sec[SS_SFLUX] + e.js - 1) * B.ng;
    pw.sflux2(o, 1. - e.fs, o + B.ng, e.fs);
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
    // Rayleigh: scalar-rayl bands carry one row filled with the scalar; band 24 lower is eta-interpolated
    if constexpr (BAND != 8) pw.rayl1(B.sec[SS_RAYL] * B.ng, p.colmol);
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
        pw.add(B.sec[SS_X1] * B.ng, p.colch4);
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
            pw.add(B.sec[SS_X1] * B.ng, p.colo3);
            selffor(pw, B, p, p.colh2o);
            if (solar) sflux_eta(pw, B, e);
            const int o = (B.sec[SS_RAYL] + e.js - 1) * B.ng;
            pw.rayl2(o, p.colmol * (1. - e.fs), o + B.ng, p.colmol * e.fs);
        } else {
            key4(pw, B, SS_ABSB, IND0B(1) + 1, IND1B(1) + 1, p.colo2, p);
            pw.add(B.sec[SS_X2] * B.ng, p.colo3);
            pw.rayl1(B.sec[SS_RAYLB] * B.ng, p.colmol);
        }
    } else if constexpr (BAND == 9) { // band 25: 16000-22650, H2O lower, O3 (:1156-1217)
        if (lower) {
            key4(pw, B, SS_ABSA, IND0A(1) + 1, IND1A(1) + 1, p.colh2o, p);
            pw.add(B.sec[SS_X1] * B.ng, p.colo3);
            if (solar) sflux_const(pw, B, 1.0);
        } else {
            pw.add(B.sec[SS_X2] * B.ng, p.colo3);
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
            pw.add(B.sec[SS_X2] * B.ng, p.colco2);
        } else {
            key4(pw, B, SS_ABSB, IND0B(1) + 1, IND1B(1) + 1, p.colco2, p);
            pw.add(B.sec[SS_X1] * B.ng, p.colh2o);
            if (solar) sflux_const(pw, B, 1.0);
        }
    }
}

template <int BAND>
__device__ __forceinline__ void sw_band(const SwTables &T, const SwPair &p, bool valid, bool lower, int lay1,
                                        const int *__restrict__ laysolfr, double *slab,
                                        double *__restrict__ taug, double *__restrict__ taur, double *sflx_col,
                                        size_t cell0, size_t colstride, unsigned vmask)
{
    constexpr int NG = sw_ng(BAND);
    const int lane = threadIdx.x & 31;
    const SwBand &B = c_sw.band[BAND];
    const int g0 = B.g0;
    if (valid) {
        BandAcc<NG> pw;
        pw.tab = T.tab + B.base;
        pw.sr = slab + (32 + lane) * TM_STRIDE;
        pw.sflx = sflx_col + g0;
        pw.clear();
        sw_band_terms<BAND>(p, lower, laysolfr[BAND] == lay1, pw);
        double *st = slab + lane * TM_STRIDE;
#pragma unroll
        for (int j = 0; j < NG / 2; ++j) reinterpret_cast<double2 *>(st)[j] = make_double2(pw.t[2 * j], pw.t[2 * j + 1]);
    }
    __syncwarp();
    constexpr int HP = NG / 2;
#pragma unroll
    for (int i = lane; i < 32 * HP; i += 32) {
        const int c = i / HP, j = i - c * HP;
        if ((vmask >> c) & 1u) {
            const double2 a = reinterpret_cast<const double2 *>(slab + c * TM_STRIDE)[j];
            const double2 b = reinterpret_cast<const double2 *>(slab + (32 + c) * TM_STRIDE)[j];
            const size_t o = cell0 + (size_t)c * colstride + g0 + 2 * j;
            *reinterpret_cast<double2 *>(taug + o) = a;
            *reinterpret_cast<double2 *>(taur + o) = b;
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(32 * TM_WARPS) sw_taumol_kernel(SwTables T, SwIn in, SwWork w)
{
    __shared__ __align__(16) double s_slab[TM_WARPS][64 * TM_STRIDE];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 32;
    const int lay = blockIdx.y * TM_WARPS + wid;
    const int nlay = w.nlay, nc = w.nc;
    if (lay >= nlay) return;                         // no block-level barrier below
    const int col = c0 + lane;
    bool valid = col < nc;
    int laytrop = 0;
    if (valid) {
        laytrop = w.laytrop[col];
        if (laytrop < 0) valid = false;              // night column: the solver never reads its staging
    }
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    if (vmask == 0u) return;
    SwPair p;
    bool lower = false;
    if (valid) {
        sw_cell(in, col, lay, p);
        lower = (lay + 1) <= laytrop;
    }
    double *slab = s_slab[wid];
    const int *ls = w.laysolfr + (size_t)(valid ? col : 0) * 14;
    double *sflx = w.sfluxzen + (size_t)(valid ? col : 0) * NGPTSW;
    const size_t colstride = (size_t)nlay * NGPTSW;
    const size_t cell0 = ((size_t)c0 * nlay + lay) * NGPTSW;
#define SW_BAND(b) sw_band<b>(T, p, valid, lower, lay + 1, ls, slab, w.taug, w.taur, sflx, cell0, colstride, vmask)
    SW_BAND(0); SW_BAND(1); SW_BAND(2); SW_BAND(3); SW_BAND(4); SW_BAND(5); SW_BAND(6);
    SW_BAND(7); SW_BAND(8); SW_BAND(9); SW_BAND(10); SW_BAND(11); SW_BAND(12); SW_BAND(13);
#undef SW_BAND
}

// =====================================================================================================
// solver: spcvrt_sw (SW/src/rrtmg_sw_spcvrt.f90:296-619) + reftra_sw (rrtmg_sw_reftra.f90:129-300, kmodts=2)
//         + vrtqdr_sw (rrtmg_sw_vrtqdr.f90:103-150) + heating (rrtmg_sw_rad.nomcica.f90:686-727).
// icld = 0 and iaer = 0: the aerosol/cloud terms of the layer assembly are exact identities
// (tau_a = 0, omega_a = 1, g = 0 => delta scaling is the identity) and the total-sky stream equals the
// clear-sky stream bit for bit, so one stream is computed and stored to both outputs.
//
// Block <-> column, thread <-> g-point.  Two sweeps instead of the reference's four loops:
//   up   (surface -> top): layer R/T (reftra) fused with the bottom-up adding recurrence (vrtqdr :103-121);
//        keeps ref, refd, tra, trad, dbt per layer and rup, rupd per level in per-thread local arrays;
//   down (top -> surface): top-down recurrence (:125-140) fused with the level fluxes (:144-150) and the
//        spectral accumulation (spcvrt :570-619).  The lowest-layer and top-layer special cases of the
//        reference are the general formulas evaluated at rup = albedo resp. tdn = 1, rdnd = 0 (bitwise).
// The direct-beam transmittance of spcvrt :519-531 is the same table look-up as reftra's exp(-tau/mu0)
// (for tau/mu0 > 500 both hit the 1e-20 floor of exp_tbl), so it is taken from there.
// Divides go through rcp_fast/sqrt_fast; zbeta is folded into zdend's denominator.
// The sum over g-points goes through shared memory in batches of 8 levels (tile_reduce16).
// =====================================================================================================
constexpr int SV_THREADS = 128;   // 112 g-points -> 3.5 warps
constexpr int SV_S = 113;         // tile row stride (odd)

// exp(-ze) by the reference's Pade-indexed table (ze > od_lo) or 2nd-order series; also returns exp(+ze)
__device__ __forceinline__ double sw_exp(const double2 *__restrict__ tb, double ze, double bpade, double &recip)
{
    if (ze <= 0.06) {
        const double em = 1. - ze + 0.5 * ze * ze;
        recip = rcp_fast(em);
        return em;
    }
    const double tblind = ze * rcp_fast(bpade + ze);
    const int itind = (int)(10000.0 * tblind + 0.5);
    const double2 e = __ldg(tb + itind);
    recip = e.y;
    return e.x;
}

template <int LMAX>
__global__ void __launch_bounds__(SV_THREADS) sw_solver_kernel(SwTables T, SwIn in, SwOut out, SwWork w)
{
    __shared__ double s_tile[16 * SV_S];
    __shared__ double s_part[16 * (SV_THREADS / 16 + 1)];
    __shared__ double s_up[LMAX + 1], s_dn[LMAX + 1];
    const int col = blockIdx.x;
    const int klev = w.nlay;
    const int g = threadIdx.x;
    const size_t old = (size_t)out.ld;

    const double prmu0 = in.coszen[col];
    if (prmu0 < ZEPZEN) {
        // night column: zero everything (rad.nomcica:502-510)
        for (int lev = threadIdx.x; lev <= klev; lev += SV_THREADS) {
            const size_t o = col + (size_t)lev * old;
            out.uflx[o] = 0.; out.dflx[o] = 0.; out.uflxc[o] = 0.; out.dflxc[o] = 0.;
            if (lev < klev) { out.hr[o] = 0.; out.hrc[o] = 0.; }
        }
        return;
    }
    const bool active = g < NGPTSW;
    const int band = active ? c_sw_ngb[g] : 0;
    const double bpade = c_sw.bpade;
    const double eps = 1.e-08, zwcrit = 0.9999995;
    const double2 *__restrict__ tb = reinterpret_cast<const double2 *>(T.exptbl);
    const double rmu0 = 1. / prmu0;

    // band albedos (rad.nomcica:565-578): bands 16-24 and 29 near-IR, 25-28 UV/visible
    const bool uvvis = band >= 9 && band <= 12;
    const double albd = uvvis ? in.asdif[col] : in.aldif[col];   // palbd: diffuse
    const double albp = uvvis ? in.asdir[col] : in.aldir[col];   // palbp: direct

    // per-thread state, index = layer / level counted from the surface
    double zref[LMAX], zrefd[LMAX], ztra[LMAX], ztrad[LMAX], zdbt[LMAX];
    double zrup[LMAX + 1], zrupd[LMAX + 1];
    double zincflx = 0.0;

    if (active) {
        zincflx = in.adjflux * w.sfluxzen[(size_t)col * NGPTSW + g] * prmu0;
        const double *taug = w.taug + (size_t)col * klev * NGPTSW + g;
        const double *taur = w.taur + (size_t)col * klev * NGPTSW + g;
        double rup = albp, rupd = albd;      // zrup(klev+1) = palbp, zrupd(klev+1) = palbd
        zrup[0] = rup;
        zrupd[0] = rupd;
#pragma unroll 2
        for (int l = 0; l < klev; ++l) {
            const double tr = taur[(size_t)l * NGPTSW];
            const double zto1 = tr + taug[(size_t)l * NGPTSW];      // ztauc
            const double zw = tr * rcp_fast(zto1);                   // zomcc
            // ---- reftra, zg = 0: gamma3 = gamma4 = 1/2, zwo = zw
            const double zgamma1 = (8. - zw * 5.) * 0.25;
            const double zgamma2 = 3. * zw * 0.25;
            const double zed = zto1 * rmu0;                          // direct-beam optical path
            double ref, refd, tra, trad, dbt;
            if (zw >= zwcrit) {
                // conservative scattering (:162-214)
                const double za1 = zgamma1 * prmu0 - 0.5;
                const double zgt = zgamma1 * zto1;
                double rcp;
                const double ze2 = sw_exp(tb, fmin(zed, 500.), bpade, rcp);
                const double rg = rcp_fast(1. + zgt);
                ref = (zgt - za1 * (1. - ze2)) * rg;
                tra = 1. - ref;
                refd = zgt * rg;
                trad = 1. - refd;
                if (ze2 == 1.0) { ref = 0.0; tra = 1.0; refd = 0.0; trad = 1.0; }
                dbt = ze2;
            } else {
                const double za1 = (zgamma1 + zgamma2) * 0.5;        // = za2
                const double zrk = sqrt_fast(zgamma1 * zgamma1 - zgamma2 * zgamma2);
                const double zrp = zrk * prmu0;
                const double zrp1 = 1. + zrp;
                const double zrm1 = 1. - zrp;
                const double zrk2 = 2. * zrk;
                const double zrpp = 1. - zrp * zrp;
                const double zrkg = zrk + zgamma1;
                const double hA = fma(zrk, 0.5, za1), hB = fma(zrk, -0.5, za1);
                const double zr1 = zrm1 * hA;
                const double zr2 = zrp1 * hB;
                const double zr3 = zrk2 * (0.5 - za1 * prmu0);
                const double zr4 = zrpp * zrkg;
                const double zr5 = zrpp * (zrk - zgamma1);
                const double zt1 = zrp1 * hA;
                const double zt2 = zrm1 * hB;
                const double zt3 = zrk2 * (0.5 + za1 * prmu0);
                double zep1, zep2;
                const double zem1 = sw_exp(tb, fmin(zrk * zto1, 500.), bpade, zep1);
                const double zem2 = sw_exp(tb, fmin(zed, 500.), bpade, zep2);
                const double zdenr = fma(zr4, zep1, zr5 * zem1);     // = zdent (zt4 = zr4, zt5 = zr5)
                if (zdenr >= -eps && zdenr <= eps) {
                    ref = eps;
                    tra = zem2;
                } else {
                    const double rd = zw * rcp_fast(zdenr);
                    ref = (zr1 * zep1 - zr2 * zem1 - zr3 * zem2) * rd;
                    tra = zem2 - zem2 * ((zt1 * zep1 - zt2 * zem1 - zt3 * zep2) * rd);
                }
                const double zemm = zem1 * zem1;
                // zdend = 1/((1 - zbeta*zemm)*zrkg), zbeta = (gamma1 - zrk)/zrkg
                const double zdend = rcp_fast(fma(-(zgamma1 - zrk), zemm, zrkg));
                refd = zgamma2 * (1. - zemm) * zdend;
                trad = zrk2 * zem1 * zdend;
                dbt = zem2;
            }
            zref[l] = ref; zrefd[l] = refd; ztra[l] = tra; ztrad[l] = trad; zdbt[l] = dbt;
            // ---- vrtqdr, bottom -> top (:103-121)
            const double zreflect = rcp_fast(1. - rupd * refd);
            const double rup_n = ref + (trad * ((tra - dbt) * rupd + dbt * rup)) * zreflect;
            const double rupd_n = refd + trad * trad * rupd * zreflect;
            rup = rup_n;
            rupd = rupd_n;
            zrup[l + 1] = rup;
            zrupd[l + 1] = rupd;
        }
    }

    // ---- top -> bottom: ztdn, prdnd, cumulative direct beam; fluxes at every level (:125-150)
    double ztdn = 1., zrdnd = 0., ztdbt = 1.;
    for (int k = 0; k <= klev; ++k) {
        const int s = klev - k;            // level counted from the surface
        const int slot = k & 7;
        if (active) {
            const double ru = zrup[s], rud = zrupd[s];
            const double zreflect = rcp_fast(1. - zrdnd * rud);
            const double dif = ztdn - ztdbt;
            const double pfu = (ztdbt * ru + dif * rud) * zreflect;
            const double pfd = ztdbt + (dif + ztdbt * ru * zrdnd) * zreflect;
            s_tile[(2 * slot) * SV_S + g] = zincflx * pfu;
            s_tile[(2 * slot + 1) * SV_S + g] = zincflx * pfd;
            if (s > 0) {
                const int l = s - 1;
                const double ref = zref[l], refd = zrefd[l], tra = ztra[l], trad = ztrad[l], dbt = zdbt[l];
                const double zr = rcp_fast(1. - refd * zrdnd);
                const double ztdn_n = ztdbt * tra + (trad * (dif + ztdbt * ref * zrdnd)) * zr;
                const double zrdnd_n = refd + trad * trad * zrdnd * zr;
                ztdbt = dbt * ztdbt;
                ztdn = ztdn_n;
                zrdnd = zrdnd_n;
            }
        }
        if (slot == 7 || k == klev) {
            const double sum = tile_reduce16<SV_THREADS, NGPTSW, SV_S>(s_tile, s_part);
            if (threadIdx.x < 16) {
                const int kk = (k & ~7) + (threadIdx.x >> 1);
                if (kk <= k) {
                    if (threadIdx.x & 1) s_dn[klev - kk] = sum;
                    else s_up[klev - kk] = sum;
                }
            }
        }
    }
    __syncthreads();
    for (int lev = threadIdx.x; lev <= klev; lev += SV_THREADS) {
        const double u = s_up[lev], d = s_dn[lev];
        const size_t o = col + (size_t)lev * old;
        out.uflx[o] = u; out.dflx[o] = d; out.uflxc[o] = u; out.dflxc[o] = d;
    }
    for (int lay = threadIdx.x; lay < klev; lay += SV_THREADS) {
        const size_t o = col + (size_t)lay * old;
        double h = 0.0;
        if (lay < klev - 1) {      // MiMA: no heating in the top layer (rad.nomcica:724-726)
            const double pdp = in.plev[col + (size_t)lay * in.ld] - in.plev[col + (size_t)(lay + 1) * in.ld];
            h = ((s_dn[lay + 1] - s_up[lay + 1]) - (s_dn[lay] - s_up[lay])) * (c_sw.heatfac / pdp);
        }
        out.hr[o] = h;
        out.hrc[o] = h;
    }
}

int sw_run_pass(const SwTables &t, const SwIn &in, const SwOut &out, SwWork &w, cudaStream_t s)
{
    ktimer_begin(K_SW_PREP, s);
    sw_prep_kernel<<<(w.nc + 127) / 128, 128, 0, s>>>(in, w);
    ktimer_end(s);
    dim3 grid((w.nc + 31) / 32, (w.nlay + TM_WARPS - 1) / TM_WARPS);
    ktimer_begin(K_SW_TAUMOL, s);
    sw_taumol_kernel<<<grid, 32 * TM_WARPS, 0, s>>>(t, in, w);
    ktimer_end(s);
    ktimer_begin(K_SW_SOLVER, s);
    {
        const size_t pad = (size_t)g_tune.sw_solver_pad_kb * 1024;
        if (w.nlay <= 64) {
            cudaFuncSetAttribute(sw_solver_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
            sw_solver_kernel<64><<<w.nc, SV_THREADS, pad, s>>>(t, in, out, w);
        } else {
            cudaFuncSetAttribute(sw_solver_kernel<MAXLAY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
            sw_solver_kernel<MAXLAY><<<w.nc, SV_THREADS, pad, s>>>(t, in, out, w);
        }
    }
    ktimer_end(s);
    return 3;
}

} // namespace rrtmg
