/*
 * sw.c -- oracle restatement of RRTMG_SW as linked by MiMA: clear sky (icld=0, iaer=0: what MiMA runs), plus the
 * input-optical-property branches of the interface, iaer=10 and icld>=1 with inflgsw=0 (overcast or clear layers only).
 * TEST INFRASTRUCTURE ONLY (see rrtmg_oracle.h).
 *
 * Follows  SW/src/rrtmg_sw_rad.nomcica.f90:78-731 (rrtmg_sw), :734-758 (earth_sun), :761-1101 (inatm_sw)
 *          SW/src/rrtmg_sw_cldprop.f90:122-130     (clear sky: taucloud=0, ssacloud=1, asmcloud=0)
 *          SW/src/rrtmg_sw_setcoef.f90:30-286      (setcoef_sw)
 *          SW/src/rrtmg_sw_taumol.f90:31-1538      (taumol_sw, taumol16..29)
 *          SW/src/rrtmg_sw_spcvrt.f90:255-626      (spcvrt_sw)
 *          SW/src/rrtmg_sw_reftra.f90:122-303      (reftra_sw)
 *          SW/src/rrtmg_sw_vrtqdr.f90:103-150      (vrtqdr_sw)
 * Both the clear and the total-sky stream are computed exactly as the Fortran does (two reftra calls,
 * two vrtqdr calls) even though they coincide for icld=0.
 */
#include "rrtmg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NL (ORC_MAXLAY + 3)
#define F2(p, n1, i, j) ((p)[((long)(j) - 1) * (n1) + ((i) - 1)])

static const int nspa[14] = {9, 9, 9, 9, 1, 9, 9, 1, 9, 1, 0, 1, 9, 1};
static const int nspb[14] = {1, 5, 1, 1, 1, 5, 1, 0, 1, 0, 0, 1, 5, 1};
static const int ngc[14] = {6, 12, 8, 8, 10, 10, 2, 10, 8, 6, 6, 8, 6, 12};
static const int ngs[14] = {6, 18, 26, 34, 44, 54, 56, 66, 74, 80, 86, 94, 100, 112};

typedef struct {
    int nlayers, laytrop, layswtch, laylow;
    double pavel[NL], tavel[NL], pz[NL], tz[NL], pdp[NL], tbound, coldry[NL];
    double wkl[8][NL], adjflux[30];
    int jp[NL], jt[NL], jt1[NL], indself[NL], indfor[NL];
    double colh2o[NL], colco2[NL], colo3[NL], coln2o[NL], colch4[NL], colo2[NL], colmol[NL], co2mult[NL];
    double fac00[NL], fac01[NL], fac10[NL], fac11[NL];
    double selffac[NL], selffrac[NL], forfac[NL], forfrac[NL];
    double ztaug[ORC_NGPTSW + 1][NL], ztaur[ORC_NGPTSW + 1][NL], zsflxzen[ORC_NGPTSW + 1];
    double zbbfd[NL], zbbfu[NL], zbbcd[NL], zbbcu[NL];
    double zbbfddir[NL], zbbcddir[NL], zuvfd[NL], zuvcd[NL], znifd[NL], znicd[NL];
    double zuvfddir[NL], zuvcddir[NL], znifddir[NL], znicddir[NL];
    double oneminus;
    /* cloud and aerosol optical properties per (layer, band 1..14) as spcvrt_sw receives them (rad.nomcica:581-640) */
    double pclfr[NL], ptauc[NL][15], pasyc[NL][15], pomgc[NL][15], ptaua[NL][15], pasya[NL][15], pomga[NL][15];
} swcol_t;

/* earth_sun (rad.nomcica:734-758) */
static double earth_sun(int idn)
{
    double pi = 2. * asin(1.);
    double gamma = 2. * pi * (idn - 1) / 365.;
    return 1.000110 + .034221 * cos(gamma) + .001289 * sin(gamma) + .000719 * cos(2. * gamma) + .000077 * sin(2. * gamma);
}

/* ---------------------------------------------------------------- inatm_sw (rad.nomcica:761-1101) */
static void inatm_sw(swcol_t *c, int iplon, int ncol, int nlay,
                     const double *play, const double *plev, const double *tlay, const double *tlev,
                     const double *tsfc, const double *h2ovmr, const double *o3vmr, const double *co2vmr,
                     const double *ch4vmr, const double *n2ovmr, const double *o2vmr,
                     double adjes, int dyofyr, double scon)
{
    const double amd = 28.9660, amw = 18.0160, amdw = 1.607793, amdo = 0.603428;
    const double grav = 9.8066, avogad = 6.02214199e+23, rrsw_scon = 1.36822e+03;
    const int nmol = 7;
    double amm, adjflx;
#define IN2(a, l) ((a)[((long)(l) - 1) * ncol + (iplon - 1)])
    c->nlayers = nlay;
    for (int m = 0; m < 8; ++m)
        for (int l = 0; l < NL; ++l) c->wkl[m][l] = 0.0;
    adjflx = adjes;
    if (dyofyr > 0) adjflx = earth_sun(dyofyr);
    for (int ib = 16; ib <= 29; ++ib) {
        double solvar = scon / rrsw_scon;
        c->adjflux[ib] = adjflx * solvar;
    }
    c->tbound = tsfc[iplon - 1];
    c->pz[0] = IN2(plev, 1);
    c->tz[0] = IN2(tlev, 1);
    for (int l = 1; l <= nlay; ++l) {
        c->pavel[l] = IN2(play, l);
        c->tavel[l] = IN2(tlay, l);
        c->pz[l] = IN2(plev, l + 1);
        c->tz[l] = IN2(tlev, l + 1);
        c->pdp[l] = c->pz[l - 1] - c->pz[l];
        c->wkl[1][l] = (IN2(h2ovmr, l) / (1. - IN2(h2ovmr, l))) * amdw;
        c->wkl[2][l] = IN2(co2vmr, l);
        c->wkl[3][l] = IN2(o3vmr, l) * amdo;
        c->wkl[4][l] = IN2(n2ovmr, l);
        c->wkl[6][l] = IN2(ch4vmr, l);
        c->wkl[7][l] = IN2(o2vmr, l);
        amm = (1. - c->wkl[1][l]) * amd + c->wkl[1][l] * amw;
        c->coldry[l] = (c->pz[l - 1] - c->pz[l]) * 1.e3 * avogad / (1.e2 * grav * amm * (1. + c->wkl[1][l]));
    }
    for (int l = 1; l <= nlay; ++l)
        for (int imol = 1; imol <= nmol; ++imol) c->wkl[imol][l] = c->coldry[l] * c->wkl[imol][l];
#undef IN2
}

/* ---------------------------------------------------------------- setcoef_sw (setcoef.f90:30-286) */
static void setcoef_sw(swcol_t *c)
{
    const orc_state_t *S = &g_orc;
    const int nlayers = c->nlayers;
    const double stpfac = 296. / 1013.;
    double plog, fp, ft, ft1, water, scalefac, factor, co2reg, compfp;
    int jp1;
    c->laytrop = 0;
    c->layswtch = 0;
    c->laylow = 0;
    /* taumol reads jp(lay-1) at lay = laytrop+1 and jp(lay+1) at lay = laytrop (taumol.f90:321, :481): outside
     * 1..nlayers only when laytrop is 0 or nlayers, where the Fortran reads undefined storage.  Defined as 0
     * here (the product does the same) so that the oracle is deterministic. */
    c->jp[0] = 0;
    c->jp[nlayers + 1] = 0;
    for (int lay = 1; lay <= nlayers; ++lay) {
        plog = log(c->pavel[lay]);
        c->jp[lay] = (int)(36. - 5 * (plog + 0.04));
        if (c->jp[lay] < 1) c->jp[lay] = 1;
        else if (c->jp[lay] > 58) c->jp[lay] = 58;
        jp1 = c->jp[lay] + 1;
        fp = 5. * (S->sw_preflog[c->jp[lay] - 1] - plog);
        c->jt[lay] = (int)(3. + (c->tavel[lay] - S->sw_tref[c->jp[lay] - 1]) / 15.);
        if (c->jt[lay] < 1) c->jt[lay] = 1;
        else if (c->jt[lay] > 4) c->jt[lay] = 4;
        ft = ((c->tavel[lay] - S->sw_tref[c->jp[lay] - 1]) / 15.) - (double)(c->jt[lay] - 3);
        c->jt1[lay] = (int)(3. + (c->tavel[lay] - S->sw_tref[jp1 - 1]) / 15.);
        if (c->jt1[lay] < 1) c->jt1[lay] = 1;
        else if (c->jt1[lay] > 4) c->jt1[lay] = 4;
        ft1 = ((c->tavel[lay] - S->sw_tref[jp1 - 1]) / 15.) - (double)(c->jt1[lay] - 3);
        water = c->wkl[1][lay] / c->coldry[lay];
        scalefac = c->pavel[lay] * stpfac / c->tavel[lay];

        if (!(plog <= 4.56)) {
            c->laytrop = c->laytrop + 1;
            if (plog >= 6.62) c->laylow = c->laylow + 1;
            c->forfac[lay] = scalefac / (1. + water);
            factor = (332.0 - c->tavel[lay]) / 36.0;
            {
                int t = (int)factor;
                c->indfor[lay] = t < 1 ? 1 : (t > 2 ? 2 : t);
            }
            c->forfrac[lay] = factor - (double)c->indfor[lay];
            c->selffac[lay] = water * c->forfac[lay];
            factor = (c->tavel[lay] - 188.0) / 7.2;
            {
                int t = (int)factor - 7;
                c->indself[lay] = t < 1 ? 1 : (t > 9 ? 9 : t);
            }
            c->selffrac[lay] = factor - (double)(c->indself[lay] + 7);
        } else {
            c->forfac[lay] = scalefac / (1. + water);
            factor = (c->tavel[lay] - 188.0) / 36.0;
            c->indfor[lay] = 3;
            c->forfrac[lay] = factor - 1.0;
        }
        c->colh2o[lay] = 1.e-20 * c->wkl[1][lay];
        c->colco2[lay] = 1.e-20 * c->wkl[2][lay];
        c->colo3[lay] = 1.e-20 * c->wkl[3][lay];
        c->coln2o[lay] = 1.e-20 * c->wkl[4][lay];
        c->colch4[lay] = 1.e-20 * c->wkl[6][lay];
        c->colo2[lay] = 1.e-20 * c->wkl[7][lay];
        c->colmol[lay] = 1.e-20 * c->coldry[lay] + c->colh2o[lay];
        if (c->colco2[lay] == 0.) c->colco2[lay] = 1.e-32 * c->coldry[lay];
        if (c->coln2o[lay] == 0.) c->coln2o[lay] = 1.e-32 * c->coldry[lay];
        if (c->colch4[lay] == 0.) c->colch4[lay] = 1.e-32 * c->coldry[lay];
        if (c->colo2[lay] == 0.) c->colo2[lay] = 1.e-32 * c->coldry[lay];
        co2reg = 3.55e-24 * c->coldry[lay];
        c->co2mult[lay] = (c->colco2[lay] - co2reg) * 272.63 * exp(-1919.4 / c->tavel[lay]) / (8.7604e-4 * c->tavel[lay]);
        if (plog <= 4.56) {
            c->selffac[lay] = 0.;
            c->selffrac[lay] = 0.;
            c->indself[lay] = 0;
        }
        compfp = 1. - fp;
        c->fac10[lay] = compfp * ft;
        c->fac00[lay] = compfp * (1. - ft);
        c->fac11[lay] = fp * ft1;
        c->fac01[lay] = fp * (1. - ft1);
    }
}

/* ---------------------------------------------------------------- taumol_sw helpers */
#define IND0A(b) (((c->jp[lay] - 1) * 5 + (c->jt[lay] - 1)) * nspa[(b) - 16])
#define IND1A(b) ((c->jp[lay] * 5 + (c->jt1[lay] - 1)) * nspa[(b) - 16])
#define IND0B(b) (((c->jp[lay] - 13) * 5 + (c->jt[lay] - 1)) * nspb[(b) - 16])
#define IND1B(b) (((c->jp[lay] - 12) * 5 + (c->jt1[lay] - 1)) * nspb[(b) - 16])

#define SELFT(K, lay, inds, ig) \
    (c->selffac[lay] * (F2((K)->selfref, 10, inds, ig) + c->selffrac[lay] * \
        (F2((K)->selfref, 10, (inds) + 1, ig) - F2((K)->selfref, 10, inds, ig))))
#define FORT(K, lay, indf, ig) \
    (c->forfac[lay] * (F2((K)->forref, (K)->nfor, indf, ig) + c->forfrac[lay] * \
        (F2((K)->forref, (K)->nfor, (indf) + 1, ig) - F2((K)->forref, (K)->nfor, indf, ig))))
#define SELFFOR(K, lay, inds, indf, ig) (SELFT(K, lay, inds, ig) + FORT(K, lay, indf, ig))
#define FORONLY(K, lay, indf, ig) \
    (c->forfac[lay] * (F2((K)->forref, (K)->nfor, indf, ig) + c->forfrac[lay] * \
        (F2((K)->forref, (K)->nfor, (indf) + 1, ig) - F2((K)->forref, (K)->nfor, indf, ig))))
/* upper-atmosphere foreign continuum of bands 17 and 21: colh2o * forfac * (...), evaluated left to right
 * (taumol.f90:451-454, :859-862) -- not colh2o * (forfac * (...)) as in the band-20 form above (:732-741) */
#define FORUP(K, lay, indf, ig) \
    (c->colh2o[lay] * c->forfac[lay] * (F2((K)->forref, (K)->nfor, indf, ig) + c->forfrac[lay] * \
        (F2((K)->forref, (K)->nfor, (indf) + 1, ig) - F2((K)->forref, (K)->nfor, indf, ig))))
#define KEY4(abs_, n_, ind0, ind1, ig) \
    (c->fac00[lay] * F2(abs_, n_, ind0, ig) + c->fac10[lay] * F2(abs_, n_, (ind0) + 1, ig) + \
     c->fac01[lay] * F2(abs_, n_, ind1, ig) + c->fac11[lay] * F2(abs_, n_, (ind1) + 1, ig))

typedef struct { double speccomb, fs; int js; double f000, f010, f100, f110, f001, f011, f101, f111; } swbin_t;

/* binary-species setup shared by all SW binary bands (e.g. taumol.f90:270-283) */
static swbin_t sw_binary(const swcol_t *c, int lay, double colA, double strrat, double colB, double mult)
{
    swbin_t b;
    double specparm, specmult;
    b.speccomb = colA + strrat * colB;
    specparm = colA / b.speccomb;
    if (specparm >= c->oneminus) specparm = c->oneminus;
    specmult = mult * (specparm);
    b.js = 1 + (int)specmult;
    b.fs = fmod(specmult, 1.);
    b.f000 = (1. - b.fs) * c->fac00[lay];
    b.f010 = (1. - b.fs) * c->fac10[lay];
    b.f100 = b.fs * c->fac00[lay];
    b.f110 = b.fs * c->fac10[lay];
    b.f001 = (1. - b.fs) * c->fac01[lay];
    b.f011 = (1. - b.fs) * c->fac11[lay];
    b.f101 = b.fs * c->fac01[lay];
    b.f111 = b.fs * c->fac11[lay];
    return b;
}
/* 8-point key-species sum; dT = 9 (lower) or 5 (upper) */
static double key8(const swbin_t *b, const double *abs_, int n, int ind0, int ind1, int dT, int ig)
{
    return b->f000 * F2(abs_, n, ind0, ig) + b->f100 * F2(abs_, n, ind0 + 1, ig) +
           b->f010 * F2(abs_, n, ind0 + dT, ig) + b->f110 * F2(abs_, n, ind0 + dT + 1, ig) +
           b->f001 * F2(abs_, n, ind1, ig) + b->f101 * F2(abs_, n, ind1 + 1, ig) +
           b->f011 * F2(abs_, n, ind1 + dT, ig) + b->f111 * F2(abs_, n, ind1 + dT + 1, ig);
}

/* ---------------------------------------------------------------- taumol_sw (taumol.f90:223-1536) */
static void taumol_sw(swcol_t *c)
{
    const orc_state_t *S = &g_orc;
    const int nlayers = c->nlayers, laytrop = c->laytrop;
    int lay, ig, ind0, ind1, inds, indf, laysolfr, layreffr;
    double tauray;
    const orc_sw_kg_t *K;
    swbin_t b;
#define TAUG(off, ig) c->ztaug[(off) + (ig)][lay]
#define TAUR(off, ig) c->ztaur[(off) + (ig)][lay]
#define SFLX(off, ig) c->zsflxzen[(off) + (ig)]

    /* ---- band 16: 2600-3250 (h2o,ch4; ch4) (:243-339) */
    K = &S->sw[0];
    {
        const double strrat1 = 252.131;
        layreffr = 18;
        for (lay = 1; lay <= laytrop; ++lay) {
            b = sw_binary(c, lay, c->colh2o[lay], strrat1, c->colch4[lay], 8.);
            ind0 = IND0A(16) + b.js; ind1 = IND1A(16) + b.js;
            inds = c->indself[lay]; indf = c->indfor[lay];
            tauray = c->colmol[lay] * K->rayl;
            for (ig = 1; ig <= ngc[0]; ++ig) {
                TAUG(0, ig) = b.speccomb * key8(&b, K->absa, 585, ind0, ind1, 9, ig) + c->colh2o[lay] * SELFFOR(K, lay, inds, indf, ig);
                TAUR(0, ig) = tauray;
            }
        }
        laysolfr = nlayers;
        for (lay = laytrop + 1; lay <= nlayers; ++lay) {
            if (c->jp[lay - 1] < layreffr && c->jp[lay] >= layreffr) laysolfr = lay;
            ind0 = IND0B(16) + 1; ind1 = IND1B(16) + 1;
            tauray = c->colmol[lay] * K->rayl;
            for (ig = 1; ig <= ngc[0]; ++ig) {
                TAUG(0, ig) = c->colch4[lay] * KEY4(K->absb, 235, ind0, ind1, ig);
                if (lay == laysolfr) SFLX(0, ig) = K->sfluxref[ig - 1];
                TAUR(0, ig) = tauray;
            }
        }
    }

    /* ---- band 17: 3250-4000 (h2o,co2; h2o,co2) (:342-462) */
    K = &S->sw[1];
    {
        const double strrat = 0.364641;
        const int o = ngs[0];
        layreffr = 30;
        for (lay = 1; lay <= laytrop; ++lay) {
            b = sw_binary(c, lay, c->colh2o[lay], strrat, c->colco2[lay], 8.);
            ind0 = IND0A(17) + b.js; ind1 = IND1A(17) + b.js;
            inds = c->indself[lay]; indf = c->indfor[lay];
            tauray = c->colmol[lay] * K->rayl;
            for (ig = 1; ig <= ngc[1]; ++ig) {
                TAUG(o, ig) = b.speccomb * key8(&b, K->absa, 585, ind0, ind1, 9, ig) + c->colh2o[lay] * SELFFOR(K, lay, inds, indf, ig);
                TAUR(o, ig) = tauray;
            }
        }
        laysolfr = nlayers;
        for (lay = laytrop + 1; lay <= nlayers; ++lay) {
            if (c->jp[lay - 1] < layreffr && c->jp[lay] >= layreffr) laysolfr = lay;
            b = sw_binary(c, lay, c->colh2o[lay], strrat, c->colco2[lay], 4.);
            ind0 = IND0B(17) + b.js; ind1 = IND1B(17) + b.js;
            indf = c->indfor[lay];
            tauray = c->colmol[lay] * K->rayl;
            for (ig = 1; ig <= ngc[1]; ++ig) {
                TAUG(o, ig) = b.speccomb * key8(&b, K->absb, 1175, ind0, ind1, 5, ig) + FORUP(K, lay, indf, ig);
                if (lay == laysolfr)
                    SFLX(o, ig) = F2(K->sfluxref, 12, ig, b.js) + b.fs * (F2(K->sfluxref, 12, ig, b.js + 1) - F2(K->sfluxref, 12, ig, b.js));
                TAUR(o, ig) = tauray;
            }
        }
    }

    /* ---- bands 18, 19: (h2o,ch4; ch4) / (h2o,co2; co2) (:465-561, :564-660) */
    for (int bb = 18; bb <= 19; ++bb) {
        K = &S->sw[bb - 16];
        const double strrat = (bb == 18) ? 38.9589 : 5.49281;
        const int o = ngs[bb - 17], n = ngc[bb - 16];
        const double *colB = (bb == 18) ? c->colch4 : c->colco2;
        layreffr = (bb == 18) ? 6 : 3;
        laysolfr = laytrop;
        for (lay = 1; lay <= laytrop; ++lay) {
            if (c->jp[lay] < layreffr && c->jp[lay + 1] >= layreffr) laysolfr = (lay + 1 < laytrop) ? lay + 1 : laytrop;
            b = sw_binary(c, lay, c->colh2o[lay], strrat, colB[lay], 8.);
            ind0 = IND0A(bb) + b.js; ind1 = IND1A(bb) + b.js;
            inds = c->indself[lay]; indf = c->indfor[lay];
            tauray = c->colmol[lay] * K->rayl;
            for (ig = 1; ig <= n; ++ig) {
                TAUG(o, ig) = b.speccomb * key8(&b, K->absa, 585, ind0, ind1, 9, ig) + c->colh2o[lay] * SELFFOR(K, lay, inds, indf, ig);
                if (lay == laysolfr)
                    SFLX(o, ig) = F2(K->sfluxref, n, ig, b.js) + b.fs * (F2(K->sfluxref, n, ig, b.js + 1) - F2(K->sfluxref, n, ig, b.js));
                TAUR(o, ig) = tauray;
            }
        }
        for (lay = laytrop + 1; lay <= nlayers; ++lay) {
            ind0 = IND0B(bb) + 1; ind1 = IND1B(bb) + 1;
            tauray = c->colmol[lay] * K->rayl;
            for (ig = 1; ig <= n; ++ig) {
                TAUG(o, ig) = colB[lay] * KEY4(K->absb, 235, ind0, ind1, ig);
                TAUR(o, ig) = tauray;
            }
        }
    }

    /* ---- band 20: 5150-6150 (h2o; h2o) + CH4 (:663-746) */
    K = &S->sw[4];
    {
        const int o = ngs[3];
        layreffr = 3;
        laysolfr = laytrop;
        for (lay = 1; lay <= laytrop; ++lay) {
            if (c->jp[lay] < layreffr && c->jp[lay + 1] >= layreffr) laysolfr = (lay + 1 < laytrop) ? lay + 1 : laytrop;
            ind0 = IND0A(20) + 1; ind1 = IND1A(20) + 1;
            inds = c->indself[lay]; indf = c->indfor[lay];
            tauray = c->colmol[lay] * K->rayl;
            for (ig = 1; ig <= ngc[4]; ++ig) {
                TAUG(o, ig) = c->colh2o[lay] * ((KEY4(K->absa, 65, ind0, ind1, ig)) + SELFT(K, lay, inds, ig) + FORT(K, lay, indf, ig))
                              + c->colch4[lay] * K->absch4[ig - 1];
                TAUR(o, ig) = tauray;
                if (lay == laysolfr) SFLX(o, ig) = K->sfluxref[ig - 1];
            }
        }
        for (lay = laytrop + 1; lay <= nlayers; ++lay) {
            ind0 = IND0B(20) + 1; ind1 = IND1B(20) + 1;
            indf = c->indfor[lay];
            tauray = c->colmol[lay] * K->rayl;
            for (ig = 1; ig <= ngc[4]; ++ig) {
                TAUG(o, ig) = c->colh2o[lay] * (c->fac00[lay] * F2(K->absb, 235, ind0, ig) + c->fac10[lay] * F2(K->absb, 235, ind0 + 1, ig) +
                                                c->fac01[lay] * F2(K->absb, 235, ind1, ig) + c->fac11[lay] * F2(K->absb, 235, ind1 + 1, ig) +
                                                FORONLY(K, lay, indf, ig))
                              + c->colch4[lay] * K->absch4[ig - 1];
                TAUR(o, ig) = tauray;
            }
        }
    }

    /* ---- band 21: 6150-7700 (h2o,co2; h2o,co2) (:749-868) */
    K = &S->sw[5];
    {
        const double strrat = 0.0045321;
        const int o = ngs[4];
        layreffr = 8;
        laysolfr = laytrop;
        for (lay = 1; lay <= laytrop; ++lay) {
            if (c->jp[lay] < layreffr && c->jp[lay + 1] >= layreffr) laysolfr = (lay + 1 < laytrop) ? lay + 1 : laytrop;
            b = sw_binary(c, lay, c->colh2o[lay], strrat, c->colco2[lay], 8.);
            ind0 = IND0A(21) + b.js; ind1 = IND1A(21) + b.js;
            inds = c->indself[lay]; indf = c->indfor[lay];
            tauray = c->colmol[lay] * K->rayl;
            for (ig = 1; ig <= ngc[5]; ++ig) {
                TAUG(o, ig) = b.speccomb * key8(&b, K->absa, 585, ind0, ind1, 9, ig) + c->colh2o[lay] * SELFFOR(K, lay, inds, indf, ig);
                if (lay == laysolfr)
                    SFLX(o, ig) = F2(K->sfluxref, 10, ig, b.js) + b.fs * (F2(K->sfluxref, 10, ig, b.js + 1) - F2(K->sfluxref, 10, ig, b.js));
                TAUR(o, ig) = tauray;
            }
        }
        for (lay = laytrop + 1; lay <= nlayers; ++lay) {
            b = sw_binary(c, lay, c->colh2o[lay], strrat, c->colco2[lay], 4.);
            ind0 = IND0B(21) + b.js; ind1 = IND1B(21) + b.js;
            indf = c->indfor[lay];
            tauray = c->colmol[lay] * K->rayl;
            for (ig = 1; ig <= ngc[5]; ++ig) {
                TAUG(o, ig) = b.speccomb * key8(&b, K->absb, 1175, ind0, ind1, 5, ig) + FORUP(K, lay, indf, ig);
                TAUR(o, ig) = tauray;
            }
        }
    }

    /* ---- band 22: 7700-8050 (h2o,o2; o2) (:871-977) */
    K = &S->sw[6];
    {
        const double o2adj = 1.6, strrat = 0.022708;
        const int o = ngs[5];
        double o2cont;
        layreffr = 2;
        laysolfr = laytrop;
        for (lay = 1; lay <= laytrop; ++lay) {
            if (c->jp[lay] < layreffr && c->jp[lay + 1] >= layreffr) laysolfr = (lay + 1 < laytrop) ? lay + 1 : laytrop;
            o2cont = 4.35e-4 * c->colo2[lay] / (350.0 * 2.0);
            b = sw_binary(c, lay, c->colh2o[lay], o2adj * strrat, c->colo2[lay], 8.);
            ind0 = IND0A(22) + b.js; ind1 = IND1A(22) + b.js;
            inds = c->indself[lay]; indf = c->indfor[lay];
            tauray = c->colmol[lay] * K->rayl;
            for (ig = 1; ig <= ngc[6]; ++ig) {
                TAUG(o, ig) = b.speccomb * key8(&b, K->absa, 585, ind0, ind1, 9, ig) + c->colh2o[lay] * SELFFOR(K, lay, inds, indf, ig) + o2cont;
                if (lay == laysolfr)
                    SFLX(o, ig) = F2(K->sfluxref, 2, ig, b.js) + b.fs * (F2(K->sfluxref, 2, ig, b.js + 1) - F2(K->sfluxref, 2, ig, b.js));
                TAUR(o, ig) = tauray;
            }
        }
        for (lay = laytrop + 1; lay <= nlayers; ++lay) {
            o2cont = 4.35e-4 * c->colo2[lay] / (350.0 * 2.0);
            ind0 = IND0B(22) + 1; ind1 = IND1B(22) + 1;
            tauray = c->colmol[lay] * K->rayl;
            for (ig = 1; ig <= ngc[6]; ++ig) {
                TAUG(o, ig) = c->colo2[lay] * o2adj * KEY4(K->absb, 235, ind0, ind1, ig) + o2cont;
                TAUR(o, ig) = tauray;
            }
        }
    }

    /* ---- band 23: 8050-12850 (h2o; nothing) (:980-1051) */
    K = &S->sw[7];
    {
        const double givfac = 1.029;
        const int o = ngs[6];
        layreffr = 6;
        laysolfr = laytrop;
        for (lay = 1; lay <= laytrop; ++lay) {
            if (c->jp[lay] < layreffr && c->jp[lay + 1] >= layreffr) laysolfr = (lay + 1 < laytrop) ? lay + 1 : laytrop;
            ind0 = IND0A(23) + 1; ind1 = IND1A(23) + 1;
            inds = c->indself[lay]; indf = c->indfor[lay];
            for (ig = 1; ig <= ngc[7]; ++ig) {
                tauray = c->colmol[lay] * K->raylv[ig - 1];
                TAUG(o, ig) = c->colh2o[lay] * (givfac * (KEY4(K->absa, 65, ind0, ind1, ig)) + SELFT(K, lay, inds, ig) + FORT(K, lay, indf, ig));
                if (lay == laysolfr) SFLX(o, ig) = K->sfluxref[ig - 1];
                TAUR(o, ig) = tauray;
            }
        }
        for (lay = laytrop + 1; lay <= nlayers; ++lay)
            for (ig = 1; ig <= ngc[7]; ++ig) {
                TAUG(o, ig) = 0.;
                TAUR(o, ig) = c->colmol[lay] * K->raylv[ig - 1];
            }
    }

    /* ---- band 24: 12850-16000 (h2o,o2; o2) + O3 (:1054-1153) */
    K = &S->sw[8];
    {
        const double strrat = 0.124692;
        const int o = ngs[7];
        layreffr = 1;
        laysolfr = laytrop;
        for (lay = 1; lay <= laytrop; ++lay) {
            if (c->jp[lay] < layreffr && c->jp[lay + 1] >= layreffr) laysolfr = (lay + 1 < laytrop) ? lay + 1 : laytrop;
            b = sw_binary(c, lay, c->colh2o[lay], strrat, c->colo2[lay], 8.);
            ind0 = IND0A(24) + b.js; ind1 = IND1A(24) + b.js;
            inds = c->indself[lay]; indf = c->indfor[lay];
            for (ig = 1; ig <= ngc[8]; ++ig) {
                tauray = c->colmol[lay] * (F2(K->rayla, 8, ig, b.js) + b.fs * (F2(K->rayla, 8, ig, b.js + 1) - F2(K->rayla, 8, ig, b.js)));
                TAUG(o, ig) = b.speccomb * key8(&b, K->absa, 585, ind0, ind1, 9, ig) + c->colo3[lay] * K->abso3a[ig - 1] +
                              c->colh2o[lay] * SELFFOR(K, lay, inds, indf, ig);
                if (lay == laysolfr)
                    SFLX(o, ig) = F2(K->sfluxref, 8, ig, b.js) + b.fs * (F2(K->sfluxref, 8, ig, b.js + 1) - F2(K->sfluxref, 8, ig, b.js));
                TAUR(o, ig) = tauray;
            }
        }
        for (lay = laytrop + 1; lay <= nlayers; ++lay) {
            ind0 = IND0B(24) + 1; ind1 = IND1B(24) + 1;
            for (ig = 1; ig <= ngc[8]; ++ig) {
                tauray = c->colmol[lay] * K->raylb[ig - 1];
                TAUG(o, ig) = c->colo2[lay] * KEY4(K->absb, 235, ind0, ind1, ig) + c->colo3[lay] * K->abso3b[ig - 1];
                TAUR(o, ig) = tauray;
            }
        }
    }

    /* ---- band 25: 16000-22650 (h2o; nothing) + O3 (:1156-1217) */
    K = &S->sw[9];
    {
        const int o = ngs[8];
        layreffr = 2;
        laysolfr = laytrop;
        for (lay = 1; lay <= laytrop; ++lay) {
            if (c->jp[lay] < layreffr && c->jp[lay + 1] >= layreffr) laysolfr = (lay + 1 < laytrop) ? lay + 1 : laytrop;
            ind0 = IND0A(25) + 1; ind1 = IND1A(25) + 1;
            for (ig = 1; ig <= ngc[9]; ++ig) {
                tauray = c->colmol[lay] * K->raylv[ig - 1];
                TAUG(o, ig) = c->colh2o[lay] * KEY4(K->absa, 65, ind0, ind1, ig) + c->colo3[lay] * K->abso3a[ig - 1];
                if (lay == laysolfr) SFLX(o, ig) = K->sfluxref[ig - 1];
                TAUR(o, ig) = tauray;
            }
        }
        for (lay = laytrop + 1; lay <= nlayers; ++lay)
            for (ig = 1; ig <= ngc[9]; ++ig) {
                tauray = c->colmol[lay] * K->raylv[ig - 1];
                TAUG(o, ig) = c->colo3[lay] * K->abso3b[ig - 1];
                TAUR(o, ig) = tauray;
            }
    }

    /* ---- band 26: 22650-29000 (nothing; nothing) (:1220-1268) */
    K = &S->sw[10];
    {
        const int o = ngs[9];
        laysolfr = laytrop;
        for (lay = 1; lay <= laytrop; ++lay)
            for (ig = 1; ig <= ngc[10]; ++ig) {
                if (lay == laysolfr) SFLX(o, ig) = K->sfluxref[ig - 1];
                TAUG(o, ig) = 0.;
                TAUR(o, ig) = c->colmol[lay] * K->raylv[ig - 1];
            }
        for (lay = laytrop + 1; lay <= nlayers; ++lay)
            for (ig = 1; ig <= ngc[10]; ++ig) {
                TAUG(o, ig) = 0.;
                TAUR(o, ig) = c->colmol[lay] * K->raylv[ig - 1];
            }
    }

    /* ---- band 27: 29000-38000 (o3; o3) (:1271-1347) */
    K = &S->sw[11];
    {
        const double scalekur = 50.15 / 48.37;
        const int o = ngs[10];
        layreffr = 32;
        for (lay = 1; lay <= laytrop; ++lay) {
            ind0 = IND0A(27) + 1; ind1 = IND1A(27) + 1;
            for (ig = 1; ig <= ngc[11]; ++ig) {
                tauray = c->colmol[lay] * K->raylv[ig - 1];
                TAUG(o, ig) = c->colo3[lay] * KEY4(K->absa, 65, ind0, ind1, ig);
                TAUR(o, ig) = tauray;
            }
        }
        laysolfr = nlayers;
        for (lay = laytrop + 1; lay <= nlayers; ++lay) {
            if (c->jp[lay - 1] < layreffr && c->jp[lay] >= layreffr) laysolfr = lay;
            ind0 = IND0B(27) + 1; ind1 = IND1B(27) + 1;
            for (ig = 1; ig <= ngc[11]; ++ig) {
                tauray = c->colmol[lay] * K->raylv[ig - 1];
                TAUG(o, ig) = c->colo3[lay] * KEY4(K->absb, 235, ind0, ind1, ig);
                if (lay == laysolfr) SFLX(o, ig) = scalekur * K->sfluxref[ig - 1];
                TAUR(o, ig) = tauray;
            }
        }
    }

    /* ---- band 28: 38000-50000 (o3,o2; o3,o2) (:1350-1455) */
    K = &S->sw[12];
    {
        const double strrat = 6.67029e-07;
        const int o = ngs[11];
        layreffr = 58;
        for (lay = 1; lay <= laytrop; ++lay) {
            b = sw_binary(c, lay, c->colo3[lay], strrat, c->colo2[lay], 8.);
            ind0 = IND0A(28) + b.js; ind1 = IND1A(28) + b.js;
            tauray = c->colmol[lay] * K->rayl;
            for (ig = 1; ig <= ngc[12]; ++ig) {
                TAUG(o, ig) = b.speccomb * key8(&b, K->absa, 585, ind0, ind1, 9, ig);
                TAUR(o, ig) = tauray;
            }
        }
        laysolfr = nlayers;
        for (lay = laytrop + 1; lay <= nlayers; ++lay) {
            if (c->jp[lay - 1] < layreffr && c->jp[lay] >= layreffr) laysolfr = lay;
            b = sw_binary(c, lay, c->colo3[lay], strrat, c->colo2[lay], 4.);
            ind0 = IND0B(28) + b.js; ind1 = IND1B(28) + b.js;
            tauray = c->colmol[lay] * K->rayl;
            for (ig = 1; ig <= ngc[12]; ++ig) {
                TAUG(o, ig) = b.speccomb * key8(&b, K->absb, 1175, ind0, ind1, 5, ig);
                if (lay == laysolfr)
                    SFLX(o, ig) = F2(K->sfluxref, 6, ig, b.js) + b.fs * (F2(K->sfluxref, 6, ig, b.js + 1) - F2(K->sfluxref, 6, ig, b.js));
                TAUR(o, ig) = tauray;
            }
        }
    }

    /* ---- band 29: 820-2600 (h2o; co2) + CO2 / H2O minor (:1458-1536) */
    K = &S->sw[13];
    {
        const int o = ngs[12];
        layreffr = 49;
        for (lay = 1; lay <= laytrop; ++lay) {
            ind0 = IND0A(29) + 1; ind1 = IND1A(29) + 1;
            inds = c->indself[lay]; indf = c->indfor[lay];
            tauray = c->colmol[lay] * K->rayl;
            for (ig = 1; ig <= ngc[13]; ++ig) {
                TAUG(o, ig) = c->colh2o[lay] * ((KEY4(K->absa, 65, ind0, ind1, ig)) + SELFT(K, lay, inds, ig) + FORT(K, lay, indf, ig))
                              + c->colco2[lay] * K->absco2[ig - 1];
                TAUR(o, ig) = tauray;
            }
        }
        laysolfr = nlayers;
        for (lay = laytrop + 1; lay <= nlayers; ++lay) {
            if (c->jp[lay - 1] < layreffr && c->jp[lay] >= layreffr) laysolfr = lay;
            ind0 = IND0B(29) + 1; ind1 = IND1B(29) + 1;
            tauray = c->colmol[lay] * K->rayl;
            for (ig = 1; ig <= ngc[13]; ++ig) {
                TAUG(o, ig) = c->colco2[lay] * KEY4(K->absb, 235, ind0, ind1, ig) + c->colh2o[lay] * K->absh2o[ig - 1];
                if (lay == laysolfr) SFLX(o, ig) = K->sfluxref[ig - 1];
                TAUR(o, ig) = tauray;
            }
        }
    }
#undef TAUG
#undef TAUR
#undef SFLX
}

/* ---------------------------------------------------------------- reftra_sw (reftra.f90:122-303) */
static void reftra_sw(int nlayers, const int *lrtchk, const double *pgg, double prmuz, const double *ptau,
                      const double *pw, double *pref, double *prefd, double *ptra, double *ptrad)
{
    const orc_state_t *S = &g_orc;
    const double eps = 1.e-08, od_lo = 0.06, tblint = 10000.0, bpade = S->sw_bpade;
    const double zwcrit = 0.9999995;
    double za, za1, za2, zbeta, zdend, zdenr, zdent, ze1, ze2, zem1, zem2, zemm, zep1, zep2;
    double zg, zg3, zgamma1, zgamma2, zgamma3, zgamma4, zgt;
    double zr1, zr2, zr3, zr4, zr5, zrk, zrk2, zrkg, zrm1, zrp, zrp1, zrpp;
    double zt1, zt2, zt3, zt4, zt5, zto1, zw, zwo, tblind;
    int itind;
    for (int jk = 1; jk <= nlayers; ++jk) {
        if (!lrtchk[jk]) {
            pref[jk] = 0.;
            ptra[jk] = 1.;
            prefd[jk] = 0.;
            ptrad[jk] = 1.;
        } else {
            zto1 = ptau[jk];
            zw = pw[jk];
            zg = pgg[jk];
            zg3 = 3. * zg;
            /* kmodts == 2: practical improved flux method */
            zgamma1 = (8. - zw * (5. + zg3)) * 0.25;
            zgamma2 = 3. * (zw * (1. - zg)) * 0.25;
            zgamma3 = (2. - zg3 * prmuz) * 0.25;
            zgamma4 = 1. - zgamma3;
            {
                double t = zg / (1. - zg);
                zwo = zw / (1. - (1. - zw) * (t * t));
            }
            if (zwo >= zwcrit) {
                /* conservative scattering */
                za = zgamma1 * prmuz;
                za1 = za - zgamma3;
                zgt = zgamma1 * zto1;
                ze1 = fmin(zto1 / prmuz, 500.);
                if (ze1 <= od_lo) {
                    ze2 = 1. - ze1 + 0.5 * ze1 * ze1;
                } else {
                    tblind = ze1 / (bpade + ze1);
                    itind = (int)(tblint * tblind + 0.5);
                    ze2 = S->sw_exp_tbl[itind];
                }
                pref[jk] = (zgt - za1 * (1. - ze2)) / (1. + zgt);
                ptra[jk] = 1. - pref[jk];
                prefd[jk] = zgt / (1. + zgt);
                ptrad[jk] = 1. - prefd[jk];
                if (ze2 == 1.0) {
                    pref[jk] = 0.0;
                    ptra[jk] = 1.0;
                    prefd[jk] = 0.0;
                    ptrad[jk] = 1.0;
                }
            } else {
                za1 = zgamma1 * zgamma4 + zgamma2 * zgamma3;
                za2 = zgamma1 * zgamma3 + zgamma2 * zgamma4;
                zrk = sqrt(zgamma1 * zgamma1 - zgamma2 * zgamma2);
                zrp = zrk * prmuz;
                zrp1 = 1. + zrp;
                zrm1 = 1. - zrp;
                zrk2 = 2. * zrk;
                zrpp = 1. - zrp * zrp;
                zrkg = zrk + zgamma1;
                zr1 = zrm1 * (za2 + zrk * zgamma3);
                zr2 = zrp1 * (za2 - zrk * zgamma3);
                zr3 = zrk2 * (zgamma3 - za2 * prmuz);
                zr4 = zrpp * zrkg;
                zr5 = zrpp * (zrk - zgamma1);
                zt1 = zrp1 * (za1 + zrk * zgamma4);
                zt2 = zrm1 * (za1 - zrk * zgamma4);
                zt3 = zrk2 * (zgamma4 + za1 * prmuz);
                zt4 = zr4;
                zt5 = zr5;
                zbeta = (zgamma1 - zrk) / zrkg;
                ze1 = fmin(zrk * zto1, 500.);
                ze2 = fmin(zto1 / prmuz, 500.);
                if (ze1 <= od_lo) {
                    zem1 = 1. - ze1 + 0.5 * ze1 * ze1;
                    zep1 = 1. / zem1;
                } else {
                    tblind = ze1 / (bpade + ze1);
                    itind = (int)(tblint * tblind + 0.5);
                    zem1 = S->sw_exp_tbl[itind];
                    zep1 = 1. / zem1;
                }
                if (ze2 <= od_lo) {
                    zem2 = 1. - ze2 + 0.5 * ze2 * ze2;
                    zep2 = 1. / zem2;
                } else {
                    tblind = ze2 / (bpade + ze2);
                    itind = (int)(tblint * tblind + 0.5);
                    zem2 = S->sw_exp_tbl[itind];
                    zep2 = 1. / zem2;
                }
                zdenr = zr4 * zep1 + zr5 * zem1;
                zdent = zt4 * zep1 + zt5 * zem1;
                if (zdenr >= -eps && zdenr <= eps) {
                    pref[jk] = eps;
                    ptra[jk] = zem2;
                } else {
                    pref[jk] = zw * (zr1 * zep1 - zr2 * zem1 - zr3 * zem2) / zdenr;
                    ptra[jk] = zem2 - zem2 * zw * (zt1 * zep1 - zt2 * zem1 - zt3 * zep2) / zdent;
                }
                zemm = zem1 * zem1;
                zdend = 1. / ((1. - zbeta * zemm) * zrkg);
                prefd[jk] = zgamma2 * (1. - zemm) * zdend;
                ptrad[jk] = zrk2 * zem1 * zdend;
            }
        }
    }
}

/* ---------------------------------------------------------------- vrtqdr_sw (vrtqdr.f90:103-150) */
static void vrtqdr_sw(int klev, const double *pref, const double *prefd, const double *ptra, const double *ptrad,
                      const double *pdbt, double *prdnd, double *prup, double *prupd, const double *ptdbt,
                      double *pfd, double *pfu)
{
    double zreflect, ztdn[NL];
    int ikp, ikx, jk;
    zreflect = 1. / (1. - prefd[klev + 1] * prefd[klev]);
    prup[klev] = pref[klev] + (ptrad[klev] * ((ptra[klev] - pdbt[klev]) * prefd[klev + 1] + pdbt[klev] * pref[klev + 1])) * zreflect;
    prupd[klev] = prefd[klev] + ptrad[klev] * ptrad[klev] * prefd[klev + 1] * zreflect;
    for (jk = 1; jk <= klev - 1; ++jk) {
        ikp = klev + 1 - jk;
        ikx = ikp - 1;
        zreflect = 1. / (1. - prupd[ikp] * prefd[ikx]);
        prup[ikx] = pref[ikx] + (ptrad[ikx] * ((ptra[ikx] - pdbt[ikx]) * prupd[ikp] + pdbt[ikx] * prup[ikp])) * zreflect;
        prupd[ikx] = prefd[ikx] + ptrad[ikx] * ptrad[ikx] * prupd[ikp] * zreflect;
    }
    ztdn[1] = 1.;
    prdnd[1] = 0.;
    ztdn[2] = ptra[1];
    prdnd[2] = prefd[1];
    for (jk = 2; jk <= klev; ++jk) {
        ikp = jk + 1;
        zreflect = 1. / (1. - prefd[jk] * prdnd[jk]);
        ztdn[ikp] = ptdbt[jk] * ptra[jk] + (ptrad[jk] * ((ztdn[jk] - ptdbt[jk]) + ptdbt[jk] * pref[jk] * prdnd[jk])) * zreflect;
        prdnd[ikp] = prefd[jk] + ptrad[jk] * ptrad[jk] * prdnd[jk] * zreflect;
    }
    for (jk = 1; jk <= klev + 1; ++jk) {
        zreflect = 1. / (1. - prdnd[jk] * prupd[jk]);
        pfu[jk] = (ptdbt[jk] * prup[jk] + (ztdn[jk] - ptdbt[jk]) * prupd[jk]) * zreflect;
        pfd[jk] = ptdbt[jk] + (ztdn[jk] - ptdbt[jk] + ptdbt[jk] * prup[jk] * prdnd[jk]) * zreflect;
    }
}

/* ---------------------------------------------------------------- spcvrt_sw (spcvrt.f90:255-626) */
static void spcvrt_sw(swcol_t *c, const double *palbd, const double *palbp, double prmu0)
{
    const orc_state_t *S = &g_orc;
    const int klev = c->nlayers;
    const double od_lo = 0.06, tblint = 10000.0, bpade = S->sw_bpade, repclc = 1.e-12;
    const int icpr = 1, idelm = 1;
    int lrtchkclr[NL], lrtchkcld[NL];
    double zdbt[NL], zdbtc[NL], zgcc[NL], zgco[NL], zomcc[NL], zomco[NL];
    double zrdnd[NL], zrdndc[NL], zref[NL], zrefc[NL], zrefo[NL], zrefd[NL], zrefdc[NL], zrefdo[NL];
    double zrup[NL], zrupd[NL], zrupc[NL], zrupdc[NL], ztauc[NL], ztauo[NL], ztdbt[NL], ztdbtc[NL];
    double ztra[NL], ztrac[NL], ztrao[NL], ztrad[NL], ztradc[NL], ztrado[NL];
    double zcd[NL], zcu[NL], zfd[NL], zfu[NL];
    double zincflx, zclear, zcloud, zdbtmc, zdbtmo, zf, zwf, ze1, tblind;
    int iw = 0, itind, ikl;

    for (int jk = 1; jk <= klev + 1; ++jk) {
        c->zbbcd[jk] = 0.; c->zbbcu[jk] = 0.; c->zbbfd[jk] = 0.; c->zbbfu[jk] = 0.;
        c->zbbcddir[jk] = 0.; c->zbbfddir[jk] = 0.; c->zuvcd[jk] = 0.; c->zuvfd[jk] = 0.;
        c->zuvcddir[jk] = 0.; c->zuvfddir[jk] = 0.; c->znicd[jk] = 0.; c->znifd[jk] = 0.;
        c->znicddir[jk] = 0.; c->znifddir[jk] = 0.;
    }
    taumol_sw(c);

    for (int jb = 16; jb <= 29; ++jb) {
        const int ibm = jb - 15;
        const int igt = ngc[ibm - 1];
        for (int jg = 1; jg <= igt; ++jg) {
            iw = iw + 1;
            zincflx = c->adjflux[jb] * c->zsflxzen[iw] * prmu0;
            ztdbtc[1] = 1.0;
            zdbtc[klev + 1] = 0.0;
            ztrac[klev + 1] = 0.0;
            ztradc[klev + 1] = 0.0;
            zrefc[klev + 1] = palbp[ibm];
            zrefdc[klev + 1] = palbd[ibm];
            zrupc[klev + 1] = palbp[ibm];
            zrupdc[klev + 1] = palbd[ibm];
            ztrao[klev + 1] = 0.0;
            ztrado[klev + 1] = 0.0;
            zrefo[klev + 1] = palbp[ibm];
            zrefdo[klev + 1] = palbd[ibm];
            ztdbt[1] = 1.0;
            zdbt[klev + 1] = 0.0;
            ztra[klev + 1] = 0.0;
            ztrad[klev + 1] = 0.0;
            zref[klev + 1] = palbp[ibm];
            zrefd[klev + 1] = palbd[ibm];
            zrup[klev + 1] = palbp[ibm];
            zrupd[klev + 1] = palbd[ibm];

            for (int jk = 1; jk <= klev; ++jk) {
                ikl = klev + 1 - jk;
                lrtchkclr[jk] = 1;
                const double pclfr = c->pclfr[ikl], ptauc = c->ptauc[ikl][ibm], pasyc = c->pasyc[ikl][ibm], pomgc = c->pomgc[ikl][ibm];
                const double ptaua = c->ptaua[ikl][ibm], pasya = c->pasya[ikl][ibm], pomga = c->pomga[ikl][ibm];
                lrtchkcld[jk] = (pclfr > repclc);
                /* clear-sky optical parameters including aerosols (:386-396) */
                ztauc[jk] = c->ztaur[iw][ikl] + c->ztaug[iw][ikl] + ptaua;
                zomcc[jk] = c->ztaur[iw][ikl] * 1.0 + ptaua * pomga;
                zgcc[jk] = pasya * pomga * ptaua / zomcc[jk];
                zomcc[jk] = zomcc[jk] / ztauc[jk];
                /* delta scaling, clear (:446-452) */
                zf = zgcc[jk] * zgcc[jk];
                zwf = zomcc[jk] * zf;
                ztauc[jk] = (1.0 - zwf) * ztauc[jk];
                zomcc[jk] = (zomcc[jk] - zwf) / (1.0 - zwf);
                zgcc[jk] = (zgcc[jk] - zf) / (1.0 - zf);
                /* total sky, icpr >= 1 (:455-463) */
                ztauo[jk] = ztauc[jk] + ptauc;
                zomco[jk] = ztauc[jk] * zomcc[jk] + ptauc * pomgc;
                zgco[jk] = (ptauc * pomgc * pasyc + ztauc[jk] * zomcc[jk] * zgcc[jk]) / zomco[jk];
                zomco[jk] = zomco[jk] / ztauo[jk];
            }
            reftra_sw(klev, lrtchkclr, zgcc, prmu0, ztauc, zomcc, zrefc, zrefdc, ztrac, ztradc);
            reftra_sw(klev, lrtchkcld, zgco, prmu0, ztauo, zomco, zrefo, zrefdo, ztrao, ztrado);

            for (int jk = 1; jk <= klev; ++jk) {
                ikl = klev + 1 - jk;
                zclear = 1.0 - c->pclfr[ikl];
                zcloud = c->pclfr[ikl];
                zref[jk] = zclear * zrefc[jk] + zcloud * zrefo[jk];
                zrefd[jk] = zclear * zrefdc[jk] + zcloud * zrefdo[jk];
                ztra[jk] = zclear * ztrac[jk] + zcloud * ztrao[jk];
                ztrad[jk] = zclear * ztradc[jk] + zcloud * ztrado[jk];
                /* direct beam transmittance, clear (:519-531) */
                ze1 = ztauc[jk] / prmu0;
                if (ze1 <= od_lo) {
                    zdbtmc = 1. - ze1 + 0.5 * ze1 * ze1;
                } else {
                    tblind = ze1 / (bpade + ze1);
                    itind = (int)(tblint * tblind + 0.5);
                    zdbtmc = S->sw_exp_tbl[itind];
                }
                zdbtc[jk] = zdbtmc;
                ztdbtc[jk + 1] = zdbtc[jk] * ztdbtc[jk];
                /* total (:535-548) */
                ze1 = ztauo[jk] / prmu0;
                if (ze1 <= od_lo) {
                    zdbtmo = 1. - ze1 + 0.5 * ze1 * ze1;
                } else {
                    tblind = ze1 / (bpade + ze1);
                    itind = (int)(tblint * tblind + 0.5);
                    zdbtmo = S->sw_exp_tbl[itind];
                }
                zdbt[jk] = zclear * zdbtmc + zcloud * zdbtmo;
                ztdbt[jk + 1] = zdbt[jk] * ztdbt[jk];
            }
            vrtqdr_sw(klev, zrefc, zrefdc, ztrac, ztradc, zdbtc, zrdndc, zrupc, zrupdc, ztdbtc, zcd, zcu);
            vrtqdr_sw(klev, zref, zrefd, ztra, ztrad, zdbt, zrdnd, zrup, zrupd, ztdbt, zfd, zfu);

            /* upwelling and downwelling fluxes at levels (:570-619); idelm == 1 */
            for (int jk = 1; jk <= klev + 1; ++jk) {
                ikl = klev + 2 - jk;
                c->zbbfu[ikl] = c->zbbfu[ikl] + zincflx * zfu[jk];
                c->zbbfd[ikl] = c->zbbfd[ikl] + zincflx * zfd[jk];
                c->zbbcu[ikl] = c->zbbcu[ikl] + zincflx * zcu[jk];
                c->zbbcd[ikl] = c->zbbcd[ikl] + zincflx * zcd[jk];
                c->zbbfddir[ikl] = c->zbbfddir[ikl] + zincflx * ztdbt[jk];
                c->zbbcddir[ikl] = c->zbbcddir[ikl] + zincflx * ztdbtc[jk];
                if (ibm >= 10 && ibm <= 13) {
                    c->zuvcd[ikl] = c->zuvcd[ikl] + zincflx * zcd[jk];
                    c->zuvfd[ikl] = c->zuvfd[ikl] + zincflx * zfd[jk];
                    c->zuvfddir[ikl] = c->zuvfddir[ikl] + zincflx * ztdbt[jk];
                    c->zuvcddir[ikl] = c->zuvcddir[ikl] + zincflx * ztdbtc[jk];
                } else if (ibm == 14 || ibm <= 9) {
                    c->znicd[ikl] = c->znicd[ikl] + zincflx * zcd[jk];
                    c->znifd[ikl] = c->znifd[ikl] + zincflx * zfd[jk];
                    c->znifddir[ikl] = c->znifddir[ikl] + zincflx * ztdbt[jk];
                    c->znicddir[ikl] = c->znicddir[ikl] + zincflx * ztdbtc[jk];
                }
            }
        }
    }
    (void)idelm; (void)icpr;
}

/* ---------------------------------------------------------------- cldprop_sw, inflag = 2 (rrtmg_sw_cldprop.f90:168-345)
   one layer; ptauc/pomgc/pasyc[1..14] out.  Returns 0 or the number of the Fortran `stop`: 1 ICE RADIUS OUT OF BOUNDS,
   2 ICE GENERALIZED EFFECTIVE SIZE OUT OF BOUNDS, 3 LIQUID EFFECTIVE RADIUS OUT OF BOUNDS, 4 one of the range checks on
   the interpolated properties (extinction < 0, single-scattering albedo or asymmetry outside [0, 1], fdelta). */
static int cldprop_sw_layer(int iceflag, int liqflag, double ciwp, double clwp, double radice, double radliq,
                            double *ptauc, double *pomgc, double *pasyc)
{
    const orc_state_t *S = &g_orc;
    const double eps = 1.e-06, cldmin = 1.e-20;
    static const double wavenum2[15] = {0., 3250., 4000., 4650., 5150., 6150., 7700., 8050., 12850., 16000., 22650., 29000.,
                                        38000., 50000., 2600.};
    double extcoice[15], gice[15], ssacoice[15], forwice[15], extcoliq[15], gliq[15], ssacoliq[15], forwliq[15], fdelta[15];
    double factor, fint;
    int index, icx = 0;
    for (int ib = 1; ib <= 14; ++ib) {
        extcoice[ib] = 0.; ssacoice[ib] = 0.; gice[ib] = 0.; forwice[ib] = 0.;
        extcoliq[ib] = 0.; ssacoliq[ib] = 0.; gliq[ib] = 0.; forwliq[ib] = 0.;
    }
    if (ciwp == 0.0) {
        /* zeros */
    } else if (iceflag == 1) {
        if (radice < 13.0 || radice > 130.) return 1;
        for (int ib = 1; ib <= 14; ++ib) {
            if (wavenum2[ib] > 1.43e04) icx = 1;
            else if (wavenum2[ib] > 7.7e03) icx = 2;
            else if (wavenum2[ib] > 5.3e03) icx = 3;
            else if (wavenum2[ib] > 4.0e03) icx = 4;
            else if (wavenum2[ib] >= 2.5e03) icx = 5;
            extcoice[ib] = S->abari[icx] + S->bbari[icx] / radice;
            ssacoice[ib] = 1. - S->cbari[icx] - S->dbari[icx] * radice;
            gice[ib] = S->ebari[icx] + S->fbari[icx] * radice;
            if (gice[ib] >= 1.0) gice[ib] = 1.0 - eps;
            forwice[ib] = gice[ib] * gice[ib];
            if (extcoice[ib] < 0.0 || ssacoice[ib] > 1.0 || ssacoice[ib] < 0.0 || gice[ib] > 1.0 || gice[ib] < 0.0) return 4;
        }
    } else if (iceflag == 2) {
        if (radice < 5.0 || radice > 131.0) return 1;
        factor = (radice - 2.) / 3.;
        index = (int)factor;
        if (index == 43) index = 42;
        fint = factor - (double)index;
        for (int ib = 1; ib <= 14; ++ib) {
            extcoice[ib] = S->extice2[index][ib] + fint * (S->extice2[index + 1][ib] - S->extice2[index][ib]);
            ssacoice[ib] = S->ssaice2[index][ib] + fint * (S->ssaice2[index + 1][ib] - S->ssaice2[index][ib]);
            gice[ib] = S->asyice2[index][ib] + fint * (S->asyice2[index + 1][ib] - S->asyice2[index][ib]);
            forwice[ib] = gice[ib] * gice[ib];
            if (extcoice[ib] < 0.0 || ssacoice[ib] > 1.0 || ssacoice[ib] < 0.0 || gice[ib] > 1.0 || gice[ib] < 0.0) return 4;
        }
    } else if (iceflag == 3) {
        if (radice < 5.0 || radice > 140.0) return 2;
        factor = (radice - 2.) / 3.;
        index = (int)factor;
        if (index == 46) index = 45;
        fint = factor - (double)index;
        for (int ib = 1; ib <= 14; ++ib) {
            extcoice[ib] = S->extice3[index][ib] + fint * (S->extice3[index + 1][ib] - S->extice3[index][ib]);
            ssacoice[ib] = S->ssaice3[index][ib] + fint * (S->ssaice3[index + 1][ib] - S->ssaice3[index][ib]);
            gice[ib] = S->asyice3[index][ib] + fint * (S->asyice3[index + 1][ib] - S->asyice3[index][ib]);
            fdelta[ib] = S->fdlice3[index][ib] + fint * (S->fdlice3[index + 1][ib] - S->fdlice3[index][ib]);
            if (fdelta[ib] < 0.0 || fdelta[ib] > 1.0) return 4;
            forwice[ib] = fdelta[ib] + 0.5 / ssacoice[ib];
            if (forwice[ib] > gice[ib]) forwice[ib] = gice[ib];
            if (extcoice[ib] < 0.0 || ssacoice[ib] > 1.0 || ssacoice[ib] < 0.0 || gice[ib] > 1.0 || gice[ib] < 0.0) return 4;
        }
    }
    if (clwp == 0.0) {
        /* zeros */
    } else if (liqflag == 1) {
        if (radliq < 2.5 || radliq > 60.) return 3;
        index = (int)(radliq - 1.5);
        if (index == 0) index = 1;
        if (index == 58) index = 57;
        fint = radliq - 1.5 - (double)index;
        for (int ib = 1; ib <= 14; ++ib) {
            extcoliq[ib] = S->extliq1[index][ib] + fint * (S->extliq1[index + 1][ib] - S->extliq1[index][ib]);
            ssacoliq[ib] = S->ssaliq1[index][ib] + fint * (S->ssaliq1[index + 1][ib] - S->ssaliq1[index][ib]);
            if (fint < 0. && ssacoliq[ib] > 1.) ssacoliq[ib] = S->ssaliq1[index][ib];
            gliq[ib] = S->asyliq1[index][ib] + fint * (S->asyliq1[index + 1][ib] - S->asyliq1[index][ib]);
            forwliq[ib] = gliq[ib] * gliq[ib];
            if (extcoliq[ib] < 0.0 || ssacoliq[ib] > 1.0 || ssacoliq[ib] < 0.0 || gliq[ib] > 1.0 || gliq[ib] < 0.0) return 4;
        }
    }
    for (int ib = 1; ib <= 14; ++ib) {
        const double tauliqorig = clwp * extcoliq[ib];
        const double tauiceorig = ciwp * extcoice[ib];
        const double ssaliq = ssacoliq[ib] * (1.0 - forwliq[ib]) / (1.0 - forwliq[ib] * ssacoliq[ib]);
        const double tauliq = (1.0 - forwliq[ib] * ssacoliq[ib]) * tauliqorig;
        const double ssaice = ssacoice[ib] * (1.0 - forwice[ib]) / (1.0 - forwice[ib] * ssacoice[ib]);
        const double tauice = (1.0 - forwice[ib] * ssacoice[ib]) * tauiceorig;
        const double scatliq = ssaliq * tauliq;
        double scatice = ssaice * tauice;
        double taucloud = tauliq + tauice;
        if (taucloud == 0.0) taucloud = cldmin;
        if (scatice == 0.0) scatice = cldmin;
        ptauc[ib] = taucloud;
        pomgc[ib] = (scatliq + scatice) / taucloud;
        if (iceflag == 3)
            pasyc[ib] = (1.0 / (scatliq + scatice)) *
                        (scatliq * (gliq[ib] - forwliq[ib]) / (1.0 - forwliq[ib]) + scatice * ((gice[ib] - forwice[ib]) / (1.0 - forwice[ib])));
        else
            pasyc[ib] = (scatliq * (gliq[ib] - forwliq[ib]) / (1.0 - forwliq[ib]) + scatice * (gice[ib] - forwice[ib]) / (1.0 - forwice[ib])) /
                        (scatliq + scatice);
    }
    return 0;
}

/* ---------------------------------------------------------------- rrtmg_sw (rad.nomcica:78-731) */
int orc_rrtmg_sw(int ncol, int nlay, int icld, int iaer,
                 const double *play, const double *plev, const double *tlay, const double *tlev,
                 const double *tsfc, const double *h2ovmr, const double *o3vmr, const double *co2vmr,
                 const double *ch4vmr, const double *n2ovmr, const double *o2vmr,
                 const double *asdir, const double *asdif, const double *aldir, const double *aldif,
                 const double *coszen, double adjes, int dyofyr, double scon,
                 int inflgsw, const double *cldfr, const double *taucld, const double *ssacld, const double *asmcld,
                 const double *fsfcld, const double *tauaer, const double *ssaaer, const double *asmaer,
                 const double *ecaer,
                 int iceflgsw, int liqflgsw, const double *cicewp, const double *cliqwp, const double *reice, const double *reliq,
                 double *swuflx, double *swdflx, double *swhr, double *swuflxc, double *swdflxc,
                 double *swhrc, const orc_sw_stages_t *st, int nthreads)
{
    /* cldfr (ncol,nlay); taucld, ssacld, asmcld, fsfcld (14,ncol,nlay); tauaer, ssaaer, asmaer (ncol,nlay,14) */
    if (!g_orc.ready) return 1;
    if (icld < 0 || icld > 3) icld = 2;                              /* :468 */
    if (iaer != 0 && iaer != 6 && iaer != 10) iaer = 0;              /* :473 */
    if (iaer == 6 && !ecaer) return 3;                               /* ecaer (ncol,nlay,6) */
    if (icld >= 1 && inflgsw != 0 && inflgsw != 2) return 2;
    if (icld >= 1 && (!cldfr || !taucld || !ssacld || !asmcld || !fsfcld)) return 3;
    if (icld >= 1 && inflgsw == 2 && (!cicewp || !cliqwp || !reice || !reliq || iceflgsw < 1 || iceflgsw > 3 || liqflgsw != 1)) return 3;
    int cld_stop = 0;
    if (iaer == 10 && (!tauaer || !ssaaer || !asmaer)) return 3;
    if (icld >= 1) /* without McICA: clear or overcast layers only (:534-539, `stop 'PARTIAL CLOUD NOT ALLOWED'`); the test sits
                      behind the night-column skip (:497-505), so only sunlit columns are looked at */
        for (long i = 0; i < (long)ncol * nlay; ++i)
            if (!(coszen[i % ncol] < 1.e-10) && cldfr[i] > 1.e-06 && cldfr[i] < 1.0 - 1.e-06) return 4;
    if (nlay < 1 || nlay > ORC_MAXLAY) return 3;
    if (nthreads < 1) nthreads = 1;
    const double zepsec = 1.e-06, zepzen = 1.e-10;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
    {
        swcol_t *c = (swcol_t *)malloc(sizeof(swcol_t));
        c->oneminus = 1.0 - zepsec;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
        for (int iplon = 1; iplon <= ncol; ++iplon) {
            const long i0 = iplon - 1;
            double albdir[15], albdif[15], cossza, swnflx[NL], swnflxc[NL], zdpgcp;
            if (coszen[i0] < zepzen) {
                /* night column (:502-510) */
                for (int k = 0; k <= nlay; ++k) {
                    swuflx[(long)k * ncol + i0] = 0.; swdflx[(long)k * ncol + i0] = 0.;
                    swuflxc[(long)k * ncol + i0] = 0.; swdflxc[(long)k * ncol + i0] = 0.;
                }
                for (int k = 0; k < nlay; ++k) {
                    swhr[(long)k * ncol + i0] = 0.; swhrc[(long)k * ncol + i0] = 0.;
                }
                if (st) {
                    if (st->laytrop) st->laytrop[i0] = -1;
                }
                continue;
            }
            memset(c->zsflxzen, 0, sizeof c->zsflxzen); /* deterministic if a band never reaches laysolfr */
            inatm_sw(c, iplon, ncol, nlay, play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr,
                     adjes, dyofyr, scon);
            /* clouds: cldprop_sw, inflag = 0 branch (cldprop.f90:120-166) and the transfer at rad.nomcica:581-597 */
            for (int lay = 1; lay <= nlay; ++lay) {
                double tauctot = 0.0;
                c->pclfr[lay] = icld >= 1 ? cldfr[(long)(lay - 1) * ncol + i0] : 0.0;
                for (int ib = 1; ib <= 14; ++ib) {
                    c->ptauc[lay][ib] = 0.0; c->pomgc[lay][ib] = 1.0; c->pasyc[lay][ib] = 0.0;
                    if (icld >= 1) tauctot = tauctot + taucld[(ib - 1) + 14 * (i0 + (long)(lay - 1) * ncol)];
                }
                const long ol = (long)(lay - 1) * ncol + i0;
                const double cwp = (icld >= 1 && inflgsw == 2) ? cicewp[ol] + cliqwp[ol] : 0.0;
                if (icld >= 1 && inflgsw == 2 && c->pclfr[lay] >= 1.e-20 && (cwp >= 1.e-20 || tauctot >= 1.e-20)) {
                    const int stop = cldprop_sw_layer(iceflgsw, liqflgsw, cicewp[ol], cliqwp[ol], reice[ol], reliq[ol],
                                                      c->ptauc[lay], c->pomgc[lay], c->pasyc[lay]);
                    if (stop) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
                        cld_stop = stop;
                    }
                } else if (icld >= 1 && inflgsw == 0 && c->pclfr[lay] >= 1.e-20 && tauctot >= 1.e-20) {   /* cwp = 0 for inflag = 0 inputs */
                    for (int ib = 1; ib <= 14; ++ib) {
                        const long o = (ib - 1) + 14 * (i0 + (long)(lay - 1) * ncol);
                        const double taucldorig_a = taucld[o];
                        const double ffp = fsfcld[o];
                        const double ffp1 = 1.0 - ffp;
                        const double ffpssa = 1.0 - ffp * ssacld[o];
                        c->pomgc[lay][ib] = ffp1 * ssacld[o] / ffpssa;
                        c->ptauc[lay][ib] = ffpssa * taucldorig_a;
                        c->pasyc[lay][ib] = (asmcld[o] - ffp) / (ffp1);
                    }
                }
                /* aerosols (rad.nomcica:599-640) */
                for (int ib = 1; ib <= 14; ++ib) {
                    if (iaer == 10) {
                        const long o = i0 + (long)ncol * ((lay - 1) + (long)nlay * (ib - 1));
                        c->ptaua[lay][ib] = tauaer[o]; c->pasya[lay][ib] = asmaer[o]; c->pomga[lay][ib] = ssaaer[o];
                    } else if (iaer == 6) { /* six ECMWF aerosol types (:608-640) */
                        const orc_state_t *S = &g_orc;
                        double ztaua = 0., zasya = 0., zomga = 0.;
                        for (int ia = 1; ia <= 6; ++ia) {
                            const double e = ecaer[i0 + (long)ncol * ((lay - 1) + (long)nlay * (ia - 1))];
                            ztaua = ztaua + S->rsrtaua[ib - 1][ia - 1] * e;
                            zomga = zomga + S->rsrtaua[ib - 1][ia - 1] * e * S->rsrpiza[ib - 1][ia - 1];
                            zasya = zasya + S->rsrtaua[ib - 1][ia - 1] * e * S->rsrpiza[ib - 1][ia - 1] * S->rsrasya[ib - 1][ia - 1];
                        }
                        if (ztaua == 0.) {
                            ztaua = 0.; zasya = 0.; zomga = 1.;
                        } else {
                            if (zomga != 0.) zasya = zasya / zomga;
                            if (ztaua != 0.) zomga = zomga / ztaua;
                        }
                        c->ptaua[lay][ib] = ztaua; c->pasya[lay][ib] = zasya; c->pomga[lay][ib] = zomga;
                    } else {
                        c->ptaua[lay][ib] = 0.0; c->pasya[lay][ib] = 0.0; c->pomga[lay][ib] = 1.0;
                    }
                }
            }
            setcoef_sw(c);
            cossza = coszen[i0];
            if (cossza < zepzen) cossza = zepzen;
            /* band albedos (:565-578) */
            for (int ib = 1; ib <= 9; ++ib) { albdir[ib] = aldir[i0]; albdif[ib] = aldif[i0]; }
            albdir[14] = aldir[i0];
            albdif[14] = aldif[i0];
            for (int ib = 10; ib <= 13; ++ib) { albdir[ib] = asdir[i0]; albdif[ib] = asdif[i0]; }

            spcvrt_sw(c, albdif, albdir, cossza);

            for (int i = 1; i <= nlay + 1; ++i) {
                swuflxc[(long)(i - 1) * ncol + i0] = c->zbbcu[i];
                swdflxc[(long)(i - 1) * ncol + i0] = c->zbbcd[i];
                swuflx[(long)(i - 1) * ncol + i0] = c->zbbfu[i];
                swdflx[(long)(i - 1) * ncol + i0] = c->zbbfd[i];
            }
            for (int i = 1; i <= nlay + 1; ++i) {
                swnflxc[i] = c->zbbcd[i] - c->zbbcu[i];
                swnflx[i] = c->zbbfd[i] - c->zbbfu[i];
            }
            for (int i = 1; i <= nlay; ++i) {
                zdpgcp = g_orc.sw_heatfac / c->pdp[i];
                swhrc[(long)(i - 1) * ncol + i0] = (swnflxc[i + 1] - swnflxc[i]) * zdpgcp;
                swhr[(long)(i - 1) * ncol + i0] = (swnflx[i + 1] - swnflx[i]) * zdpgcp;
            }
            /* MiMA modification: no heating in the top layer (:724-726) */
            swhrc[(long)(nlay - 1) * ncol + i0] = 0.;
            swhr[(long)(nlay - 1) * ncol + i0] = 0.;

            if (st) {
#define PUT(dst, src) if (st->dst) for (int l = 1; l <= nlay; ++l) st->dst[(long)(l - 1) * ncol + i0] = c->src[l]
                if (st->laytrop) st->laytrop[i0] = c->laytrop;
                PUT(jp, jp); PUT(jt, jt); PUT(jt1, jt1); PUT(indself, indself); PUT(indfor, indfor);
                PUT(fac00, fac00); PUT(fac01, fac01); PUT(fac10, fac10); PUT(fac11, fac11);
                PUT(colh2o, colh2o); PUT(colco2, colco2); PUT(colo3, colo3); PUT(coln2o, coln2o);
                PUT(colch4, colch4); PUT(colo2, colo2); PUT(colmol, colmol);
                PUT(selffac, selffac); PUT(selffrac, selffrac); PUT(forfac, forfac); PUT(forfrac, forfrac);
#undef PUT
                for (int ig = 1; ig <= ORC_NGPTSW; ++ig) {
                    if (st->sfluxzen) st->sfluxzen[(long)(ig - 1) * ncol + i0] = c->zsflxzen[ig];
                    for (int l = 1; l <= nlay; ++l) {
                        if (st->taug) st->taug[((long)(ig - 1) * nlay + (l - 1)) * ncol + i0] = c->ztaug[ig][l];
                        if (st->taur) st->taur[((long)(ig - 1) * nlay + (l - 1)) * ncol + i0] = c->ztaur[ig][l];
                    }
                }
            }
        }
        free(c);
    }
    return cld_stop ? 10 + cld_stop : 0;
}
