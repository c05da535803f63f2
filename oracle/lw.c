/*
 * lw.c -- oracle restatement of RRTMG_LW as linked by MiMA (clear sky, icld=0; idrv = 0 and 1).
 * TEST INFRASTRUCTURE ONLY (see rrtmg_oracle.h).
 *
 * Follows  LW/src/rrtmg_lw_rad.nomcica.f90:80-569  (rrtmg_lw), :572-901 (inatm)
 *          LW/src/rrtmg_lw_cldprop.f90:154-162     (clear sky: ncbands=1, taucloud=0)
 *          LW/src/rrtmg_lw_setcoef.f90:31-415      (setcoef)
 *          LW/src/rrtmg_lw_taumol.f90:31-3149      (taumol, taugb1..16)
 *          LW/src/rrtmg_lw_rtrnmr.f90:259-280, 481-777 (rtrnmr, "Clear layer" branch)
 * One column at a time, loop order and arithmetic order as in the Fortran.  All arrays 1-based
 * like the Fortran (index 0 unused unless the Fortran array starts at 0).
 */
#include "rrtmg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NL (ORC_MAXLAY + 2)

/* Fortran-style accessors (1-based, column-major) */
#define F2(p, n1, i, j) ((p)[((long)(j) - 1) * (n1) + ((i) - 1)])
#define F3(p, n1, n2, i, j, k) ((p)[(((long)(k) - 1) * (n2) + ((j) - 1)) * (n1) + ((i) - 1)])
#define CHI(m, j) (S->chi_mls[((j) - 1) * 7 + ((m) - 1)])

static const int nspa[16] = {1, 1, 9, 9, 9, 1, 9, 1, 9, 1, 1, 9, 9, 1, 9, 9};
static const int nspb[16] = {1, 1, 5, 5, 5, 0, 1, 1, 1, 1, 1, 0, 0, 1, 0, 0};
static const int ngs[16] = {10, 22, 38, 52, 68, 76, 88, 96, 108, 114, 122, 130, 134, 136, 138, 140};
static const int ngc[16] = {10, 12, 16, 14, 16, 8, 12, 8, 12, 6, 8, 8, 4, 2, 2, 2};
static const double delwave[16] = {340., 150., 130., 70., 120., 160., 100., 100.,
                                   210., 90., 320., 280., 170., 130., 220., 650.};

typedef struct {
    int nlayers, laytrop;
    double pavel[NL], tavel[NL], pz[NL], tz[NL], tbound, coldry[NL], wbrodl[NL];
    double wkl[8][NL], wx[5][NL], pwvcm, semiss[17], taua[NL][17];
    int jp[NL], jt[NL], jt1[NL], indself[NL], indfor[NL], indminor[NL];
    double planklay[NL][17], planklev[NL][17], plankbnd[17];
    int idrv;
    double dplankbnd_dt[17], dtotuflux_dt[NL], dtotuclfl_dt[NL];   /* idrv = 1 (setcoef.f90:197-201, rtrnmr.f90:629-742) */
    double colh2o[NL], colco2[NL], colo3[NL], coln2o[NL], colco[NL], colch4[NL], colo2[NL], colbrd[NL];
    double fac00[NL], fac01[NL], fac10[NL], fac11[NL];
    double rat_h2oco2[NL], rat_h2oco2_1[NL], rat_h2oo3[NL], rat_h2oo3_1[NL];
    double rat_h2on2o[NL], rat_h2on2o_1[NL], rat_h2och4[NL], rat_h2och4_1[NL];
    double rat_n2oco2[NL], rat_n2oco2_1[NL], rat_o3co2[NL], rat_o3co2_1[NL];
    double selffac[NL], selffrac[NL], forfac[NL], forfrac[NL];
    double minorfrac[NL], scaleminor[NL], scaleminorn2[NL];
    double taug[ORC_NGPTLW + 1][NL], fracs[ORC_NGPTLW + 1][NL], taut[ORC_NGPTLW + 1][NL];
    double totuflux[NL], totdflux[NL], fnet[NL], htr[NL];
    double totuclfl[NL], totdclfl[NL], fnetc[NL], htrc[NL];
    double oneminus, fluxfac;
    /* clouds (icld >= 1): inatm :868-887, cldprop inflag = 0 (cldprop.f90:154-176); cldfrac[0] and cldfrac[nlayers+1]
       are the out-of-range elements rtrnmr.f90:403-404, 472-473 multiply by factors that are zero there */
    double cldfrac[NL + 2], taucloud[NL][17];
    int ncbands;   /* 1, 5 or 16: what cldprop leaves behind for the column; selects ipat in rtrn / rtrnmr */
} lwcol_t;

/* ---------------------------------------------------------------- inatm (rad.nomcica:572-901) */
static void inatm(lwcol_t *c, int iplon, int ncol, int nlay, int iaer,
                  const double *play, const double *plev, const double *tlay, const double *tlev,
                  const double *tsfc, const double *h2ovmr, const double *o3vmr, const double *co2vmr,
                  const double *ch4vmr, const double *n2ovmr, const double *o2vmr,
                  const double *cfc11vmr, const double *cfc12vmr, const double *cfc22vmr,
                  const double *ccl4vmr, const double *emis, const double *tauaer)
{
    const double amd = 28.9660, amw = 18.0160, amdw = 1.607793, amdo = 0.603428;
    const double grav = 9.8066, avogad = 6.02214199e+23;
    const int nmol = 7, maxxsec = 4;
    static const int ixindx[4] = {1, 2, 3, 4};
    double amm, amttl, wvttl, wvsh, summol;
#define IN2(a, l) ((a)[((long)(l) - 1) * ncol + (iplon - 1)])

    c->nlayers = nlay;
    for (int m = 0; m < 8; ++m)
        for (int l = 0; l < NL; ++l) c->wkl[m][l] = 0.0;
    for (int m = 0; m < 5; ++m)
        for (int l = 0; l < NL; ++l) c->wx[m][l] = 0.0;
    for (int l = 0; l < NL; ++l)
        for (int ib = 0; ib < 17; ++ib) c->taua[l][ib] = 0.0;
    amttl = 0.0;
    wvttl = 0.0;

    c->tbound = tsfc[iplon - 1];
    c->pz[0] = IN2(plev, 1);
    c->tz[0] = IN2(tlev, 1);
    for (int l = 1; l <= nlay; ++l) {
        c->pavel[l] = IN2(play, l);
        c->tavel[l] = IN2(tlay, l);
        c->pz[l] = IN2(plev, l + 1);
        c->tz[l] = IN2(tlev, l + 1);
        /* MiMA: h2o input is specific humidity, o3 is mass mixing ratio (:775-778) */
        c->wkl[1][l] = (IN2(h2ovmr, l) / (1.0 - IN2(h2ovmr, l))) * amdw;
        c->wkl[2][l] = IN2(co2vmr, l);
        c->wkl[3][l] = IN2(o3vmr, l) * amdo;
        c->wkl[4][l] = IN2(n2ovmr, l);
        c->wkl[6][l] = IN2(ch4vmr, l);
        c->wkl[7][l] = IN2(o2vmr, l);
        amm = (1.0 - c->wkl[1][l]) * amd + c->wkl[1][l] * amw;
        c->coldry[l] = (c->pz[l - 1] - c->pz[l]) * 1.e3 * avogad /
                       (1.e2 * grav * amm * (1.0 + c->wkl[1][l]));
    }
    for (int l = 1; l <= nlay; ++l) {
        c->wx[1][l] = IN2(ccl4vmr, l);
        c->wx[2][l] = IN2(cfc11vmr, l);
        c->wx[3][l] = IN2(cfc12vmr, l);
        c->wx[4][l] = IN2(cfc22vmr, l);
    }
    for (int l = 1; l <= nlay; ++l) {
        summol = 0.0;
        for (int imol = 2; imol <= nmol; ++imol) summol = summol + c->wkl[imol][l];
        c->wbrodl[l] = c->coldry[l] * (1.0 - summol);
        for (int imol = 1; imol <= nmol; ++imol) c->wkl[imol][l] = c->coldry[l] * c->wkl[imol][l];
        amttl = amttl + c->coldry[l] + c->wkl[1][l];
        wvttl = wvttl + c->wkl[1][l];
        for (int ix = 1; ix <= maxxsec; ++ix) {
            if (ixindx[ix - 1] != 0)
                c->wx[ixindx[ix - 1]][l] = c->coldry[l] * c->wx[ix][l] * 1.e-20;
        }
    }
    wvsh = (amw * wvttl) / (amd * amttl);
    c->pwvcm = wvsh * (1.e3 * c->pz[0]) / (1.e2 * grav);

    for (int n = 1; n <= 16; ++n) c->semiss[n] = emis[((long)n - 1) * ncol + (iplon - 1)];
    if (iaer >= 1) {
        for (int l = 1; l <= nlay; ++l)
            for (int ib = 1; ib <= 16; ++ib)
                c->taua[l][ib] = tauaer[(((long)ib - 1) * nlay + (l - 1)) * ncol + (iplon - 1)];
    }
#undef IN2
}

/* ---------------------------------------------------------------- setcoef (setcoef.f90:31-415) */
static void setcoef(lwcol_t *c, int istart)
{
    const orc_state_t *S = &g_orc;
    const int nlayers = c->nlayers;
    int indbound, indlev0, indlay, indlev, jp1;
    double stpfac, tbndfrac, t0frac, tlayfrac, tlevfrac, dbdtlev, dbdtlay;
    double plog, fp, ft, ft1, water, scalefac, factor, compfp;
#define TOTPLNK(i, b) F2(S->totplnk, 181, i, b)
#define TOTPLNKDERIV(i, b) F2(S->totplnkderiv, 181, i, b)
    (void)istart;
    stpfac = 296. / 1013.;

    indbound = (int)(c->tbound - 159.);
    if (indbound < 1) indbound = 1;
    else if (indbound > 180) indbound = 180;
    tbndfrac = c->tbound - 159. - (double)indbound;
    indlev0 = (int)(c->tz[0] - 159.);
    if (indlev0 < 1) indlev0 = 1;
    else if (indlev0 > 180) indlev0 = 180;
    t0frac = c->tz[0] - 159. - (double)indlev0;
    c->laytrop = 0;

    for (int lay = 1; lay <= nlayers; ++lay) {
        indlay = (int)(c->tavel[lay] - 159.);
        if (indlay < 1) indlay = 1;
        else if (indlay > 180) indlay = 180;
        tlayfrac = c->tavel[lay] - 159. - (double)indlay;
        indlev = (int)(c->tz[lay] - 159.);
        if (indlev < 1) indlev = 1;
        else if (indlev > 180) indlev = 180;
        tlevfrac = c->tz[lay] - 159. - (double)indlev;

        /* bands 1-15, then band 16 through the istart/=16 branch: identical formulas (:170-249) */
        for (int iband = 1; iband <= 16; ++iband) {
            if (lay == 1) {
                dbdtlev = TOTPLNK(indbound + 1, iband) - TOTPLNK(indbound, iband);
                c->plankbnd[iband] = c->semiss[iband] * (TOTPLNK(indbound, iband) + tbndfrac * dbdtlev);
                dbdtlev = TOTPLNK(indlev0 + 1, iband) - TOTPLNK(indlev0, iband);
                c->planklev[0][iband] = TOTPLNK(indlev0, iband) + t0frac * dbdtlev;
                if (c->idrv == 1) { /* :197-201, :238-242 */
                    dbdtlev = TOTPLNKDERIV(indbound + 1, iband) - TOTPLNKDERIV(indbound, iband);
                    c->dplankbnd_dt[iband] = c->semiss[iband] * (TOTPLNKDERIV(indbound, iband) + tbndfrac * dbdtlev);
                }
            }
            dbdtlev = TOTPLNK(indlev + 1, iband) - TOTPLNK(indlev, iband);
            dbdtlay = TOTPLNK(indlay + 1, iband) - TOTPLNK(indlay, iband);
            c->planklay[lay][iband] = TOTPLNK(indlay, iband) + tlayfrac * dbdtlay;
            c->planklev[lay][iband] = TOTPLNK(indlev, iband) + tlevfrac * dbdtlev;
        }

        plog = log(c->pavel[lay]);
        c->jp[lay] = (int)(36. - 5 * (plog + 0.04));
        if (c->jp[lay] < 1) c->jp[lay] = 1;
        else if (c->jp[lay] > 58) c->jp[lay] = 58;
        jp1 = c->jp[lay] + 1;
        fp = 5. * (S->lw_preflog[c->jp[lay] - 1] - plog);

        c->jt[lay] = (int)(3. + (c->tavel[lay] - S->lw_tref[c->jp[lay] - 1]) / 15.);
        if (c->jt[lay] < 1) c->jt[lay] = 1;
        else if (c->jt[lay] > 4) c->jt[lay] = 4;
        ft = ((c->tavel[lay] - S->lw_tref[c->jp[lay] - 1]) / 15.) - (double)(c->jt[lay] - 3);
        c->jt1[lay] = (int)(3. + (c->tavel[lay] - S->lw_tref[jp1 - 1]) / 15.);
        if (c->jt1[lay] < 1) c->jt1[lay] = 1;
        else if (c->jt1[lay] > 4) c->jt1[lay] = 4;
        ft1 = ((c->tavel[lay] - S->lw_tref[jp1 - 1]) / 15.) - (double)(c->jt1[lay] - 3);
        water = c->wkl[1][lay] / c->coldry[lay];
        scalefac = c->pavel[lay] * stpfac / c->tavel[lay];

        if (!(plog <= 4.56)) {
            /* below the tropopause (:293-348) */
            c->laytrop = c->laytrop + 1;
            c->forfac[lay] = scalefac / (1. + water);
            factor = (332.0 - c->tavel[lay]) / 36.0;
            c->indfor[lay] = (int)fmin(2, fmax(1, (int)factor));
            c->forfrac[lay] = factor - (double)c->indfor[lay];
            c->selffac[lay] = water * c->forfac[lay];
            factor = (c->tavel[lay] - 188.0) / 7.2;
            {
                int t = (int)factor - 7;
                c->indself[lay] = t < 1 ? 1 : (t > 9 ? 9 : t);
            }
            c->selffrac[lay] = factor - (double)(c->indself[lay] + 7);
            c->scaleminor[lay] = c->pavel[lay] / c->tavel[lay];
            c->scaleminorn2[lay] = (c->pavel[lay] / c->tavel[lay]) *
                                   (c->wbrodl[lay] / (c->coldry[lay] + c->wkl[1][lay]));
            factor = (c->tavel[lay] - 180.8) / 7.2;
            {
                int t = (int)factor;
                c->indminor[lay] = t < 1 ? 1 : (t > 18 ? 18 : t);
            }
            c->minorfrac[lay] = factor - (double)c->indminor[lay];
            {
                const int j = c->jp[lay];
                c->rat_h2oco2[lay] = CHI(1, j) / CHI(2, j);
                c->rat_h2oco2_1[lay] = CHI(1, j + 1) / CHI(2, j + 1);
                c->rat_h2oo3[lay] = CHI(1, j) / CHI(3, j);
                c->rat_h2oo3_1[lay] = CHI(1, j + 1) / CHI(3, j + 1);
                c->rat_h2on2o[lay] = CHI(1, j) / CHI(4, j);
                c->rat_h2on2o_1[lay] = CHI(1, j + 1) / CHI(4, j + 1);
                c->rat_h2och4[lay] = CHI(1, j) / CHI(6, j);
                c->rat_h2och4_1[lay] = CHI(1, j + 1) / CHI(6, j + 1);
                c->rat_n2oco2[lay] = CHI(4, j) / CHI(2, j);
                c->rat_n2oco2_1[lay] = CHI(4, j + 1) / CHI(2, j + 1);
            }
        } else {
            /* above the tropopause (:350-398) */
            c->forfac[lay] = scalefac / (1. + water);
            factor = (c->tavel[lay] - 188.0) / 36.0;
            c->indfor[lay] = 3;
            c->forfrac[lay] = factor - 1.0;
            c->selffac[lay] = water * c->forfac[lay];
            c->scaleminor[lay] = c->pavel[lay] / c->tavel[lay];
            c->scaleminorn2[lay] = (c->pavel[lay] / c->tavel[lay]) *
                                   (c->wbrodl[lay] / (c->coldry[lay] + c->wkl[1][lay]));
            factor = (c->tavel[lay] - 180.8) / 7.2;
            {
                int t = (int)factor;
                c->indminor[lay] = t < 1 ? 1 : (t > 18 ? 18 : t);
            }
            c->minorfrac[lay] = factor - (double)c->indminor[lay];
            {
                const int j = c->jp[lay];
                c->rat_h2oco2[lay] = CHI(1, j) / CHI(2, j);
                c->rat_h2oco2_1[lay] = CHI(1, j + 1) / CHI(2, j + 1);
                c->rat_o3co2[lay] = CHI(3, j) / CHI(2, j);
                c->rat_o3co2_1[lay] = CHI(3, j + 1) / CHI(2, j + 1);
            }
            /* not assigned by the Fortran above laytrop; defined here so stage dumps are deterministic */
            c->indself[lay] = 0;
            c->selffrac[lay] = 0.0;
        }
        /* column amounts (:335-348 / :378-391) -- same in both branches */
        c->colh2o[lay] = 1.e-20 * c->wkl[1][lay];
        c->colco2[lay] = 1.e-20 * c->wkl[2][lay];
        c->colo3[lay] = 1.e-20 * c->wkl[3][lay];
        c->coln2o[lay] = 1.e-20 * c->wkl[4][lay];
        c->colco[lay] = 1.e-20 * c->wkl[5][lay];
        c->colch4[lay] = 1.e-20 * c->wkl[6][lay];
        c->colo2[lay] = 1.e-20 * c->wkl[7][lay];
        if (c->colco2[lay] == 0.) c->colco2[lay] = 1.e-32 * c->coldry[lay];
        if (c->colo3[lay] == 0.) c->colo3[lay] = 1.e-32 * c->coldry[lay];
        if (c->coln2o[lay] == 0.) c->coln2o[lay] = 1.e-32 * c->coldry[lay];
        if (c->colco[lay] == 0.) c->colco[lay] = 1.e-32 * c->coldry[lay];
        if (c->colch4[lay] == 0.) c->colch4[lay] = 1.e-32 * c->coldry[lay];
        c->colbrd[lay] = 1.e-20 * c->wbrodl[lay];

        compfp = 1. - fp;
        c->fac10[lay] = compfp * ft;
        c->fac00[lay] = compfp * (1. - ft);
        c->fac11[lay] = fp * ft1;
        c->fac01[lay] = fp * (1. - ft1);
        c->selffac[lay] = c->colh2o[lay] * c->selffac[lay];
        c->forfac[lay] = c->colh2o[lay] * c->forfac[lay];
    }
#undef TOTPLNK
}

/* ---------------------------------------------------------------- taumol helpers */
/* The 3-point / 2-point binary-species stencil block that is textually identical in
 * taugb3,4,5,7,9,12,13,15,16 (template: taumol.f90:548-606 and :622-680). */
typedef struct { double f0, f1, f2, g0, g1, g2; int mode; } stencil_t;

static stencil_t stencil(double specparm, double fs, double facA, double facB)
{
    stencil_t s;
    double p, p4, fk0, fk1, fk2;
    if (specparm < 0.125) {
        p = fs - 1;
        p4 = (p * p) * (p * p);
        fk0 = p4;
        fk1 = 1 - p - 2.0 * p4;
        fk2 = p + p4;
        s.f0 = fk0 * facA; s.f1 = fk1 * facA; s.f2 = fk2 * facA;
        s.g0 = fk0 * facB; s.g1 = fk1 * facB; s.g2 = fk2 * facB;
        s.mode = 0;
    } else if (specparm > 0.875) {
        p = -fs;
        p4 = (p * p) * (p * p);
        fk0 = p4;
        fk1 = 1 - p - 2.0 * p4;
        fk2 = p + p4;
        s.f0 = fk0 * facA; s.f1 = fk1 * facA; s.f2 = fk2 * facA;
        s.g0 = fk0 * facB; s.g1 = fk1 * facB; s.g2 = fk2 * facB;
        s.mode = 1;
    } else {
        s.f0 = (1. - fs) * facA; s.g0 = (1. - fs) * facB;
        s.f1 = fs * facA;        s.g1 = fs * facB;
        s.f2 = 0.0; s.g2 = 0.0;
        s.mode = 2;
    }
    return s;
}

/* tau_major = speccomb * (...), with fac000=f0, fac100=f1, fac200=f2, fac010=g0, fac110=g1, fac210=g2 */
static double tau_major(const stencil_t *s, double speccomb, const double *absa, int na, int ind, int ig)
{
#define ABSA(i) F2(absa, na, i, ig)
    if (s->mode == 0)
        return speccomb * (s->f0 * ABSA(ind) + s->f1 * ABSA(ind + 1) + s->f2 * ABSA(ind + 2) +
                           s->g0 * ABSA(ind + 9) + s->g1 * ABSA(ind + 10) + s->g2 * ABSA(ind + 11));
    else if (s->mode == 1)
        return speccomb * (s->f2 * ABSA(ind - 1) + s->f1 * ABSA(ind) + s->f0 * ABSA(ind + 1) +
                           s->g2 * ABSA(ind + 8) + s->g1 * ABSA(ind + 9) + s->g0 * ABSA(ind + 10));
    else
        return speccomb * (s->f0 * ABSA(ind) + s->f1 * ABSA(ind + 1) +
                           s->g0 * ABSA(ind + 9) + s->g1 * ABSA(ind + 10));
#undef ABSA
}

/* eta = colA/(colA + rat*colB) clamped, js = 1+int(mult*eta), fs = mod(mult*eta,1) */
static void binary(double colA, double rat, double colB, double mult, double oneminus,
                   double *speccomb, double *specparm, int *js, double *fs)
{
    double specmult;
    *speccomb = colA + rat * colB;
    *specparm = colA / *speccomb;
    if (*specparm >= oneminus) *specparm = oneminus;
    specmult = mult * (*specparm);
    *js = 1 + (int)specmult;
    *fs = fmod(specmult, 1.0);
}

#define SELF(K, lay, inds, ig) (c->selffac[lay] * (F2((K)->selfref, 10, inds, ig) + c->selffrac[lay] * \
                                (F2((K)->selfref, 10, (inds) + 1, ig) - F2((K)->selfref, 10, inds, ig))))
#define FORN(K, lay, indf, ig) (c->forfac[lay] * (F2((K)->forref, 4, indf, ig) + c->forfrac[lay] * \
                                (F2((K)->forref, 4, (indf) + 1, ig) - F2((K)->forref, 4, indf, ig))))
#define MINOR1(tab, indm, ig) (F2(tab, 19, indm, ig) + c->minorfrac[lay] * \
                               (F2(tab, 19, (indm) + 1, ig) - F2(tab, 19, indm, ig)))
/* minor gas with eta dimension: (neta,19,ng) */
static double minor_eta(const double *tab, int neta, int jm, double fm, int indm, double minorfrac, int ig)
{
    double m1 = F3(tab, neta, 19, jm, indm, ig) + fm * (F3(tab, neta, 19, jm + 1, indm, ig) - F3(tab, neta, 19, jm, indm, ig));
    double m2 = F3(tab, neta, 19, jm, indm + 1, ig) + fm * (F3(tab, neta, 19, jm + 1, indm + 1, ig) - F3(tab, neta, 19, jm, indm + 1, ig));
    return m1 + minorfrac * (m2 - m1);
}
/* 4-point single-species key interpolation */
#define KEY4(abs_, n_, ind0, ind1, ig) (c->fac00[lay] * F2(abs_, n_, ind0, ig) + c->fac10[lay] * F2(abs_, n_, (ind0) + 1, ig) + \
                                        c->fac01[lay] * F2(abs_, n_, ind1, ig) + c->fac11[lay] * F2(abs_, n_, (ind1) + 1, ig))

#define IND0A(b) (((c->jp[lay] - 1) * 5 + (c->jt[lay] - 1)) * nspa[(b) - 1])
#define IND1A(b) ((c->jp[lay] * 5 + (c->jt1[lay] - 1)) * nspa[(b) - 1])
#define IND0B(b) (((c->jp[lay] - 13) * 5 + (c->jt[lay] - 1)) * nspb[(b) - 1])
#define IND1B(b) (((c->jp[lay] - 12) * 5 + (c->jt1[lay] - 1)) * nspb[(b) - 1])

/* upper-atmosphere binary band (2-point eta stencil): taumol.f90:681-757 template */
static double upper_binary(const lwcol_t *c, int lay, const double *absb, int nb, int ind0, int ind1,
                           double speccomb, double fs, double speccomb1, double fs1, int ig)
{
    double fac000 = (1. - fs) * c->fac00[lay], fac010 = (1. - fs) * c->fac10[lay];
    double fac100 = fs * c->fac00[lay], fac110 = fs * c->fac10[lay];
    double fac001 = (1. - fs1) * c->fac01[lay], fac011 = (1. - fs1) * c->fac11[lay];
    double fac101 = fs1 * c->fac01[lay], fac111 = fs1 * c->fac11[lay];
    return speccomb * (fac000 * F2(absb, nb, ind0, ig) + fac100 * F2(absb, nb, ind0 + 1, ig) +
                       fac010 * F2(absb, nb, ind0 + 5, ig) + fac110 * F2(absb, nb, ind0 + 6, ig)) +
           speccomb1 * (fac001 * F2(absb, nb, ind1, ig) + fac101 * F2(absb, nb, ind1 + 1, ig) +
                        fac011 * F2(absb, nb, ind1 + 5, ig) + fac111 * F2(absb, nb, ind1 + 6, ig));
}

/* ---------------------------------------------------------------- taumol (taumol.f90:260-3147) */
static void taumol(lwcol_t *c)
{
    const orc_state_t *S = &g_orc;
    const int nlayers = c->nlayers, laytrop = c->laytrop;
    const double oneminus = c->oneminus;
    int lay, ig, ind0, ind1, inds, indf, indm, js, js1, jpl, jm, jm2;
    double speccomb, specparm, fs, speccomb1, specparm1, fs1, sc, sp, fpl, fm, fm2;
    double tauself, taufor, corradj, pp, scalen2, adjfac, adjcol, chi, rat, ratx;
    const orc_lw_kg_t *K;

    /* ---- band 1: 10-350 cm-1 (low key h2o; high key h2o), N2 minor (:280-373) */
    K = &S->lw[0];
    for (lay = 1; lay <= laytrop; ++lay) {
        ind0 = IND0A(1) + 1; ind1 = IND1A(1) + 1;
        inds = c->indself[lay]; indf = c->indfor[lay]; indm = c->indminor[lay];
        pp = c->pavel[lay];
        corradj = 1.;
        if (pp < 250.) corradj = 1. - 0.15 * (250. - pp) / 154.4;
        scalen2 = c->colbrd[lay] * c->scaleminorn2[lay];
        for (ig = 1; ig <= ngc[0]; ++ig) {
            tauself = SELF(K, lay, inds, ig);
            taufor = FORN(K, lay, indf, ig);
            double taun2 = scalen2 * MINOR1(K->ka_mn2, indm, ig);
            c->taug[ig][lay] = corradj * (c->colh2o[lay] * KEY4(K->absa, 65, ind0, ind1, ig) + tauself + taufor + taun2);
            c->fracs[ig][lay] = K->fracrefa[ig - 1];
        }
    }
    for (lay = laytrop + 1; lay <= nlayers; ++lay) {
        ind0 = IND0B(1) + 1; ind1 = IND1B(1) + 1;
        indf = c->indfor[lay]; indm = c->indminor[lay];
        pp = c->pavel[lay];
        corradj = 1. - 0.15 * (pp / 95.6);
        scalen2 = c->colbrd[lay] * c->scaleminorn2[lay];
        for (ig = 1; ig <= ngc[0]; ++ig) {
            taufor = FORN(K, lay, indf, ig);
            double taun2 = scalen2 * MINOR1(K->kb_mn2, indm, ig);
            c->taug[ig][lay] = corradj * (c->colh2o[lay] * KEY4(K->absb, 235, ind0, ind1, ig) + taufor + taun2);
            c->fracs[ig][lay] = K->fracrefb[ig - 1];
        }
    }

    /* ---- band 2: 350-500 (h2o; h2o) (:376-445) */
    K = &S->lw[1];
    for (lay = 1; lay <= laytrop; ++lay) {
        ind0 = IND0A(2) + 1; ind1 = IND1A(2) + 1;
        inds = c->indself[lay]; indf = c->indfor[lay];
        pp = c->pavel[lay];
        corradj = 1. - .05 * (pp - 100.) / 900.;
        for (ig = 1; ig <= ngc[1]; ++ig) {
            tauself = SELF(K, lay, inds, ig);
            taufor = FORN(K, lay, indf, ig);
            c->taug[ngs[0] + ig][lay] = corradj * (c->colh2o[lay] * KEY4(K->absa, 65, ind0, ind1, ig) + tauself + taufor);
            c->fracs[ngs[0] + ig][lay] = K->fracrefa[ig - 1];
        }
    }
    for (lay = laytrop + 1; lay <= nlayers; ++lay) {
        ind0 = IND0B(2) + 1; ind1 = IND1B(2) + 1;
        indf = c->indfor[lay];
        for (ig = 1; ig <= ngc[1]; ++ig) {
            taufor = FORN(K, lay, indf, ig);
            c->taug[ngs[0] + ig][lay] = c->colh2o[lay] * KEY4(K->absb, 235, ind0, ind1, ig) + taufor;
            c->fracs[ngs[0] + ig][lay] = K->fracrefb[ig - 1];
        }
    }

    /* ---- band 3: 500-630 (h2o,co2; h2o,co2), N2O minor (:448-760) */
    K = &S->lw[2];
    {
        const double refrat_planck_a = CHI(1, 9) / CHI(2, 9);
        const double refrat_planck_b = CHI(1, 13) / CHI(2, 13);
        const double refrat_m_a = CHI(1, 3) / CHI(2, 3);
        const double refrat_m_b = CHI(1, 13) / CHI(2, 13);
        for (lay = 1; lay <= laytrop; ++lay) {
            binary(c->colh2o[lay], c->rat_h2oco2[lay], c->colco2[lay], 8., oneminus, &speccomb, &specparm, &js, &fs);
            binary(c->colh2o[lay], c->rat_h2oco2_1[lay], c->colco2[lay], 8., oneminus, &speccomb1, &specparm1, &js1, &fs1);
            binary(c->colh2o[lay], refrat_m_a, c->colco2[lay], 8., oneminus, &sc, &sp, &jm, &fm);
            chi = c->coln2o[lay] / c->coldry[lay];
            rat = 1.e20 * chi / CHI(4, c->jp[lay] + 1);
            if (rat > 1.5) {
                adjfac = 0.5 + pow(rat - 0.5, 0.65);
                adjcol = adjfac * CHI(4, c->jp[lay] + 1) * c->coldry[lay] * 1.e-20;
            } else {
                adjcol = c->coln2o[lay];
            }
            binary(c->colh2o[lay], refrat_planck_a, c->colco2[lay], 8., oneminus, &sc, &sp, &jpl, &fpl);
            ind0 = IND0A(3) + js; ind1 = IND1A(3) + js1;
            inds = c->indself[lay]; indf = c->indfor[lay]; indm = c->indminor[lay];
            stencil_t s0 = stencil(specparm, fs, c->fac00[lay], c->fac10[lay]);
            stencil_t s1 = stencil(specparm1, fs1, c->fac01[lay], c->fac11[lay]);
            for (ig = 1; ig <= ngc[2]; ++ig) {
                tauself = SELF(K, lay, inds, ig);
                taufor = FORN(K, lay, indf, ig);
                double absn2o = minor_eta(K->ka_mn2o, 9, jm, fm, indm, c->minorfrac[lay], ig);
                double tm = tau_major(&s0, speccomb, K->absa, 585, ind0, ig);
                double tm1 = tau_major(&s1, speccomb1, K->absa, 585, ind1, ig);
                c->taug[ngs[1] + ig][lay] = tm + tm1 + tauself + taufor + adjcol * absn2o;
                c->fracs[ngs[1] + ig][lay] = F2(K->fracrefa, 16, ig, jpl) + fpl * (F2(K->fracrefa, 16, ig, jpl + 1) - F2(K->fracrefa, 16, ig, jpl));
            }
        }
        for (lay = laytrop + 1; lay <= nlayers; ++lay) {
            binary(c->colh2o[lay], c->rat_h2oco2[lay], c->colco2[lay], 4., oneminus, &speccomb, &specparm, &js, &fs);
            binary(c->colh2o[lay], c->rat_h2oco2_1[lay], c->colco2[lay], 4., oneminus, &speccomb1, &specparm1, &js1, &fs1);
            binary(c->colh2o[lay], refrat_m_b, c->colco2[lay], 4., oneminus, &sc, &sp, &jm, &fm);
            chi = c->coln2o[lay] / c->coldry[lay];
            rat = 1.e20 * chi / CHI(4, c->jp[lay] + 1);
            if (rat > 1.5) {
                adjfac = 0.5 + pow(rat - 0.5, 0.65);
                adjcol = adjfac * CHI(4, c->jp[lay] + 1) * c->coldry[lay] * 1.e-20;
            } else {
                adjcol = c->coln2o[lay];
            }
            binary(c->colh2o[lay], refrat_planck_b, c->colco2[lay], 4., oneminus, &sc, &sp, &jpl, &fpl);
            ind0 = IND0B(3) + js; ind1 = IND1B(3) + js1;
            indf = c->indfor[lay]; indm = c->indminor[lay];
            for (ig = 1; ig <= ngc[2]; ++ig) {
                taufor = FORN(K, lay, indf, ig);
                double absn2o = minor_eta(K->kb_mn2o, 5, jm, fm, indm, c->minorfrac[lay], ig);
                c->taug[ngs[1] + ig][lay] = upper_binary(c, lay, K->absb, 1175, ind0, ind1, speccomb, fs, speccomb1, fs1, ig)
                                            + taufor + adjcol * absn2o;
                c->fracs[ngs[1] + ig][lay] = F2(K->fracrefb, 16, ig, jpl) + fpl * (F2(K->fracrefb, 16, ig, jpl + 1) - F2(K->fracrefb, 16, ig, jpl));
            }
        }
    }

    /* ---- band 4: 630-700 (h2o,co2; o3,co2) (:763-1019) */
    K = &S->lw[3];
    {
        const double refrat_planck_a = CHI(1, 11) / CHI(2, 11);
        const double refrat_planck_b = CHI(3, 13) / CHI(2, 13);
        const int n4 = 14;
        for (lay = 1; lay <= laytrop; ++lay) {
            binary(c->colh2o[lay], c->rat_h2oco2[lay], c->colco2[lay], 8., oneminus, &speccomb, &specparm, &js, &fs);
            binary(c->colh2o[lay], c->rat_h2oco2_1[lay], c->colco2[lay], 8., oneminus, &speccomb1, &specparm1, &js1, &fs1);
            binary(c->colh2o[lay], refrat_planck_a, c->colco2[lay], 8., oneminus, &sc, &sp, &jpl, &fpl);
            ind0 = IND0A(4) + js; ind1 = IND1A(4) + js1;
            inds = c->indself[lay]; indf = c->indfor[lay];
            stencil_t s0 = stencil(specparm, fs, c->fac00[lay], c->fac10[lay]);
            stencil_t s1 = stencil(specparm1, fs1, c->fac01[lay], c->fac11[lay]);
            for (ig = 1; ig <= n4; ++ig) {
                tauself = SELF(K, lay, inds, ig);
                taufor = FORN(K, lay, indf, ig);
                double tm = tau_major(&s0, speccomb, K->absa, 585, ind0, ig);
                double tm1 = tau_major(&s1, speccomb1, K->absa, 585, ind1, ig);
                c->taug[ngs[2] + ig][lay] = tm + tm1 + tauself + taufor;
                c->fracs[ngs[2] + ig][lay] = F2(K->fracrefa, n4, ig, jpl) + fpl * (F2(K->fracrefa, n4, ig, jpl + 1) - F2(K->fracrefa, n4, ig, jpl));
            }
        }
        for (lay = laytrop + 1; lay <= nlayers; ++lay) {
            binary(c->colo3[lay], c->rat_o3co2[lay], c->colco2[lay], 4., oneminus, &speccomb, &specparm, &js, &fs);
            binary(c->colo3[lay], c->rat_o3co2_1[lay], c->colco2[lay], 4., oneminus, &speccomb1, &specparm1, &js1, &fs1);
            binary(c->colo3[lay], refrat_planck_b, c->colco2[lay], 4., oneminus, &sc, &sp, &jpl, &fpl);
            ind0 = IND0B(4) + js; ind1 = IND1B(4) + js1;
            for (ig = 1; ig <= n4; ++ig) {
                c->taug[ngs[2] + ig][lay] = upper_binary(c, lay, K->absb, 1175, ind0, ind1, speccomb, fs, speccomb1, fs1, ig);
                c->fracs[ngs[2] + ig][lay] = F2(K->fracrefb, n4, ig, jpl) + fpl * (F2(K->fracrefb, n4, ig, jpl + 1) - F2(K->fracrefb, n4, ig, jpl));
            }
            /* empirical stratospheric CO2 correction (:1009-1015) */
            c->taug[ngs[2] + 8][lay] = c->taug[ngs[2] + 8][lay] * 0.92;
            c->taug[ngs[2] + 9][lay] = c->taug[ngs[2] + 9][lay] * 0.88;
            c->taug[ngs[2] + 10][lay] = c->taug[ngs[2] + 10][lay] * 1.07;
            c->taug[ngs[2] + 11][lay] = c->taug[ngs[2] + 11][lay] * 1.1;
            c->taug[ngs[2] + 12][lay] = c->taug[ngs[2] + 12][lay] * 0.99;
            c->taug[ngs[2] + 13][lay] = c->taug[ngs[2] + 13][lay] * 0.88;
            c->taug[ngs[2] + 14][lay] = c->taug[ngs[2] + 14][lay] * 0.943;
        }
    }

    /* ---- band 5: 700-820 (h2o,co2; o3,co2), O3 minor, CCl4 (:1022-1294) */
    K = &S->lw[4];
    {
        const double refrat_planck_a = CHI(1, 5) / CHI(2, 5);
        const double refrat_planck_b = CHI(3, 43) / CHI(2, 43);
        const double refrat_m_a = CHI(1, 7) / CHI(2, 7);
        for (lay = 1; lay <= laytrop; ++lay) {
            binary(c->colh2o[lay], c->rat_h2oco2[lay], c->colco2[lay], 8., oneminus, &speccomb, &specparm, &js, &fs);
            binary(c->colh2o[lay], c->rat_h2oco2_1[lay], c->colco2[lay], 8., oneminus, &speccomb1, &specparm1, &js1, &fs1);
            binary(c->colh2o[lay], refrat_m_a, c->colco2[lay], 8., oneminus, &sc, &sp, &jm, &fm);
            binary(c->colh2o[lay], refrat_planck_a, c->colco2[lay], 8., oneminus, &sc, &sp, &jpl, &fpl);
            ind0 = IND0A(5) + js; ind1 = IND1A(5) + js1;
            inds = c->indself[lay]; indf = c->indfor[lay]; indm = c->indminor[lay];
            stencil_t s0 = stencil(specparm, fs, c->fac00[lay], c->fac10[lay]);
            stencil_t s1 = stencil(specparm1, fs1, c->fac01[lay], c->fac11[lay]);
            for (ig = 1; ig <= 16; ++ig) {
                tauself = SELF(K, lay, inds, ig);
                taufor = FORN(K, lay, indf, ig);
                double abso3 = minor_eta(K->ka_mo3, 9, jm, fm, indm, c->minorfrac[lay], ig);
                double tm = tau_major(&s0, speccomb, K->absa, 585, ind0, ig);
                double tm1 = tau_major(&s1, speccomb1, K->absa, 585, ind1, ig);
                c->taug[ngs[3] + ig][lay] = tm + tm1 + tauself + taufor + abso3 * c->colo3[lay] + c->wx[1][lay] * K->ccl4[ig - 1];
                c->fracs[ngs[3] + ig][lay] = F2(K->fracrefa, 16, ig, jpl) + fpl * (F2(K->fracrefa, 16, ig, jpl + 1) - F2(K->fracrefa, 16, ig, jpl));
            }
        }
        for (lay = laytrop + 1; lay <= nlayers; ++lay) {
            binary(c->colo3[lay], c->rat_o3co2[lay], c->colco2[lay], 4., oneminus, &speccomb, &specparm, &js, &fs);
            binary(c->colo3[lay], c->rat_o3co2_1[lay], c->colco2[lay], 4., oneminus, &speccomb1, &specparm1, &js1, &fs1);
            binary(c->colo3[lay], refrat_planck_b, c->colco2[lay], 4., oneminus, &sc, &sp, &jpl, &fpl);
            ind0 = IND0B(5) + js; ind1 = IND1B(5) + js1;
            for (ig = 1; ig <= 16; ++ig) {
                c->taug[ngs[3] + ig][lay] = upper_binary(c, lay, K->absb, 1175, ind0, ind1, speccomb, fs, speccomb1, fs1, ig)
                                            + c->wx[1][lay] * K->ccl4[ig - 1];
                c->fracs[ngs[3] + ig][lay] = F2(K->fracrefb, 16, ig, jpl) + fpl * (F2(K->fracrefb, 16, ig, jpl + 1) - F2(K->fracrefb, 16, ig, jpl));
            }
        }
    }

    /* ---- band 6: 820-980 (h2o; nothing), CO2 minor, CFC11, CFC12 (:1297-1380) */
    K = &S->lw[5];
    for (lay = 1; lay <= laytrop; ++lay) {
        chi = c->colco2[lay] / (c->coldry[lay]);
        rat = 1.e20 * chi / CHI(2, c->jp[lay] + 1);
        if (rat > 3.0) {
            adjfac = 2.0 + pow(rat - 2.0, 0.77);
            adjcol = adjfac * CHI(2, c->jp[lay] + 1) * c->coldry[lay] * 1.e-20;
        } else {
            adjcol = c->colco2[lay];
        }
        ind0 = IND0A(6) + 1; ind1 = IND1A(6) + 1;
        inds = c->indself[lay]; indf = c->indfor[lay]; indm = c->indminor[lay];
        for (ig = 1; ig <= 8; ++ig) {
            tauself = SELF(K, lay, inds, ig);
            taufor = FORN(K, lay, indf, ig);
            double absco2 = MINOR1(K->ka_mco2, indm, ig);
            c->taug[ngs[4] + ig][lay] = c->colh2o[lay] * KEY4(K->absa, 65, ind0, ind1, ig) + tauself + taufor
                                        + adjcol * absco2 + c->wx[2][lay] * K->cfc11adj[ig - 1] + c->wx[3][lay] * K->cfc12[ig - 1];
            c->fracs[ngs[4] + ig][lay] = K->fracrefa[ig - 1];
        }
    }
    for (lay = laytrop + 1; lay <= nlayers; ++lay) {
        for (ig = 1; ig <= 8; ++ig) {
            c->taug[ngs[4] + ig][lay] = 0.0 + c->wx[2][lay] * K->cfc11adj[ig - 1] + c->wx[3][lay] * K->cfc12[ig - 1];
            c->fracs[ngs[4] + ig][lay] = K->fracrefa[ig - 1];
        }
    }

    /* ---- band 7: 980-1080 (h2o,o3; o3), CO2 minor (:1383-1654) */
    K = &S->lw[6];
    {
        const double refrat_planck_a = CHI(1, 3) / CHI(3, 3);
        const double refrat_m_a = CHI(1, 3) / CHI(3, 3);
        for (lay = 1; lay <= laytrop; ++lay) {
            binary(c->colh2o[lay], c->rat_h2oo3[lay], c->colo3[lay], 8., oneminus, &speccomb, &specparm, &js, &fs);
            binary(c->colh2o[lay], c->rat_h2oo3_1[lay], c->colo3[lay], 8., oneminus, &speccomb1, &specparm1, &js1, &fs1);
            binary(c->colh2o[lay], refrat_m_a, c->colo3[lay], 8., oneminus, &sc, &sp, &jm, &fm);
            chi = c->colco2[lay] / (c->coldry[lay]);
            rat = 1.e20 * chi / CHI(2, c->jp[lay] + 1);
            if (rat > 3.0) {
                adjfac = 3.0 + pow(rat - 3.0, 0.79);
                adjcol = adjfac * CHI(2, c->jp[lay] + 1) * c->coldry[lay] * 1.e-20;
            } else {
                adjcol = c->colco2[lay];
            }
            binary(c->colh2o[lay], refrat_planck_a, c->colo3[lay], 8., oneminus, &sc, &sp, &jpl, &fpl);
            ind0 = IND0A(7) + js; ind1 = IND1A(7) + js1;
            inds = c->indself[lay]; indf = c->indfor[lay]; indm = c->indminor[lay];
            stencil_t s0 = stencil(specparm, fs, c->fac00[lay], c->fac10[lay]);
            stencil_t s1 = stencil(specparm1, fs1, c->fac01[lay], c->fac11[lay]);
            for (ig = 1; ig <= 12; ++ig) {
                tauself = SELF(K, lay, inds, ig);
                taufor = FORN(K, lay, indf, ig);
                double absco2 = minor_eta(K->ka_mco2, 9, jm, fm, indm, c->minorfrac[lay], ig);
                double tm = tau_major(&s0, speccomb, K->absa, 585, ind0, ig);
                double tm1 = tau_major(&s1, speccomb1, K->absa, 585, ind1, ig);
                c->taug[ngs[5] + ig][lay] = tm + tm1 + tauself + taufor + adjcol * absco2;
                c->fracs[ngs[5] + ig][lay] = F2(K->fracrefa, 12, ig, jpl) + fpl * (F2(K->fracrefa, 12, ig, jpl + 1) - F2(K->fracrefa, 12, ig, jpl));
            }
        }
        for (lay = laytrop + 1; lay <= nlayers; ++lay) {
            chi = c->colco2[lay] / (c->coldry[lay]);
            rat = 1.e20 * chi / CHI(2, c->jp[lay] + 1);
            if (rat > 3.0) {
                adjfac = 2.0 + pow(rat - 2.0, 0.79);
                adjcol = adjfac * CHI(2, c->jp[lay] + 1) * c->coldry[lay] * 1.e-20;
            } else {
                adjcol = c->colco2[lay];
            }
            ind0 = IND0B(7) + 1; ind1 = IND1B(7) + 1;
            indm = c->indminor[lay];
            for (ig = 1; ig <= 12; ++ig) {
                double absco2 = MINOR1(K->kb_mco2, indm, ig);
                c->taug[ngs[5] + ig][lay] = c->colo3[lay] * KEY4(K->absb, 235, ind0, ind1, ig) + adjcol * absco2;
                c->fracs[ngs[5] + ig][lay] = K->fracrefb[ig - 1];
            }
            /* empirical stratospheric CO2 correction (:1645-1650) */
            c->taug[ngs[5] + 6][lay] = c->taug[ngs[5] + 6][lay] * 0.92;
            c->taug[ngs[5] + 7][lay] = c->taug[ngs[5] + 7][lay] * 0.88;
            c->taug[ngs[5] + 8][lay] = c->taug[ngs[5] + 8][lay] * 1.07;
            c->taug[ngs[5] + 9][lay] = c->taug[ngs[5] + 9][lay] * 1.1;
            c->taug[ngs[5] + 10][lay] = c->taug[ngs[5] + 10][lay] * 0.99;
            c->taug[ngs[5] + 11][lay] = c->taug[ngs[5] + 11][lay] * 0.855;
        }
    }

    /* ---- band 8: 1080-1180 (h2o; o3), CO2/O3/N2O minor, CFC12, CFC22 (:1657-1777) */
    K = &S->lw[7];
    for (lay = 1; lay <= laytrop; ++lay) {
        chi = c->colco2[lay] / (c->coldry[lay]);
        rat = 1.e20 * chi / CHI(2, c->jp[lay] + 1);
        if (rat > 3.0) {
            adjfac = 2.0 + pow(rat - 2.0, 0.65);
            adjcol = adjfac * CHI(2, c->jp[lay] + 1) * c->coldry[lay] * 1.e-20;
        } else {
            adjcol = c->colco2[lay];
        }
        ind0 = IND0A(8) + 1; ind1 = IND1A(8) + 1;
        inds = c->indself[lay]; indf = c->indfor[lay]; indm = c->indminor[lay];
        for (ig = 1; ig <= 8; ++ig) {
            tauself = SELF(K, lay, inds, ig);
            taufor = FORN(K, lay, indf, ig);
            double absco2 = MINOR1(K->ka_mco2, indm, ig);
            double abso3 = MINOR1(K->ka_mo3, indm, ig);
            double absn2o = MINOR1(K->ka_mn2o, indm, ig);
            c->taug[ngs[6] + ig][lay] = c->colh2o[lay] * KEY4(K->absa, 65, ind0, ind1, ig) + tauself + taufor
                                        + adjcol * absco2 + c->colo3[lay] * abso3 + c->coln2o[lay] * absn2o
                                        + c->wx[3][lay] * K->cfc12[ig - 1] + c->wx[4][lay] * K->cfc22adj[ig - 1];
            c->fracs[ngs[6] + ig][lay] = K->fracrefa[ig - 1];
        }
    }
    for (lay = laytrop + 1; lay <= nlayers; ++lay) {
        chi = c->colco2[lay] / c->coldry[lay];
        rat = 1.e20 * chi / CHI(2, c->jp[lay] + 1);
        if (rat > 3.0) {
            adjfac = 2.0 + pow(rat - 2.0, 0.65);
            adjcol = adjfac * CHI(2, c->jp[lay] + 1) * c->coldry[lay] * 1.e-20;
        } else {
            adjcol = c->colco2[lay];
        }
        ind0 = IND0B(8) + 1; ind1 = IND1B(8) + 1;
        indm = c->indminor[lay];
        for (ig = 1; ig <= 8; ++ig) {
            double absco2 = MINOR1(K->kb_mco2, indm, ig);
            double absn2o = MINOR1(K->kb_mn2o, indm, ig);
            c->taug[ngs[6] + ig][lay] = c->colo3[lay] * KEY4(K->absb, 235, ind0, ind1, ig)
                                        + adjcol * absco2 + c->coln2o[lay] * absn2o
                                        + c->wx[3][lay] * K->cfc12[ig - 1] + c->wx[4][lay] * K->cfc22adj[ig - 1];
            c->fracs[ngs[6] + ig][lay] = K->fracrefb[ig - 1];
        }
    }

    /* ---- band 9: 1180-1390 (h2o,ch4; ch4), N2O minor (:1780-2040) */
    K = &S->lw[8];
    {
        const double refrat_planck_a = CHI(1, 9) / CHI(6, 9);
        const double refrat_m_a = CHI(1, 3) / CHI(6, 3);
        for (lay = 1; lay <= laytrop; ++lay) {
            binary(c->colh2o[lay], c->rat_h2och4[lay], c->colch4[lay], 8., oneminus, &speccomb, &specparm, &js, &fs);
            binary(c->colh2o[lay], c->rat_h2och4_1[lay], c->colch4[lay], 8., oneminus, &speccomb1, &specparm1, &js1, &fs1);
            binary(c->colh2o[lay], refrat_m_a, c->colch4[lay], 8., oneminus, &sc, &sp, &jm, &fm);
            chi = c->coln2o[lay] / (c->coldry[lay]);
            rat = 1.e20 * chi / CHI(4, c->jp[lay] + 1);
            if (rat > 1.5) {
                adjfac = 0.5 + pow(rat - 0.5, 0.65);
                adjcol = adjfac * CHI(4, c->jp[lay] + 1) * c->coldry[lay] * 1.e-20;
            } else {
                adjcol = c->coln2o[lay];
            }
            binary(c->colh2o[lay], refrat_planck_a, c->colch4[lay], 8., oneminus, &sc, &sp, &jpl, &fpl);
            ind0 = IND0A(9) + js; ind1 = IND1A(9) + js1;
            inds = c->indself[lay]; indf = c->indfor[lay]; indm = c->indminor[lay];
            stencil_t s0 = stencil(specparm, fs, c->fac00[lay], c->fac10[lay]);
            stencil_t s1 = stencil(specparm1, fs1, c->fac01[lay], c->fac11[lay]);
            for (ig = 1; ig <= 12; ++ig) {
                tauself = SELF(K, lay, inds, ig);
                taufor = FORN(K, lay, indf, ig);
                double absn2o = minor_eta(K->ka_mn2o, 9, jm, fm, indm, c->minorfrac[lay], ig);
                double tm = tau_major(&s0, speccomb, K->absa, 585, ind0, ig);
                double tm1 = tau_major(&s1, speccomb1, K->absa, 585, ind1, ig);
                c->taug[ngs[7] + ig][lay] = tm + tm1 + tauself + taufor + adjcol * absn2o;
                c->fracs[ngs[7] + ig][lay] = F2(K->fracrefa, 12, ig, jpl) + fpl * (F2(K->fracrefa, 12, ig, jpl + 1) - F2(K->fracrefa, 12, ig, jpl));
            }
        }
        for (lay = laytrop + 1; lay <= nlayers; ++lay) {
            chi = c->coln2o[lay] / (c->coldry[lay]);
            rat = 1.e20 * chi / CHI(4, c->jp[lay] + 1);
            if (rat > 1.5) {
                adjfac = 0.5 + pow(rat - 0.5, 0.65);
                adjcol = adjfac * CHI(4, c->jp[lay] + 1) * c->coldry[lay] * 1.e-20;
            } else {
                adjcol = c->coln2o[lay];
            }
            ind0 = IND0B(9) + 1; ind1 = IND1B(9) + 1;
            indm = c->indminor[lay];
            for (ig = 1; ig <= 12; ++ig) {
                double absn2o = MINOR1(K->kb_mn2o, indm, ig);
                c->taug[ngs[7] + ig][lay] = c->colch4[lay] * KEY4(K->absb, 235, ind0, ind1, ig) + adjcol * absn2o;
                c->fracs[ngs[7] + ig][lay] = K->fracrefb[ig - 1];
            }
        }
    }

    /* ---- band 10: 1390-1480 (h2o; h2o) (:2043-2107) */
    K = &S->lw[9];
    for (lay = 1; lay <= laytrop; ++lay) {
        ind0 = IND0A(10) + 1; ind1 = IND1A(10) + 1;
        inds = c->indself[lay]; indf = c->indfor[lay];
        for (ig = 1; ig <= 6; ++ig) {
            tauself = SELF(K, lay, inds, ig);
            taufor = FORN(K, lay, indf, ig);
            c->taug[ngs[8] + ig][lay] = c->colh2o[lay] * KEY4(K->absa, 65, ind0, ind1, ig) + tauself + taufor;
            c->fracs[ngs[8] + ig][lay] = K->fracrefa[ig - 1];
        }
    }
    for (lay = laytrop + 1; lay <= nlayers; ++lay) {
        ind0 = IND0B(10) + 1; ind1 = IND1B(10) + 1;
        indf = c->indfor[lay];
        for (ig = 1; ig <= 6; ++ig) {
            taufor = FORN(K, lay, indf, ig);
            c->taug[ngs[8] + ig][lay] = c->colh2o[lay] * KEY4(K->absb, 235, ind0, ind1, ig) + taufor;
            c->fracs[ngs[8] + ig][lay] = K->fracrefb[ig - 1];
        }
    }

    /* ---- band 11: 1480-1800 (h2o; h2o), O2 minor (:2110-2187) */
    K = &S->lw[10];
    for (lay = 1; lay <= laytrop; ++lay) {
        ind0 = IND0A(11) + 1; ind1 = IND1A(11) + 1;
        inds = c->indself[lay]; indf = c->indfor[lay]; indm = c->indminor[lay];
        double scaleo2 = c->colo2[lay] * c->scaleminor[lay];
        for (ig = 1; ig <= 8; ++ig) {
            tauself = SELF(K, lay, inds, ig);
            taufor = FORN(K, lay, indf, ig);
            double tauo2 = scaleo2 * MINOR1(K->ka_mo2, indm, ig);
            c->taug[ngs[9] + ig][lay] = c->colh2o[lay] * KEY4(K->absa, 65, ind0, ind1, ig) + tauself + taufor + tauo2;
            c->fracs[ngs[9] + ig][lay] = K->fracrefa[ig - 1];
        }
    }
    for (lay = laytrop + 1; lay <= nlayers; ++lay) {
        ind0 = IND0B(11) + 1; ind1 = IND1B(11) + 1;
        indf = c->indfor[lay]; indm = c->indminor[lay];
        double scaleo2 = c->colo2[lay] * c->scaleminor[lay];
        for (ig = 1; ig <= 8; ++ig) {
            taufor = FORN(K, lay, indf, ig);
            double tauo2 = scaleo2 * MINOR1(K->kb_mo2, indm, ig);
            c->taug[ngs[9] + ig][lay] = c->colh2o[lay] * KEY4(K->absb, 235, ind0, ind1, ig) + taufor + tauo2;
            c->fracs[ngs[9] + ig][lay] = K->fracrefb[ig - 1];
        }
    }

    /* ---- band 12: 1800-2080 (h2o,co2; nothing) (:2190-2392) */
    K = &S->lw[11];
    {
        const double refrat_planck_a = CHI(1, 10) / CHI(2, 10);
        for (lay = 1; lay <= laytrop; ++lay) {
            binary(c->colh2o[lay], c->rat_h2oco2[lay], c->colco2[lay], 8., oneminus, &speccomb, &specparm, &js, &fs);
            binary(c->colh2o[lay], c->rat_h2oco2_1[lay], c->colco2[lay], 8., oneminus, &speccomb1, &specparm1, &js1, &fs1);
            binary(c->colh2o[lay], refrat_planck_a, c->colco2[lay], 8., oneminus, &sc, &sp, &jpl, &fpl);
            ind0 = IND0A(12) + js; ind1 = IND1A(12) + js1;
            inds = c->indself[lay]; indf = c->indfor[lay];
            stencil_t s0 = stencil(specparm, fs, c->fac00[lay], c->fac10[lay]);
            stencil_t s1 = stencil(specparm1, fs1, c->fac01[lay], c->fac11[lay]);
            for (ig = 1; ig <= 8; ++ig) {
                tauself = SELF(K, lay, inds, ig);
                taufor = FORN(K, lay, indf, ig);
                double tm = tau_major(&s0, speccomb, K->absa, 585, ind0, ig);
                double tm1 = tau_major(&s1, speccomb1, K->absa, 585, ind1, ig);
                c->taug[ngs[10] + ig][lay] = tm + tm1 + tauself + taufor;
                c->fracs[ngs[10] + ig][lay] = F2(K->fracrefa, 8, ig, jpl) + fpl * (F2(K->fracrefa, 8, ig, jpl + 1) - F2(K->fracrefa, 8, ig, jpl));
            }
        }
        for (lay = laytrop + 1; lay <= nlayers; ++lay)
            for (ig = 1; ig <= 8; ++ig) {
                c->taug[ngs[10] + ig][lay] = 0.0;
                c->fracs[ngs[10] + ig][lay] = 0.0;
            }
    }

    /* ---- band 13: 2080-2250 (h2o,n2o; nothing), CO2+CO minor low, O3 minor high (:2395-2652) */
    K = &S->lw[12];
    {
        const double refrat_planck_a = CHI(1, 5) / CHI(4, 5);
        const double refrat_m_a = CHI(1, 1) / CHI(4, 1);
        const double refrat_m_a3 = CHI(1, 3) / CHI(4, 3);
        for (lay = 1; lay <= laytrop; ++lay) {
            binary(c->colh2o[lay], c->rat_h2on2o[lay], c->coln2o[lay], 8., oneminus, &speccomb, &specparm, &js, &fs);
            binary(c->colh2o[lay], c->rat_h2on2o_1[lay], c->coln2o[lay], 8., oneminus, &speccomb1, &specparm1, &js1, &fs1);
            binary(c->colh2o[lay], refrat_m_a, c->coln2o[lay], 8., oneminus, &sc, &sp, &jm, &fm);
            chi = c->colco2[lay] / (c->coldry[lay]);
            ratx = 1.e20 * chi / 3.55e-4;
            if (ratx > 3.0) {
                adjfac = 2.0 + pow(ratx - 2.0, 0.68);
                adjcol = adjfac * 3.55e-4 * c->coldry[lay] * 1.e-20;
            } else {
                adjcol = c->colco2[lay];
            }
            binary(c->colh2o[lay], refrat_m_a3, c->coln2o[lay], 8., oneminus, &sc, &sp, &jm2, &fm2);
            binary(c->colh2o[lay], refrat_planck_a, c->coln2o[lay], 8., oneminus, &sc, &sp, &jpl, &fpl);
            ind0 = IND0A(13) + js; ind1 = IND1A(13) + js1;
            inds = c->indself[lay]; indf = c->indfor[lay]; indm = c->indminor[lay];
            stencil_t s0 = stencil(specparm, fs, c->fac00[lay], c->fac10[lay]);
            stencil_t s1 = stencil(specparm1, fs1, c->fac01[lay], c->fac11[lay]);
            for (ig = 1; ig <= 4; ++ig) {
                tauself = SELF(K, lay, inds, ig);
                taufor = FORN(K, lay, indf, ig);
                double absco2 = minor_eta(K->ka_mco2, 9, jm, fm, indm, c->minorfrac[lay], ig);
                double absco = minor_eta(K->ka_mco, 9, jm2, fm2, indm, c->minorfrac[lay], ig);
                double tm = tau_major(&s0, speccomb, K->absa, 585, ind0, ig);
                double tm1 = tau_major(&s1, speccomb1, K->absa, 585, ind1, ig);
                c->taug[ngs[11] + ig][lay] = tm + tm1 + tauself + taufor + adjcol * absco2 + c->colco[lay] * absco;
                c->fracs[ngs[11] + ig][lay] = F2(K->fracrefa, 4, ig, jpl) + fpl * (F2(K->fracrefa, 4, ig, jpl + 1) - F2(K->fracrefa, 4, ig, jpl));
            }
        }
        for (lay = laytrop + 1; lay <= nlayers; ++lay) {
            indm = c->indminor[lay];
            for (ig = 1; ig <= 4; ++ig) {
                double abso3 = MINOR1(K->kb_mo3, indm, ig);
                c->taug[ngs[11] + ig][lay] = c->colo3[lay] * abso3;
                c->fracs[ngs[11] + ig][lay] = K->fracrefb[ig - 1];
            }
        }
    }

    /* ---- band 14: 2250-2380 (co2; co2) (:2655-2713) */
    K = &S->lw[13];
    for (lay = 1; lay <= laytrop; ++lay) {
        ind0 = IND0A(14) + 1; ind1 = IND1A(14) + 1;
        inds = c->indself[lay]; indf = c->indfor[lay];
        for (ig = 1; ig <= 2; ++ig) {
            tauself = SELF(K, lay, inds, ig);
            taufor = FORN(K, lay, indf, ig);
            c->taug[ngs[12] + ig][lay] = c->colco2[lay] * KEY4(K->absa, 65, ind0, ind1, ig) + tauself + taufor;
            c->fracs[ngs[12] + ig][lay] = K->fracrefa[ig - 1];
        }
    }
    for (lay = laytrop + 1; lay <= nlayers; ++lay) {
        ind0 = IND0B(14) + 1; ind1 = IND1B(14) + 1;
        for (ig = 1; ig <= 2; ++ig) {
            c->taug[ngs[12] + ig][lay] = c->colco2[lay] * KEY4(K->absb, 235, ind0, ind1, ig);
            c->fracs[ngs[12] + ig][lay] = K->fracrefb[ig - 1];
        }
    }

    /* ---- band 15: 2380-2600 (n2o,co2; nothing), N2 minor (:2716-2938) */
    K = &S->lw[14];
    {
        const double refrat_planck_a = CHI(4, 1) / CHI(2, 1);
        const double refrat_m_a = CHI(4, 1) / CHI(2, 1);
        for (lay = 1; lay <= laytrop; ++lay) {
            binary(c->coln2o[lay], c->rat_n2oco2[lay], c->colco2[lay], 8., oneminus, &speccomb, &specparm, &js, &fs);
            binary(c->coln2o[lay], c->rat_n2oco2_1[lay], c->colco2[lay], 8., oneminus, &speccomb1, &specparm1, &js1, &fs1);
            binary(c->coln2o[lay], refrat_m_a, c->colco2[lay], 8., oneminus, &sc, &sp, &jm, &fm);
            binary(c->coln2o[lay], refrat_planck_a, c->colco2[lay], 8., oneminus, &sc, &sp, &jpl, &fpl);
            ind0 = IND0A(15) + js; ind1 = IND1A(15) + js1;
            inds = c->indself[lay]; indf = c->indfor[lay]; indm = c->indminor[lay];
            scalen2 = c->colbrd[lay] * c->scaleminor[lay];
            stencil_t s0 = stencil(specparm, fs, c->fac00[lay], c->fac10[lay]);
            stencil_t s1 = stencil(specparm1, fs1, c->fac01[lay], c->fac11[lay]);
            for (ig = 1; ig <= 2; ++ig) {
                tauself = SELF(K, lay, inds, ig);
                taufor = FORN(K, lay, indf, ig);
                double taun2 = scalen2 * minor_eta(K->ka_mn2, 9, jm, fm, indm, c->minorfrac[lay], ig);
                double tm = tau_major(&s0, speccomb, K->absa, 585, ind0, ig);
                double tm1 = tau_major(&s1, speccomb1, K->absa, 585, ind1, ig);
                c->taug[ngs[13] + ig][lay] = tm + tm1 + tauself + taufor + taun2;
                c->fracs[ngs[13] + ig][lay] = F2(K->fracrefa, 2, ig, jpl) + fpl * (F2(K->fracrefa, 2, ig, jpl + 1) - F2(K->fracrefa, 2, ig, jpl));
            }
        }
        for (lay = laytrop + 1; lay <= nlayers; ++lay)
            for (ig = 1; ig <= 2; ++ig) {
                c->taug[ngs[13] + ig][lay] = 0.0;
                c->fracs[ngs[13] + ig][lay] = 0.0;
            }
    }

    /* ---- band 16: 2600-3250 (h2o,ch4; ch4) (:2941-3147) */
    K = &S->lw[15];
    {
        const double refrat_planck_a = CHI(1, 6) / CHI(6, 6);
        for (lay = 1; lay <= laytrop; ++lay) {
            binary(c->colh2o[lay], c->rat_h2och4[lay], c->colch4[lay], 8., oneminus, &speccomb, &specparm, &js, &fs);
            binary(c->colh2o[lay], c->rat_h2och4_1[lay], c->colch4[lay], 8., oneminus, &speccomb1, &specparm1, &js1, &fs1);
            binary(c->colh2o[lay], refrat_planck_a, c->colch4[lay], 8., oneminus, &sc, &sp, &jpl, &fpl);
            ind0 = IND0A(16) + js; ind1 = IND1A(16) + js1;
            inds = c->indself[lay]; indf = c->indfor[lay];
            stencil_t s0 = stencil(specparm, fs, c->fac00[lay], c->fac10[lay]);
            stencil_t s1 = stencil(specparm1, fs1, c->fac01[lay], c->fac11[lay]);
            for (ig = 1; ig <= 2; ++ig) {
                tauself = SELF(K, lay, inds, ig);
                taufor = FORN(K, lay, indf, ig);
                double tm = tau_major(&s0, speccomb, K->absa, 585, ind0, ig);
                double tm1 = tau_major(&s1, speccomb1, K->absa, 585, ind1, ig);
                c->taug[ngs[14] + ig][lay] = tm + tm1 + tauself + taufor;
                c->fracs[ngs[14] + ig][lay] = F2(K->fracrefa, 2, ig, jpl) + fpl * (F2(K->fracrefa, 2, ig, jpl + 1) - F2(K->fracrefa, 2, ig, jpl));
            }
        }
        for (lay = laytrop + 1; lay <= nlayers; ++lay) {
            ind0 = IND0B(16) + 1; ind1 = IND1B(16) + 1;
            for (ig = 1; ig <= 2; ++ig) {
                c->taug[ngs[14] + ig][lay] = c->colch4[lay] * KEY4(K->absb, 235, ind0, ind1, ig);
                c->fracs[ngs[14] + ig][lay] = K->fracrefb[ig - 1];
            }
        }
    }
}

/* ---------------------------------------------------------------- rtrnmr, clear branch (rtrnmr.f90) */
static void rtrnmr_clear(lwcol_t *c)
{
    const orc_state_t *S = &g_orc;
    const int nlayers = c->nlayers;
    const double tblint = 10000.0, bpade = S->lw_bpade;
    const double wtdiff = 0.5, rec_6 = 0.166667;
    static const double a0[16] = {1.66, 1.55, 1.58, 1.66, 1.54, 1.454, 1.89, 1.33, 1.668, 1.66, 1.66, 1.66, 1.66, 1.66, 1.66, 1.66};
    static const double a1[16] = {0.00, 0.25, 0.22, 0.00, 0.13, 0.446, -0.10, 0.40, -0.006, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00};
    static const double a2[16] = {0.00, -12.0, -11.7, 0.00, -0.72, -0.243, 0.19, -0.062, 0.414, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00};
    double secdiff[17], atrans[NL], bbugas[NL], urad[NL], drad[NL], clrurad[NL], clrdrad[NL];
    double d_urad_dt[NL], d_clrurad_dt[NL];
    const int idrv = c->idrv;
    double uflux, dflux, uclfl, dclfl;
    int igc, itr, lev, iband, ibnd, l;

    /* :259-280 */
    for (ibnd = 1; ibnd <= 16; ++ibnd) {
        if (ibnd == 1 || ibnd == 4 || ibnd >= 10) {
            secdiff[ibnd] = 1.66;
        } else {
            secdiff[ibnd] = a0[ibnd - 1] + a1[ibnd - 1] * exp(a2[ibnd - 1] * c->pwvcm);
            if (secdiff[ibnd] > 1.80) secdiff[ibnd] = 1.80;
            if (secdiff[ibnd] < 1.50) secdiff[ibnd] = 1.50;
        }
    }
    for (lev = 0; lev <= nlayers; ++lev) {
        urad[lev] = 0.0; drad[lev] = 0.0; clrurad[lev] = 0.0; clrdrad[lev] = 0.0;
        c->totuflux[lev] = 0.0; c->totdflux[lev] = 0.0; c->totuclfl[lev] = 0.0; c->totdclfl[lev] = 0.0;
        if (idrv == 1) { /* :292-315 */
            d_urad_dt[lev] = 0.0; d_clrurad_dt[lev] = 0.0;
            c->dtotuflux_dt[lev] = 0.0; c->dtotuclfl_dt[lev] = 0.0;
        }
    }
    igc = 1;
    for (iband = 1; iband <= 16; ++iband) {
        do { /* g-point loop: "1000 continue ... if (igc .le. ngs(iband)) go to 1000" */
            double radld = 0., radclrd = 0., radlu, radclru, rad0, reflect;
            /* downward loop (:505-618), clear-layer branch :589-607 and iclddn=0 branch :614-616 */
            for (lev = nlayers; lev >= 1; --lev) {
                double plfrac = c->fracs[igc][lev];
                double blay = c->planklay[lev][iband];
                double dplankup = c->planklev[lev][iband] - blay;
                double dplankdn = c->planklev[lev - 1][iband] - blay;
                double odepth = secdiff[iband] * c->taut[igc][lev];
                double bbd;
                if (odepth < 0.0) odepth = 0.0;
                if (odepth <= 0.06) {
                    atrans[lev] = odepth - 0.5 * odepth * odepth;
                    odepth = rec_6 * odepth;
                    bbd = plfrac * (blay + dplankdn * odepth);
                    bbugas[lev] = plfrac * (blay + dplankup * odepth);
                } else {
                    double tblind = odepth / (bpade + odepth);
                    itr = (int)(tblint * tblind + 0.5);
                    double transc = S->exp_tbl[itr];
                    atrans[lev] = 1. - transc;
                    double tausfac = S->tfn_tbl[itr];
                    bbd = plfrac * (blay + tausfac * dplankdn);
                    bbugas[lev] = plfrac * (blay + tausfac * dplankup);
                }
                radld = radld + (bbd - radld) * atrans[lev];
                drad[lev - 1] = drad[lev - 1] + radld;
                radclrd = radld;
                clrdrad[lev - 1] = drad[lev - 1];
            }
            /* surface (:628-636) */
            rad0 = c->fracs[igc][1] * c->plankbnd[iband];
            reflect = 1. - c->semiss[iband];
            radlu = rad0 + reflect * radld;
            radclru = rad0 + reflect * radclrd;
            double d_rad0_dt = 0., d_radlu_dt = 0., d_radclru_dt = 0.;
            if (idrv == 1) d_rad0_dt = c->fracs[igc][1] * c->dplankbnd_dt[iband];   /* :629-631 */
            urad[0] = urad[0] + radlu;
            clrurad[0] = clrurad[0] + radclru;
            if (idrv == 1) { /* :642-647 */
                d_radlu_dt = d_rad0_dt;
                d_urad_dt[0] = d_urad_dt[0] + d_radlu_dt;
                d_radclru_dt = d_rad0_dt;
                d_clrurad_dt[0] = d_clrurad_dt[0] + d_radclru_dt;
            }
            /* upward loop (:649-711), clear-layer branch :682-701 */
            for (lev = 1; lev <= nlayers; ++lev) {
                radlu = radlu + (bbugas[lev] - radlu) * atrans[lev];
                urad[lev] = urad[lev] + radlu;
                if (idrv == 1) { /* clear layer :686-689 */
                    d_radlu_dt = d_radlu_dt * (1.0 - atrans[lev]);
                    d_urad_dt[lev] = d_urad_dt[lev] + d_radlu_dt;
                }
                radclru = radlu;
                clrurad[lev] = urad[lev];
                if (idrv == 1) { /* iclddn = 0 branch :706-709 */
                    d_radclru_dt = d_radlu_dt;
                    d_clrurad_dt[lev] = d_urad_dt[lev];
                }
            }
            (void)d_radclru_dt;
            igc = igc + 1;
        } while (igc <= ngs[iband - 1]);

        /* band totals (:720-733) */
        for (lev = nlayers; lev >= 0; --lev) {
            uflux = urad[lev] * wtdiff;
            dflux = drad[lev] * wtdiff;
            urad[lev] = 0.0;
            drad[lev] = 0.0;
            c->totuflux[lev] = c->totuflux[lev] + uflux * delwave[iband - 1];
            c->totdflux[lev] = c->totdflux[lev] + dflux * delwave[iband - 1];
            uclfl = clrurad[lev] * wtdiff;
            dclfl = clrdrad[lev] * wtdiff;
            clrurad[lev] = 0.0;
            clrdrad[lev] = 0.0;
            c->totuclfl[lev] = c->totuclfl[lev] + uclfl * delwave[iband - 1];
            c->totdclfl[lev] = c->totdclfl[lev] + dclfl * delwave[iband - 1];
        }
        if (idrv == 1) { /* :736-746 */
            for (lev = nlayers; lev >= 0; --lev) {
                double duflux_dt = d_urad_dt[lev] * wtdiff;
                d_urad_dt[lev] = 0.0;
                c->dtotuflux_dt[lev] = c->dtotuflux_dt[lev] + duflux_dt * delwave[iband - 1] * c->fluxfac;
                double duclfl_dt = d_clrurad_dt[lev] * wtdiff;
                d_clrurad_dt[lev] = 0.0;
                c->dtotuclfl_dt[lev] = c->dtotuclfl_dt[lev] + duclfl_dt * delwave[iband - 1] * c->fluxfac;
            }
        }
    }
    /* fluxes and heating rates (:751-777) */
    c->totuflux[0] = c->totuflux[0] * c->fluxfac;
    c->totdflux[0] = c->totdflux[0] * c->fluxfac;
    c->fnet[0] = c->totuflux[0] - c->totdflux[0];
    c->totuclfl[0] = c->totuclfl[0] * c->fluxfac;
    c->totdclfl[0] = c->totdclfl[0] * c->fluxfac;
    c->fnetc[0] = c->totuclfl[0] - c->totdclfl[0];
    for (lev = 1; lev <= nlayers; ++lev) {
        c->totuflux[lev] = c->totuflux[lev] * c->fluxfac;
        c->totdflux[lev] = c->totdflux[lev] * c->fluxfac;
        c->fnet[lev] = c->totuflux[lev] - c->totdflux[lev];
        c->totuclfl[lev] = c->totuclfl[lev] * c->fluxfac;
        c->totdclfl[lev] = c->totdclfl[lev] * c->fluxfac;
        c->fnetc[lev] = c->totuclfl[lev] - c->totdclfl[lev];
        l = lev - 1;
        c->htr[l] = S->lw_heatfac * (c->fnet[l] - c->fnet[lev]) / (c->pz[l] - c->pz[lev]);
        c->htrc[l] = S->lw_heatfac * (c->fnetc[l] - c->fnetc[lev]) / (c->pz[l] - c->pz[lev]);
    }
    c->htr[nlayers] = 0.0;
    c->htrc[nlayers] = 0.0;
}

/* ---------------------------------------------------------------- cldprop (rrtmg_lw_cldprop.f90:31-276)
   returns 0, or the number of the Fortran `stop`: 1 ICE RADIUS TOO SMALL (:193), 2 ICE RADIUS OUT OF BOUNDS (:198, :209),
   3 ICE GENERALIZED EFFECTIVE SIZE OUT OF BOUNDS (:225), 4 LIQUID EFFECTIVE RADIUS OUT OF BOUNDS (:253) */
static int cldprop(lwcol_t *c, int inflag, int iceflag, int liqflag, const double *tauc /* [lay][ib] 1-based */,
                   const double *ciwp, const double *clwp, const double *rei, const double *rel)
{
    const orc_state_t *S = &g_orc;
    const int nlayers = c->nlayers;
    const double cldmin = 1.e-20;
    /* icb (:147-149): spectral region of band ib for iceind / liqind = 0, 1, 2 */
    static const int icb[3][17] = {{0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1},
                                   {0, 1, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5, 5, 5, 5, 5, 5},
                                   {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16}};
    double abscoice[17], abscoliq[17], tauctot[NL];
    int ncbands = 1, iceind = 0, liqind = 0, index;
    double factor, fint, radice, radliq, cwp;
    for (int ib = 0; ib <= 16; ++ib) { abscoice[ib] = 0.; abscoliq[ib] = 0.; }
    for (int lay = 1; lay <= nlayers; ++lay) {
        tauctot[lay] = 0.;
        for (int ib = 1; ib <= 16; ++ib) {
            c->taucloud[lay][ib] = 0.0;
            tauctot[lay] = tauctot[lay] + tauc[lay * 17 + ib];
        }
    }
    for (int lay = 1; lay <= nlayers; ++lay) {
        cwp = ciwp[lay] + clwp[lay];
        if (c->cldfrac[lay] >= cldmin && (cwp >= cldmin || tauctot[lay] >= cldmin)) {
            if (inflag == 0) {
                ncbands = 16;
                for (int ib = 1; ib <= ncbands; ++ib) c->taucloud[lay][ib] = tauc[lay * 17 + ib];
            } else if (inflag == 1) {
                ncbands = 16;
                for (int ib = 1; ib <= ncbands; ++ib) c->taucloud[lay][ib] = S->abscld1 * cwp;
            } else if (inflag == 2) {
                radice = rei[lay];
                if (ciwp[lay] == 0.0) {
                    abscoice[1] = 0.0;
                    iceind = 0;
                } else if (iceflag == 0) {
                    if (radice < 10.0) return 1;
                    abscoice[1] = S->absice0[1] + S->absice0[2] / radice;
                    iceind = 0;
                } else if (iceflag == 1) {
                    if (radice < 13.0 || radice > 130.) return 2;
                    ncbands = 5;
                    for (int ib = 1; ib <= ncbands; ++ib) abscoice[ib] = S->absice1[1][ib] + S->absice1[2][ib] / radice;
                    iceind = 1;
                } else if (iceflag == 2) {
                    if (radice < 5.0 || radice > 131.0) return 2;
                    ncbands = 16;
                    factor = (radice - 2.) / 3.;
                    index = (int)factor;
                    if (index == 43) index = 42;
                    fint = factor - (double)index;
                    for (int ib = 1; ib <= ncbands; ++ib)
                        abscoice[ib] = S->absice2[index][ib] + fint * (S->absice2[index + 1][ib] - (S->absice2[index][ib]));
                    iceind = 2;
                } else if (iceflag == 3) {
                    if (radice < 5.0 || radice > 140.0) return 3;
                    ncbands = 16;
                    factor = (radice - 2.) / 3.;
                    index = (int)factor;
                    if (index == 46) index = 45;
                    fint = factor - (double)index;
                    for (int ib = 1; ib <= ncbands; ++ib)
                        abscoice[ib] = S->absice3[index][ib] + fint * (S->absice3[index + 1][ib] - (S->absice3[index][ib]));
                    iceind = 2;
                }
                if (clwp[lay] == 0.0) {
                    abscoliq[1] = 0.0;
                    liqind = 0;
                    if (iceind == 1) iceind = 2;
                } else if (liqflag == 0) {
                    abscoliq[1] = S->absliq0;
                    liqind = 0;
                    if (iceind == 1) iceind = 2;
                } else if (liqflag == 1) {
                    radliq = rel[lay];
                    if (radliq < 2.5 || radliq > 60.) return 4;
                    index = (int)(radliq - 1.5);
                    if (index == 0) index = 1;
                    if (index == 58) index = 57;
                    fint = radliq - 1.5 - (double)index;
                    ncbands = 16;
                    for (int ib = 1; ib <= ncbands; ++ib)
                        abscoliq[ib] = S->absliq1[index][ib] + fint * (S->absliq1[index + 1][ib] - (S->absliq1[index][ib]));
                    liqind = 2;
                }
                for (int ib = 1; ib <= ncbands; ++ib)
                    c->taucloud[lay][ib] = ciwp[lay] * abscoice[icb[iceind][ib]] + clwp[lay] * abscoliq[icb[liqind][ib]];
            }
        }
    }
    c->ncbands = ncbands;
    return 0;
}

/* ---------------------------------------------------------------- cloudy sky: rtrn (random overlap, icld = 1,
   rrtmg_lw_rtrn.f90:262-588) and rtrnmr (maximum/random overlap, icld = 2, 3, rrtmg_lw_rtrnmr.f90:259-779) in one body:
   the two share the layer optics (:514-567) and differ in how the cloudy and clear parts of a level are carried. */
static void rtrn_cloudy(lwcol_t *c, int maxrandom)
{
    const orc_state_t *S = &g_orc;
    const int nlayers = c->nlayers;
    const double tblint = 10000.0, bpade = S->lw_bpade;
    const double wtdiff = 0.5, rec_6 = 0.166667;
    static const double a0[16] = {1.66, 1.55, 1.58, 1.66, 1.54, 1.454, 1.89, 1.33, 1.668, 1.66, 1.66, 1.66, 1.66, 1.66, 1.66, 1.66};
    static const double a1[16] = {0.00, 0.25, 0.22, 0.00, 0.13, 0.446, -0.10, 0.40, -0.006, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00};
    static const double a2[16] = {0.00, -12.0, -11.7, 0.00, -0.72, -0.243, 0.19, -0.062, 0.414, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00};
    double secdiff[17], atrans[NL], atot[NL], bbugas[NL], bbutot[NL], urad[NL], drad[NL], clrurad[NL], clrdrad[NL];
    double d_urad_dt[NL], d_clrurad_dt[NL];
    static __thread double odcld[NL][17], abscld[NL][17], efclfrac[NL][17];
    double faccld1[NL + 2], faccld2[NL + 2], facclr1[NL + 2], facclr2[NL + 2], faccmb1[NL + 2], faccmb2[NL + 2];
    double faccld1d[NL + 2], faccld2d[NL + 2], facclr1d[NL + 2], facclr2d[NL + 2], faccmb1d[NL + 2], faccmb2d[NL + 2];
    int icldlyr[NL + 2], istcld[NL + 2], istcldd[NL + 2];
    const double *cldfrac = c->cldfrac;
    const int idrv = c->idrv;
    double rat1 = 0., rat2 = 0., fmx, fmn;
    int igc, lev, iband, ibnd, l;

    for (ibnd = 1; ibnd <= 16; ++ibnd) {
        if (ibnd == 1 || ibnd == 4 || ibnd >= 10) {
            secdiff[ibnd] = 1.66;
        } else {
            secdiff[ibnd] = a0[ibnd - 1] + a1[ibnd - 1] * exp(a2[ibnd - 1] * c->pwvcm);
            if (secdiff[ibnd] > 1.80) secdiff[ibnd] = 1.80;
            if (secdiff[ibnd] < 1.50) secdiff[ibnd] = 1.50;
        }
    }
    for (lev = 0; lev <= nlayers + 1; ++lev) {
        faccld1[lev] = faccld2[lev] = facclr1[lev] = facclr2[lev] = faccmb1[lev] = faccmb2[lev] = 0.;
        faccld1d[lev] = faccld2d[lev] = facclr1d[lev] = facclr2d[lev] = faccmb1d[lev] = faccmb2d[lev] = 0.;
        icldlyr[lev] = 0; istcld[lev] = 0; istcldd[lev] = 0;
    }
    for (lev = 0; lev <= nlayers; ++lev) {
        urad[lev] = 0.0; drad[lev] = 0.0; clrurad[lev] = 0.0; clrdrad[lev] = 0.0;
        c->totuflux[lev] = 0.0; c->totdflux[lev] = 0.0; c->totuclfl[lev] = 0.0; c->totdclfl[lev] = 0.0;
        d_urad_dt[lev] = 0.0; d_clrurad_dt[lev] = 0.0;
        c->dtotuflux_dt[lev] = 0.0; c->dtotuclfl_dt[lev] = 0.0;
    }
    /* cloud optical depth along the diffusivity angle (rtrnmr :316-324, rtrn :302-316), for the column's ncbands */
    for (int lay = 1; lay <= nlayers; ++lay) {
        icldlyr[lay] = cldfrac[lay] >= 1.e-6;          /* set inside the ib loop in the Fortran (ncbands >= 1) */
        for (int ib = 1; ib <= 16; ++ib) { odcld[lay][ib] = 0.0; abscld[lay][ib] = 0.0; efclfrac[lay][ib] = 0.0; }
        for (int ib = 1; ib <= c->ncbands; ++ib) {
            if (cldfrac[lay] >= 1.e-6) {
                odcld[lay][ib] = secdiff[ib] * c->taucloud[lay][ib];
                if (!maxrandom) {
                    double transcld = exp(-odcld[lay][ib]);
                    abscld[lay][ib] = 1. - transcld;
                    efclfrac[lay][ib] = abscld[lay][ib] * cldfrac[lay];
                }
                icldlyr[lay] = 1;
            } else {
                odcld[lay][ib] = 0.0;
                abscld[lay][ib] = 0.0;
                efclfrac[lay][ib] = 0.0;
                icldlyr[lay] = 0;
            }
        }
    }
    if (maxrandom) {
        /* maximum/random overlap factors, upward (:328-407) */
        istcld[1] = 1;
        istcldd[nlayers] = 1;
        for (lev = 1; lev <= nlayers; ++lev) {
            if (icldlyr[lev] == 1) {
                istcld[lev + 1] = 0;
                if (lev == nlayers) {
                    faccld1[lev + 1] = 0.; faccld2[lev + 1] = 0.; facclr1[lev + 1] = 0.;
                    facclr2[lev + 1] = 0.; faccmb1[lev + 1] = 0.; faccmb2[lev + 1] = 0.;
                } else if (cldfrac[lev + 1] >= cldfrac[lev]) {
                    faccld1[lev + 1] = 0.;
                    faccld2[lev + 1] = 0.;
                    if (istcld[lev] == 1) {
                        facclr1[lev + 1] = 0.;
                        facclr2[lev + 1] = 0.;
                        if (cldfrac[lev] < 1.) facclr2[lev + 1] = (cldfrac[lev + 1] - cldfrac[lev]) / (1. - cldfrac[lev]);
                        facclr2[lev] = 0.;
                        faccld2[lev] = 0.;
                    } else {
                        fmx = fmax(cldfrac[lev], cldfrac[lev - 1]);
                        if (cldfrac[lev + 1] > fmx) {
                            facclr1[lev + 1] = rat2;
                            facclr2[lev + 1] = (cldfrac[lev + 1] - fmx) / (1. - fmx);
                        } else if (cldfrac[lev + 1] < fmx) {
                            facclr1[lev + 1] = (cldfrac[lev + 1] - cldfrac[lev]) / (cldfrac[lev - 1] - cldfrac[lev]);
                            facclr2[lev + 1] = 0.;
                        } else {
                            facclr1[lev + 1] = rat2;
                            facclr2[lev + 1] = 0.;
                        }
                    }
                    if (facclr1[lev + 1] > 0. || facclr2[lev + 1] > 0.) { rat1 = 1.; rat2 = 0.; }
                    else { rat1 = 0.; rat2 = 0.; }
                } else {
                    facclr1[lev + 1] = 0.;
                    facclr2[lev + 1] = 0.;
                    if (istcld[lev] == 1) {
                        faccld1[lev + 1] = 0.;
                        faccld2[lev + 1] = (cldfrac[lev] - cldfrac[lev + 1]) / cldfrac[lev];
                        facclr2[lev] = 0.;
                        faccld2[lev] = 0.;
                    } else {
                        fmn = fmin(cldfrac[lev], cldfrac[lev - 1]);
                        if (cldfrac[lev + 1] <= fmn) {
                            faccld1[lev + 1] = rat1;
                            faccld2[lev + 1] = (fmn - cldfrac[lev + 1]) / fmn;
                        } else {
                            faccld1[lev + 1] = (cldfrac[lev] - cldfrac[lev + 1]) / (cldfrac[lev] - fmn);
                            faccld2[lev + 1] = 0.;
                        }
                    }
                    if (faccld1[lev + 1] > 0. || faccld2[lev + 1] > 0.) { rat1 = 0.; rat2 = 1.; }
                    else { rat1 = 0.; rat2 = 0.; }
                }
                faccmb1[lev + 1] = facclr1[lev + 1] * faccld2[lev] * cldfrac[lev - 1];
                faccmb2[lev + 1] = faccld1[lev + 1] * facclr2[lev] * (1. - cldfrac[lev - 1]);
            } else {
                istcld[lev + 1] = 1;
            }
        }
        /* downward (:409-479); rat1, rat2 carry over from the upward pass as in the Fortran */
        for (lev = nlayers; lev >= 1; --lev) {
            if (icldlyr[lev] == 1) {
                istcldd[lev - 1] = 0;
                if (lev == 1) {
                    faccld1d[lev - 1] = 0.; faccld2d[lev - 1] = 0.; facclr1d[lev - 1] = 0.;
                    facclr2d[lev - 1] = 0.; faccmb1d[lev - 1] = 0.; faccmb2d[lev - 1] = 0.;
                } else if (cldfrac[lev - 1] >= cldfrac[lev]) {
                    faccld1d[lev - 1] = 0.;
                    faccld2d[lev - 1] = 0.;
                    if (istcldd[lev] == 1) {
                        facclr1d[lev - 1] = 0.;
                        facclr2d[lev - 1] = 0.;
                        if (cldfrac[lev] < 1.) facclr2d[lev - 1] = (cldfrac[lev - 1] - cldfrac[lev]) / (1. - cldfrac[lev]);
                        facclr2d[lev] = 0.;
                        faccld2d[lev] = 0.;
                    } else {
                        fmx = fmax(cldfrac[lev], cldfrac[lev + 1]);
                        if (cldfrac[lev - 1] > fmx) {
                            facclr1d[lev - 1] = rat2;
                            facclr2d[lev - 1] = (cldfrac[lev - 1] - fmx) / (1. - fmx);
                        } else if (cldfrac[lev - 1] < fmx) {
                            facclr1d[lev - 1] = (cldfrac[lev - 1] - cldfrac[lev]) / (cldfrac[lev + 1] - cldfrac[lev]);
                            facclr2d[lev - 1] = 0.;
                        } else {
                            facclr1d[lev - 1] = rat2;
                            facclr2d[lev - 1] = 0.;
                        }
                    }
                    if (facclr1d[lev - 1] > 0. || facclr2d[lev - 1] > 0.) { rat1 = 1.; rat2 = 0.; }
                    else { rat1 = 0.; rat2 = 0.; }
                } else {
                    facclr1d[lev - 1] = 0.;
                    facclr2d[lev - 1] = 0.;
                    if (istcldd[lev] == 1) {
                        faccld1d[lev - 1] = 0.;
                        faccld2d[lev - 1] = (cldfrac[lev] - cldfrac[lev - 1]) / cldfrac[lev];
                        facclr2d[lev] = 0.;
                        faccld2d[lev] = 0.;
                    } else {
                        fmn = fmin(cldfrac[lev], cldfrac[lev + 1]);
                        if (cldfrac[lev - 1] <= fmn) {
                            faccld1d[lev - 1] = rat1;
                            faccld2d[lev - 1] = (fmn - cldfrac[lev - 1]) / fmn;
                        } else {
                            faccld1d[lev - 1] = (cldfrac[lev] - cldfrac[lev - 1]) / (cldfrac[lev] - fmn);
                            faccld2d[lev - 1] = 0.;
                        }
                    }
                    if (faccld1d[lev - 1] > 0. || faccld2d[lev - 1] > 0.) { rat1 = 0.; rat2 = 1.; }
                    else { rat1 = 0.; rat2 = 0.; }
                }
                faccmb1d[lev - 1] = facclr1d[lev - 1] * faccld2d[lev] * cldfrac[lev + 1];
                faccmb2d[lev - 1] = faccld1d[lev - 1] * facclr2d[lev] * (1. - cldfrac[lev + 1]);
            } else {
                istcldd[lev - 1] = 1;
            }
        }
    }

    igc = 1;
    for (iband = 1; iband <= 16; ++iband) {
        /* ipat (rtrnmr.f90:243-245): cloud band of spectral band iband for ncbands = 1, 5, 16 */
        static const int ipat5[17] = {0, 1, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5, 5, 5, 5, 5, 5};
        const int ib = c->ncbands == 16 ? iband : (c->ncbands == 5 ? ipat5[iband] : 1);
        do {
            double radld = 0., radclrd = 0., radlu, radclru, rad0, reflect;
            double cldradd = 0., clrradd = 0., cldradu = 0., clrradu = 0., oldcld, oldclr, rad = 0., radmod;
            int iclddn = 0;
            for (lev = nlayers; lev >= 1; --lev) {
                double plfrac = c->fracs[igc][lev];
                double blay = c->planklay[lev][iband];
                double dplankup = c->planklev[lev][iband] - blay;
                double dplankdn = c->planklev[lev - 1][iband] - blay;
                double odepth = secdiff[iband] * c->taut[igc][lev];
                double bbd, gassrc = 0., bbdtot = 0.;
                if (odepth < 0.0) odepth = 0.0;
                if (icldlyr[lev] == 1) {
                    iclddn = 1;
                    double odtot = odepth + odcld[lev][ib];
                    if (odtot < 0.06) {
                        atrans[lev] = odepth - 0.5 * odepth * odepth;
                        double odepth_rec = rec_6 * odepth;
                        gassrc = plfrac * (blay + dplankdn * odepth_rec) * atrans[lev];
                        atot[lev] = odtot - 0.5 * odtot * odtot;
                        double odtot_rec = rec_6 * odtot;
                        bbdtot = plfrac * (blay + dplankdn * odtot_rec);
                        bbd = plfrac * (blay + dplankdn * odepth_rec);
                        bbugas[lev] = plfrac * (blay + dplankup * odepth_rec);
                        bbutot[lev] = plfrac * (blay + dplankup * odtot_rec);
                    } else if (odepth <= 0.06) {
                        atrans[lev] = odepth - 0.5 * odepth * odepth;
                        double odepth_rec = rec_6 * odepth;
                        gassrc = plfrac * (blay + dplankdn * odepth_rec) * atrans[lev];
                        odtot = odepth + odcld[lev][ib];
                        double tblind = odtot / (bpade + odtot);
                        int ittot = (int)(tblint * tblind + 0.5);
                        double tfactot = S->tfn_tbl[ittot];
                        bbdtot = plfrac * (blay + tfactot * dplankdn);
                        bbd = plfrac * (blay + dplankdn * odepth_rec);
                        atot[lev] = 1. - S->exp_tbl[ittot];
                        bbugas[lev] = plfrac * (blay + dplankup * odepth_rec);
                        bbutot[lev] = plfrac * (blay + tfactot * dplankup);
                    } else {
                        double tblind = odepth / (bpade + odepth);
                        int itgas = (int)(tblint * tblind + 0.5);
                        odepth = S->tau_tbl[itgas];
                        atrans[lev] = 1. - S->exp_tbl[itgas];
                        double tfacgas = S->tfn_tbl[itgas];
                        gassrc = atrans[lev] * plfrac * (blay + tfacgas * dplankdn);
                        odtot = odepth + odcld[lev][ib];
                        tblind = odtot / (bpade + odtot);
                        int ittot = (int)(tblint * tblind + 0.5);
                        double tfactot = S->tfn_tbl[ittot];
                        bbdtot = plfrac * (blay + tfactot * dplankdn);
                        bbd = plfrac * (blay + tfacgas * dplankdn);
                        atot[lev] = 1. - S->exp_tbl[ittot];
                        bbugas[lev] = plfrac * (blay + tfacgas * dplankup);
                        bbutot[lev] = plfrac * (blay + tfactot * dplankup);
                    }
                    if (maxrandom) { /* rtrnmr :569-588 */
                        if (istcldd[lev] == 1) {
                            cldradd = cldfrac[lev] * radld;
                            clrradd = radld - cldradd;
                            oldcld = cldradd;
                            oldclr = clrradd;
                            rad = 0.;
                        }
                        double ttot = 1. - atot[lev];
                        double cldsrc = bbdtot * atot[lev];
                        cldradd = cldradd * ttot + cldfrac[lev] * cldsrc;
                        clrradd = clrradd * (1. - atrans[lev]) + (1. - cldfrac[lev]) * gassrc;
                        radld = cldradd + clrradd;
                        drad[lev - 1] = drad[lev - 1] + radld;
                        radmod = rad * (facclr1d[lev - 1] * (1. - atrans[lev]) + faccld1d[lev - 1] * ttot) -
                                 faccmb1d[lev - 1] * gassrc + faccmb2d[lev - 1] * cldsrc;
                        oldcld = cldradd - radmod;
                        oldclr = clrradd + radmod;
                        rad = -radmod + facclr2d[lev - 1] * oldclr - faccld2d[lev - 1] * oldcld;
                        cldradd = cldradd + rad;
                        clrradd = clrradd - rad;
                    } else { /* rtrn :371-375 (the same statement in its three branches) */
                        radld = radld - radld * (atrans[lev] + efclfrac[lev][ib] * (1. - atrans[lev])) + gassrc +
                                cldfrac[lev] * (bbdtot * atot[lev] - gassrc);
                        drad[lev - 1] = drad[lev - 1] + radld;
                    }
                } else {
                    if (odepth <= 0.06) {
                        atrans[lev] = odepth - 0.5 * odepth * odepth;
                        odepth = rec_6 * odepth;
                        bbd = plfrac * (blay + dplankdn * odepth);
                        bbugas[lev] = plfrac * (blay + dplankup * odepth);
                    } else {
                        double tblind = odepth / (bpade + odepth);
                        int itr = (int)(tblint * tblind + 0.5);
                        double transc = S->exp_tbl[itr];
                        atrans[lev] = 1. - transc;
                        double tausfac = S->tfn_tbl[itr];
                        bbd = plfrac * (blay + tausfac * dplankdn);
                        bbugas[lev] = plfrac * (blay + tausfac * dplankup);
                    }
                    radld = radld + (bbd - radld) * atrans[lev];
                    drad[lev - 1] = drad[lev - 1] + radld;
                }
                if (iclddn == 1) {
                    radclrd = radclrd + (bbd - radclrd) * atrans[lev];
                    clrdrad[lev - 1] = clrdrad[lev - 1] + radclrd;
                } else {
                    radclrd = radld;
                    clrdrad[lev - 1] = drad[lev - 1];
                }
            }
            rad0 = c->fracs[igc][1] * c->plankbnd[iband];
            reflect = 1. - c->semiss[iband];
            radlu = rad0 + reflect * radld;
            radclru = rad0 + reflect * radclrd;
            double d_rad0_dt = 0., d_radlu_dt = 0., d_radclru_dt = 0.;
            if (idrv == 1) d_rad0_dt = c->fracs[igc][1] * c->dplankbnd_dt[iband];
            urad[0] = urad[0] + radlu;
            clrurad[0] = clrurad[0] + radclru;
            if (idrv == 1) {
                d_radlu_dt = d_rad0_dt;
                d_urad_dt[0] = d_urad_dt[0] + d_radlu_dt;
                d_radclru_dt = d_rad0_dt;
                d_clrurad_dt[0] = d_clrurad_dt[0] + d_radclru_dt;
            }
            for (lev = 1; lev <= nlayers; ++lev) {
                if (icldlyr[lev] == 1) {
                    double gassrc = bbugas[lev] * atrans[lev];
                    if (maxrandom) { /* rtrnmr :653-674 */
                        if (istcld[lev] == 1) {
                            cldradu = cldfrac[lev] * radlu;
                            clrradu = radlu - cldradu;
                            oldcld = cldradu;
                            oldclr = clrradu;
                            rad = 0.;
                        }
                        double ttot = 1. - atot[lev];
                        double cldsrc = bbutot[lev] * atot[lev];
                        cldradu = cldradu * ttot + cldfrac[lev] * cldsrc;
                        clrradu = clrradu * (1.0 - atrans[lev]) + (1. - cldfrac[lev]) * gassrc;
                        radlu = cldradu + clrradu;
                        urad[lev] = urad[lev] + radlu;
                        radmod = rad * (facclr1[lev + 1] * (1.0 - atrans[lev]) + faccld1[lev + 1] * ttot) -
                                 faccmb1[lev + 1] * gassrc + faccmb2[lev + 1] * cldsrc;
                        oldcld = cldradu - radmod;
                        oldclr = clrradu + radmod;
                        rad = -radmod + facclr2[lev + 1] * oldclr - faccld2[lev + 1] * oldcld;
                        cldradu = cldradu + rad;
                        clrradu = clrradu - rad;
                    } else { /* rtrn :480-485 */
                        radlu = radlu - radlu * (atrans[lev] + efclfrac[lev][ib] * (1. - atrans[lev])) + gassrc +
                                cldfrac[lev] * (bbutot[lev] * atot[lev] - gassrc);
                        urad[lev] = urad[lev] + radlu;
                    }
                    if (idrv == 1) {
                        d_radlu_dt = d_radlu_dt * cldfrac[lev] * (1.0 - atot[lev]) +
                                     d_radlu_dt * (1.0 - cldfrac[lev]) * (1.0 - atrans[lev]);
                        d_urad_dt[lev] = d_urad_dt[lev] + d_radlu_dt;
                    }
                } else {
                    radlu = radlu + (bbugas[lev] - radlu) * atrans[lev];
                    urad[lev] = urad[lev] + radlu;
                    if (idrv == 1) {
                        d_radlu_dt = d_radlu_dt * (1.0 - atrans[lev]);
                        d_urad_dt[lev] = d_urad_dt[lev] + d_radlu_dt;
                    }
                }
                if (iclddn == 1) {
                    radclru = radclru + (bbugas[lev] - radclru) * atrans[lev];
                    clrurad[lev] = clrurad[lev] + radclru;
                } else {
                    radclru = radlu;
                    clrurad[lev] = urad[lev];
                }
                if (idrv == 1) {
                    if (iclddn == 1) {
                        d_radclru_dt = d_radclru_dt * (1.0 - atrans[lev]);
                        d_clrurad_dt[lev] = d_clrurad_dt[lev] + d_radclru_dt;
                    } else {
                        d_radclru_dt = d_radlu_dt;
                        d_clrurad_dt[lev] = d_urad_dt[lev];
                    }
                }
            }
            (void)oldcld; (void)oldclr;
            igc = igc + 1;
        } while (igc <= ngs[iband - 1]);

        for (lev = nlayers; lev >= 0; --lev) {
            double uflux = urad[lev] * wtdiff;
            double dflux = drad[lev] * wtdiff;
            urad[lev] = 0.0;
            drad[lev] = 0.0;
            c->totuflux[lev] = c->totuflux[lev] + uflux * delwave[iband - 1];
            c->totdflux[lev] = c->totdflux[lev] + dflux * delwave[iband - 1];
            double uclfl = clrurad[lev] * wtdiff;
            double dclfl = clrdrad[lev] * wtdiff;
            clrurad[lev] = 0.0;
            clrdrad[lev] = 0.0;
            c->totuclfl[lev] = c->totuclfl[lev] + uclfl * delwave[iband - 1];
            c->totdclfl[lev] = c->totdclfl[lev] + dclfl * delwave[iband - 1];
        }
        if (idrv == 1) {
            for (lev = nlayers; lev >= 0; --lev) {
                double duflux_dt = d_urad_dt[lev] * wtdiff;
                d_urad_dt[lev] = 0.0;
                c->dtotuflux_dt[lev] = c->dtotuflux_dt[lev] + duflux_dt * delwave[iband - 1] * c->fluxfac;
                double duclfl_dt = d_clrurad_dt[lev] * wtdiff;
                d_clrurad_dt[lev] = 0.0;
                c->dtotuclfl_dt[lev] = c->dtotuclfl_dt[lev] + duclfl_dt * delwave[iband - 1] * c->fluxfac;
            }
        }
    }
    c->totuflux[0] = c->totuflux[0] * c->fluxfac;
    c->totdflux[0] = c->totdflux[0] * c->fluxfac;
    c->fnet[0] = c->totuflux[0] - c->totdflux[0];
    c->totuclfl[0] = c->totuclfl[0] * c->fluxfac;
    c->totdclfl[0] = c->totdclfl[0] * c->fluxfac;
    c->fnetc[0] = c->totuclfl[0] - c->totdclfl[0];
    for (lev = 1; lev <= nlayers; ++lev) {
        c->totuflux[lev] = c->totuflux[lev] * c->fluxfac;
        c->totdflux[lev] = c->totdflux[lev] * c->fluxfac;
        c->fnet[lev] = c->totuflux[lev] - c->totdflux[lev];
        c->totuclfl[lev] = c->totuclfl[lev] * c->fluxfac;
        c->totdclfl[lev] = c->totdclfl[lev] * c->fluxfac;
        c->fnetc[lev] = c->totuclfl[lev] - c->totdclfl[lev];
        l = lev - 1;
        c->htr[l] = S->lw_heatfac * (c->fnet[l] - c->fnet[lev]) / (c->pz[l] - c->pz[lev]);
        c->htrc[l] = S->lw_heatfac * (c->fnetc[l] - c->fnetc[lev]) / (c->pz[l] - c->pz[lev]);
    }
    c->htr[nlayers] = 0.0;
    c->htrc[nlayers] = 0.0;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---------------------------------------------------------------- rrtmg_lw (rad.nomcica:80-569) */
int orc_rrtmg_lw(int ncol, int nlay, int icld, int idrv,
                 const double *play, const double *plev, const double *tlay, const double *tlev,
                 const double *tsfc, const double *h2ovmr, const double *o3vmr, const double *co2vmr,
                 const double *ch4vmr, const double *n2ovmr, const double *o2vmr,
                 const double *cfc11vmr, const double *cfc12vmr, const double *cfc22vmr,
                 const double *ccl4vmr, const double *emis, const double *tauaer,
                 int inflglw, const double *cldfr, const double *taucld,
                 int iceflglw, int liqflglw, const double *cicewp, const double *cliqwp, const double *reice, const double *reliq,
                 double *uflx, double *dflx, double *hr, double *uflxc, double *dflxc, double *hrc,
                 double *duflx_dt, double *duflxc_dt,
                 const orc_lw_stages_t *st, int nthreads)
{
    if (!g_orc.ready) return 1;
    if (icld < 0 || icld > 3) icld = 2; /* :437 */
    if (idrv < 0 || idrv > 1) return 2;
    if (icld >= 1 && (inflglw < 0 || inflglw > 2)) return 2;
    if (icld >= 1 && (!cldfr || !taucld)) return 3;
    if (icld >= 1 && inflglw >= 1 && (!cicewp || !cliqwp)) return 3;
    if (icld >= 1 && inflglw == 2 && (!reice || !reliq)) return 3;
    int cld_stop = 0; /* number of the cldprop `stop` hit by some column (10 + n is returned) */
    if (idrv == 1 && (!duflx_dt || !duflxc_dt)) return 3;
    if (nlay < 1 || nlay > ORC_MAXLAY) return 3;
    if (nthreads < 1) nthreads = 1;
    const int iaer = 10; /* forced (:442) */
    const int istart = 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
    {
        lwcol_t *c = (lwcol_t *)malloc(sizeof(lwcol_t));
        /* :419-421 */
        c->oneminus = 1. - 1.e-6;
        {
            double pi = 2. * asin(1.);
            c->fluxfac = pi * 2.e4;
        }
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int iplon = 1; iplon <= ncol; ++iplon) {
            c->idrv = idrv;
            inatm(c, iplon, ncol, nlay, iaer, play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr,
                  n2ovmr, o2vmr, cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr, emis, tauaer);
            /* cldprop with cldfrac=0: ncbands=1, taucloud=0 -- nothing to compute */
            if (icld >= 1) {
                /* inatm :868-887, then cldprop */
                static __thread double tauc[NL * 17], ciwp[NL], clwp[NL], rei[NL], rel[NL];
                c->cldfrac[0] = 0.; c->cldfrac[nlay + 1] = 0.;
                for (int l = 1; l <= nlay; ++l) {
                    const long o = (long)(l - 1) * ncol + iplon - 1;
                    c->cldfrac[l] = cldfr[o];
                    ciwp[l] = cicewp ? cicewp[o] : 0.; clwp[l] = cliqwp ? cliqwp[o] : 0.;
                    rei[l] = reice ? reice[o] : 0.; rel[l] = reliq ? reliq[o] : 0.;
                    for (int ib = 1; ib <= 16; ++ib) tauc[l * 17 + ib] = taucld[(ib - 1) + 16 * ((iplon - 1) + (long)ncol * (l - 1))];
                }
                const int stop = cldprop(c, inflglw, iceflglw, liqflglw, tauc, ciwp, clwp, rei, rel);
                if (stop) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
                    cld_stop = stop;
                    continue;
                }
            }
            setcoef(c, istart);
            taumol(c);
            /* :514-519, iaer=10 */
            for (int k = 1; k <= nlay; ++k)
                for (int ig = 1; ig <= ORC_NGPTLW; ++ig) {
                    int b = 0;
                    while (ig > ngs[b]) ++b; /* ngb(ig) */
                    c->taut[ig][k] = c->taug[ig][k] + c->taua[k][b + 1];
                }
            if (icld == 0) rtrnmr_clear(c);
            else rtrn_cloudy(c, icld != 1); /* rad.nomcica:527-541 */
            const long i0 = iplon - 1;
            for (int k = 0; k <= nlay; ++k) {
                uflx[(long)k * ncol + i0] = c->totuflux[k];
                dflx[(long)k * ncol + i0] = c->totdflux[k];
                uflxc[(long)k * ncol + i0] = c->totuclfl[k];
                dflxc[(long)k * ncol + i0] = c->totdclfl[k];
            }
            for (int k = 0; k <= nlay - 1; ++k) {
                hr[(long)k * ncol + i0] = c->htr[k];
                hrc[(long)k * ncol + i0] = c->htrc[k];
            }
            if (idrv == 1) /* rad.nomcica:559-564 */
                for (int k = 0; k <= nlay; ++k) {
                    duflx_dt[(long)k * ncol + i0] = c->dtotuflux_dt[k];
                    duflxc_dt[(long)k * ncol + i0] = c->dtotuclfl_dt[k];
                }
            if (st) {
#define PUT(dst, src) if (st->dst) for (int l = 1; l <= nlay; ++l) st->dst[(long)(l - 1) * ncol + i0] = c->src[l]
                if (st->laytrop) st->laytrop[i0] = c->laytrop;
                if (st->pwvcm) st->pwvcm[i0] = c->pwvcm;
                PUT(jp, jp); PUT(jt, jt); PUT(jt1, jt1); PUT(indself, indself); PUT(indfor, indfor); PUT(indminor, indminor);
                PUT(fac00, fac00); PUT(fac01, fac01); PUT(fac10, fac10); PUT(fac11, fac11);
                PUT(colh2o, colh2o); PUT(colco2, colco2); PUT(colo3, colo3); PUT(coln2o, coln2o);
                PUT(colco, colco); PUT(colch4, colch4); PUT(colo2, colo2); PUT(colbrd, colbrd);
                PUT(selffac, selffac); PUT(selffrac, selffrac); PUT(forfac, forfac); PUT(forfrac, forfrac);
                PUT(minorfrac, minorfrac); PUT(scaleminor, scaleminor); PUT(scaleminorn2, scaleminorn2);
                PUT(coldry, coldry);
#undef PUT
                for (int ib = 1; ib <= 16; ++ib) {
                    if (st->plankbnd) st->plankbnd[(long)(ib - 1) * ncol + i0] = c->plankbnd[ib];
                    if (st->planklay)
                        for (int l = 1; l <= nlay; ++l)
                            st->planklay[((long)(ib - 1) * nlay + (l - 1)) * ncol + i0] = c->planklay[l][ib];
                    if (st->planklev)
                        for (int l = 0; l <= nlay; ++l)
                            st->planklev[((long)(ib - 1) * (nlay + 1) + l) * ncol + i0] = c->planklev[l][ib];
                }
                for (int ig = 1; ig <= ORC_NGPTLW; ++ig)
                    for (int l = 1; l <= nlay; ++l) {
                        if (st->taug) st->taug[((long)(ig - 1) * nlay + (l - 1)) * ncol + i0] = c->taug[ig][l];
                        if (st->fracs) st->fracs[((long)(ig - 1) * nlay + (l - 1)) * ncol + i0] = c->fracs[ig][l];
                    }
            }
        }
        free(c);
    }
    return cld_stop ? 10 + cld_stop : 0;
}
