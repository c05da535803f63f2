/*
 * tables.c -- oracle initialisation: blob reader, lookup tables, 16 -> ngc g-point reduction.
 * TEST INFRASTRUCTURE ONLY (see rrtmg_oracle.h).
 *
 * Follows  LW/src/rrtmg_lw_init.f90:28-175 (rrtmg_lw_ini), :178-281 (lwdatinit), :284-363 (lwcmbdat),
 *          :366-2659 (cmbgb1..16);
 *          SW/src/rrtmg_sw_init.f90:28-154 (rrtmg_sw_ini), :157-241 (swdatinit), :244-367 (swcmbdat),
 *          :473-1516 (cmbgb16s..29).
 */
#include "rrtmg_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

orc_state_t g_orc;

/* ------------------------------------------------------------------ blob reader */
int orc_blob_load(const char *path, orc_blob_t *out)
{
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    unsigned char *raw = (unsigned char *)malloc((size_t)sz);
    if (!raw || fread(raw, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(raw); return -2; }
    fclose(f);
    if (memcmp(raw, "RRTMGTB1", 8) != 0) { free(raw); return -3; }
    uint32_t n;
    memcpy(&n, raw + 8, 4);
    const size_t recsz = 32 + 4 + 16 + 8;
    const unsigned char *p = raw + 12;
    const double *data = (const double *)(raw + 12 + (size_t)n * recsz);
    out->n = (int)n;
    out->arr = (orc_array_t *)calloc(n, sizeof(orc_array_t));
    out->raw = raw;
    for (uint32_t i = 0; i < n; ++i, p += recsz) {
        orc_array_t *a = &out->arr[i];
        memcpy(a->name, p, 32);
        a->name[31] = 0;
        uint32_t nd, d[4];
        uint64_t off;
        memcpy(&nd, p + 32, 4);
        memcpy(d, p + 36, 16);
        memcpy(&off, p + 52, 8);
        a->ndim = (int)nd;
        for (int k = 0; k < 4; ++k) a->dims[k] = (int)d[k];
        a->data = data + off;
    }
    return 0;
}

void orc_blob_free(orc_blob_t *b)
{
    free(b->arr);
    free(b->raw);
    b->arr = NULL;
    b->raw = NULL;
    b->n = 0;
}

const orc_array_t *orc_blob_find(const orc_blob_t *b, const char *name)
{
    for (int i = 0; i < b->n; ++i)
        if (strcmp(b->arr[i].name, name) == 0) return &b->arr[i];
    return NULL;
}

/* ------------------------------------------------------------------ g-point maps */
/* LW/src/rrtmg_lw_init.f90:306-361 */
static const int lw_ngc[16] = {10, 12, 16, 14, 16, 8, 12, 8, 12, 6, 8, 8, 4, 2, 2, 2};
static const int lw_ngs[16] = {10, 22, 38, 52, 68, 76, 88, 96, 108, 114, 122, 130, 134, 136, 138, 140};
static const int lw_ngm[256] = {
    1, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 10,
    1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 10, 10, 11, 11, 12, 12,
    1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16,
    1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 14, 14,
    1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16,
    1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8,
    1, 1, 2, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 11, 12, 12,
    1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8,
    1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 10, 10, 11, 11, 12, 12,
    1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6,
    1, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 7, 8, 8, 8,
    1, 2, 3, 4, 5, 5, 6, 6, 7, 7, 7, 7, 8, 8, 8, 8,
    1, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4,
    1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2,
    1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2,
    1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2};
static const int lw_ngn[140] = {
    1, 1, 2, 2, 2, 2, 2, 2, 1, 1,
    1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2,
    1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
    1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 3,
    1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
    2, 2, 2, 2, 2, 2, 2, 2,
    2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2,
    2, 2, 2, 2, 2, 2, 2, 2,
    1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2,
    2, 2, 2, 2, 4, 4,
    1, 1, 2, 2, 2, 2, 3, 3,
    1, 1, 1, 1, 2, 2, 4, 4,
    3, 3, 4, 6,
    8, 8,
    8, 8,
    4, 12};
/* SW/src/rrtmg_sw_init.f90:267-365 */
static const int sw_ngc[14] = {6, 12, 8, 8, 10, 10, 2, 10, 8, 6, 6, 8, 6, 12};
static const int sw_ngs[14] = {6, 18, 26, 34, 44, 54, 56, 66, 74, 80, 86, 94, 100, 112};
static const int sw_ngm[224] = {
    1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6,
    1, 2, 3, 4, 5, 6, 6, 7, 8, 8, 9, 10, 10, 11, 12, 12,
    1, 2, 3, 4, 5, 5, 6, 6, 7, 7, 7, 7, 8, 8, 8, 8,
    1, 2, 3, 4, 5, 5, 6, 6, 7, 7, 7, 7, 8, 8, 8, 8,
    1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 10, 10, 10, 10, 10, 10,
    1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 10, 10, 10, 10, 10, 10,
    1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2,
    1, 1, 2, 2, 3, 4, 5, 6, 7, 8, 9, 9, 10, 10, 10, 10,
    1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8,
    1, 2, 3, 3, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6,
    1, 2, 3, 3, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6,
    1, 2, 3, 4, 5, 6, 7, 7, 7, 7, 8, 8, 8, 8, 8, 8,
    1, 2, 3, 3, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6,
    1, 2, 3, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 10, 11, 12};
static const int sw_ngn[112] = {
    2, 2, 2, 2, 4, 4,
    1, 1, 1, 1, 1, 2, 1, 2, 1, 2, 1, 2,
    1, 1, 1, 1, 2, 2, 4, 4,
    1, 1, 1, 1, 2, 2, 4, 4,
    1, 1, 1, 1, 1, 1, 1, 1, 2, 6,
    1, 1, 1, 1, 1, 1, 1, 1, 2, 6,
    8, 8,
    2, 2, 1, 1, 1, 1, 1, 1, 2, 4,
    2, 2, 2, 2, 2, 2, 2, 2,
    1, 1, 2, 2, 4, 6,
    1, 1, 2, 2, 4, 6,
    1, 1, 1, 1, 1, 1, 4, 6,
    1, 1, 2, 2, 4, 6,
    1, 1, 1, 1, 2, 2, 2, 2, 1, 1, 1, 1};
/* Gaussian weights of the 16 original g-points (LW init :356-361, SW init :361-366) */
static const double wt16[16] = {
    0.1527534276, 0.1491729617, 0.1420961469, 0.1316886544, 0.1181945205, 0.1019300893,
    0.0832767040, 0.0626720116, 0.0424925000, 0.0046269894, 0.0038279891, 0.0030260086,
    0.0022199750, 0.0014140010, 0.0005330000, 0.0000750000};

/* rwgt: LW/src/rrtmg_lw_init.f90:130-154, SW/src/rrtmg_sw_init.f90:112-136 */
static void compute_rwgt(int nbnd, const int *ngc, const int *ngn, const int *ngm, double *rwgt)
{
    const int mg = 16;
    int igcsm = 0;
    double wtsm[16];
    for (int ibnd = 1; ibnd <= nbnd; ++ibnd) {
        int iprsm = 0;
        if (ngc[ibnd - 1] < mg) {
            for (int igc = 1; igc <= ngc[ibnd - 1]; ++igc) {
                igcsm = igcsm + 1;
                double wtsum = 0.0;
                for (int ipr = 1; ipr <= ngn[igcsm - 1]; ++ipr) {
                    iprsm = iprsm + 1;
                    wtsum = wtsum + wt16[iprsm - 1];
                }
                wtsm[igc - 1] = wtsum;
            }
            for (int ig = 1; ig <= 16; ++ig) {
                int ind = (ibnd - 1) * mg + ig;
                rwgt[ind - 1] = wt16[ig - 1] / wtsm[ngm[ind - 1] - 1];
            }
        } else {
            for (int ig = 1; ig <= 16; ++ig) {
                igcsm = igcsm + 1;
                int ind = (ibnd - 1) * mg + ig;
                rwgt[ind - 1] = 1.0;
            }
        }
    }
}

/* One cmbgbNN inner pattern: reduce the 16-wide g axis of `src` into ngc groups.
 *   glast=1 : g is the last (slowest) dimension, `inner` contiguous elements per g  (k tables)
 *   glast=0 : g is the first (fastest) dimension, `outer` slabs of 16             (fracref, sfluxref, rayla)
 *   weighted: multiply by rwgt (k-like) or plain sum (Planck fractions, solar source). */
static double *reduce_g(const double *src, long inner, long outer, int glast, int weighted,
                        int ngc, const int *ngn_band, const double *rwgt_band)
{
    double *dst;
    if (glast) {
        dst = (double *)malloc(sizeof(double) * (size_t)(inner * ngc));
        for (long e = 0; e < inner; ++e) {
            int iprsm = 0;
            for (int igc = 1; igc <= ngc; ++igc) {
                double sumk = 0.0;
                for (int ipr = 1; ipr <= ngn_band[igc - 1]; ++ipr) {
                    iprsm = iprsm + 1;
                    if (weighted) sumk = sumk + src[(long)(iprsm - 1) * inner + e] * rwgt_band[iprsm - 1];
                    else sumk = sumk + src[(long)(iprsm - 1) * inner + e];
                }
                dst[(long)(igc - 1) * inner + e] = sumk;
            }
        }
    } else {
        dst = (double *)malloc(sizeof(double) * (size_t)(outer * ngc));
        for (long o = 0; o < outer; ++o) {
            int iprsm = 0;
            for (int igc = 1; igc <= ngc; ++igc) {
                double sumf = 0.0;
                for (int ipr = 1; ipr <= ngn_band[igc - 1]; ++ipr) {
                    iprsm = iprsm + 1;
                    if (weighted) sumf = sumf + src[o * 16 + (iprsm - 1)] * rwgt_band[iprsm - 1];
                    else sumf = sumf + src[o * 16 + (iprsm - 1)];
                }
                dst[o * ngc + (igc - 1)] = sumf;
            }
        }
    }
    return dst;
}

/* registry of reduced tables, for export and for freeing */
typedef struct { char name[32]; long n; double *data; } reg_t;
static reg_t g_reg[512];
static int g_nreg = 0;

static const double *reduce_named(const orc_blob_t *blob, const char *prefix, const char *oname,
                                  const char *rname, int ngc, const int *ngn_band, const double *rwgt_band)
{
    char full[64];
    snprintf(full, sizeof full, "%s.%s", prefix, oname);
    const orc_array_t *a = orc_blob_find(blob, full);
    if (!a) return NULL;
    long total = 1;
    for (int k = 0; k < a->ndim; ++k) total *= a->dims[k];
    int plain = (strcmp(oname, "fracrefao") == 0 || strcmp(oname, "fracrefbo") == 0 ||
                 strcmp(oname, "sfluxrefo") == 0);
    int gfirst = plain || strcmp(oname, "raylao") == 0 || a->ndim == 1;
    double *dst;
    if (gfirst) dst = reduce_g(a->data, 0, total / 16, 0, !plain, ngc, ngn_band, rwgt_band);
    else dst = reduce_g(a->data, total / 16, 0, 1, 1, ngc, ngn_band, rwgt_band);
    reg_t *r = &g_reg[g_nreg++];
    snprintf(r->name, sizeof r->name, "%s.%s", prefix, rname);
    r->n = total / 16 * ngc;
    r->data = dst;
    return dst;
}

long orc_get_table(const char *name, const double **data)
{
    for (int i = 0; i < g_nreg; ++i)
        if (strcmp(g_reg[i].name, name) == 0) { *data = g_reg[i].data; return g_reg[i].n; }
    return -1;
}

static int copy_arr(const orc_blob_t *b, const char *name, double *dst, long n)
{
    const orc_array_t *a = orc_blob_find(b, name);
    if (!a) return -1;
    long total = 1;
    for (int k = 0; k < a->ndim; ++k) total *= a->dims[k];
    if (total != n) return -2;
    memcpy(dst, a->data, sizeof(double) * (size_t)n);
    return 0;
}

void orc_finalize(void)
{
    for (int i = 0; i < g_nreg; ++i) free(g_reg[i].data);
    g_nreg = 0;
    memset(&g_orc, 0, sizeof g_orc);
}

int orc_init(const char *lw_ref_blob, const char *lw_kg_blob, const char *sw_kg_blob, double cpdair)
{
    orc_finalize();
    orc_state_t *S = &g_orc;
    orc_blob_t bref, blw, bsw;
    if (orc_blob_load(lw_ref_blob, &bref)) return 1;
    if (orc_blob_load(lw_kg_blob, &blw)) return 2;
    if (orc_blob_load(sw_kg_blob, &bsw)) return 3;

    /* lwdatinit / swdatinit: heatfac = grav*secdy/(cpdair*1.e2)  (LW init :279, SW init :239) */
    const double grav = 9.8066, secdy = 8.6400e4;
    S->lw_heatfac = grav * secdy / (cpdair * 1.e2);
    S->sw_heatfac = grav * secdy / (cpdair * 1.e2);

    int rc = 0;
    rc |= copy_arr(&bref, "lwref.pref", S->lw_pref, 59);
    rc |= copy_arr(&bref, "lwref.preflog", S->lw_preflog, 59);
    rc |= copy_arr(&bref, "lwref.tref", S->lw_tref, 59);
    rc |= copy_arr(&bref, "lwref.chi_mls", S->chi_mls, 7 * 59);
    rc |= copy_arr(&bref, "lwref.totplnk", S->totplnk, 181 * 16);
    rc |= copy_arr(&bref, "lwref.totplnkderiv", S->totplnkderiv, 181 * 16);
    rc |= copy_arr(&bref, "lwref.totplk16", S->totplk16, 181);
    rc |= copy_arr(&bsw, "swref.pref", S->sw_pref, 59);
    rc |= copy_arr(&bsw, "swref.preflog", S->sw_preflog, 59);
    rc |= copy_arr(&bsw, "swref.tref", S->sw_tref, 59);
    if (rc) return 4;

    /* LW lookup tables: LW/src/rrtmg_lw_init.f90:106-123 */
    const double expeps = 1.e-20, pade = 0.278;
    S->tau_tbl[0] = 0.0;
    S->tau_tbl[ORC_NTBL] = 1.e10;
    S->exp_tbl[0] = 1.0;
    S->exp_tbl[ORC_NTBL] = expeps;
    S->tfn_tbl[0] = 0.0;
    S->tfn_tbl[ORC_NTBL] = 1.0;
    S->lw_bpade = 1.0 / pade;
    for (int itr = 1; itr <= ORC_NTBL - 1; ++itr) {
        double tfn = (double)itr / (double)ORC_NTBL;
        S->tau_tbl[itr] = S->lw_bpade * tfn / (1.0 - tfn);
        S->exp_tbl[itr] = exp(-S->tau_tbl[itr]);
        if (S->exp_tbl[itr] <= expeps) S->exp_tbl[itr] = expeps;
        if (S->tau_tbl[itr] < 0.06)
            S->tfn_tbl[itr] = S->tau_tbl[itr] / 6.0;
        else
            S->tfn_tbl[itr] = 1.0 - 2.0 * ((1.0 / S->tau_tbl[itr]) - (S->exp_tbl[itr] / (1. - S->exp_tbl[itr])));
    }
    /* SW lookup table: SW/src/rrtmg_sw_init.f90:96-105 */
    S->sw_exp_tbl[0] = 1.0;
    S->sw_exp_tbl[ORC_NTBL] = expeps;
    S->sw_bpade = 1.0 / pade;
    for (int itr = 1; itr <= ORC_NTBL - 1; ++itr) {
        double tfn = (double)itr / (double)ORC_NTBL;
        double tau_tbl = S->sw_bpade * tfn / (1.0 - tfn);
        S->sw_exp_tbl[itr] = exp(-tau_tbl);
        if (S->sw_exp_tbl[itr] <= expeps) S->sw_exp_tbl[itr] = expeps;
    }

    compute_rwgt(16, lw_ngc, lw_ngn, lw_ngm, S->lw_rwgt);
    compute_rwgt(14, sw_ngc, sw_ngn, sw_ngm, S->sw_rwgt);

    /* cmbgb1..16 */
    for (int ib = 0; ib < 16; ++ib) {
        orc_lw_kg_t *K = &S->lw[ib];
        char pfx[16];
        snprintf(pfx, sizeof pfx, "lw%02d", ib + 1);
        const int ngc = lw_ngc[ib];
        const int *ngn = lw_ngn + (ib ? lw_ngs[ib - 1] : 0);
        const double *rw = S->lw_rwgt + 16 * ib;
        K->ng = ngc;
#define RED(o, r) reduce_named(&blw, pfx, o, r, ngc, ngn, rw)
        K->absa = RED("kao", "absa");
        K->absb = RED("kbo", "absb");
        K->selfref = RED("selfrefo", "selfref");
        K->forref = RED("forrefo", "forref");
        K->fracrefa = RED("fracrefao", "fracrefa");
        K->fracrefb = RED("fracrefbo", "fracrefb");
        K->ka_mn2 = RED("kao_mn2", "ka_mn2");
        K->kb_mn2 = RED("kbo_mn2", "kb_mn2");
        K->ka_mn2o = RED("kao_mn2o", "ka_mn2o");
        K->kb_mn2o = RED("kbo_mn2o", "kb_mn2o");
        K->ka_mo3 = RED("kao_mo3", "ka_mo3");
        K->kb_mo3 = RED("kbo_mo3", "kb_mo3");
        K->ka_mco2 = RED("kao_mco2", "ka_mco2");
        K->kb_mco2 = RED("kbo_mco2", "kb_mco2");
        K->ka_mco = RED("kao_mco", "ka_mco");
        K->ka_mo2 = RED("kao_mo2", "ka_mo2");
        K->kb_mo2 = RED("kbo_mo2", "kb_mo2");
        K->ccl4 = RED("ccl4o", "ccl4");
        K->cfc11adj = RED("cfc11adjo", "cfc11adj");
        K->cfc12 = RED("cfc12o", "cfc12");
        K->cfc22adj = RED("cfc22adjo", "cfc22adj");
#undef RED
        if (!K->absa || !K->selfref || !K->forref || !K->fracrefa) return 5;
    }
    /* cmbgb16s..29 */
    for (int ib = 0; ib < 14; ++ib) {
        orc_sw_kg_t *K = &S->sw[ib];
        char pfx[16], nm[40];
        snprintf(pfx, sizeof pfx, "sw%d", ib + 16);
        const int ngc = sw_ngc[ib];
        const int *ngn = sw_ngn + (ib ? sw_ngs[ib - 1] : 0);
        const double *rw = S->sw_rwgt + 16 * ib;
        K->ng = ngc;
#define RED(o, r) reduce_named(&bsw, pfx, o, r, ngc, ngn, rw)
        K->absa = RED("kao", "absa");
        K->absb = RED("kbo", "absb");
        K->selfref = RED("selfrefo", "selfref");
        K->forref = RED("forrefo", "forref");
        K->sfluxref = RED("sfluxrefo", "sfluxref");
        K->raylv = RED("raylo", "rayl");
        K->rayla = RED("raylao", "rayla");
        K->raylb = RED("raylbo", "raylb");
        K->abso3a = RED("abso3ao", "abso3a");
        K->abso3b = RED("abso3bo", "abso3b");
        K->absch4 = RED("absch4o", "absch4");
        K->absh2o = RED("absh2oo", "absh2o");
        K->absco2 = RED("absco2o", "absco2");
#undef RED
        snprintf(nm, sizeof nm, "%s.forrefo", pfx);
        const orc_array_t *fa = orc_blob_find(&bsw, nm);
        K->nfor = fa ? fa->dims[0] : 0;
        snprintf(nm, sizeof nm, "%s.rayl", pfx);
        const orc_array_t *ra = orc_blob_find(&bsw, nm);
        K->rayl = ra ? ra->data[0] : 0.0;
        if (!K->sfluxref) return 6;
    }
    /* SW cloud optical properties (swcldpr) */
    {
        struct { const char *nm; double *dst; int n; } t2[10] = {
            {"swcld.extliq1", &S->extliq1[0][0], 58}, {"swcld.ssaliq1", &S->ssaliq1[0][0], 58}, {"swcld.asyliq1", &S->asyliq1[0][0], 58},
            {"swcld.extice2", &S->extice2[0][0], 43}, {"swcld.ssaice2", &S->ssaice2[0][0], 43}, {"swcld.asyice2", &S->asyice2[0][0], 43},
            {"swcld.extice3", &S->extice3[0][0], 46}, {"swcld.ssaice3", &S->ssaice3[0][0], 46}, {"swcld.asyice3", &S->asyice3[0][0], 46},
            {"swcld.fdlice3", &S->fdlice3[0][0], 46}};
        for (int q = 0; q < 10; ++q) {
            const orc_array_t *a = orc_blob_find(&bsw, t2[q].nm);
            if (!a) return 8;
            for (int i = 1; i <= t2[q].n; ++i)
                for (int ib = 1; ib <= 14; ++ib) t2[q].dst[i * 15 + ib] = a->data[(i - 1) + t2[q].n * (ib - 1)];
            reg_t *r = &g_reg[g_nreg++];               /* visible through orc_get_table, column-major as in the blob */
            snprintf(r->name, sizeof r->name, "%s", t2[q].nm);
            r->n = t2[q].n * 14;
            r->data = (double *)malloc(sizeof(double) * r->n);
            memcpy(r->data, a->data, sizeof(double) * r->n);
        }
        struct { const char *nm; double *dst; } t1[6] = {{"swcld.abari", S->abari}, {"swcld.bbari", S->bbari}, {"swcld.cbari", S->cbari},
                                                           {"swcld.dbari", S->dbari}, {"swcld.ebari", S->ebari}, {"swcld.fbari", S->fbari}};
        for (int q = 0; q < 6; ++q) {
            const orc_array_t *a = orc_blob_find(&bsw, t1[q].nm);
            if (!a) return 8;
            for (int i = 1; i <= 5; ++i) t1[q].dst[i] = a->data[i - 1];
        }
    }
    /* LW cloud absorption coefficients (lwcldpr) */
    {
        const orc_array_t *a;
        if (!(a = orc_blob_find(&bref, "lwcld.abscld1"))) return 7;
        S->abscld1 = a->data[0];
        if (!(a = orc_blob_find(&bref, "lwcld.absliq0"))) return 7;
        S->absliq0 = a->data[0];
        if (!(a = orc_blob_find(&bref, "lwcld.absice0"))) return 7;
        for (int i = 1; i <= 2; ++i) S->absice0[i] = a->data[i - 1];
        if (!(a = orc_blob_find(&bref, "lwcld.absice1"))) return 7;
        for (int i = 1; i <= 2; ++i)
            for (int ib = 1; ib <= 5; ++ib) S->absice1[i][ib] = a->data[(i - 1) + 2 * (ib - 1)];
        if (!(a = orc_blob_find(&bref, "lwcld.absice2"))) return 7;
        for (int i = 1; i <= 43; ++i)
            for (int ib = 1; ib <= 16; ++ib) S->absice2[i][ib] = a->data[(i - 1) + 43 * (ib - 1)];
        if (!(a = orc_blob_find(&bref, "lwcld.absice3"))) return 7;
        for (int i = 1; i <= 46; ++i)
            for (int ib = 1; ib <= 16; ++ib) S->absice3[i][ib] = a->data[(i - 1) + 46 * (ib - 1)];
        if (!(a = orc_blob_find(&bref, "lwcld.absliq1"))) return 7;
        for (int i = 1; i <= 58; ++i)
            for (int ib = 1; ib <= 16; ++ib) S->absliq1[i][ib] = a->data[(i - 1) + 58 * (ib - 1)];
    }
    /* ECMWF aerosol optical properties (swaerpr), (nbndsw, naerec) column-major in the blob */
    {
        const char *nm3[3] = {"swaer.rsrtaua", "swaer.rsrpiza", "swaer.rsrasya"};
        double (*dst[3])[6] = {S->rsrtaua, S->rsrpiza, S->rsrasya};
        for (int q = 0; q < 3; ++q) {
            const orc_array_t *a = orc_blob_find(&bsw, nm3[q]);
            if (!a) return 6;
            for (int ib = 0; ib < 14; ++ib)
                for (int ia = 0; ia < 6; ++ia) dst[q][ib][ia] = a->data[ib + 14 * ia];
            reg_t *r = &g_reg[g_nreg++];               /* visible through orc_get_table, column-major as in the blob */
            snprintf(r->name, sizeof r->name, "%s", nm3[q]);
            r->n = 84;
            r->data = (double *)malloc(sizeof(double) * 84);
            memcpy(r->data, a->data, sizeof(double) * 84);
        }
    }
    /* export the lookup tables through the same registry (copies, so finalize can free them) */
    {
        const struct { const char *name; const double *src; } lut[3] = {
            {"lw.exp_tbl", S->exp_tbl}, {"lw.tfn_tbl", S->tfn_tbl}, {"sw.exp_tbl", S->sw_exp_tbl}};
        for (int i = 0; i < 3; ++i) {
            reg_t *r = &g_reg[g_nreg++];
            snprintf(r->name, sizeof r->name, "%s", lut[i].name);
            r->n = ORC_NTBL + 1;
            r->data = (double *)malloc(sizeof(double) * (ORC_NTBL + 1));
            memcpy(r->data, lut[i].src, sizeof(double) * (ORC_NTBL + 1));
        }
    }
    orc_blob_free(&bref);
    orc_blob_free(&blw);
    orc_blob_free(&bsw);
    S->ready = 1;
    return 0;
}
