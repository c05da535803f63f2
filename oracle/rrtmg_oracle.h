/*
 * rrtmg_oracle.h -- CPU oracle for the RRTMG LW+SW hot path of mjucker/MiMA.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a line-faithful C restatement of the reference Fortran
 * (column loop, same loop order, same lookup tables, same truncations).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (mima_b200/csrc) never includes, links or calls anything in this directory.
 *
 * PARITY STATUS ("pinning"):
 *   - The reference cannot be compiled in the authoring container (no Fortran compiler), so the
 *     oracle is pinned by restatement plus the known-answer checks the reference data allows
 *     (tests/test_oracle_tables.py): Kurucz solar source sums to rrsw_scon, Planck table vs
 *     sigma T^4, g-point weight normalisation, reference-atmosphere monotonicity.
 *   - SW runs on the REAL coefficient tables (SW/src/rrtmg_sw_k_g.f90).  No golden flux file exists
 *     in the reference for SW: parity unpinned beyond the restatement.
 *   - LW runs on SYNTHETIC coefficient tables of the declared shapes because
 *     LW/src/rrtmg_lw_k_g.f90 is stripped from the reference checkout: LW coefficients unpinned;
 *     the doc_rrtm/runs_std_atm known answers cannot be reproduced without that file.
 *
 * Reference abbreviations used in citations:
 *   LW/ = src/atmos_param/rrtm_radiation/rrtmg_lw/gcm_model/
 *   SW/ = src/atmos_param/rrtm_radiation/rrtmg_sw/gcm_model/
 */
#ifndef RRTMG_ORACLE_H
#define RRTMG_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NBNDLW 16
#define ORC_NGPTLW 140
#define ORC_NBNDSW 14
#define ORC_NGPTSW 112
#define ORC_MAXLAY 128
#define ORC_NTBL 10000

/* ---- blob reader (format written by tools/build_tables.py) ---- */
typedef struct {
    char name[32];
    int ndim;
    int dims[4];
    const double *data; /* column-major */
} orc_array_t;

typedef struct {
    int n;
    orc_array_t *arr;
    void *raw;
} orc_blob_t;

int orc_blob_load(const char *path, orc_blob_t *out);
void orc_blob_free(orc_blob_t *b);
const orc_array_t *orc_blob_find(const orc_blob_t *b, const char *name);

/* ---- reduced LW tables for one band (names as in LW/modules/rrlw_kgNN.f90) ---- */
typedef struct {
    int ng;
    const double *absa, *absb, *selfref, *forref, *fracrefa, *fracrefb;
    const double *ka_mn2, *kb_mn2, *ka_mn2o, *kb_mn2o, *ka_mo3, *kb_mo3;
    const double *ka_mco2, *kb_mco2, *ka_mco, *ka_mo2, *kb_mo2;
    const double *ccl4, *cfc11adj, *cfc12, *cfc22adj;
} orc_lw_kg_t;

/* ---- reduced SW tables for one band (names as in SW/modules/rrsw_kgNN.f90) ---- */
typedef struct {
    int ng;
    int nfor; /* leading dimension of forref: 3 or 4 */
    double rayl; /* scalar Rayleigh coefficient where the band has one */
    const double *absa, *absb, *selfref, *forref, *sfluxref;
    const double *raylv, *rayla, *raylb; /* rayl(ng) / rayla(ng,9) / raylb(ng) */
    const double *abso3a, *abso3b, *absch4, *absh2o, *absco2;
} orc_sw_kg_t;

typedef struct {
    int ready;
    /* LW */
    double lw_heatfac;
    double lw_pref[59], lw_preflog[59], lw_tref[59], chi_mls[7 * 59];
    double totplnk[181 * 16], totplk16[181];
    double totplnkderiv[181 * 16];   /* rrtmg_lw_setcoef.f90 lwavplankderiv: d(totplnk)/dT, for idrv = 1 */
    double lw_rwgt[16 * 16];
    double tau_tbl[ORC_NTBL + 1], exp_tbl[ORC_NTBL + 1], tfn_tbl[ORC_NTBL + 1];
    double lw_bpade;
    orc_lw_kg_t lw[ORC_NBNDLW];
    /* SW */
    double sw_heatfac;
    double sw_pref[59], sw_preflog[59], sw_tref[59];
    double sw_rwgt[14 * 16];
    double sw_exp_tbl[ORC_NTBL + 1];
    double sw_bpade;
    double rsrtaua[14][6], rsrpiza[14][6], rsrasya[14][6];   /* ECMWF aerosol types (iaer = 6), rrtmg_sw_init.f90:370-470 */
    /* SW cloud optical properties (swcldpr, rrtmg_sw_init.f90:1519-3341): [radius index 1-based][band 1..14] */
    double extliq1[59][15], ssaliq1[59][15], asyliq1[59][15], extice2[44][15], ssaice2[44][15], asyice2[44][15];
    double extice3[47][15], ssaice3[47][15], asyice3[47][15], fdlice3[47][15];
    double abari[6], bbari[6], cbari[6], dbari[6], ebari[6], fbari[6];
    /* LW cloud absorption coefficients (lwcldpr, rrtmg_lw_init.f90:2018-2656), Fortran 1-based indices kept */
    double abscld1, absliq0, absice0[3], absice1[3][6], absice2[44][17], absice3[47][17], absliq1[59][17];
    orc_sw_kg_t sw[ORC_NBNDSW];
} orc_state_t;

extern orc_state_t g_orc;

/* rrtmg_lw_ini + rrtmg_sw_ini (LW/src/rrtmg_lw_init.f90:28-175, SW/src/rrtmg_sw_init.f90:28-154) */
int orc_init(const char *lw_ref_blob, const char *lw_kg_blob, const char *sw_kg_blob, double cpdair);
void orc_finalize(void);

/* reduced-table export for bit-exact comparison with the product's reduction.
 * name e.g. "lw03.absa", "sw24.rayla"; returns element count or -1. */
long orc_get_table(const char *name, const double **data);

/* Per-(column,layer) / per-(column,layer,g) intermediates; any pointer may be NULL.
 * Layouts are column-major with the column index fastest, like the interface arrays. */
typedef struct {
    int *laytrop;                 /* (ncol) */
    int *jp, *jt, *jt1;           /* (ncol,nlay) */
    int *indself, *indfor, *indminor;
    double *fac00, *fac01, *fac10, *fac11;
    double *colh2o, *colco2, *colo3, *coln2o, *colco, *colch4, *colo2, *colbrd;
    double *selffac, *selffrac, *forfac, *forfrac, *minorfrac, *scaleminor, *scaleminorn2;
    double *coldry, *pwvcm;       /* (ncol,nlay), (ncol) */
    double *planklay, *planklev, *plankbnd; /* (ncol,nlay,16), (ncol,nlay+1,16), (ncol,16) */
    double *taug, *fracs;         /* (ncol,nlay,140) */
} orc_lw_stages_t;

typedef struct {
    int *laytrop;
    int *jp, *jt, *jt1, *indself, *indfor;
    double *fac00, *fac01, *fac10, *fac11;
    double *colh2o, *colco2, *colo3, *coln2o, *colch4, *colo2, *colmol;
    double *selffac, *selffrac, *forfac, *forfrac;
    double *taug, *taur;          /* (ncol,nlay,112) */
    double *sfluxzen;             /* (ncol,112) */
} orc_sw_stages_t;

/* rrtmg_lw (LW/src/rrtmg_lw_rad.nomcica.f90:80-569); arrays column-major, leading dimension ncol.
 * Only icld=0, idrv=0 are restated (the MiMA configuration); returns non-zero otherwise. */
int orc_rrtmg_lw(int ncol, int nlay, int icld, int idrv,
                 const double *play, const double *plev, const double *tlay, const double *tlev,
                 const double *tsfc, const double *h2ovmr, const double *o3vmr, const double *co2vmr,
                 const double *ch4vmr, const double *n2ovmr, const double *o2vmr,
                 const double *cfc11vmr, const double *cfc12vmr, const double *cfc22vmr,
                 const double *ccl4vmr, const double *emis, const double *tauaer,
                 int inflglw, const double *cldfr, const double *taucld,
                 int iceflglw, int liqflglw, const double *cicewp, const double *cliqwp, const double *reice, const double *reliq,
                 double *uflx, double *dflx, double *hr, double *uflxc, double *dflxc, double *hrc,
                 double *duflx_dt, double *duflxc_dt, /* (ncol, nlay+1), written when idrv == 1; may be NULL otherwise */
                 const orc_lw_stages_t *stages, int nthreads);

/* rrtmg_sw (SW/src/rrtmg_sw_rad.nomcica.f90:78-731); icld=0, iaer=0 only. */
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
                 double *swhrc, const orc_sw_stages_t *stages, int nthreads);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
