/* ref_harness.c -- C entry points around the machine-translated reference (oracle/_ref/src, tools/f90_to_c.py).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/Makefile, target _ref).  Everything that computes lives in the translated
 * sources: rrtmg_sw_ini / rrtmg_lw_ini (SW/src/rrtmg_sw_init.f90:28, LW/src/rrtmg_lw_init.f90:28) and
 * rrtmg_sw / rrtmg_lw (SW/src/rrtmg_sw_rad.nomcica.f90:78, LW/src/rrtmg_lw_rad.nomcica.f90:80).  This file only
 *   - gives them plain-pointer entry points (the translated routines take pointer + extent per assumed-shape array),
 *   - cuts a batch into column blocks and runs the blocks on a few threads (the reference's own column loop is
 *     serial; blocks are copied to contiguous temporaries because an assumed-shape array has no leading dimension),
 *   - turns the Fortran `stop 'message'` into a return code, and
 *   - exposes the module variables by name (f90_vars, generated) so that tests can read the reduced tables.
 */
#include <pthread.h>
#include <setjmp.h>
#include <stdio.h>
#include "f90ref.h"

static _Thread_local jmp_buf *t_stop;
static _Thread_local const char *t_stop_msg;

void f90_stop(const char *msg)
{
    t_stop_msg = msg;
    if (t_stop) longjmp(*t_stop, 1);
    fprintf(stderr, "f90_stop outside a harness call: %s\n", msg);
    abort();
}

/* ------------------------------------------------------------------------------------ module variables by name */
const struct f90_var *ref_find(const char *name)
{
    for (const struct f90_var *v = f90_vars; v->name; ++v)
        if (!strcmp(v->name, name)) return v;
    return 0;
}
long ref_get_var(const char *name, void **ptr, int *is_int)
{
    const struct f90_var *v = ref_find(name);
    if (!v) return -1;
    *ptr = v->ptr;
    *is_int = v->is_int;
    return v->n;
}
int ref_var_count(void) { int n = 0; while (f90_vars[n].name) ++n; return n; }
const char *ref_var_name(int i) { return f90_vars[i].name; }

/* ------------------------------------------------------------------------------------ init */
int ref_sw_init(double cpdair)
{
    jmp_buf jb;
    if (setjmp(jb)) { t_stop = 0; return 10; }
    t_stop = &jb;
    rrtmg_sw_init__rrtmg_sw_ini(cpdair);
    t_stop = 0;
    return 0;
}
int ref_lw_init(double cpdair)
{
    jmp_buf jb;
    if (setjmp(jb)) { t_stop = 0; return 10; }
    t_stop = &jb;
    rrtmg_lw_init__rrtmg_lw_ini(cpdair);
    t_stop = 0;
    return 0;
}

/* ------------------------------------------------------------------------------------ column blocks on threads */
typedef struct { const double *p; int nlev; } in2d_t;       /* (ncol, nlev) column-major, may be null */
typedef struct { double *p; int nlev; } out2d_t;

static double *take(const double *src, int ncol, int c0, int nc, int nlev)
{
    if (!src) return 0;
    double *d = malloc(sizeof(double) * (size_t)nc * nlev);
    for (int k = 0; k < nlev; ++k) memcpy(d + (size_t)k * nc, src + (size_t)k * ncol + c0, sizeof(double) * nc);
    return d;
}
static void give(double *dst, const double *blk, int ncol, int c0, int nc, int nlev)
{
    if (!dst) return;
    for (int k = 0; k < nlev; ++k) memcpy(dst + (size_t)k * ncol + c0, blk + (size_t)k * nc, sizeof(double) * nc);
}
/* (nb, ncol, nlay) -> block (nb, nc, nlay) */
static double *take_b(const double *src, int nb, int ncol, int c0, int nc, int nlay)
{
    if (!src) return 0;
    double *d = malloc(sizeof(double) * (size_t)nb * nc * nlay);
    for (int k = 0; k < nlay; ++k)
        memcpy(d + (size_t)k * nb * nc, src + ((size_t)k * ncol + c0) * nb, sizeof(double) * nb * nc);
    return d;
}

typedef struct {
    int lw;                     /* 1: rrtmg_lw, 0: rrtmg_sw */
    int ncol, nlay, icld, idrv, iaer, inflg, iceflg, liqflg, dyofyr;
    double adjes, scon;
    const double *play, *plev, *tlay, *tlev, *tsfc, *h2o, *o3, *co2, *ch4, *n2o, *o2;
    const double *cfc11, *cfc12, *cfc22, *ccl4, *emis;                  /* LW */
    const double *asdir, *asdif, *aldir, *aldif, *coszen;                /* SW */
    const double *cldfr, *taucld, *ssacld, *asmcld, *fsfcld, *cicewp, *cliqwp, *reice, *reliq;
    const double *tauaer, *ssaaer, *asmaer, *ecaer;
    double *uflx, *dflx, *hr, *uflxc, *dflxc, *hrc, *duflx_dt, *duflxc_dt;
    int block, nblock;
    int next;
    pthread_mutex_t mu;
    int rc;
} job_t;

static double *zeros(size_t n) { return calloc(n ? n : 1, sizeof(double)); }

static int run_block(job_t *J, int c0, int nc)
{
    const int L = J->nlay, N = J->ncol;
    const int nb = J->lw ? 16 : 14;
    double *play = take(J->play, N, c0, nc, L), *plev = take(J->plev, N, c0, nc, L + 1);
    double *tlay = take(J->tlay, N, c0, nc, L), *tlev = take(J->tlev, N, c0, nc, L + 1);
    double *tsfc = take(J->tsfc, N, c0, nc, 1);
    double *h2o = take(J->h2o, N, c0, nc, L), *o3 = take(J->o3, N, c0, nc, L), *co2 = take(J->co2, N, c0, nc, L);
    double *ch4 = take(J->ch4, N, c0, nc, L), *n2o = take(J->n2o, N, c0, nc, L), *o2 = take(J->o2, N, c0, nc, L);
    /* optional inputs the caller left out: what MiMA passes (zeros, rrtm_radiation.f90:346-371) */
    double *cldfr = J->cldfr ? take(J->cldfr, N, c0, nc, L) : zeros((size_t)nc * L);
    double *taucld = J->taucld ? take_b(J->taucld, nb, N, c0, nc, L) : zeros((size_t)nb * nc * L);
    double *cicewp = J->cicewp ? take(J->cicewp, N, c0, nc, L) : zeros((size_t)nc * L);
    double *cliqwp = J->cliqwp ? take(J->cliqwp, N, c0, nc, L) : zeros((size_t)nc * L);
    double *reice = J->reice ? take(J->reice, N, c0, nc, L) : zeros((size_t)nc * L);
    double *reliq = J->reliq ? take(J->reliq, N, c0, nc, L) : zeros((size_t)nc * L);
    double *uflx = zeros((size_t)nc * (L + 1)), *dflx = zeros((size_t)nc * (L + 1)), *hr = zeros((size_t)nc * L);
    double *uflxc = zeros((size_t)nc * (L + 1)), *dflxc = zeros((size_t)nc * (L + 1)), *hrc = zeros((size_t)nc * L);
    int icld = J->icld, rc = 0;
    jmp_buf jb;
    if (setjmp(jb)) {
        rc = 10;
    } else {
        t_stop = &jb;
        if (J->lw) {
            double *cfc11 = J->cfc11 ? take(J->cfc11, N, c0, nc, L) : zeros((size_t)nc * L);
            double *cfc12 = J->cfc12 ? take(J->cfc12, N, c0, nc, L) : zeros((size_t)nc * L);
            double *cfc22 = J->cfc22 ? take(J->cfc22, N, c0, nc, L) : zeros((size_t)nc * L);
            double *ccl4 = J->ccl4 ? take(J->ccl4, N, c0, nc, L) : zeros((size_t)nc * L);
            double *emis = take(J->emis, N, c0, nc, 16);
            double *tauaer = J->tauaer ? take(J->tauaer, N, c0, nc, L * 16) : zeros((size_t)nc * L * 16);
            double *du = J->idrv ? zeros((size_t)nc * (L + 1)) : 0, *duc = J->idrv ? zeros((size_t)nc * (L + 1)) : 0;
            rrtmg_lw_rad__rrtmg_lw(nc, L, &icld, J->idrv, play, nc, L, plev, nc, L + 1, tlay, nc, L, tlev, nc, L + 1, tsfc, nc,
                                   h2o, nc, L, o3, nc, L, co2, nc, L, ch4, nc, L, n2o, nc, L, o2, nc, L,
                                   cfc11, nc, L, cfc12, nc, L, cfc22, nc, L, ccl4, nc, L, emis, nc, 16,
                                   J->inflg, J->iceflg, J->liqflg, cldfr, nc, L, taucld, 16, nc, L,
                                   cicewp, nc, L, cliqwp, nc, L, reice, nc, L, reliq, nc, L, tauaer, nc, L, 16,
                                   uflx, nc, L + 1, dflx, nc, L + 1, hr, nc, L, uflxc, nc, L + 1, dflxc, nc, L + 1, hrc, nc, L,
                                   du, du ? nc : 0, du ? L + 1 : 0, duc, duc ? nc : 0, duc ? L + 1 : 0);
            if (J->idrv) { give(J->duflx_dt, du, N, c0, nc, L + 1); give(J->duflxc_dt, duc, N, c0, nc, L + 1); }
            free(cfc11); free(cfc12); free(cfc22); free(ccl4); free(emis); free(tauaer); free(du); free(duc);
        } else {
            int iaer = J->iaer;
            double *asdir = take(J->asdir, N, c0, nc, 1), *asdif = take(J->asdif, N, c0, nc, 1);
            double *aldir = take(J->aldir, N, c0, nc, 1), *aldif = take(J->aldif, N, c0, nc, 1);
            double *coszen = take(J->coszen, N, c0, nc, 1);
            double *ssacld = J->ssacld ? take_b(J->ssacld, 14, N, c0, nc, L) : zeros((size_t)14 * nc * L);
            double *asmcld = J->asmcld ? take_b(J->asmcld, 14, N, c0, nc, L) : zeros((size_t)14 * nc * L);
            double *fsfcld = J->fsfcld ? take_b(J->fsfcld, 14, N, c0, nc, L) : zeros((size_t)14 * nc * L);
            double *tauaer = J->tauaer ? take(J->tauaer, N, c0, nc, L * 14) : zeros((size_t)nc * L * 14);
            double *ssaaer = J->ssaaer ? take(J->ssaaer, N, c0, nc, L * 14) : zeros((size_t)nc * L * 14);
            double *asmaer = J->asmaer ? take(J->asmaer, N, c0, nc, L * 14) : zeros((size_t)nc * L * 14);
            double *ecaer = J->ecaer ? take(J->ecaer, N, c0, nc, L * 6) : zeros((size_t)nc * L * 6);
            rrtmg_sw_rad__rrtmg_sw(nc, L, &icld, &iaer, play, nc, L, plev, nc, L + 1, tlay, nc, L, tlev, nc, L + 1, tsfc, nc,
                                   h2o, nc, L, o3, nc, L, co2, nc, L, ch4, nc, L, n2o, nc, L, o2, nc, L,
                                   asdir, nc, asdif, nc, aldir, nc, aldif, nc, coszen, nc, J->adjes, J->dyofyr, J->scon,
                                   J->inflg, J->iceflg, J->liqflg, cldfr, nc, L, taucld, 14, nc, L, ssacld, 14, nc, L,
                                   asmcld, 14, nc, L, fsfcld, 14, nc, L, cicewp, nc, L, cliqwp, nc, L, reice, nc, L, reliq, nc, L,
                                   tauaer, nc, L, 14, ssaaer, nc, L, 14, asmaer, nc, L, 14, ecaer, nc, L, 6,
                                   uflx, nc, L + 1, dflx, nc, L + 1, hr, nc, L, uflxc, nc, L + 1, dflxc, nc, L + 1, hrc, nc, L);
            free(asdir); free(asdif); free(aldir); free(aldif); free(coszen); free(ssacld); free(asmcld); free(fsfcld);
            free(tauaer); free(ssaaer); free(asmaer); free(ecaer);
        }
    }
    t_stop = 0;
    if (!rc) {
        give(J->uflx, uflx, N, c0, nc, L + 1); give(J->dflx, dflx, N, c0, nc, L + 1); give(J->hr, hr, N, c0, nc, L);
        give(J->uflxc, uflxc, N, c0, nc, L + 1); give(J->dflxc, dflxc, N, c0, nc, L + 1); give(J->hrc, hrc, N, c0, nc, L);
    }
    free(play); free(plev); free(tlay); free(tlev); free(tsfc); free(h2o); free(o3); free(co2); free(ch4); free(n2o); free(o2);
    free(cldfr); free(taucld); free(cicewp); free(cliqwp); free(reice); free(reliq);
    free(uflx); free(dflx); free(hr); free(uflxc); free(dflxc); free(hrc);
    return rc;
}

static void *worker(void *arg)
{
    job_t *J = arg;
    for (;;) {
        pthread_mutex_lock(&J->mu);
        const int b = J->next++;
        pthread_mutex_unlock(&J->mu);
        if (b >= J->nblock) break;
        const int c0 = b * J->block, nc = c0 + J->block <= J->ncol ? J->block : J->ncol - c0;
        const int rc = run_block(J, c0, nc);
        if (rc) { pthread_mutex_lock(&J->mu); J->rc = rc; pthread_mutex_unlock(&J->mu); }
    }
    return 0;
}

static int run_job(job_t *J, int nthreads)
{
    if (J->ncol <= 0) return 0;
    if (nthreads < 1) nthreads = 1;
    J->block = 64;
    J->nblock = (J->ncol + J->block - 1) / J->block;
    if (nthreads > J->nblock) nthreads = J->nblock;
    J->next = 0;
    J->rc = 0;
    pthread_mutex_init(&J->mu, 0);
    pthread_attr_t at;
    pthread_attr_init(&at);
    pthread_attr_setstacksize(&at, (size_t)256 << 20);      /* the translated routines keep their automatic arrays on the stack */
    pthread_t th[256];
    if (nthreads > 256) nthreads = 256;
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], &at, worker, J);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], 0);
    pthread_attr_destroy(&at);
    pthread_mutex_destroy(&J->mu);
    return J->rc;
}

/* Arguments in the order of the Fortran dummy lists (LW/src/rrtmg_lw_rad.nomcica.f90:80-89); optional arrays may be null. */
int ref_rrtmg_lw(int ncol, int nlay, int icld, int idrv,
                 const double *play, const double *plev, const double *tlay, const double *tlev, const double *tsfc,
                 const double *h2o, const double *o3, const double *co2, const double *ch4, const double *n2o, const double *o2,
                 const double *cfc11, const double *cfc12, const double *cfc22, const double *ccl4, const double *emis,
                 int inflglw, int iceflglw, int liqflglw, const double *cldfr, const double *taucld, const double *cicewp,
                 const double *cliqwp, const double *reice, const double *reliq, const double *tauaer,
                 double *uflx, double *dflx, double *hr, double *uflxc, double *dflxc, double *hrc,
                 double *duflx_dt, double *duflxc_dt, int nthreads)
{
    job_t J;
    memset(&J, 0, sizeof J);
    J.lw = 1; J.ncol = ncol; J.nlay = nlay; J.icld = icld; J.idrv = idrv;
    J.inflg = inflglw; J.iceflg = iceflglw; J.liqflg = liqflglw;
    J.play = play; J.plev = plev; J.tlay = tlay; J.tlev = tlev; J.tsfc = tsfc;
    J.h2o = h2o; J.o3 = o3; J.co2 = co2; J.ch4 = ch4; J.n2o = n2o; J.o2 = o2;
    J.cfc11 = cfc11; J.cfc12 = cfc12; J.cfc22 = cfc22; J.ccl4 = ccl4; J.emis = emis;
    J.cldfr = cldfr; J.taucld = taucld; J.cicewp = cicewp; J.cliqwp = cliqwp; J.reice = reice; J.reliq = reliq; J.tauaer = tauaer;
    J.uflx = uflx; J.dflx = dflx; J.hr = hr; J.uflxc = uflxc; J.dflxc = dflxc; J.hrc = hrc;
    J.duflx_dt = duflx_dt; J.duflxc_dt = duflxc_dt;
    return run_job(&J, nthreads);
}

/* SW/src/rrtmg_sw_rad.nomcica.f90:78-88 */
int ref_rrtmg_sw(int ncol, int nlay, int icld, int iaer,
                 const double *play, const double *plev, const double *tlay, const double *tlev, const double *tsfc,
                 const double *h2o, const double *o3, const double *co2, const double *ch4, const double *n2o, const double *o2,
                 const double *asdir, const double *asdif, const double *aldir, const double *aldif, const double *coszen,
                 double adjes, int dyofyr, double scon, int inflgsw, int iceflgsw, int liqflgsw,
                 const double *cldfr, const double *taucld, const double *ssacld, const double *asmcld, const double *fsfcld,
                 const double *cicewp, const double *cliqwp, const double *reice, const double *reliq,
                 const double *tauaer, const double *ssaaer, const double *asmaer, const double *ecaer,
                 double *swuflx, double *swdflx, double *swhr, double *swuflxc, double *swdflxc, double *swhrc, int nthreads)
{
    job_t J;
    memset(&J, 0, sizeof J);
    J.lw = 0; J.ncol = ncol; J.nlay = nlay; J.icld = icld; J.iaer = iaer;
    J.inflg = inflgsw; J.iceflg = iceflgsw; J.liqflg = liqflgsw; J.adjes = adjes; J.dyofyr = dyofyr; J.scon = scon;
    J.play = play; J.plev = plev; J.tlay = tlay; J.tlev = tlev; J.tsfc = tsfc;
    J.h2o = h2o; J.o3 = o3; J.co2 = co2; J.ch4 = ch4; J.n2o = n2o; J.o2 = o2;
    J.asdir = asdir; J.asdif = asdif; J.aldir = aldir; J.aldif = aldif; J.coszen = coszen;
    J.cldfr = cldfr; J.taucld = taucld; J.ssacld = ssacld; J.asmcld = asmcld; J.fsfcld = fsfcld;
    J.cicewp = cicewp; J.cliqwp = cliqwp; J.reice = reice; J.reliq = reliq;
    J.tauaer = tauaer; J.ssaaer = ssaaer; J.asmaer = asmaer; J.ecaer = ecaer;
    J.uflx = swuflx; J.dflx = swdflx; J.hr = swhr; J.uflxc = swuflxc; J.dflxc = swdflxc; J.hrc = swhrc;
    return run_job(&J, nthreads);
}

/* ------------------------------------------------------------------------------------ stage hooks (one column) */
/* taumol_sw (SW/src/rrtmg_sw_taumol.f90:31) on the setcoef_sw state of one column; taug/taur are (nlay, 112) column-major */
int ref_sw_taumol(int nlay, double *colh2o, double *colco2, double *colch4, double *colo2, double *colo3, double *colmol,
                  int laytrop, int *jp, int *jt, int *jt1, double *fac00, double *fac01, double *fac10, double *fac11,
                  double *selffac, double *selffrac, int *indself, double *forfac, double *forfrac, int *indfor,
                  double *sfluxzen, double *taug, double *taur)
{
    rrtmg_sw_taumol__taumol_sw(nlay, colh2o, nlay, colco2, nlay, colch4, nlay, colo2, nlay, colo3, nlay, colmol, nlay, laytrop,
                               jp, nlay, jt, nlay, jt1, nlay, fac00, nlay, fac01, nlay, fac10, nlay, fac11, nlay,
                               selffac, nlay, selffrac, nlay, indself, nlay, forfac, nlay, forfrac, nlay, indfor, nlay,
                               sfluxzen, 112, taug, nlay, 112, taur, nlay, 112);
    return 0;
}
