/*
 * rrtmg_b200.h -- C ABI of the B200-native RRTMG LW+SW radiation path (librrtmg_b200.so).
 *
 * Drop-in boundary: these entry points are what a Fortran ISO_C_BINDING shim for MiMA binds in place
 * of the four module procedures `run_rrtmg`/`physics_driver_init` call today (the reference has no C
 * ABI; see INTEGRATION.md for the shim):
 *
 *   rrtmg_b200_lw_init  <- subroutine rrtmg_lw_ini(cpdair)   LW/src/rrtmg_lw_init.f90:28
 *   rrtmg_b200_sw_init  <- subroutine rrtmg_sw_ini(cpdair)   SW/src/rrtmg_sw_init.f90:28
 *   rrtmg_b200_lw       <- subroutine rrtmg_lw(...)          LW/src/rrtmg_lw_rad.nomcica.f90:80-89
 *                          called at src/atmos_param/rrtm_radiation/rrtm_radiation.f90:722-748
 *   rrtmg_b200_sw       <- subroutine rrtmg_sw(...)          SW/src/rrtmg_sw_rad.nomcica.f90:78-88
 *                          called at src/atmos_param/rrtm_radiation/rrtm_radiation.f90:686-712
 *   (LW/ = src/atmos_param/rrtm_radiation/rrtmg_lw/gcm_model/, SW/ = .../rrtmg_sw/gcm_model/)
 *
 * Conventions (identical to the Fortran interface):
 *   - every array is contiguous, column-major, FP64; leading dimension = ncol unless stated;
 *     level index 1 = surface; pressures hPa, temperatures K;
 *   - h2ovmr is SPECIFIC HUMIDITY, o3vmr is MASS MIXING RATIO, all other gases volume mixing ratio
 *     (MiMA-specific, LW rad.nomcica:775-778 / SW rad.nomcica:1004-1007);
 *   - argument order and meaning are those of the Fortran dummy lists; scalars are passed by value
 *     except icld (intent(inout): clamped to 2 when outside 0..3, LW rad.nomcica:437).
 *   - outputs are fully overwritten for every column (night columns in SW get zeros).
 *
 * Extensions relative to the Fortran interface (all optional):
 *   - an input array pointer that is NULL means "all zeros" for: ch4vmr, n2ovmr, o2vmr, cfc11vmr,
 *     cfc12vmr, cfc22vmr, ccl4vmr, tauaer; emis == NULL means emissivity 1 in all bands.
 *     (MiMA passes compile-time zeros/ones for these, rrtm_radiation.f90:700-712,737-748; the shim
 *     can skip the PCIe transfer.)
 *   - cloud and SW aerosol arrays are never dereferenced when icld == 0 / iaer == 0 (as in the
 *     reference) and may be NULL or of any extent; with icld > 0 / iaer = 10 the arrays listed under
 *     "built" below are required (RRTMG_B200_ERR_BAD_ARGUMENT when NULL).
 *   - the clear-sky outputs of the host-pointer entry points (uflxc, dflxc, hrc, duflxc_dt; swuflxc, swdflxc,
 *     swhrc) may be NULL: they are then not copied back.  MiMA never reads them
 *     (rrtm_radiation.f90:716, 752, 790-791 use the total-sky arrays only), so the shim passes NULL and halves
 *     the device-to-host volume of the call.
 *   - *_device variants take device pointers plus a cudaStream_t and run asynchronously.
 *
 * Branches beyond MiMA's configuration (icld = 0, iaer = 0, idrv = 0):
 *   - built: idrv = 1 (LW flux derivative); LW icld = 1 (random overlap, rtrn) and icld = 2, 3
 *     (maximum/random overlap, rtrnmr) with inflglw = 0 (cloud optical depth taucld given per band),
 *     inflglw = 1 (one absorption coefficient times cicewp + cliqwp) and inflglw = 2 (cldprop's ice options
 *     iceflglw = 0..3 in reice and liquid options liqflglw = 0..1 in reliq; a radius outside an option's range
 *     is the Fortran `stop` and returns RRTMG_B200_ERR_CLOUD_INPUT);
 *     SW icld = 1..3 with inflgsw = 0 (taucld, ssacld, asmcld, fsfcld given per band) or inflgsw = 2
 *     (cldprop_sw's parameterisations: cicewp, cliqwp, reice, reliq with iceflgsw = 1..3 and liqflgsw = 1;
 *     a radius outside its range returns RRTMG_B200_ERR_CLOUD_INPUT); layers clear or overcast, as the
 *     reference requires, tested like the Fortran on sunlit columns only; SW iaer = 10 (tauaer, ssaaer, asmaer given per band) and
 *     iaer = 6 (ecaer: optical depth at 0.55 micron of the six ECMWF aerosol types; spectral properties from
 *     the swaer.rsrtaua/rsrpiza/rsrasya tables of the SW coefficient blob).
 *     These run in general kernels that are correct but not tuned like the clear-sky path.
 *   - RRTMG_B200_ERR_UNSUPPORTED: inflgsw = 1, for which cldprop_sw has no branch.
 *
 * Error behaviour: every function returns 0 on success or one of RRTMG_B200_ERR_*; nothing is ever
 * computed on the CPU and there is no fallback path.  The Fortran `stop 'PARTIAL CLOUD NOT ALLOWED'`
 * (SW rad.nomcica:537) maps to RRTMG_B200_ERR_PARTIAL_CLOUD (the general SW path synchronises the
 * stream to read that flag back, also in the *_device variant).
 */
#ifndef RRTMG_B200_H
#define RRTMG_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define RRTMG_B200_OK 0
#define RRTMG_B200_ERR_NOT_INITIALIZED 1 /* init not called (cf. FATAL at rrtm_radiation.f90:527-528) */
#define RRTMG_B200_ERR_UNSUPPORTED 2     /* an option the reference itself has no branch for (inflgsw = 1) */
#define RRTMG_B200_ERR_PARTIAL_CLOUD 3   /* SW rad.nomcica:537 */
#define RRTMG_B200_ERR_BAD_ARGUMENT 4
#define RRTMG_B200_ERR_CUDA 5            /* see rrtmg_b200_last_error() */
#define RRTMG_B200_ERR_TABLES 6          /* a coefficient array is missing or has the wrong shape */
#define RRTMG_B200_ERR_CLOUD_INPUT 7     /* a Fortran `stop` of cldprop / cldprop_sw: effective radius out of range */

#define RRTMG_B200_NBNDLW 16
#define RRTMG_B200_NGPTLW 140
#define RRTMG_B200_NBNDSW 14
#define RRTMG_B200_NGPTSW 112

/* Select the GPU for this process: cudaSetDevice(local_rank % device_count).  One MiMA MPI rank (one
 * block of latitude rows, src/atmos_spectral/tools/spec_mpp.f90:42-44) drives one GPU. */
int rrtmg_b200_set_device(int local_rank);

/* Coefficient ingestion.  Arrays are registered by name with the dimensions the Fortran modules declare
 * (16 g-points per band, before the cmbgbNN reduction), e.g. "lw03.kao" (9,5,13,16), "sw24.raylao"
 * (16,9), "lwref.totplnk" (181,16), "lwref.chi_mls" (7,59).  load_tables registers every array of a
 * blob written by tools/build_tables.py; set_table registers one array from host memory (what the
 * Fortran shim does with the rrlw_kgNN / rrsw_kgNN module arrays after the stock *_ini has run). */
int rrtmg_b200_load_tables(const char *blob_path);
int rrtmg_b200_set_table(const char *name, const double *data, int ndim, const int *dims);

/* rrtmg_lw_ini / rrtmg_sw_ini: constants, lookup tables, 16 -> ngc g-point reduction, upload. */
int rrtmg_b200_lw_init(double cpdair);
int rrtmg_b200_sw_init(double cpdair);
/* Releases every device buffer and forgets the registered coefficient arrays (register them again before a new init). */
int rrtmg_b200_finalize(void);

/* Export a reduced table as the device sees it, in the Fortran (column-major) order of the reduced
 * module array, e.g. "lw03.absa" (585,16).  Returns the element count, or -1. (test hook) */
long rrtmg_b200_get_table(const char *name, double *out, long capacity);
/* State of the coefficient registry.  *lw_synthetic = 1 when the registered LW k-distribution is the packaged synthetic
 * stand-in (marker array "lwmeta.synthetic" of mima_b200/data/rrtmg_lw_kg_synth.bin): every LW flux is then an exercise
 * of the algorithm, not physics.  The reference fills these tables in rrtmg_lw_ini from rrtmg_lw_k_g.f90
 * (LW/src/rrtmg_lw_init.f90:80-95), which the MiMA checkout does not carry.  Any pointer may be NULL. */
int rrtmg_b200_tables_info(int *lw_synthetic, int *lw_ready, int *sw_ready);

const char *rrtmg_b200_last_error(void);

/* Number of kernels launched by this library since load (bench.py's gpu_launches). */
long rrtmg_b200_launch_count(void);

/* ---- rrtmg_lw ------------------------------------------------------------------------------
 * play,tlay,h2ovmr..ccl4vmr,cldfr,cicewp,cliqwp,reice,reliq (ncol,nlay); plev,tlev (ncol,nlay+1);
 * tsfc (ncol); emis (ncol,16); taucld (16,ncol,nlay); tauaer (ncol,nlay,16);
 * uflx,dflx,uflxc,dflxc (ncol,nlay+1) W/m2; hr,hrc (ncol,nlay) K/day;
 * duflx_dt,duflxc_dt (ncol,nlay+1) [W/m2/K]: written when idrv == 1 (then required), ignored for idrv == 0. */
int rrtmg_b200_lw(int ncol, int nlay, int *icld, int idrv,
                  const double *play, const double *plev, const double *tlay, const double *tlev,
                  const double *tsfc,
                  const double *h2ovmr, const double *o3vmr, const double *co2vmr, const double *ch4vmr,
                  const double *n2ovmr, const double *o2vmr,
                  const double *cfc11vmr, const double *cfc12vmr, const double *cfc22vmr,
                  const double *ccl4vmr, const double *emis,
                  int inflglw, int iceflglw, int liqflglw, const double *cldfr,
                  const double *taucld, const double *cicewp, const double *cliqwp,
                  const double *reice, const double *reliq,
                  const double *tauaer,
                  double *uflx, double *dflx, double *hr, double *uflxc, double *dflxc, double *hrc,
                  double *duflx_dt, double *duflxc_dt);

/* Same arguments, device pointers, asynchronous on `stream` (a cudaStream_t; NULL = default stream). */
int rrtmg_b200_lw_device(int ncol, int nlay, int *icld, int idrv,
                         const double *play, const double *plev, const double *tlay, const double *tlev,
                         const double *tsfc,
                         const double *h2ovmr, const double *o3vmr, const double *co2vmr,
                         const double *ch4vmr, const double *n2ovmr, const double *o2vmr,
                         const double *cfc11vmr, const double *cfc12vmr, const double *cfc22vmr,
                         const double *ccl4vmr, const double *emis,
                         int inflglw, int iceflglw, int liqflglw, const double *cldfr,
                         const double *taucld, const double *cicewp, const double *cliqwp,
                         const double *reice, const double *reliq,
                         const double *tauaer,
                         double *uflx, double *dflx, double *hr, double *uflxc, double *dflxc, double *hrc,
                         double *duflx_dt, double *duflxc_dt, void *stream);

/* ---- rrtmg_sw ------------------------------------------------------------------------------
 * asdir,asdif,aldir,aldif,coszen (ncol); adjes, scon scalars; dyofyr day of year (0 = use adjes);
 * taucld,ssacld,asmcld,fsfcld (14,ncol,nlay); tauaer,ssaaer,asmaer (ncol,nlay,14); ecaer (ncol,nlay,6);
 * swuflx,swdflx,swuflxc,swdflxc (ncol,nlay+1) W/m2; swhr,swhrc (ncol,nlay) K/day. */
int rrtmg_b200_sw(int ncol, int nlay, int *icld, int *iaer,
                  const double *play, const double *plev, const double *tlay, const double *tlev,
                  const double *tsfc,
                  const double *h2ovmr, const double *o3vmr, const double *co2vmr, const double *ch4vmr,
                  const double *n2ovmr, const double *o2vmr,
                  const double *asdir, const double *asdif, const double *aldir, const double *aldif,
                  const double *coszen, double adjes, int dyofyr, double scon,
                  int inflgsw, int iceflgsw, int liqflgsw, const double *cldfr,
                  const double *taucld, const double *ssacld, const double *asmcld, const double *fsfcld,
                  const double *cicewp, const double *cliqwp, const double *reice, const double *reliq,
                  const double *tauaer, const double *ssaaer, const double *asmaer, const double *ecaer,
                  double *swuflx, double *swdflx, double *swhr, double *swuflxc, double *swdflxc,
                  double *swhrc);

int rrtmg_b200_sw_device(int ncol, int nlay, int *icld, int *iaer,
                         const double *play, const double *plev, const double *tlay, const double *tlev,
                         const double *tsfc,
                         const double *h2ovmr, const double *o3vmr, const double *co2vmr,
                         const double *ch4vmr, const double *n2ovmr, const double *o2vmr,
                         const double *asdir, const double *asdif, const double *aldir, const double *aldif,
                         const double *coszen, double adjes, int dyofyr, double scon,
                         int inflgsw, int iceflgsw, int liqflgsw, const double *cldfr,
                         const double *taucld, const double *ssacld, const double *asmcld,
                         const double *fsfcld,
                         const double *cicewp, const double *cliqwp, const double *reice,
                         const double *reliq,
                         const double *tauaer, const double *ssaaer, const double *asmaer,
                         const double *ecaer,
                         double *swuflx, double *swdflx, double *swhr, double *swuflxc, double *swdflxc,
                         double *swhrc, void *stream);

/* ---- stage dumps (test hooks; device-side intermediates of the most recent *_device / host call,
 *      valid only when the call fitted one column chunk -- see rrtmg_b200_set_chunk) -------------
 * which: "lw.taug","lw.fracs" (ncol,nlay,140); "sw.taug","sw.taur" (ncol,nlay,112); "sw.sfluxzen"
 * (ncol,112); "lw.jp","lw.jt",... packed indices unpacked to double (ncol,nlay); "lw.fac00",... ;
 * "lw.planklay" (ncol,nlay,16), "lw.planklev" (ncol,nlay+1,16), "lw.plankbnd" (ncol,16),
 * "lw.laytrop","sw.laytrop" (ncol).  Output is column-major with ncol leading, FP64. */
long rrtmg_b200_get_stage(const char *which, double *out, long capacity);

/* Columns per device pass (0 = automatic: 131072).  The workspace of one pass is ~ chunk * nlay * 3.2 KB (LW) +
 * chunk * nlay * 4.3 KB (SW): 59 GB at 131072 columns x 60 layers.  A pass whose workspace does not fit into the free device
 * memory is halved until it does (several MPI ranks may share one GPU: MiMA's shipped job size is 32 ranks). */
int rrtmg_b200_set_chunk(int ncol_per_pass);

/* Generic options:
 *   "chunk"          as rrtmg_b200_set_chunk.
 *   "host_chunk"     columns per block of the host-pointer entry points (default 16384): blocks flow through a pipeline of
 *                    "host_slots" slots, so that the H2D copies of the next blocks and the D2H copies of the previous ones
 *                    overlap the kernels of block i.
 *   "host_slots"     slots (streams + sets of device buffers) of those pipelines, 2..6, default 4.  A slot's copy-in, kernels
 *                    and copy-out run in stream order; with two slots the copy-in of block i+2 waits for the copy-out of
 *                    block i (T170L60 end to end: 2 slots 29.8 ms, 3: 25.0, 4: 24.1, against 21.8 ms of kernels).  Each
 *                    slot holds the workspace of one block (about 7 GB at 16384 columns x 60 layers, LW + SW).
 *   "run_chunk"      the same for rrtmg_b200_run_rrtmg (RRTMG columns per block of latitude rows).
 *   "share_inputs"   1: rrtmg_b200_sw keeps its device copies of the eleven arrays both codes read (play, plev, tlay, tlev,
 *                    tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr), and the rrtmg_b200_lw call that FOLLOWS it with
 *                    the same host pointers, ncol and nlay uses them instead of uploading again -- run_rrtmg passes the same
 *                    arrays to both calls (rrtm_radiation.f90:686-712, 722-748), so half of the host-to-device traffic of a
 *                    radiation step is redundant.  The caller promises not to change those arrays between the two calls.
 *                    One shot: the copies are forgotten after the LW call; an LW call with other pointers or with
 *                    LW-only array inputs (CFCs, emis, tauaer, clouds) uploads everything itself.  Default 0.
 *   "capture_stages" 1: keep a copy of lw.taug / lw.fracs, which the LW solver otherwise overwrites in place (test hook).
 *   "lw_fused"       1 (default): clear-sky longwave calls without derivatives (icld = 0, idrv = 0) run the fused column
 *                    kernel; 0: they run the staged kernels that every other longwave configuration uses.
 *   "kernel_timing"  see rrtmg_b200_kernel_times.
 *   "sw_solver_variant", "lw_rtrn_variant", "x0".."x7"   development builds only (RRTMG_B200_DEV_VARIANTS=1 python -m mima_b200.build
 *                    --force): the earlier and the experimental forms of the two solvers, kept for comparison (sw_solver.cu,
 *                    lw_solver.cu; measurements in profiles/).  The default library carries one clear-sky form of each.
 * Thread safety: the library keeps one set of workspaces and one error string per process (like the reference's module
 * variables): ONE call in flight per process -- the *_device entry points are asynchronous on the caller's stream, but a
 * second call must not be issued on another stream before the first has finished. */
int rrtmg_b200_set_option(const char *key, long value);

/* With option "kernel_timing" = 1 every kernel launch is bracketed by CUDA events on its stream.  Returns the
 * accumulated device time [ms] and launch count per kernel since the last reset, in the order
 * lw_prep, lw_taumol, lw_rtrn, sw_prep, sw_taumol, sw_solver, lw_column (arrays of 7).  lw_column is the fused clear-sky
 * longwave kernel (taumol + rtrn per column, and its flux kernel); it replaces lw_taumol and lw_rtrn for icld = 0, idrv = 0
 * unless option "lw_fused" = 0 selects the staged kernels for those calls too. */
int rrtmg_b200_kernel_times(double *ms, long *launches, int reset);

/* ---- the radiation driver around the two calls ------------------
 * Device-side replacement of the marshaling MiMA's run_rrtmg does on the host
 * (src/atmos_param/rrtm_radiation/rrtm_radiation.f90:585-808), of interp_temp (:422-461) and of
 * compute_zenith (src/atmos_param/rrtm_radiation/astro.f90:59-248).  The Fortran shim calls
 * rrtmg_b200_run_rrtmg from run_rrtmg once the alarm has decided that this is a radiation step; the alarm,
 * the netCDF interpolators and the diag manager stay in the model (FMS control plane).
 *
 * Fields are FMS-ordered (lon, lat, lev), column-major, level 1 = top, pressures in Pa -- exactly the dummy
 * arguments of run_rrtmg (:471).  The struct mirrors the entries of rrtm_radiation_nml (:104-197) and
 * astro_nml (astro.f90:24-33) that this part of run_rrtmg reads; rrtmg_b200_rad_config_default() sets the
 * Fortran defaults. */
typedef struct rrtmg_b200_rad_config {
    int include_secondary_gases;   /* pass ch4_val..ccl4_val instead of zeros (:679-712, 721-748) */
    int do_fixed_water;            /* :638-646 */
    int do_zm_tracers;             /* feed the zonal mean of q (:622-626) */
    int do_rad_time_avg;           /* average cos(zenith) over dt_rad_avg (:562-566) */
    int dt_rad_avg;                /* seconds; already resolved as in rrtm_radiation_init :336-340 (< 0 is an error here) */
    int lonstep;                   /* sub-sample longitudes (:163, :652) */
    int do_zm_rad;                 /* zonal-mean heating and surface fluxes (:768-769, :800-802) */
    int use_dyofyr;                /* astro_nml: let RRTMG compute the Earth-Sun distance from the day of year; needs
                                    * days_per_year = 365 (astro.f90:99-104 is FATAL otherwise; an error here) */
    int solday;                    /* astro_nml: perpetual day if > 0 */
    int days_per_year;             /* length_of_year() of the model calendar (360 for MiMA's thirty_day_months) */
    double scale_ozone, o3_val;
    double ch4_val, n2o_val, o2_val, cfc11_val, cfc12_val, cfc22_val, ccl4_val;
    double h2o_lower_limit, temp_lower_limit, temp_upper_limit;
    double co2ppmv;
    double fixed_water, fixed_water_pres, fixed_water_lat;
    double slowdown_rad;
    double obliq, solr_cnst, solrad, equinox_day;   /* astro_nml */
} rrtmg_b200_rad_config;

void rrtmg_b200_rad_config_default(rrtmg_b200_rad_config *cfg);

/* compute_zenith(Time, equinox_day, dt, lat, lon, cosz, dyofyr): Time as (seconds, days) of get_time();
 * lat, lon, cosz (n) host arrays in radians; dt = 0 instantaneous, 0 < dt < 86400 average over dt seconds,
 * dt >= 86400 daily mean. */
int rrtmg_b200_compute_zenith(const rrtmg_b200_rad_config *cfg, int seconds, int days, int dt, int n,
                              const double *lat, const double *lon, double *cosz, int *dyofyr);

/* interp_temp(z_full, z_half, t_surf_rad, t): t_half (si, sj, sk+1), host arrays. */
int rrtmg_b200_interp_temp(int si, int sj, int sk, const double *z_full, const double *z_half,
                           const double *t_surf_rad, const double *t, double *t_half);

/* The radiation step of run_rrtmg.  (seconds, days) = get_time(Time).
 *   in : lat, lon, albedo, t_surf_rad (si, sj); p_full, q, t (si, sj, sk); p_half (si, sj, sk+1);
 *        either t_half_in (si, sj, sk+1) or z_full (si, sj, sk) + z_half (si, sj, sk+1) for interp_temp;
 *        o3f (si, sj, sk) = the ozone field of the interpolator (do_read_ozone) or NULL for o3_val.
 *   inout: tdt (si, sj, sk) [K/s] += radiative heating.
 *   out: coszen (si, sj); any of flux_sw, flux_lw, tdt_rad, tdt_sw, tdt_lw, olr, isr, t_half_out may be NULL.
 * Host pointers; the *_device variant takes device pointers and a stream and is asynchronous. */
int rrtmg_b200_run_rrtmg(const rrtmg_b200_rad_config *cfg, int si, int sj, int sk, int seconds, int days,
                         const double *lat, const double *lon, const double *p_full, const double *p_half,
                         const double *albedo, const double *q, const double *t, const double *t_surf_rad,
                         const double *z_full, const double *z_half, const double *t_half_in, const double *o3f,
                         double *tdt, double *coszen, double *flux_sw, double *flux_lw,
                         double *tdt_rad, double *tdt_sw, double *tdt_lw, double *olr, double *isr,
                         double *t_half_out);
int rrtmg_b200_run_rrtmg_device(const rrtmg_b200_rad_config *cfg, int si, int sj, int sk, int seconds, int days,
                                const double *lat, const double *lon, const double *p_full, const double *p_half,
                                const double *albedo, const double *q, const double *t, const double *t_surf_rad,
                                const double *z_full, const double *z_half, const double *t_half_in,
                                const double *o3f,
                                double *tdt, double *coszen, double *flux_sw, double *flux_lw,
                                double *tdt_rad, double *tdt_sw, double *tdt_lw, double *olr, double *isr,
                                double *t_half_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif
