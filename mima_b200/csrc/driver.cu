// driver.cu -- the radiation driver around the two RRTMG calls, on the device (sm_100a).
//
// What MiMA's run_rrtmg does on the host between the GCM state and rrtmg_sw / rrtmg_lw, and back
// (RR/ = src/atmos_param/rrtm_radiation/):
//   zenith_kernel       compute_zenith            RR/astro.f90:59-248 (instantaneous, dt-averaged, daily mean)
//   interp_temp_kernel  interp_temp               RR/rrtm_radiation.f90:422-461
//   zonal_mean_kernel   do_zm_tracers             RR/rrtm_radiation.f90:622-626
//   top_flag_kernel     minval(phalf(:,sk+1))<=0  RR/rrtm_radiation.f90:655
//   pack_kernel         lon sub-sampling, vertical flip, Pa -> hPa, top-interface fix, ozone scaling, fixed
//                       water, clamps             RR/rrtm_radiation.f90:603-605, 638-677
//   unpack_kernel       K/day -> K/s, flip back, lon re-interpolation, tdt += tdt_rrtm, surface fluxes, OLR,
//                       incoming SW               RR/rrtm_radiation.f90:715-716, 751-752, 759-808
// With these the PCIe traffic of a radiation step is the GCM state in (p, T, q, [O3], Ts, albedo) and the
// heating rate plus four 2-D fields out, instead of the 36-/45-argument RRTMG interfaces.
//
// FMS arrays are (lon, lat, lev) column-major with level 1 = top; RRTMG arrays are (ncols_rrt, nlay) with
// level 1 = surface and column = sub-sampled lon fastest, then lat.  All elementwise and HBM-bound; lanes
// run along lon (contiguous in both layouts when lonstep = 1).
// Compiled with -fmad=false so the arithmetic order is the Fortran's (oracle/run_rrtmg.py is the checker).
#include "rrtmg_dev.cuh"

namespace rrtmg {

constexpr double DRV_PI = 3.14159265358979323846;      // constants_mod PI

// ------------------------------------------------------------------------------------------------
__global__ void zenith_kernel(int n, const double *__restrict__ lat, const double *__restrict__ lon,
                              double *__restrict__ cosz, ZenithArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double twopi = 2 * DRV_PI;
    const double la = lat[i];
    // local time in [-pi, pi) (astro.f90:108-113); Fortran modulo(a, p) = a - floor(a/p)*p
    const double x = a.radsec + lon[i];
    double time_pi = (x - floor(x / twopi) * twopi) - DRV_PI;
    if (time_pi >= DRV_PI) time_pi = time_pi - twopi;
    if (time_pi < -DRV_PI) time_pi = time_pi + twopi;
    // half-day length (:129-137)
    const double eps = 1.0E-05;
    double lat_h = la;
    if (lat_h == 0.5 * DRV_PI) lat_h = la - eps;
    if (lat_h == -0.5 * DRV_PI) lat_h = la + eps;
    const double cos_h = -tan(lat_h) * a.dec_tan;
    double h;
    if (cos_h <= -1.0) h = DRV_PI;
    else if (cos_h >= 1.0) h = 0.0;
    else h = acos(cos_h);
    const double aa = sin(la) * a.dec_sin;
    const double bb = cos(la) * a.dec_cos;
    double c;
    if (a.dt > 0 && a.dt < 86400) {
        // average over dt (:146-222): the `where` statements in source order, later ones overriding
        const double tt = time_pi + a.dt_pi;
        const double st = sin(time_pi), stt = sin(tt), sh = sin(h);
        c = 0.0;
        if (time_pi < -h && tt < -h) c = 0.0;
        if ((tt + h) != 0.0 && time_pi < -h && fabs(tt) <= h) c = aa + bb * (stt + sh) / (tt + h);
        if (time_pi < -h && h != 0.0 && h < tt) c = aa + bb * (sh + sh) / (h + h);
        if (fabs(time_pi) <= h && fabs(tt) <= h) c = aa + bb * (stt - st) / (tt - time_pi);
        if ((h - time_pi) != 0.0 && fabs(time_pi) <= h && h < tt) c = aa + bb * (sh - st) / (h - time_pi);
        if (twopi - h < tt && (tt + h - twopi) != 0.0 && time_pi <= h)
            c = (c * (h - time_pi) + (aa * (tt + h - twopi) + bb * (stt + sh))) / ((h - time_pi) + (tt + h - twopi));
        if (h < time_pi && twopi - h >= tt) c = 0.0;
        if (h < time_pi && twopi - h < tt) c = aa + bb * (stt + sh) / (tt + h - twopi);
        const double dt = (double)a.dt;
        double fracday = 0.0;
        if (time_pi < -h && tt < -h) fracday = 0.0;
        if (time_pi < -h && fabs(tt) <= h) fracday = (tt + h) / dt;
        if (time_pi < -h && h < tt) fracday = (h + h) / dt;
        if (fabs(time_pi) <= h && fabs(tt) <= h) fracday = (tt - time_pi) / dt;
        if (fabs(time_pi) <= h && h < tt) fracday = (h - time_pi) / dt;
        if (h < time_pi) fracday = 0.0;
        if (twopi - h < tt) fracday = fracday + (tt + h - twopi) / dt;
        c = c * fracday / a.radpersec;
    } else if (a.dt >= 86400) {
        c = (aa * h + bb * sin(h)) / DRV_PI;                 // daily mean (:226-227)
    } else {
        c = fabs(time_pi) <= h ? aa + bb * cos(time_pi) : 0.0;   // instantaneous (:232-238)
    }
    cosz[i] = fmax(0.0, c);
}

// ------------------------------------------------------------------------------------------------
// t_half (si, sj, sk+1); thread <-> (point, interface)
__global__ void interp_temp_kernel(int np, int sk, const double *__restrict__ z_full, const double *__restrict__ z_half,
                                   const double *__restrict__ t_surf, const double *__restrict__ t,
                                   double *__restrict__ t_half)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)np * (sk + 1)) return;
    const int k = (int)(i / np);                             // 0-based interface, 0 = top
    const size_t p = i - (size_t)k * np;
    double v;
    if (k == 0) {
        v = 0.5 * (3 * t[p] - t[p + np]);
    } else if (k == sk) {
        v = t_surf[p];
    } else {
        const size_t a = p + (size_t)k * np, b = p + (size_t)(k - 1) * np;
        const double dzk2 = 1. / (z_full[b] - z_full[a]);
        const double dzk = (z_half[a] - z_full[a]) * dzk2;
        const double dzk1 = (z_full[b] - z_half[a]) * dzk2;
        v = t[a] * dzk1 + t[b] * dzk;
    }
    t_half[i] = v;
}

// zonal mean of q(:, j, k), replicated along lon: sum(q,1)/size(q,1) in index order (one thread per (j, k):
// the sum order is the Fortran intrinsic's only up to compiler choice; sequential here and in the oracle)
__global__ void zonal_mean_kernel(int si, int nrow, const double *__restrict__ q, double *__restrict__ qm)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrow) return;
    const double *src = q + (size_t)r * si;
    double s = 0.0;
    for (int i = 0; i < si; ++i) s = s + src[i];
    qm[r] = s / si;
}

// flag = any(p_half(1:si:ls, :, 1) <= 0): run_rrtmg replaces the top interface of EVERY column when the
// minimum over the rank's columns is not positive (:655)
__global__ void top_flag_kernel(RadGeom g, const double *__restrict__ p_half, int *flag)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.ncols) return;
    const int ii = c % g.ni, j = c / g.ni;
    if (p_half[(size_t)ii * g.ls + (size_t)g.si * j] * 0.01 <= 0.0) atomicOr(flag, 1);
}

// thread <-> (rrtmg column, rrtmg level l = 0..sk); l = 0 is the surface
__global__ void pack_kernel(RadGeom g, PackArgs a)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)g.ncols * (g.sk + 1)) return;
    const int l = (int)(i / g.ncols);
    const int c = (int)(i - (size_t)l * g.ncols);
    const int ii = c % g.ni, j = c / g.ni;
    const size_t p2 = (size_t)ii * g.ls + (size_t)g.si * j;          // (lon, lat) offset in an FMS field
    const size_t np = (size_t)g.si * g.sj;
    // interfaces: FMS index sk - l (0 = top)
    {
        const size_t s = p2 + np * (size_t)(g.sk - l);
        double ph = a.p_half[s] * 0.01;
        if (l == g.sk && *a.top_flag) ph = (a.p_full[p2] * 0.01) * 0.5;     // pfull(:, sk) * 0.5, top full level
        a.phalf[i] = ph;
        double th = a.t_half[s];
        th = fmax(th, a.tmin);
        th = fmin(th, a.tmax);
        a.thalf[i] = th;
    }
    if (l < g.sk) {
        const size_t s = p2 + np * (size_t)(g.sk - 1 - l);
        const size_t o = (size_t)c + (size_t)g.ncols * l;
        const double pf = a.p_full[s];
        a.pfull[o] = pf * 0.01;
        double tf = a.t[s];
        tf = fmax(tf, a.tmin);
        tf = fmin(tf, a.tmax);
        a.tfull[o] = tf;
        double q = a.qzm ? a.qzm[(size_t)j + (size_t)g.sj * (g.sk - 1 - l)] : a.q[s];
        if (a.do_fixed_water && fabs(a.lat[p2]) <= a.fixed_water_lat && pf <= a.fixed_water_pres * 100.) q = a.fixed_water;
        a.h2o[o] = fmax(q, a.qmin);
        if (a.o3f) a.o3[o] = fmax(0.0, a.o3f[s] * a.scale_ozone);
        else a.o3[o] = a.o3_val;
    }
    if (l == 0) {
        a.cosz_rr[c] = a.coszen[p2];
        a.albedo_rr[c] = a.albedo[p2];
        a.tsrf[c] = a.t_surf[p2];
    }
}

__global__ void fill_kernel(size_t n, double *p, double v)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// zonal means for do_zm_rad: one thread per (lat, level) row of the sub-sampled grid
__global__ void zm_rad_kernel(RadGeom g, UnpackArgs a, double *__restrict__ zm_tdt, double *__restrict__ zm_fsw,
                              double *__restrict__ zm_flw)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int nrow = g.sj * g.sk;
    const double daypersec = 1. / 86400;
    if (r < nrow) {
        const int j = r % g.sj, k = r / g.sj;                 // FMS level k (0 = top)
        const int l = g.sk - 1 - k;
        double s = 0.0;
        for (int ii = 0; ii < g.ni; ++ii) {
            const size_t o = (size_t)(ii + g.ni * j) + (size_t)g.ncols * l;
            s = s + (a.swhr[o] * daypersec + a.lwhr[o] * daypersec);
        }
        zm_tdt[r] = s / max(1, g.ni);
    }
    if (r < g.sj) {
        double s1 = 0.0, s2 = 0.0;
        for (int ii = 0; ii < g.ni; ++ii) {
            const size_t c = (size_t)(ii + g.ni * r);
            s1 = s1 + (a.swdflx[c] - a.swuflx[c]);
            s2 = s2 + a.lwdflx[c];
        }
        zm_fsw[r] = s1 / max(1, g.ni);
        zm_flw[r] = s2 / max(1, g.ni);
    }
}

// thread <-> (lon, lat, level) of the GCM grid; level index k = 0 is the top
__global__ void unpack_kernel(RadGeom g, UnpackArgs a, const double *__restrict__ zm_tdt, const double *__restrict__ zm_fsw,
                              const double *__restrict__ zm_flw)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t np = (size_t)g.si * g.sj;
    if (idx >= np * g.sk) return;
    const int k = (int)(idx / np);
    const size_t p2 = idx - (size_t)k * np;
    const int ij1 = (int)(p2 % g.si), j = (int)(p2 / g.si);
    const int i = ij1 / g.ls, ij = ij1 - i * g.ls;
    int i1 = i + 1;
    if (i1 > g.ni - 1) i1 = 0;                                 // close toroidally (:764)
    const double dlon = 1. / g.ls;
    const double di = ij * dlon;
    const double daypersec = 1. / 86400;
    const int l = g.sk - 1 - k;                               // RRTMG level (0 = surface)
    const size_t c0 = (size_t)(i + g.ni * j), c1 = (size_t)(i1 + g.ni * j);
    const size_t o0 = c0 + (size_t)g.ncols * l, o1 = c1 + (size_t)g.ncols * l;
    const double sw0 = a.swhr[o0] * daypersec, sw1 = a.swhr[o1] * daypersec;      // swijk (:715)
    const double lw0 = a.lwhr[o0] * daypersec, lw1 = a.lwhr[o1] * daypersec;      // lwijk (:751)
    double tr;
    if (zm_tdt) tr = zm_tdt[(size_t)j + (size_t)g.sj * k];
    else tr = di * (sw1 + lw1) + (1. - di) * (sw0 + lw0);                         // :769-771
    if (a.tdt) a.tdt[idx] = a.tdt[idx] + tr;                                      // :782
    if (a.tdt_rad) a.tdt_rad[idx] = tr;
    if (a.tdt_sw) a.tdt_sw[idx] = di * sw1 + (1. - di) * sw0;
    if (a.tdt_lw) a.tdt_lw[idx] = di * lw1 + (1. - di) * lw0;
    if (k == 0) {
        const size_t top = (size_t)g.ncols * g.sk;            // level sk+1 of the (ncols, sk+1) flux arrays
        if (a.flux_sw) {
            const double f0 = a.swdflx[c0] - a.swuflx[c0], f1 = a.swdflx[c1] - a.swuflx[c1];   // net down SW (:791)
            a.flux_sw[p2] = zm_fsw ? zm_fsw[j] : di * f1 + (1. - di) * f0;
        }
        if (a.flux_lw) a.flux_lw[p2] = zm_flw ? zm_flw[j] : di * a.lwdflx[c1] + (1. - di) * a.lwdflx[c0];
        if (a.olr) a.olr[p2] = di * a.lwuflx[c1 + top] + (1. - di) * a.lwuflx[c0 + top];
        if (a.isr) {
            const double s0 = a.swdflx[c0 + top] - a.swuflx[c0 + top], s1 = a.swdflx[c1 + top] - a.swuflx[c1 + top];
            a.isr[p2] = di * s1 + (1. - di) * s0;
        }
    }
}

// ------------------------------------------------------------------------------------------------ launchers
static inline unsigned nblk(size_t n, int t) { return (unsigned)((n + t - 1) / t); }

int drv_zenith(int n, const double *lat, const double *lon, double *cosz, const ZenithArgs &a, cudaStream_t s)
{
    if (n > 0) zenith_kernel<<<nblk(n, 256), 256, 0, s>>>(n, lat, lon, cosz, a);
    return 1;
}
int drv_interp_temp(int np, int sk, const double *z_full, const double *z_half, const double *t_surf, const double *t,
                    double *t_half, cudaStream_t s)
{
    interp_temp_kernel<<<nblk((size_t)np * (sk + 1), 256), 256, 0, s>>>(np, sk, z_full, z_half, t_surf, t, t_half);
    return 1;
}
int drv_pack(const RadGeom &g, const PackArgs &a, double *qzm_buf, int *flag, cudaStream_t s, int top_flag)
{
    int n = 0;
    PackArgs b = a;
    // top_flag >= 0: the caller has evaluated the test over the rank's whole field (a row block must not decide it alone)
    cudaMemsetAsync(flag, top_flag > 0 ? 1 : 0, sizeof(int), s);
    if (top_flag < 0) { top_flag_kernel<<<nblk(g.ncols, 256), 256, 0, s>>>(g, a.p_half, flag); ++n; }
    b.top_flag = flag;
    if (qzm_buf) {
        zonal_mean_kernel<<<nblk((size_t)g.sj * g.sk, 128), 128, 0, s>>>(g.si, g.sj * g.sk, a.q, qzm_buf); ++n;
        b.qzm = qzm_buf;
    }
    pack_kernel<<<nblk((size_t)g.ncols * (g.sk + 1), 256), 256, 0, s>>>(g, b); ++n;
    return n;
}
int drv_fill(double *p, size_t n, double v, cudaStream_t s)
{
    fill_kernel<<<nblk(n, 256), 256, 0, s>>>(n, p, v);
    return 1;
}
int drv_unpack(const RadGeom &g, const UnpackArgs &a, double *zm_buf, cudaStream_t s)
{
    int n = 0;
    double *zt = nullptr, *zs = nullptr, *zl = nullptr;
    if (zm_buf) {
        zt = zm_buf; zs = zm_buf + (size_t)g.sj * g.sk; zl = zs + g.sj;
        zm_rad_kernel<<<nblk((size_t)g.sj * g.sk, 128), 128, 0, s>>>(g, a, zt, zs, zl); ++n;
    }
    unpack_kernel<<<nblk((size_t)g.si * g.sj * g.sk, 256), 256, 0, s>>>(g, a, zt, zs, zl); ++n;
    return n;
}

} // namespace rrtmg
