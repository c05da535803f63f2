// sw_twostream.cuh -- reftra_sw for one clear-sky (g, layer) cell (SW/src/rrtmg_sw_reftra.f90:122-303) and the table
// exponential it uses, shared by the staged solver (sw_solver.cu) and the fused column kernel (sw_column.cu).  Written for
// translation units compiled with FMA contraction on (build.py).
#pragma once
#include "rrtmg_dev.cuh"

namespace rrtmg {

// exp(-ze) by the reference's Pade-indexed table (ze > od_lo) or 2nd-order series; also returns exp(+ze)
template <bool SMEM = false>
__device__ __forceinline__ double sw_exp(const double2 *__restrict__ tb, double ze, double bpade, double &recip)
{
    if (ze <= 0.06) {
        const double em = 1. - ze + 0.5 * ze * ze;
        recip = rcp_fast(em);
        return em;
    }
    const double tblind = ze * rcp_fast(bpade + ze);
    const int itind = (int)(10000.0 * tblind + 0.5);
    const double2 e = SMEM ? tb[itind] : ld_tbl(tb + itind);     // SMEM: tb is the block's shared-memory copy of the table
    recip = e.y;
    return e.x;
}

// reftra_sw for one clear-sky (g, layer) cell with asymmetry 0 (gamma3 = gamma4 = 1/2, zwo = zw), plus the
// direct-beam transmittance dbt = exp(-tau/mu0).
// one Newton step on the MUFU seed: relative error ~1e-12, enough wherever no table index depends on it
__device__ __forceinline__ double rcp_1n(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return fma(y, fma(-x, y, 1.0), y);
}
template <bool R1> __device__ __forceinline__ double rcp_sel(double x) { return R1 ? rcp_1n(x) : rcp_fast(x); }

template <bool R1, bool SMEM = false>
__device__ __forceinline__ void sw_reftra(const double2 *__restrict__ tb, double bpade, double prmu0, double rmu0,
                                          double tr, double tg, double &ref, double &refd, double &tra,
                                          double &trad, double &dbt)
{
    const double eps = 1.e-08, zwcrit = 0.9999995;
    const double zto1 = tr + tg;                             // ztauc
    const double zw = tr * rcp_fast(zto1);                   // zomcc
    const double zgamma1 = (8. - zw * 5.) * 0.25;
    const double zgamma2 = 3. * zw * 0.25;
    const double zed = zto1 * rmu0;                          // direct-beam optical path
    // exp(-tau/mu0) is needed by both branches; a warp usually holds lanes of both (two thirds of the warps of the
    // bench workload enter the conservative branch), so it is looked up once, before the branch
    double zep2;
    const double zem2 = sw_exp<SMEM>(tb, fmin(zed, 500.), bpade, zep2);
    if (zw >= zwcrit) {
        // conservative scattering (:162-214)
        const double za1 = zgamma1 * prmu0 - 0.5;
        const double zgt = zgamma1 * zto1;
        const double ze2 = zem2;
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
        double zep1;
        const double zem1 = sw_exp<SMEM>(tb, fmin(zrk * zto1, 500.), bpade, zep1);
        const double zdenr = fma(zr4, zep1, zr5 * zem1);     // = zdent (zt4 = zr4, zt5 = zr5)
        if (zdenr >= -eps && zdenr <= eps) {
            ref = eps;
            tra = zem2;
        } else {
            const double rd = zw * rcp_sel<R1>(zdenr);
            ref = (zr1 * zep1 - zr2 * zem1 - zr3 * zem2) * rd;
            tra = zem2 - zem2 * ((zt1 * zep1 - zt2 * zem1 - zt3 * zep2) * rd);
        }
        const double zemm = zem1 * zem1;
        // zdend = 1/((1 - zbeta*zemm)*zrkg), zbeta = (gamma1 - zrk)/zrkg
        const double zdend = rcp_sel<R1>(fma(-(zgamma1 - zrk), zemm, zrkg));
        refd = zgamma2 * (1. - zemm) * zdend;
        trad = zrk2 * zem1 * zdend;
        dbt = zem2;
    }
}

} // namespace rrtmg
