// sw_solver.cu -- RRTMG shortwave two-stream solvers on sm_100a: the STAGED kernels (they read the [col][lay][g] staging that
// sw_taumol_kernel writes).  The general kernel serves clouds and aerosols; the clear-sky one serves stage capture and option
// sw_fused = 0 -- clear-sky calls run the fused column kernel of sw_column.cu by default.
//
// spcvrt_sw (SW/src/rrtmg_sw_spcvrt.f90:296-619) + reftra_sw (rrtmg_sw_reftra.f90:129-300, kmodts=2)
//         + vrtqdr_sw (rrtmg_sw_vrtqdr.f90:103-150) + heating (rrtmg_sw_rad.nomcica.f90:686-727).
// icld = 0 and iaer = 0: the aerosol/cloud terms of the layer assembly are exact identities
// (tau_a = 0, omega_a = 1, g = 0 => delta scaling is the identity) and the total-sky stream equals the
// clear-sky stream bit for bit, so one stream is computed and stored to both outputs.
//
// Kernels in this file:
//   sw_solver_warp_kernel + sw_finish_kernel  the staged clear-sky solver (option sw_solver_variant = 4): one warp per block, the
//        reference's top-down recurrence first, then a two-term upward flux recurrence on three stored values per
//        cell -- reftra is evaluated once per cell.  Derivation at "Top-down first" in sw_solver_kernel, layout and
//        the reason for one-warp blocks at sw_solver_warp_kernel.
//   sw_solver_kernel<LMAX, STORE, OPT>  the earlier forms, selectable for comparison and run by the tests:
//        block <-> two columns (2 x 112 g-points = 7 full warps), thread <-> (column, g-point);
//        OPT bit 4 (variant 3)  the default scheme in 7-warp blocks;
//        OPT bit 3 (variant 2)  "flux propagation", bottom-up first: the up sweep (reftra fused with vrtqdr :103-121,
//            keeps rup, rupd per level) also keeps, per layer, the coefficients of
//              D_below = fa * D_above + fb * S_above,  S_below = dbt * S_above   (D diffuse, S direct downward flux)
//            with fa = trad * zreflect, fb = ((tra - dbt) + refd * rup * dbt) * zreflect, zreflect = 1/(1 - refd * rupd)
//            being the factor the bottom-up recurrence has just computed (interaction principle at the lower boundary
//            of the layer); the upward flux at a level is rupd*D + rup*S.  Algebraically vrtqdr's result
//            ((tdbt*rup + (tdn - tdbt)*rupd)/(1 - rdnd*rupd) and its pfd companion) without the (ztdn, prdnd)
//            recurrence: five stored values and six FP64 operations per cell in the down sweep;
//        variants 1, 0 and STORE  the reference's two recurrences literally: up sweep as above, down sweep = top-down
//            recurrence (:125-140) fused with the level fluxes (:144-150) and the spectral accumulation (spcvrt
//            :570-619); the lowest-layer and top-layer special cases of the reference are the general formulas at
//            rup = albedo resp. tdn = 1, rdnd = 0 (bitwise).  The five layer properties (ref, refd, tra, trad, dbt)
//            are needed by both sweeps: STORE = true keeps them in per-thread local arrays (40 B written + 40 B read
//            per cell, which at full occupancy streams through HBM), STORE = false evaluates reftra again (16 B
//            re-read per cell, ~90 more FP64 operations).
//   sw_solver_gen_kernel  aerosols / clouds (not MiMA's configuration), further down.
// All forms agree with the oracle's literal formulas to 1e-13 relative on the fluxes (tests assert 1e-9).
// The direct-beam transmittance of spcvrt :519-531 is the same table look-up as reftra's exp(-tau/mu0)
// (for tau/mu0 > 500 both hit the 1e-20 floor of exp_tbl), so it is taken from there.
// Divides go through rcp_fast/sqrt_fast; zbeta is folded into zdend's denominator.
// The sum over g-points goes through shared memory in batches of 8 levels (tile_reduce, 32 rows).
//
// This translation unit is compiled with FMA contraction on (build.py).
#include "sw_twostream.cuh"
#include <algorithm>

namespace rrtmg {

struct SwSolverConst {
    double heatfac, bpade;
    int g0[NBNDSW], rs[NBNDSW], rayl[NBNDSW];   // first g-point, table row stride, element offset of the Rayleigh row
    unsigned char ngb[NGPTSW];
};
__constant__ SwSolverConst c_ss;

int sw_solver_upload_const(const SwConst &c, const unsigned char *ngb)
{
    SwSolverConst h;
    h.heatfac = c.heatfac; h.bpade = c.bpade;
    for (int b = 0; b < NBNDSW; ++b) {
        h.g0[b] = c.band[b].g0; h.rs[b] = c.band[b].rs;
        h.rayl[b] = c.band[b].base + c.band[b].sec[SS_RAYL] * c.band[b].rs;
    }
    for (int g = 0; g < NGPTSW; ++g) h.ngb[g] = ngb[g];
    return cudaMemcpyToSymbol(c_ss, &h, sizeof h) == cudaSuccess ? 0 : -1;
}

constexpr double ZEPZEN = 1.e-10;
constexpr int SV_COLS = 2;                      // columns per block: 2 x 112 g-points = 7 full warps
constexpr int SV_THREADS = SV_COLS * NGPTSW;    // 224
constexpr int SV_S = 113;                       // tile row stride (odd)

constexpr int SV_U = 2;                         // layers per load group
constexpr int SV_WARPS = SV_THREADS / 32;       // 7
constexpr int SV_WS = 34;                       // row stride of the warp-local tile
constexpr int SV_HPC = NGPTSW / 16;             // half-warps per column: 7

#ifdef RRTMG_B200_DEV_VARIANTS       // the earlier forms of the clear-sky solver (option sw_solver_variant = 0..3): development builds only
// OPT bit 0: one-Newton reciprocals outside the table-index paths; bit 1: block-level g-sum in batches of 4 levels;
// bit 2: warp-local g-sums (no block barrier inside the sweep)
template <int LMAX, bool STORE, int OPT>
__global__ void __launch_bounds__(SV_THREADS, 4) sw_solver_kernel(SwTables T, SwIn in, SwOut out, SwWork w)
{
    constexpr bool R1 = (OPT & 1) != 0;
    constexpr bool WR = (OPT & 4) != 0;            // warp-local g-sums (no block barrier inside the sweep)
    constexpr bool FP = (OPT & 8) != 0;            // flux propagation: the down sweep needs no layer properties
    constexpr bool F2 = (OPT & 16) != 0;           // top-down first: three stored values per cell (see below)
    static_assert(!(FP && STORE) && (!FP || WR), "flux propagation keeps its own per-layer coefficients and uses the warp-local sums");
    static_assert(!(F2 && (STORE || FP)) && (!F2 || WR), "the top-down-first mode is a mode of its own");
    constexpr int NB = (WR || (OPT & 2)) ? 4 : 8;  // levels per reduction batch
    constexpr int NR = NB * 2 * SV_COLS;           // tile rows (block-level reduction)
    __shared__ double s_tile[WR ? SV_WARPS * 8 * SV_WS : NR * SV_S];
    __shared__ double s_part[WR ? SV_COLS * SV_HPC * 2 * (LMAX + 1) : NR * (SV_THREADS / NR + 1)];
    __shared__ double s_up[SV_COLS][LMAX + 1], s_dn[SV_COLS][LMAX + 1];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // warp-local layout: row = 2*slot + dir, 34 doubles per row, the upper half-warp shifted by one slot so that
    // the four lanes that add up one row (two per half-warp) hit distinct banks
    double *wt = s_tile + (WR ? wid * 8 * SV_WS + lane + (lane >> 4) : 0);
    const int klev = w.nlay;
    const int cb = threadIdx.x / NGPTSW;           // column of the block this thread works on
    const int g = threadIdx.x - cb * NGPTSW;
    const int col = blockIdx.x * SV_COLS + cb;
    const size_t old = (size_t)out.ld;
    const bool incol = col < w.nc;

    const double prmu0 = incol ? in.coszen[col] : 0.0;
    const bool active = incol && !(prmu0 < ZEPZEN);      // night columns: zeros (rad.nomcica:502-510)
    const int colr = incol ? col : 0;
    const int band = c_ss.ngb[g];
    const double bpade = c_ss.bpade;
    const double2 *__restrict__ tb = reinterpret_cast<const double2 *>(T.exptbl);
    const double mu0 = active ? prmu0 : 1.0;
    const double rmu0 = 1. / mu0;

    // band albedos (rad.nomcica:565-578): bands 16-24 and 29 near-IR, 25-28 UV/visible
    const bool uvvis = band >= 9 && band <= 12;
    // (the top-down-first mode needs them after its first sweep only and loads them there)
    const double albd = F2 ? 0. : (uvvis ? in.asdif[colr] : in.aldif[colr]);   // palbd: diffuse
    const double albp = F2 ? 0. : (uvvis ? in.asdir[colr] : in.aldir[colr]);   // palbp: direct

    // per-thread state, index = layer / level counted from the surface
    constexpr int LP = STORE ? LMAX : 1;
    double zref[LP], zrefd[LP], ztra[LP], ztrad[LP], zdbt[LP];
    double zrup[LMAX + 1], zrupd[LMAX + 1];
    constexpr int L2A = F2 ? LMAX : 1;
    double zp[L2A], zq[L2A], zr[L2A + 1];           // F2: u_above = zp*u_below + zq; rdnd per level
    constexpr int LF = FP ? LMAX : 1;
    double zfa[LF], zfb[LF], zfs[LF];               // FP: D_below = zfa*D_above + zfb*S_above, S_below = zfs*S_above
    const double *__restrict__ taug = w.taug + (size_t)colr * klev * NGPTSW + g;
    // Rayleigh optical depth: colmol(col, lay) * rayl(g) (same product as taumol_sw's `taur = colmol * rayl`);
    // band 24 reads its materialised value from taur24 (then raylg = 1).  One load per level either way.
    const bool b24 = band == 8;
    const double raylg = b24 ? 1.0 : __ldg(T.tab + c_ss.rayl[band] + g - c_ss.g0[band]);
    const double *__restrict__ taur = b24 ? w.taur24 + (size_t)colr * klev * 8 + (g - c_ss.g0[band])
                                          : w.colmol + (size_t)colr * klev;
    const int trs = b24 ? 8 : 1;              // stride per layer
    const double zincflx = active ? in.adjflux * w.sfluxzen[(size_t)col * NGPTSW + g] * prmu0 : 0.0;

    // ---- up sweep: reftra + vrtqdr bottom -> top (:103-121); the loads of the next layer group are issued
    //      before the arithmetic of the current one
    if (F2) {
        // ---- Top-down first (OPT bit 4).  Pass 1 runs vrtqdr's top-down recurrence (:125-140) fused with reftra; at
        // the upper boundary a of a layer the direct beam S_a = tdbt_a and the diffuse flux of the atmosphere above
        // over a black lower half-space E_a = tdn_a - tdbt_a are then known numbers, and with the diffuse downward
        // flux D_a = E_a + rdnd_a*U_a the layer equation U_a = ref*S_a + refd*D_a + trad*U_b becomes
        //     U_a = zp*U_b + zq,  zp = trad*zreflect,  zq = (ref*tdbt_a + refd*(tdn_a - tdbt_a))*zreflect,
        // zreflect = 1/(1 - refd*rdnd_a) being the factor the top-down recurrence computes anyway.  At the surface
        // U_0 = (albp*tdbt_0 + albd*(tdn_0 - tdbt_0))/(1 - albd*rdnd_0) (= the reference's pfu there), and the total
        // downward flux is tdn_a + rdnd_a*U_a.  The sum over g of incflx*tdn is taken in pass 1, so pass 2 (bottom-up)
        // reads three stored values per cell (zp, zq scaled by the incident flux, rdnd) and does three FP64
        // operations: algebraically vrtqdr's fluxes (:144-150), no second reftra, 24 B per cell kept.
        auto warp_rows = [&](auto store) {
            // lanes 4r..4r+3 add up row r of the warp tile (see the classic sweep below); store(row, half-warp, sum)
            __syncwarp();
            const int row = lane >> 2, q = lane & 3, half = q >> 1;
            const double *src = s_tile + (wid * 8 + row) * SV_WS + 17 * half + 8 * (q & 1);
            double acc = src[0];
#pragma unroll
            for (int j = 1; j < 8; ++j) acc += src[j];
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            if ((q & 1) == 0) store(row, 2 * wid + half, acc);
            __syncwarp();
        };
        // The gas optical depths are loaded one group of SV_U layers ahead and not touched until their group comes
        // up; the Rayleigh term (colmol: one value per layer and column, L1-resident) is loaded where it is used.
        double tdn = 1., rdnd = 0., tdbt = 1.;
        double tgn[SV_U];
#pragma unroll
        for (int j = 0; j < SV_U; ++j) tgn[j] = active ? __ldcs(taug + (size_t)max(klev - 1 - j, 0) * NGPTSW) : 0.;
        for (int kg = 0; kg <= klev; kg += SV_U) {
            double tg[SV_U];
#pragma unroll
            for (int j = 0; j < SV_U; ++j) tg[j] = tgn[j];
            if (active && kg + SV_U < klev) {      // night columns have no staged optical depths
#pragma unroll
                for (int j = 0; j < SV_U; ++j) tgn[j] = __ldcs(taug + (size_t)max(klev - 1 - (kg + SV_U + j), 0) * NGPTSW);
            }
#pragma unroll
            for (int j = 0; j < SV_U; ++j) {
                const int k = kg + j, s = klev - k;      // level s counted from the surface, layer s - 1 below it
                double row = 0.;
                if (active && s >= 0) {
                    row = zincflx * tdn;
                    zr[s] = rdnd;
                    const double dif = tdn - tdbt;
                    if (s > 0) {
                        double ref, refd, tra, trad, dbt;
                        const double trj = __ldg(taur + (s - 1) * trs) * raylg;
                        sw_reftra<R1>(tb, bpade, mu0, rmu0, trj, tg[j], ref, refd, tra, trad, dbt);
                        const double zreflect = rcp_sel<R1>(1. - refd * rdnd);
                        zp[s - 1] = trad * zreflect;
                        zq[s - 1] = zincflx * ((ref * tdbt + refd * dif) * zreflect);
                        const double tdn_n = tdbt * tra + (trad * (dif + tdbt * ref * rdnd)) * zreflect;
                        const double rdnd_n = refd + trad * trad * rdnd * zreflect;
                        tdbt = dbt * tdbt;
                        tdn = tdn_n;
                        rdnd = rdnd_n;
                    }
                }
                wt[(k & 7) * SV_WS] = row;
            }
            const int kl = min(kg + SV_U - 1, klev);
            if ((kl & 7) == 7 || kl == klev) {
                const int kb = kl & ~7;
                warp_rows([&](int row, int hw, double acc) {
                    if (kb + row <= kl) s_part[(hw * 2 + 1) * (LMAX + 1) + (klev - kb - row)] = acc;
                });
            }
        }
        // surface (level 0): the state of pass 1 is now that of the lowest level
        double u = 0.;
        if (active) {
            const double sd = uvvis ? in.asdif[colr] : in.aldif[colr];   // palbd: diffuse
            const double sp = uvvis ? in.asdir[colr] : in.aldir[colr];   // palbp: direct
            u = zincflx * ((sp * tdbt + sd * (tdn - tdbt)) * rcp_sel<R1>(1. - sd * rdnd));
        }
        // pass 2, bottom-up: u = incflx * U; the loads of F2_G levels are issued together, sums in batches of four
        constexpr int F2_G = 8;
        for (int s0 = 0; s0 <= klev; s0 += F2_G) {
            double p[F2_G], q[F2_G], r[F2_G];
            if (active) {
#pragma unroll
                for (int j = 0; j < F2_G; ++j) {
                    const int sj = min(s0 + j, klev), l = max(sj - 1, 0);
                    p[j] = zp[l]; q[j] = zq[l]; r[j] = zr[sj];
                }
            }
#pragma unroll
            for (int h = 0; h < F2_G; h += 4) {
                if (s0 + h <= klev) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int sj = s0 + h + j;
                        double pu = 0., pd = 0.;
                        if (active && sj <= klev) {
                            if (sj > 0) u = fma(p[h + j], u, q[h + j]);
                            pu = u;
                            pd = r[h + j] * u;
                        }
                        wt[(2 * j) * SV_WS] = pu;
                        wt[(2 * j + 1) * SV_WS] = pd;
                    }
                    warp_rows([&](int row, int hw, double acc) {
                        const int lev = s0 + h + (row >> 1);
                        if (lev <= klev) {
                            double *dst = s_part + (hw * 2 + (row & 1)) * (LMAX + 1) + lev;
                            *dst = (row & 1) ? *dst + acc : acc;
                        }
                    });
                }
            }
        }
    } else if (active) {
        double rup = albp, rupd = albd;      // zrup(klev+1) = palbp, zrupd(klev+1) = palbd
        zrup[0] = rup;
        zrupd[0] = rupd;
        double trn[SV_U], tgn[SV_U];
#pragma unroll
        for (int j = 0; j < SV_U; ++j) {
            const int l = min(j, klev - 1);
            trn[j] = __ldg(taur + l * trs) * raylg;
            tgn[j] = __ldcs(taug + (size_t)l * NGPTSW);
        }
        for (int l0 = 0; l0 < klev; l0 += SV_U) {
            double tr[SV_U], tg[SV_U];
#pragma unroll
            for (int j = 0; j < SV_U; ++j) { tr[j] = trn[j]; tg[j] = tgn[j]; }
            if (l0 + SV_U < klev) {
#pragma unroll
                for (int j = 0; j < SV_U; ++j) {
                    const int l = min(l0 + SV_U + j, klev - 1);
                    trn[j] = __ldg(taur + l * trs) * raylg;
                    tgn[j] = __ldcs(taug + (size_t)l * NGPTSW);
                }
            }
#pragma unroll
            for (int j = 0; j < SV_U; ++j) {
                const int l = l0 + j;
                if (l < klev) {
                    double ref, refd, tra, trad, dbt;
                    sw_reftra<R1>(tb, bpade, mu0, rmu0, tr[j], tg[j], ref, refd, tra, trad, dbt);
                    if (STORE) { zref[l] = ref; zrefd[l] = refd; ztra[l] = tra; ztrad[l] = trad; zdbt[l] = dbt; }
                    const double zreflect = rcp_sel<R1>(1. - rupd * refd);
                    const double rup_n = ref + (trad * ((tra - dbt) * rupd + dbt * rup)) * zreflect;
                    const double rupd_n = refd + trad * trad * rupd * zreflect;
                    if (FP) {
                        zfa[l] = trad * zreflect;
                        zfb[l] = ((tra - dbt) + refd * (rup * dbt)) * zreflect;
                        zfs[l] = dbt;
                    }
                    rup = rup_n;
                    rupd = rupd_n;
                    zrup[l + 1] = rup;
                    zrupd[l + 1] = rupd;
                }
            }
        }
    }

    // ---- down sweep: ztdn, prdnd, cumulative direct beam; fluxes at every level (:125-150)
    double ztdn = 1., zrdnd = 0., ztdbt = 1.;
    double trn = 0., tgn = 0.;
    if (active && !STORE && !FP) { trn = __ldg(taur + (klev - 1) * trs) * raylg; tgn = __ldcs(taug + (size_t)(klev - 1) * NGPTSW); }
    if (FP) {
        // Flux propagation (OPT bit 3).  With zreflect = 1/(1 - refd*rupd(below)) of the up sweep, the diffuse and
        // direct downward fluxes for unit incidence obey D_below = zfa*D_above + zfb*S_above, S_below = dbt*S_above
        // (interaction principle at the lower boundary of the layer), and the upward flux is rupd*D + rup*S:
        // algebraically the vrtqdr result (:125-150) without the top-down (ztdn, prdnd) recurrence, so the down
        // sweep needs no layer property -- five stored values and six FP64 operations per cell.  The loads of a
        // batch of four levels are issued together.
        double fD = 0., fS = 1.;
        for (int k0 = 0; k0 <= klev; k0 += 4) {
            double ru[4], rud[4], fa[4], fb[4], fs[4];
            if (active) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int s = max(klev - k0 - j, 0), l = max(s - 1, 0);
                    ru[j] = zrup[s]; rud[j] = zrupd[s];
                    fa[j] = zfa[l]; fb[j] = zfb[l]; fs[j] = zfs[l];
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int s = klev - k0 - j;
                double pfu = 0., pfd = 0.;
                if (active && s >= 0) {
                    pfu = zincflx * fma(rud[j], fD, ru[j] * fS);
                    pfd = zincflx * (fD + fS);
                    if (s > 0) {
                        fD = fma(fa[j], fD, fb[j] * fS);
                        fS = fs[j] * fS;
                    }
                }
                wt[(2 * j) * SV_WS] = pfu;
                wt[(2 * j + 1) * SV_WS] = pfd;
            }
            {
                const int k = min(k0 + 3, klev);
                __syncwarp();
                const int row = lane >> 2, q = lane & 3, half = q >> 1;
                const double *src = s_tile + (wid * 8 + row) * SV_WS + 17 * half + 8 * (q & 1);
                double acc = src[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) acc += src[j];
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                const int kk = k0 + (row >> 1);
                if ((q & 1) == 0 && kk <= k) {
                    const int hw = 2 * wid + half;                   // half-warp of the block: 0..13
                    const int c = hw / SV_HPC, i = hw - c * SV_HPC;
                    s_part[((c * SV_HPC + i) * 2 + (row & 1)) * (LMAX + 1) + (klev - kk)] = acc;
                }
                __syncwarp();
            }
        }
    }
    for (int k = (FP || F2) ? klev + 1 : 0; k <= klev; ++k) {
        const int s = klev - k;            // level counted from the surface
        const int slot = k & (NB - 1);
        if (active) {
            const double ru = zrup[s], rud = zrupd[s];
            const double tr = trn, tg = tgn;
            if (!STORE && s > 1) { trn = __ldg(taur + (s - 2) * trs) * raylg; tgn = __ldcs(taug + (size_t)(s - 2) * NGPTSW); }
            const double zreflect = rcp_sel<R1>(1. - zrdnd * rud);
            const double dif = ztdn - ztdbt;
            const double pfu = (ztdbt * ru + dif * rud) * zreflect;
            const double pfd = ztdbt + (dif + ztdbt * ru * zrdnd) * zreflect;
            if (WR) {
                wt[(2 * slot) * SV_WS] = zincflx * pfu;
                wt[(2 * slot + 1) * SV_WS] = zincflx * pfd;
            } else {
                s_tile[((2 * slot) * SV_COLS + cb) * SV_S + g] = zincflx * pfu;
                s_tile[((2 * slot + 1) * SV_COLS + cb) * SV_S + g] = zincflx * pfd;
            }
            if (s > 0) {
                const int l = s - 1;
                double ref, refd, tra, trad, dbt;
                if (STORE) {
                    ref = zref[l]; refd = zrefd[l]; tra = ztra[l]; trad = ztrad[l]; dbt = zdbt[l];
                } else {
                    sw_reftra<R1>(tb, bpade, mu0, rmu0, tr, tg, ref, refd, tra, trad, dbt);
                }
                const double zr = rcp_sel<R1>(1. - refd * zrdnd);
                const double ztdn_n = ztdbt * tra + (trad * (dif + ztdbt * ref * zrdnd)) * zr;
                const double zrdnd_n = refd + trad * trad * zrdnd * zr;
                ztdbt = dbt * ztdbt;
                ztdn = ztdn_n;
                zrdnd = zrdnd_n;
            }
        } else if (WR) {
            wt[(2 * slot) * SV_WS] = 0.0;
            wt[(2 * slot + 1) * SV_WS] = 0.0;
        } else {
            s_tile[((2 * slot) * SV_COLS + cb) * SV_S + g] = 0.0;
            s_tile[((2 * slot + 1) * SV_COLS + cb) * SV_S + g] = 0.0;
        }
        if (WR) {
            if (slot == NB - 1 || k == klev) {
                // lanes 4r..4r+3 add up row r: two lanes per half-warp, eight values each, then one exchange;
                // a half-warp never mixes columns (112 = 7 half-warps), so half-warp sums are the unit kept
                __syncwarp();
                const int row = lane >> 2, q = lane & 3, half = q >> 1;
                const double *src = s_tile + (wid * 8 + row) * SV_WS + 17 * half + 8 * (q & 1);
                double acc = src[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) acc += src[j];
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                const int kk = (k & ~3) + (row >> 1);
                if ((q & 1) == 0 && kk <= k) {
                    const int hw = 2 * wid + half;                   // half-warp of the block: 0..13
                    const int c = hw / SV_HPC, i = hw - c * SV_HPC;
                    s_part[((c * SV_HPC + i) * 2 + (row & 1)) * (LMAX + 1) + (klev - kk)] = acc;
                }
                __syncwarp();
            }
        } else if (slot == NB - 1 || k == klev) {
            const double sum = tile_reduce<SV_THREADS, NR, NGPTSW, SV_S>(s_tile, s_part);
            if (threadIdx.x < NR) {
                // row = ((2*slot + dir) * SV_COLS + column)
                const int c = threadIdx.x % SV_COLS, sd = threadIdx.x / SV_COLS;
                const int kk = (k & ~(NB - 1)) + (sd >> 1);
                if (kk <= k) {
                    if (sd & 1) s_dn[c][klev - kk] = sum;
                    else s_up[c][klev - kk] = sum;
                }
            }
        }
    }
    __syncthreads();
    if (WR) {
        for (int lev = g; lev <= klev; lev += NGPTSW) {
            double u = 0.0, d = 0.0;
#pragma unroll
            for (int i = 0; i < SV_HPC; ++i) {
                u += s_part[((cb * SV_HPC + i) * 2) * (LMAX + 1) + lev];
                d += s_part[((cb * SV_HPC + i) * 2 + 1) * (LMAX + 1) + lev];
            }
            s_up[cb][lev] = u;
            s_dn[cb][lev] = d;
        }
        __syncthreads();
    }
    if (incol) {
        for (int lev = g; lev <= klev; lev += NGPTSW) {
            const double u = s_up[cb][lev], d = s_dn[cb][lev];
            const size_t o = col + (size_t)lev * old;
            out.uflx[o] = u; out.dflx[o] = d; out.uflxc[o] = u; out.dflxc[o] = d;
        }
        for (int lay = g; lay < klev; lay += NGPTSW) {
            const size_t o = col + (size_t)lay * old;
            double h = 0.0;
            if (active && lay < klev - 1) {      // MiMA: no heating in the top layer (rad.nomcica:724-726)
                const double pdp = in.plev[col + (size_t)lay * in.ld] - in.plev[col + (size_t)(lay + 1) * in.ld];
                h = ((s_dn[cb][lay + 1] - s_up[cb][lay + 1]) - (s_dn[cb][lay] - s_up[cb][lay])) * (c_ss.heatfac / pdp);
            }
            out.hr[o] = h;
            out.hrc[o] = h;
        }
    }
}


#endif  // RRTMG_B200_DEV_VARIANTS

// =====================================================================================================
// The default clear-sky solver: the top-down-first scheme of sw_solver_kernel (OPT bit 4, derivation there) with one
// warp per block.  Thread <-> (column, g-point) in the flat order col * 112 + g, so a warp holds two half-warps of
// 16 g-points, each inside one column (112 = 7 x 16).  Registers and shared memory of a block are only released
// when its slowest warp is done, and the warps of a column differ by +-20 % in instructions (bands differ in how
// their lanes split between the branches of reftra): with 7-warp blocks 18 % of the stall samples sat at the final
// block barrier.  One-warp blocks retire individually.  The half-warp sums over g go to w.part
// ([col][half-warp][up, down][lev]); sw_finish_kernel adds the seven partials of a column in the fixed order the
// block-level kernel used (bitwise the same sums), writes the fluxes and the heating rates.
// =====================================================================================================
// WPB: resident one-warp blocks per SM the register allocation aims at (28 -> 72 registers, 32 -> 64);
// U: layers per load group (measured at T170L60: U = 1 11.16 ms, 2 10.79 ms, 4 11.41 ms with 108 B of spills)
// (the per-lane loop invariants kept in shared memory instead of registers -- 32 blocks per SM at 64 registers -- measured
// 10.64 against 10.71 ms: within noise, not kept)
template <int LMAX, int WPB, int U = SV_U>
__global__ void __launch_bounds__(32, WPB) sw_solver_warp_kernel(SwTables T, SwIn in, SwWork w)
{
    constexpr bool R1 = true;
    __shared__ double s_tile[8 * SV_WS];
    __shared__ double s_dn0[2][LMAX + 1];
    const int lane = threadIdx.x;
    double *wt = s_tile + lane + (lane >> 4);
    const int klev = w.nlay;
    const long long t = (long long)blockIdx.x * 32 + lane;
    const int col = (int)(t / NGPTSW);
    const int g = (int)(t - (long long)col * NGPTSW);
    const bool incol = col < w.nc;
    const double prmu0 = incol ? in.coszen[col] : 0.0;
    const bool active = incol && !(prmu0 < ZEPZEN);      // night columns: zeros (rad.nomcica:502-510)
    const int colr = incol ? col : 0;
    const int band = c_ss.ngb[g];
    const double bpade = c_ss.bpade;
    const double2 *__restrict__ tb = reinterpret_cast<const double2 *>(T.exptbl);
    const double mu0 = active ? prmu0 : 1.0;
    const double rmu0 = 1. / mu0;
    const bool uvvis = band >= 9 && band <= 12;          // bands 25-28 take the UV/visible albedos (rad.nomcica:565-578)
    double zp[LMAX], zq[LMAX], zr[LMAX + 1];             // u_above = zp*u_below + zq; rdnd per level
    const double *__restrict__ taug = w.taug + (size_t)colr * klev * NGPTSW + g;
    const bool b24 = band == 8;
    const double raylg = b24 ? 1.0 : __ldg(T.tab + c_ss.rayl[band] + g - c_ss.g0[band]);
    const double *__restrict__ taur = b24 ? w.taur24 + (size_t)colr * klev * 8 + (g - c_ss.g0[band])
                                          : w.colmol + (size_t)colr * klev;
    const int trs = b24 ? 8 : 1;
    const double zincflx = active ? in.adjflux * w.sfluxzen[(size_t)colr * NGPTSW + g] * prmu0 : 0.0;

    auto warp_rows = [&](auto store) {
        // lanes 4r..4r+3 add up row r of the tile: two lanes per half-warp, eight values each, then one exchange
        __syncwarp();
        const int row = lane >> 2, q = lane & 3, half = q >> 1;
        const double *src = s_tile + row * SV_WS + 17 * half + 8 * (q & 1);
        double acc = src[0];
#pragma unroll
        for (int j = 1; j < 8; ++j) acc += src[j];
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if ((q & 1) == 0) store(row, half, acc);
        __syncwarp();
    };

    // ---- pass 1, top -> surface: reftra + vrtqdr's top-down recurrence; sum over g of incflx * tdn
    double tdn = 1., rdnd = 0., tdbt = 1.;
    double trn[U], tgn[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
        const int l = max(klev - 1 - j, 0);
        trn[j] = active ? __ldg(taur + l * trs) : 0.;          // night columns have no staged optical depths
        tgn[j] = active ? __ldcs(taug + (size_t)l * NGPTSW) : 0.;
    }
    for (int kg = 0; kg <= klev; kg += U) {
        double tr[U], tg[U];
#pragma unroll
        for (int j = 0; j < U; ++j) { tr[j] = trn[j]; tg[j] = tgn[j]; }
        if (active && kg + U < klev) {
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const int l = max(klev - 1 - (kg + U + j), 0);
                trn[j] = __ldg(taur + l * trs);
                tgn[j] = __ldcs(taug + (size_t)l * NGPTSW);
            }
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int k = kg + j, s = klev - k;          // level s counted from the surface, layer s - 1 below it
            double row = 0.;
            if (active && s >= 0) {
                row = zincflx * tdn;
                zr[s] = rdnd;
                if (s > 0) {
                    const double dif = tdn - tdbt;
                    double ref, refd, tra, trad, dbt;
                    sw_reftra<R1>(tb, bpade, mu0, rmu0, tr[j] * raylg, tg[j], ref, refd, tra, trad, dbt);
                    const double zreflect = rcp_sel<R1>(1. - refd * rdnd);
                    zp[s - 1] = trad * zreflect;
                    zq[s - 1] = zincflx * ((ref * tdbt + refd * dif) * zreflect);
                    const double tdn_n = tdbt * tra + (trad * (dif + tdbt * ref * rdnd)) * zreflect;
                    const double rdnd_n = refd + trad * trad * rdnd * zreflect;
                    tdbt = dbt * tdbt;
                    tdn = tdn_n;
                    rdnd = rdnd_n;
                }
            }
            wt[(k & 7) * SV_WS] = row;
        }
        const int kl = min(kg + U - 1, klev);
        if ((kl & 7) == 7 || kl == klev) {
            const int kb = kl & ~7;
            warp_rows([&](int row, int half, double acc) {
                if (kb + row <= kl) s_dn0[half][klev - kb - row] = acc;
            });
        }
    }
    // surface (level 0): upward flux from the albedos (the reference's pfu there)
    double u = 0.;
    if (active) {
        const double sd = uvvis ? in.asdif[colr] : in.aldif[colr];   // palbd: diffuse
        const double sp = uvvis ? in.asdir[colr] : in.aldir[colr];   // palbp: direct
        u = zincflx * ((sp * tdbt + sd * (tdn - tdbt)) * rcp_sel<R1>(1. - sd * rdnd));
    }
    // ---- pass 2, surface -> top: u = incflx * U; loads of eight levels issued together, sums in batches of four
    const long long hw0 = (long long)blockIdx.x * 2;                 // first half-warp of this warp, = col * 7 + i
    const long long nhw = (long long)w.nc * SV_HPC;
    constexpr int G8 = 8;
    for (int s0 = 0; s0 <= klev; s0 += G8) {
        double p[G8], q[G8], r[G8];
        if (active) {
#pragma unroll
            for (int j = 0; j < G8; ++j) {
                const int sj = min(s0 + j, klev), l = max(sj - 1, 0);
                p[j] = zp[l]; q[j] = zq[l]; r[j] = zr[sj];
            }
        }
#pragma unroll
        for (int h = 0; h < G8; h += 4) {
            if (s0 + h <= klev) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int sj = s0 + h + j;
                    double pu = 0., pd = 0.;
                    if (active && sj <= klev) {
                        if (sj > 0) u = fma(p[h + j], u, q[h + j]);
                        pu = u;
                        pd = r[h + j] * u;
                    }
                    wt[(2 * j) * SV_WS] = pu;
                    wt[(2 * j + 1) * SV_WS] = pd;
                }
                warp_rows([&](int row, int half, double acc) {
                    const int lev = s0 + h + (row >> 1);
                    if (lev <= klev && hw0 + half < nhw) {
                        double *dst = w.part + ((hw0 + half) * 2 + (row & 1)) * (klev + 1) + lev;
                        *dst = (row & 1) ? s_dn0[half][lev] + acc : acc;
                    }
                });
            }
        }
    }
}

#ifdef RRTMG_B200_DEV_VARIANTS       // variants 5 and 6 (per-cell stack in an L2-resident scratch): measured, not faster -- development builds only
// =====================================================================================================
// sw_solver_l2_kernel (variant 5): the scheme of sw_solver_warp_kernel with the per-cell stack (rdnd, zp, zq: 24 B)
// kept on chip.  The local arrays of the one-warp kernel make 28 warps x 46.8 KB per SM = 194 MB of live-or-dead stack
// lines, more than the 126 MB L2: every cell's 24 B go to HBM and come back (339 KB per column against 110 KB of
// algorithmic traffic).  Here
//   * blocks are persistent (grid = SMs x WPB one-warp blocks, each walks the warp tiles blockIdx.x, +gridDim.x, ...),
//     so a block owns ONE stack slot in an explicit global scratch, [slot][level][rdnd, zp, zq][lane], for its life;
//   * the levels next to the surface (written last, read first) live in shared memory (NS levels, as many as fit);
//   * the rest is written with an L2 evict_last policy and, once pass 2 has read a group of levels, its lines are
//     dropped with discard.L2 -- they are dead until the next tile overwrites them, and a discarded line is never
//     written back to HBM.
// Knobs (Tuning::x): x0 = WPB (12, 16, 20, 24, 28), x1 bit 0 = evict_last policy, bit 1 = discard, x2 = shared levels
// (-1: as many as fit).
// =====================================================================================================
__device__ __forceinline__ unsigned long long l2_policy_evict_last()
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void st_hint(double *p, double v, unsigned long long pol)
{
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" :: "l"(p), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ double ld_hint(const double *p, unsigned long long pol)
{
    double v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ double ld_na(const double *p)
{
    double v;
    asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void l2_discard128(const void *p)
{
    asm volatile("discard.global.L2 [%0], 128;" :: "l"(p) : "memory");
}

template <int LMAX, int WPB, int PG = 4, int U = SV_U>
__global__ void __launch_bounds__(32, WPB) sw_solver_l2_kernel(SwTables T, SwIn in, SwWork w, long long nitems, int ns, int flags)
{
    constexpr bool R1 = true;
    extern __shared__ __align__(16) double s_stk[];      // [ns][3][32]
    __shared__ double s_tile[8 * SV_WS];
    __shared__ double s_dn0[2][LMAX + 1];
    const int lane = threadIdx.x;
    double *wt = s_tile + lane + (lane >> 4);
    const int klev = w.nlay;
    const bool pol_on = (flags & 1) != 0, disc_on = (flags & 2) != 0;
    const unsigned long long pol = l2_policy_evict_last();
    double *const gstk = w.stack + (size_t)blockIdx.x * (size_t)(klev + 1) * 96 + lane;    // this block's slot
    double *const sstk = s_stk + lane;
    const double bpade = c_ss.bpade;
    const double2 *__restrict__ tb = reinterpret_cast<const double2 *>(T.exptbl);
    const long long nhw = (long long)w.nc * SV_HPC;

    auto put = [&](int s, double r, double p, double q) {
        if (s < ns) {
            double *d = sstk + s * 96;
            d[0] = r; d[32] = p; d[64] = q;
        } else {
            double *d = gstk + (size_t)s * 96;
            if (pol_on) { st_hint(d, r, pol); st_hint(d + 32, p, pol); st_hint(d + 64, q, pol); }
            else { d[0] = r; d[32] = p; d[64] = q; }
        }
    };
    auto get = [&](int s, double &r, double &p, double &q) {
        if (s < ns) {
            const double *d = sstk + s * 96;
            r = d[0]; p = d[32]; q = d[64];
        } else {
            const double *d = gstk + (size_t)s * 96;
            if (pol_on) { r = ld_hint(d, pol); p = ld_hint(d + 32, pol); q = ld_hint(d + 64, pol); }
            else { r = ld_na(d); p = ld_na(d + 32); q = ld_na(d + 64); }
        }
    };
    auto warp_rows = [&](auto store) {
        __syncwarp();
        const int row = lane >> 2, q = lane & 3, half = q >> 1;
        const double *src = s_tile + row * SV_WS + 17 * half + 8 * (q & 1);
        double acc = src[0];
#pragma unroll
        for (int j = 1; j < 8; ++j) acc += src[j];
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if ((q & 1) == 0) store(row, half, acc);
        __syncwarp();
    };

    for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
        const long long t = item * 32 + lane;
        const int col = (int)(t / NGPTSW);
        const int g = (int)(t - (long long)col * NGPTSW);
        const bool incol = col < w.nc;
        const double prmu0 = incol ? in.coszen[col] : 0.0;
        const bool active = incol && !(prmu0 < ZEPZEN);      // night columns: zeros (rad.nomcica:502-510)
        const int colr = incol ? col : 0;
        const int band = c_ss.ngb[g];
        const double mu0 = active ? prmu0 : 1.0;
        const double rmu0 = 1. / mu0;
        const bool uvvis = band >= 9 && band <= 12;          // bands 25-28 take the UV/visible albedos
        const double *__restrict__ taug = w.taug + (size_t)colr * klev * NGPTSW + g;
        const bool b24 = band == 8;
        const double raylg = b24 ? 1.0 : __ldg(T.tab + c_ss.rayl[band] + g - c_ss.g0[band]);
        const double *__restrict__ taur = b24 ? w.taur24 + (size_t)colr * klev * 8 + (g - c_ss.g0[band])
                                              : w.colmol + (size_t)colr * klev;
        const int trs = b24 ? 8 : 1;
        const double zincflx = active ? in.adjflux * w.sfluxzen[(size_t)colr * NGPTSW + g] * prmu0 : 0.0;

        // ---- pass 1, top -> surface (see sw_solver_warp_kernel)
        double tdn = 1., rdnd = 0., tdbt = 1.;
        double trn[U], tgn[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int l = max(klev - 1 - j, 0);
            trn[j] = active ? __ldg(taur + l * trs) : 0.;
            tgn[j] = active ? __ldcs(taug + (size_t)l * NGPTSW) : 0.;
        }
        for (int kg = 0; kg <= klev; kg += U) {
            double tr[U], tg[U];
#pragma unroll
            for (int j = 0; j < U; ++j) { tr[j] = trn[j]; tg[j] = tgn[j]; }
            if (active && kg + U < klev) {
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    const int l = max(klev - 1 - (kg + U + j), 0);
                    trn[j] = __ldg(taur + l * trs);
                    tgn[j] = __ldcs(taug + (size_t)l * NGPTSW);
                }
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const int k = kg + j, s = klev - k;
                double row = 0.;
                if (active && s >= 0) {
                    row = zincflx * tdn;
                    double zpv = 0., zqv = 0.;
                    const double rd0 = rdnd;
                    if (s > 0) {
                        const double dif = tdn - tdbt;
                        double ref, refd, tra, trad, dbt;
                        sw_reftra<R1>(tb, bpade, mu0, rmu0, tr[j] * raylg, tg[j], ref, refd, tra, trad, dbt);
                        const double zreflect = rcp_sel<R1>(1. - refd * rdnd);
                        zpv = trad * zreflect;
                        zqv = zincflx * ((ref * tdbt + refd * dif) * zreflect);
                        const double tdn_n = tdbt * tra + (trad * (dif + tdbt * ref * rdnd)) * zreflect;
                        const double rdnd_n = refd + trad * trad * rdnd * zreflect;
                        tdbt = dbt * tdbt;
                        tdn = tdn_n;
                        rdnd = rdnd_n;
                    }
                    put(s, rd0, zpv, zqv);
                }
                wt[(k & 7) * SV_WS] = row;
            }
            const int kl = min(kg + U - 1, klev);
            if ((kl & 7) == 7 || kl == klev) {
                const int kb = kl & ~7;
                warp_rows([&](int row, int half, double acc) {
                    if (kb + row <= kl) s_dn0[half][klev - kb - row] = acc;
                });
            }
        }
        double u = 0.;
        if (active) {
            const double sd = uvvis ? in.asdif[colr] : in.aldif[colr];
            const double sp = uvvis ? in.asdir[colr] : in.aldir[colr];
            u = zincflx * ((sp * tdbt + sd * (tdn - tdbt)) * rcp_sel<R1>(1. - sd * rdnd));
        }
        // ---- pass 2, surface -> top
        const long long hw0 = item * 2;
        constexpr int G8 = PG;
        for (int s0 = 0; s0 <= klev; s0 += G8) {
            double p[G8], q[G8], r[G8];
            if (active) {
#pragma unroll
                for (int j = 0; j < G8; ++j) get(min(s0 + j, klev), r[j], p[j], q[j]);
            }
#pragma unroll
            for (int h = 0; h < G8; h += 4) {
                if (s0 + h <= klev) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int sj = s0 + h + j;
                        double pu = 0., pd = 0.;
                        if (active && sj <= klev) {
                            if (sj > 0) u = fma(p[h + j], u, q[h + j]);
                            pu = u;
                            pd = r[h + j] * u;
                        }
                        wt[(2 * j) * SV_WS] = pu;
                        wt[(2 * j + 1) * SV_WS] = pd;
                    }
                    warp_rows([&](int row, int half, double acc) {
                        const int lev = s0 + h + (row >> 1);
                        if (lev <= klev && hw0 + half < nhw) {
                            double *dst = w.part + ((hw0 + half) * 2 + (row & 1)) * (klev + 1) + lev;
                            *dst = (row & 1) ? s_dn0[half][lev] + acc : acc;
                        }
                    });
                }
            }
            if (disc_on) {
                // the values of levels s0 .. s0+7 are in use (warp_rows synchronised the warp after the last of them):
                // drop their lines, 6 x 128 B per level
                const int sa = max(s0, ns), sb = min(s0 + G8 - 1, klev);
                const int nline = (sb - sa + 1) * 6;
                const char *base = reinterpret_cast<const char *>(gstk - lane + (size_t)sa * 96);
                for (int i = lane; i < nline; i += 32) l2_discard128(base + (size_t)i * 128);
            }
        }
        __syncwarp();
    }
}

// =====================================================================================================
// sw_solver_sm_kernel (variant 6): variant 5 with ONE persistent block of W warps per SM, so that the warps of an SM can
// share a copy of the {exp, 1/exp} table in shared memory (160 KB): the two table look-ups of reftra sit on the critical
// path of every cell, and from L1/L2 (the 160 KB table does not survive in L1 next to the streaming loads: hit rate
// 13-25 %) they held a third of the stall samples.  Each warp still walks its own warp tiles; there is no block barrier
// after the table copy.  The per-cell stack goes to the L2 scratch as in variant 5 (no shared-memory tier: the table
// takes the room); pass 2 issues the loads of the next four levels before it works on the current four.  The downward
// sum of pass 1 is parked in w.part instead of shared memory.
// Knobs: x0 = warps per SM (16, 20, 24, 28), x1 bit 0 = evict_last policy, bit 1 = discard.
// =====================================================================================================
template <int LMAX, int W, bool LOC = false, int U = SV_U>
__global__ void __launch_bounds__(32 * W, 1) sw_solver_sm_kernel(SwTables T, SwIn in, SwWork w, long long nitems, int flags)
{
    constexpr bool R1 = true;
    extern __shared__ __align__(16) double2 s_exp[];      // [NTBL + 1] {exp, 1/exp}, then W warp tiles
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double *s_tile = reinterpret_cast<double *>(s_exp + (NTBL + 1)) + wid * (8 * SV_WS);
    double *wt = s_tile + lane + (lane >> 4);
    {
        const double2 *__restrict__ src = reinterpret_cast<const double2 *>(T.exptbl);
        for (int i = threadIdx.x; i <= NTBL; i += 32 * W) s_exp[i] = __ldg(src + i);
    }
    __syncthreads();
    const int klev = w.nlay;
    const bool pol_on = (flags & 1) != 0, disc_on = (flags & 2) != 0;
    const unsigned long long pol = l2_policy_evict_last();
    const long long slot = (long long)blockIdx.x * W + wid;
    double *const gstk = w.stack + (size_t)slot * (size_t)(klev + 1) * 96 + lane;    // this warp's slot
    const double bpade = c_ss.bpade;
    const long long nhw = (long long)w.nc * SV_HPC;
    const long long stride = (long long)gridDim.x * W;

    // LOC (variant 7): the stack in per-thread local memory as in sw_solver_warp_kernel; only the table moves to shared memory
    constexpr int LL = LOC ? LMAX + 1 : 1;
    double lr[LL], lp[LL], lq[LL];
    auto put = [&](int s, double r, double p, double q) {
        if (LOC) { lr[LOC ? s : 0] = r; lp[LOC ? s : 0] = p; lq[LOC ? s : 0] = q; return; }
        double *d = gstk + (size_t)s * 96;
        if (pol_on) { st_hint(d, r, pol); st_hint(d + 32, p, pol); st_hint(d + 64, q, pol); }
        else { d[0] = r; d[32] = p; d[64] = q; }
    };
    auto get = [&](int s, double &r, double &p, double &q) {
        if (LOC) { r = lr[LOC ? s : 0]; p = lp[LOC ? s : 0]; q = lq[LOC ? s : 0]; return; }
        const double *d = gstk + (size_t)s * 96;
        if (pol_on) { r = ld_hint(d, pol); p = ld_hint(d + 32, pol); q = ld_hint(d + 64, pol); }
        else { r = ld_na(d); p = ld_na(d + 32); q = ld_na(d + 64); }
    };
    auto warp_rows = [&](auto store) {
        __syncwarp();
        const int row = lane >> 2, q = lane & 3, half = q >> 1;
        const double *src = s_tile + row * SV_WS + 17 * half + 8 * (q & 1);
        double acc = src[0];
#pragma unroll
        for (int j = 1; j < 8; ++j) acc += src[j];
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if ((q & 1) == 0) store(row, half, acc);
        __syncwarp();
    };

    for (long long item = slot; item < nitems; item += stride) {
        const long long t = item * 32 + lane;
        const int col = (int)(t / NGPTSW);
        const int g = (int)(t - (long long)col * NGPTSW);
        const bool incol = col < w.nc;
        const double prmu0 = incol ? in.coszen[col] : 0.0;
        const bool active = incol && !(prmu0 < ZEPZEN);      // night columns: zeros (rad.nomcica:502-510)
        const int colr = incol ? col : 0;
        const int band = c_ss.ngb[g];
        const double mu0 = active ? prmu0 : 1.0;
        const double rmu0 = 1. / mu0;
        const bool uvvis = band >= 9 && band <= 12;          // bands 25-28 take the UV/visible albedos
        const double *__restrict__ taug = w.taug + (size_t)colr * klev * NGPTSW + g;
        const bool b24 = band == 8;
        const double raylg = b24 ? 1.0 : __ldg(T.tab + c_ss.rayl[band] + g - c_ss.g0[band]);
        const double *__restrict__ taur = b24 ? w.taur24 + (size_t)colr * klev * 8 + (g - c_ss.g0[band])
                                              : w.colmol + (size_t)colr * klev;
        const int trs = b24 ? 8 : 1;
        const double zincflx = active ? in.adjflux * w.sfluxzen[(size_t)colr * NGPTSW + g] * prmu0 : 0.0;
        const long long hw0 = item * 2;                      // first half-warp of this warp tile, = col * 7 + i

        // ---- pass 1, top -> surface (see sw_solver_warp_kernel)
        double tdn = 1., rdnd = 0., tdbt = 1.;
        double trn[U], tgn[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int l = max(klev - 1 - j, 0);
            trn[j] = active ? __ldg(taur + l * trs) : 0.;
            tgn[j] = active ? __ldcs(taug + (size_t)l * NGPTSW) : 0.;
        }
        for (int kg = 0; kg <= klev; kg += U) {
            double tr[U], tg[U];
#pragma unroll
            for (int j = 0; j < U; ++j) { tr[j] = trn[j]; tg[j] = tgn[j]; }
            if (active && kg + U < klev) {
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    const int l = max(klev - 1 - (kg + U + j), 0);
                    trn[j] = __ldg(taur + l * trs);
                    tgn[j] = __ldcs(taug + (size_t)l * NGPTSW);
                }
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const int k = kg + j, s = klev - k;
                double row = 0.;
                if (active && s >= 0) {
                    row = zincflx * tdn;
                    double zpv = 0., zqv = 0.;
                    const double rd0 = rdnd;
                    if (s > 0) {
                        const double dif = tdn - tdbt;
                        double ref, refd, tra, trad, dbt;
                        sw_reftra<R1, true>(s_exp, bpade, mu0, rmu0, tr[j] * raylg, tg[j], ref, refd, tra, trad, dbt);
                        const double zreflect = rcp_sel<R1>(1. - refd * rdnd);
                        zpv = trad * zreflect;
                        zqv = zincflx * ((ref * tdbt + refd * dif) * zreflect);
                        const double tdn_n = tdbt * tra + (trad * (dif + tdbt * ref * rdnd)) * zreflect;
                        const double rdnd_n = refd + trad * trad * rdnd * zreflect;
                        tdbt = dbt * tdbt;
                        tdn = tdn_n;
                        rdnd = rdnd_n;
                    }
                    put(s, rd0, zpv, zqv);
                }
                wt[(k & 7) * SV_WS] = row;
            }
            const int kl = min(kg + U - 1, klev);
            if ((kl & 7) == 7 || kl == klev) {
                const int kb = kl & ~7;
                warp_rows([&](int row, int half, double acc) {
                    if (kb + row <= kl && hw0 + half < nhw)
                        w.part[((hw0 + half) * 2 + 1) * (klev + 1) + (klev - kb - row)] = acc;     // sum of incflx * tdn
                });
            }
        }
        double u = 0.;
        if (active) {
            const double sd = uvvis ? in.asdif[colr] : in.aldif[colr];
            const double sp = uvvis ? in.asdir[colr] : in.aldir[colr];
            u = zincflx * ((sp * tdbt + sd * (tdn - tdbt)) * rcp_sel<R1>(1. - sd * rdnd));
        }
        // ---- pass 2, surface -> top: the loads of levels s0+4 .. s0+7 are in flight while s0 .. s0+3 are worked on
        double pn[4], qn[4], rn[4];
        if (active) {
#pragma unroll
            for (int j = 0; j < 4; ++j) get(min(j, klev), rn[j], pn[j], qn[j]);
        }
        for (int s0 = 0; s0 <= klev; s0 += 4) {
            double p[4], q[4], r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { p[j] = pn[j]; q[j] = qn[j]; r[j] = rn[j]; }
            if (active && s0 + 4 <= klev) {
#pragma unroll
                for (int j = 0; j < 4; ++j) get(min(s0 + 4 + j, klev), rn[j], pn[j], qn[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int sj = s0 + j;
                double pu = 0., pd = 0.;
                if (active && sj <= klev) {
                    if (sj > 0) u = fma(p[j], u, q[j]);
                    pu = u;
                    pd = r[j] * u;
                }
                wt[(2 * j) * SV_WS] = pu;
                wt[(2 * j + 1) * SV_WS] = pd;
            }
            warp_rows([&](int row, int half, double acc) {
                const int lev = s0 + (row >> 1);
                if (lev <= klev && hw0 + half < nhw) {
                    double *dst = w.part + ((hw0 + half) * 2 + (row & 1)) * (klev + 1) + lev;
                    *dst = (row & 1) ? *dst + acc : acc;
                }
            });
            if (!LOC && disc_on && s0 >= 4) {
                // levels s0-4 .. s0-1 have been consumed (warp_rows synchronised the warp after their use): drop their lines
                const int nline = 4 * 6;
                const char *base = reinterpret_cast<const char *>(gstk - lane + (size_t)(s0 - 4) * 96);
                if (lane < nline) l2_discard128(base + (size_t)lane * 128);
            }
        }
        __syncwarp();
    }
}

#endif  // RRTMG_B200_DEV_VARIANTS

// Adds the seven half-warp partials of a column and level, writes the fluxes and heating rates
// (rrtmg_sw_rad.nomcica.f90:686-727; clear == total for icld = 0).  A block takes TC columns: the partials are read
// with the level index fastest, the outputs written with the column index fastest (transpose through shared memory).
template <int LMAX, int TC>
__global__ void __launch_bounds__(256) sw_finish_kernel(SwIn in, SwOut out, SwWork w)
{
    __shared__ double s_u[LMAX + 1][TC + 1], s_d[LMAX + 1][TC + 1];
    const int klev = w.nlay, nlev = klev + 1;
    const int c0 = blockIdx.x * TC;
    for (int i = threadIdx.x; i < TC * nlev; i += 256) {
        const int c = i / nlev, lev = i - c * nlev;
        double u = 0.0, d = 0.0;
        if (c0 + c < w.nc) {
            const double *src = w.part + (size_t)(c0 + c) * SV_HPC * 2 * nlev + lev;
#pragma unroll
            for (int h = 0; h < SV_HPC; ++h) {
                u += src[(size_t)(2 * h) * nlev];
                d += src[(size_t)(2 * h + 1) * nlev];
            }
        }
        s_u[lev][c] = u;
        s_d[lev][c] = d;
    }
    __syncthreads();
    const size_t old = (size_t)out.ld;
    for (int i = threadIdx.x; i < TC * nlev; i += 256) {
        const int lev = i / TC, c = i - lev * TC, col = c0 + c;
        if (col >= w.nc) continue;
        const double u = s_u[lev][c], d = s_d[lev][c];
        const size_t o = col + (size_t)lev * old;
        out.uflx[o] = u; out.dflx[o] = d; out.uflxc[o] = u; out.dflxc[o] = d;
        if (lev < klev) {
            double h = 0.0;
            const bool active = !(in.coszen[col] < ZEPZEN);
            if (active && lev < klev - 1) {      // MiMA: no heating in the top layer (rad.nomcica:724-726)
                const double pdp = in.plev[col + (size_t)lev * in.ld] - in.plev[col + (size_t)(lev + 1) * in.ld];
                h = ((s_d[lev + 1][c] - s_u[lev + 1][c]) - (d - u)) * (c_ss.heatfac / pdp);
            }
            out.hr[o] = h;
            out.hrc[o] = h;
        }
    }
}


// =====================================================================================================
// General path of spcvrt_sw (SW/src/rrtmg_sw_spcvrt.f90:296-619): aerosols (iaer = 10) and clouds (icld >= 1, optical
// properties given, layers clear or overcast).  Not MiMA's configuration and not tuned: the layout of
// sw_solver_kernel (two columns per block, block-level g-sums) with reftra_sw in full (asymmetry != 0, delta
// scaling :446-463), a clear-sky and a total-sky stream (CLOUD), and both direct-beam look-ups (:519-548).
// =====================================================================================================
__device__ __forceinline__ double sw_tbl_exp(const double2 *__restrict__ tb, double ze1, double bpade)
{
    if (ze1 <= 0.06) return 1. - ze1 + 0.5 * ze1 * ze1;
    const double tblind = ze1 / (bpade + ze1);
    const int itind = (int)(10000.0 * tblind + 0.5);
    return __ldg(tb + itind).x;
}
// reftra_sw, kmodts = 2, one layer (reftra.f90:129-300); lrtchk false -> R = 0, T = 1 (:138-142)
__device__ __noinline__ void sw_reftra_gen(const double2 *__restrict__ tb, double bpade, bool lrtchk, double zg, double prmuz,
                                           double zto1, double zw, double &pref, double &prefd, double &ptra, double &ptrad)
{
    const double eps = 1.e-08, zwcrit = 0.9999995;
    if (!lrtchk) { pref = 0.; ptra = 1.; prefd = 0.; ptrad = 1.; return; }
    const double zg3 = 3. * zg;
    const double zgamma1 = (8. - zw * (5. + zg3)) * 0.25;
    const double zgamma2 = 3. * (zw * (1. - zg)) * 0.25;
    const double zgamma3 = (2. - zg3 * prmuz) * 0.25;
    const double zgamma4 = 1. - zgamma3;
    const double t = zg / (1. - zg);
    const double zwo = zw / (1. - (1. - zw) * (t * t));
    if (zwo >= zwcrit) {
        const double za = zgamma1 * prmuz;
        const double za1 = za - zgamma3;
        const double zgt = zgamma1 * zto1;
        const double ze2 = sw_tbl_exp(tb, fmin(zto1 / prmuz, 500.), bpade);
        pref = (zgt - za1 * (1. - ze2)) / (1. + zgt);
        ptra = 1. - pref;
        prefd = zgt / (1. + zgt);
        ptrad = 1. - prefd;
        if (ze2 == 1.0) { pref = 0.0; ptra = 1.0; prefd = 0.0; ptrad = 1.0; }
        return;
    }
    const double za1 = zgamma1 * zgamma4 + zgamma2 * zgamma3;
    const double za2 = zgamma1 * zgamma3 + zgamma2 * zgamma4;
    const double zrk = sqrt(zgamma1 * zgamma1 - zgamma2 * zgamma2);
    const double zrp = zrk * prmuz;
    const double zrp1 = 1. + zrp;
    const double zrm1 = 1. - zrp;
    const double zrk2 = 2. * zrk;
    const double zrpp = 1. - zrp * zrp;
    const double zrkg = zrk + zgamma1;
    const double zr1 = zrm1 * (za2 + zrk * zgamma3);
    const double zr2 = zrp1 * (za2 - zrk * zgamma3);
    const double zr3 = zrk2 * (zgamma3 - za2 * prmuz);
    const double zr4 = zrpp * zrkg;
    const double zr5 = zrpp * (zrk - zgamma1);
    const double zt1 = zrp1 * (za1 + zrk * zgamma4);
    const double zt2 = zrm1 * (za1 - zrk * zgamma4);
    const double zt3 = zrk2 * (zgamma4 + za1 * prmuz);
    const double zbeta = (zgamma1 - zrk) / zrkg;
    const double zem1 = sw_tbl_exp(tb, fmin(zrk * zto1, 500.), bpade);
    const double zep1 = 1. / zem1;
    const double zem2 = sw_tbl_exp(tb, fmin(zto1 / prmuz, 500.), bpade);
    const double zep2 = 1. / zem2;
    const double zdenr = zr4 * zep1 + zr5 * zem1;
    const double zdent = zdenr;                    // zt4 = zr4, zt5 = zr5
    if (zdenr >= -eps && zdenr <= eps) {
        pref = eps;
        ptra = zem2;
    } else {
        pref = zw * (zr1 * zep1 - zr2 * zem1 - zr3 * zem2) / zdenr;
        ptra = zem2 - zem2 * zw * (zt1 * zep1 - zt2 * zem1 - zt3 * zep2) / zdent;
    }
    const double zemm = zem1 * zem1;
    const double zdend = 1. / ((1. - zbeta * zemm) * zrkg);
    prefd = zgamma2 * (1. - zemm) * zdend;
    ptrad = zrk2 * zem1 * zdend;
}

struct SwGenLayer { double refc, refdc, trac, tradc, dbtc, ref, refd, tra, trad, dbt; };

// layer assembly of spcvrt (:386-463, 474-548) for one (g, layer) cell
template <bool CLOUD>
__device__ __forceinline__ void sw_gen_layer(const double2 *__restrict__ tb, double bpade, double prmu0, double taur, double taug,
                                             const double *__restrict__ o, double pclfr, SwGenLayer &L)
{
    const double repclc = 1.e-12;
    const double ptauc = o[0], pomgc = o[1], pasyc = o[2], ptaua = o[3], pomga = o[4], pasya = o[5];
    double ztauc = taur + taug + ptaua;
    double zomcc = taur * 1.0 + ptaua * pomga;
    double zgcc = pasya * pomga * ptaua / zomcc;
    zomcc = zomcc / ztauc;
    double zf = zgcc * zgcc;
    double zwf = zomcc * zf;
    ztauc = (1.0 - zwf) * ztauc;
    zomcc = (zomcc - zwf) / (1.0 - zwf);
    zgcc = (zgcc - zf) / (1.0 - zf);
    sw_reftra_gen(tb, bpade, true, zgcc, prmu0, ztauc, zomcc, L.refc, L.refdc, L.trac, L.tradc);
    L.dbtc = sw_tbl_exp(tb, ztauc / prmu0, bpade);
    if (CLOUD) {
        const double ztauo = ztauc + ptauc;
        double zomco = ztauc * zomcc + ptauc * pomgc;
        const double zgco = (ptauc * pomgc * pasyc + ztauc * zomcc * zgcc) / zomco;
        zomco = zomco / ztauo;
        double refo, refdo, trao, trado;
        sw_reftra_gen(tb, bpade, pclfr > repclc, zgco, prmu0, ztauo, zomco, refo, refdo, trao, trado);
        const double zclear = 1.0 - pclfr, zcloud = pclfr;
        L.ref = zclear * L.refc + zcloud * refo;
        L.refd = zclear * L.refdc + zcloud * refdo;
        L.tra = zclear * L.trac + zcloud * trao;
        L.trad = zclear * L.tradc + zcloud * trado;
        const double zdbtmo = sw_tbl_exp(tb, ztauo / prmu0, bpade);
        L.dbt = zclear * L.dbtc + zcloud * zdbtmo;
    } else {
        L.ref = L.refc; L.refd = L.refdc; L.tra = L.trac; L.trad = L.tradc; L.dbt = L.dbtc;
    }
}

template <int LMAX, bool CLOUD>
__global__ void __launch_bounds__(SV_THREADS) sw_solver_gen_kernel(SwTables T, SwIn in, SwOut out, SwWork w)
{
    constexpr int NV = CLOUD ? 4 : 2;              // values per level: {up, down} x {total[, clear]}
    constexpr int NB = CLOUD ? 4 : 8;              // levels per reduction batch
    constexpr int NR = 32;                         // tile rows = NB * NV * SV_COLS
    static_assert(NB * NV * SV_COLS == NR, "tile geometry");
    constexpr int SG_S = 113;
    __shared__ double s_tile[NR * SG_S];
    __shared__ double s_part[NR * (SV_THREADS / NR + 1)];
    __shared__ double s_flux[SV_COLS][4][LMAX + 1];      // up, down, upc, downc
    const int klev = w.nlay;
    const int cb = threadIdx.x / NGPTSW;
    const int g = threadIdx.x - cb * NGPTSW;
    const int col = blockIdx.x * SV_COLS + cb;
    const size_t old = (size_t)out.ld;
    const bool incol = col < w.nc;
    const double prmu0 = incol ? in.coszen[col] : 0.0;
    const bool active = incol && !(prmu0 < ZEPZEN);
    const int colr = incol ? col : 0;
    const int band = c_ss.ngb[g];
    const double bpade = c_ss.bpade;
    const double2 *__restrict__ tb = reinterpret_cast<const double2 *>(T.exptbl);
    const double mu0 = active ? prmu0 : 1.0;
    const bool uvvis = band >= 9 && band <= 12;
    const double albd = uvvis ? in.asdif[colr] : in.aldif[colr];
    const double albp = uvvis ? in.asdir[colr] : in.aldir[colr];
    double zrup[LMAX + 1], zrupd[LMAX + 1];
    double zrupc[CLOUD ? LMAX + 1 : 1], zrupdc[CLOUD ? LMAX + 1 : 1];
    const double *__restrict__ taug = w.taug + (size_t)colr * klev * NGPTSW + g;
    const bool b24 = band == 8;
    const double raylg = b24 ? 1.0 : __ldg(T.tab + c_ss.rayl[band] + g - c_ss.g0[band]);
    const double *__restrict__ taur = b24 ? w.taur24 + (size_t)colr * klev * 8 + (g - c_ss.g0[band])
                                          : w.colmol + (size_t)colr * klev;
    const int trs = b24 ? 8 : 1;
    const double *__restrict__ opt = w.opt + ((size_t)colr * klev * 14 + band) * 6;
    const double *__restrict__ clfr = w.clfr + (size_t)colr * klev;
    const double zincflx = active ? in.adjflux * w.sfluxzen[(size_t)col * NGPTSW + g] * prmu0 : 0.0;

    // ---- up sweep (vrtqdr :103-121), both streams
    if (active) {
        double rup = albp, rupd = albd, rupc = albp, rupdc = albd;
        zrup[0] = rup; zrupd[0] = rupd;
        if (CLOUD) { zrupc[0] = rupc; zrupdc[0] = rupdc; }
        for (int l = 0; l < klev; ++l) {
            SwGenLayer L;
            sw_gen_layer<CLOUD>(tb, bpade, mu0, __ldg(taur + l * trs) * raylg, taug[(size_t)l * NGPTSW], opt + (size_t)l * 14 * 6,
                                CLOUD ? clfr[l] : 0.0, L);
            {
                const double zreflect = 1. / (1. - rupd * L.refd);
                const double rup_n = L.ref + (L.trad * ((L.tra - L.dbt) * rupd + L.dbt * rup)) * zreflect;
                const double rupd_n = L.refd + L.trad * L.trad * rupd * zreflect;
                rup = rup_n; rupd = rupd_n;
                zrup[l + 1] = rup; zrupd[l + 1] = rupd;
            }
            if (CLOUD) {
                const double zreflect = 1. / (1. - rupdc * L.refdc);
                const double rup_n = L.refc + (L.tradc * ((L.trac - L.dbtc) * rupdc + L.dbtc * rupc)) * zreflect;
                const double rupd_n = L.refdc + L.tradc * L.tradc * rupdc * zreflect;
                rupc = rup_n; rupdc = rupd_n;
                zrupc[l + 1] = rupc; zrupdc[l + 1] = rupdc;
            }
        }
    }
    // ---- down sweep (:125-150) and spectral sums (spcvrt :570-619)
    double ztdn = 1., zrdnd = 0., ztdbt = 1., ztdnc = 1., zrdndc = 0., ztdbtc = 1.;
    for (int k = 0; k <= klev; ++k) {
        const int s = klev - k;
        const int slot = k & (NB - 1);
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        if (active) {
            {
                const double ru = zrup[s], rud = zrupd[s];
                const double zreflect = 1. / (1. - zrdnd * rud);
                const double dif = ztdn - ztdbt;
                v[0] = zincflx * ((ztdbt * ru + dif * rud) * zreflect);
                v[1] = zincflx * (ztdbt + (dif + ztdbt * ru * zrdnd) * zreflect);
            }
            if (CLOUD) {
                const double ru = zrupc[s], rud = zrupdc[s];
                const double zreflect = 1. / (1. - zrdndc * rud);
                const double dif = ztdnc - ztdbtc;
                v[2] = zincflx * ((ztdbtc * ru + dif * rud) * zreflect);
                v[3] = zincflx * (ztdbtc + (dif + ztdbtc * ru * zrdndc) * zreflect);
            }
            if (s > 0) {
                const int l = s - 1;
                SwGenLayer L;
                sw_gen_layer<CLOUD>(tb, bpade, mu0, __ldg(taur + l * trs) * raylg, taug[(size_t)l * NGPTSW], opt + (size_t)l * 14 * 6,
                                    CLOUD ? clfr[l] : 0.0, L);
                {
                    const double zr = 1. / (1. - L.refd * zrdnd);
                    const double dif = ztdn - ztdbt;
                    const double ztdn_n = ztdbt * L.tra + (L.trad * (dif + ztdbt * L.ref * zrdnd)) * zr;
                    const double zrdnd_n = L.refd + L.trad * L.trad * zrdnd * zr;
                    ztdbt = L.dbt * ztdbt; ztdn = ztdn_n; zrdnd = zrdnd_n;
                }
                if (CLOUD) {
                    const double zr = 1. / (1. - L.refdc * zrdndc);
                    const double dif = ztdnc - ztdbtc;
                    const double ztdn_n = ztdbtc * L.trac + (L.tradc * (dif + ztdbtc * L.refc * zrdndc)) * zr;
                    const double zrdnd_n = L.refdc + L.tradc * L.tradc * zrdndc * zr;
                    ztdbtc = L.dbtc * ztdbtc; ztdnc = ztdn_n; zrdndc = zrdnd_n;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < NV; ++q) s_tile[((slot * NV + q) * SV_COLS + cb) * SG_S + g] = v[q];
        if (slot == NB - 1 || k == klev) {
            const double sum = tile_reduce<SV_THREADS, NR, NGPTSW, SG_S>(s_tile, s_part);
            if (threadIdx.x < NR) {
                const int c = threadIdx.x % SV_COLS, sq = threadIdx.x / SV_COLS;      // row = (slot*NV + q)*SV_COLS + c
                const int q = sq % NV, sl = sq / NV;
                const int kk = (k & ~(NB - 1)) + sl;
                if (kk <= k) s_flux[c][q][klev - kk] = sum;
            }
        }
    }
    __syncthreads();
    if (incol) {
        for (int lev = g; lev <= klev; lev += NGPTSW) {
            const double u = s_flux[cb][0][lev], d = s_flux[cb][1][lev];
            const double uc = CLOUD ? s_flux[cb][2][lev] : u, dc = CLOUD ? s_flux[cb][3][lev] : d;
            const size_t o = col + (size_t)lev * old;
            out.uflx[o] = u; out.dflx[o] = d; out.uflxc[o] = uc; out.dflxc[o] = dc;
        }
        for (int lay = g; lay < klev; lay += NGPTSW) {
            const size_t o = col + (size_t)lay * old;
            double h = 0.0, hc = 0.0;
            if (active && lay < klev - 1) {
                const double pdp = in.plev[col + (size_t)lay * in.ld] - in.plev[col + (size_t)(lay + 1) * in.ld];
                h = ((s_flux[cb][1][lay + 1] - s_flux[cb][0][lay + 1]) - (s_flux[cb][1][lay] - s_flux[cb][0][lay])) * (c_ss.heatfac / pdp);
                hc = CLOUD ? ((s_flux[cb][3][lay + 1] - s_flux[cb][2][lay + 1]) - (s_flux[cb][3][lay] - s_flux[cb][2][lay])) * (c_ss.heatfac / pdp) : h;
            }
            out.hr[o] = h;
            out.hrc[o] = hc;
        }
    }
}

#ifdef RRTMG_B200_DEV_VARIANTS
template <int LMAX, bool STORE, int OPT>
static void launch(const SwTables &t, const SwIn &in, const SwOut &out, SwWork &w, cudaStream_t s)
{
    const size_t pad = (size_t)g_tune.sw_solver_pad_kb * 1024;
    if (pad) cudaFuncSetAttribute(sw_solver_kernel<LMAX, STORE, OPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
    sw_solver_kernel<LMAX, STORE, OPT><<<(w.nc + SV_COLS - 1) / SV_COLS, SV_THREADS, pad, s>>>(t, in, out, w);
}
#endif
template <int LMAX>
static void launch_opt(const SwTables &t, const SwIn &in, const SwOut &out, SwWork &w, cudaStream_t s)
{
    // Default (variant 4): sw_solver_warp_kernel + sw_finish_kernel.  Development builds add 3: the same scheme in 7-warp
    // blocks (OPT 21); 2: bottom-up first, five stored values (OPT 13); 1: reftra recomputed in the second sweep, the
    // reference's recurrences literally (OPT 5); 0: the first version (OPT 0); 5, 6: the per-cell stack in an L2 scratch.
#ifdef RRTMG_B200_DEV_VARIANTS
    if (g_tune.sw_solver_store) { launch<LMAX, true, 0>(t, in, out, w, s); return; }
    if (g_tune.sw_solver_variant == 0) { launch<LMAX, false, 0>(t, in, out, w, s); return; }
    if (g_tune.sw_solver_variant == 1) { launch<LMAX, false, 5>(t, in, out, w, s); return; }
    if (g_tune.sw_solver_variant == 2) { launch<LMAX, false, 13>(t, in, out, w, s); return; }
    if (g_tune.sw_solver_variant == 3) { launch<LMAX, false, 21>(t, in, out, w, s); return; }
    if (g_tune.sw_solver_variant == 6 || g_tune.sw_solver_variant == 7) {
        constexpr int TC = LMAX <= 64 ? 32 : 16;
        const long long nitems = ((long long)w.nc * NGPTSW + 31) / 32;
        static int nsm6 = 0;
        if (!nsm6) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&nsm6, cudaDevAttrMultiProcessorCount, dev); }
        const int wps = g_tune.x[0] > 0 ? g_tune.x[0] : 20;
        auto go = [&](auto kern, int W) {
            const size_t smem = (size_t)(NTBL + 1) * sizeof(double2) + (size_t)W * 8 * SV_WS * sizeof(double);
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            const long long grid = std::min<long long>((nitems + W - 1) / W, (long long)nsm6);
            kern<<<(unsigned)grid, 32 * W, smem, s>>>(t, in, w, nitems, g_tune.x[1]);
        };
        if (g_tune.sw_solver_variant == 7) {        // table in shared memory, stack in local memory
            if (wps <= 20) go(sw_solver_sm_kernel<LMAX, 20, true>, 20);
            else if (wps <= 24) go(sw_solver_sm_kernel<LMAX, 24, true>, 24);
            else go(sw_solver_sm_kernel<LMAX, 28, true>, 28);
        }
        else if (wps <= 16) go(sw_solver_sm_kernel<LMAX, 16>, 16);
        else if (wps <= 20) go(sw_solver_sm_kernel<LMAX, 20>, 20);
        else if (wps <= 24) go(sw_solver_sm_kernel<LMAX, 24>, 24);
        else go(sw_solver_sm_kernel<LMAX, 28>, 28);
        sw_finish_kernel<LMAX, TC><<<(w.nc + TC - 1) / TC, 256, 0, s>>>(in, out, w);
        return;
    }
    if (g_tune.sw_solver_variant == 5) {
        constexpr int TC = LMAX <= 64 ? 32 : 16;
        const long long nitems = ((long long)w.nc * NGPTSW + 31) / 32;
        static int nsm = 0;
        if (!nsm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev); }
        const int wpb = g_tune.x[0] > 0 ? g_tune.x[0] : 20;
        const int flags = g_tune.x[1];
        auto go = [&](auto kern, int W) {
            // shared levels: what fits beside the static tiles at W blocks per SM (1 KB per block is reserved by the system)
            const int fixed = (int)(sizeof(double) * (8 * SV_WS + 2 * (LMAX + 1))) + 1024 + 256;
            int ns = (227 * 1024 / W - fixed) / 768;
            if (g_tune.x[2] > 0 && g_tune.x[2] - 1 < ns) ns = g_tune.x[2] - 1;      // x2 = shared levels + 1 (0: as many as fit)
            ns = ns < 0 ? 0 : (ns > w.nlay + 1 ? w.nlay + 1 : ns);
            const size_t smem = (size_t)ns * 768;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            const long long grid = std::min<long long>(nitems, (long long)nsm * W);
            kern<<<(unsigned)grid, 32, smem, s>>>(t, in, w, nitems, ns, flags);
        };
        if (wpb <= 12) go(sw_solver_l2_kernel<LMAX, 12>, 12);
        else if (wpb <= 16) go(sw_solver_l2_kernel<LMAX, 16>, 16);
        else if (wpb <= 20) go(sw_solver_l2_kernel<LMAX, 20>, 20);
        else if (wpb <= 24) go(sw_solver_l2_kernel<LMAX, 24>, 24);
        else go(sw_solver_l2_kernel<LMAX, 28>, 28);
        sw_finish_kernel<LMAX, TC><<<(w.nc + TC - 1) / TC, 256, 0, s>>>(in, out, w);
        return;
    }
#endif
    {
        // variant 4: variant 3 with one warp per block + sw_finish_kernel
        constexpr int TC = LMAX <= 64 ? 32 : 16;
        const long long nthr = (long long)w.nc * NGPTSW;
        sw_solver_warp_kernel<LMAX, 28><<<(unsigned)((nthr + 31) / 32), 32, 0, s>>>(t, in, w);
        sw_finish_kernel<LMAX, TC><<<(w.nc + TC - 1) / TC, 256, 0, s>>>(in, out, w);
    }
}

int sw_launch_solver(const SwTables &t, const SwIn &in, const SwOut &out, SwWork &w, cudaStream_t s)
{
    if (w.opt) {            // aerosols and/or clouds: the general kernel
        const unsigned nb = (w.nc + SV_COLS - 1) / SV_COLS;
        const bool cloud = in.icld >= 1;                // iaer = 6 / 10: aerosol properties per band in w.opt
        if (w.nlay <= 64) {
            if (cloud) sw_solver_gen_kernel<64, true><<<nb, SV_THREADS, 0, s>>>(t, in, out, w);
            else sw_solver_gen_kernel<64, false><<<nb, SV_THREADS, 0, s>>>(t, in, out, w);
        } else {
            if (cloud) sw_solver_gen_kernel<MAXLAY, true><<<nb, SV_THREADS, 0, s>>>(t, in, out, w);
            else sw_solver_gen_kernel<MAXLAY, false><<<nb, SV_THREADS, 0, s>>>(t, in, out, w);
        }
        return 1;
    }
    if (w.nlay <= 64) launch_opt<64>(t, in, out, w, s);
    else launch_opt<MAXLAY>(t, in, out, w, s);
#ifdef RRTMG_B200_DEV_VARIANTS
    return (!g_tune.sw_solver_store && g_tune.sw_solver_variant >= 4) ? 2 : 1;
#else
    return 2;
#endif
}

} // namespace rrtmg
