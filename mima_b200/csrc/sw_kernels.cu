// sw_kernels.cu -- RRTMG shortwave on sm_100a: prep (inatm_sw + setcoef_sw), the staged taumol_sw kernel, cloud / aerosol
// optics.  The fused clear-sky kernel is in sw_column.cu, the staged two-stream solvers (spcvrt + reftra + vrtqdr) in
// sw_solver.cu; the per-cell arithmetic they share in sw_bands.cuh and sw_twostream.cuh.
//
//   sw_prep_cell_kernel  thread <-> (column, layer): the cell's reference-pressure index and "below 100 hPa" flag; for the
//                        column kernel the whole setcoef state into the tile-major field;
//   sw_prep_kernel       thread <-> column: night marker, laytrop, and per band the layer whose binary-species parameter
//                        selects the solar source (laysolfr), replaying the sequential loops of taumol16..29;
//   sw_taumol_kernel     (staged path: clouds, aerosols, stage capture) thread <-> (column, layer) cell as in the LW kernel:
//                        taug per g-point, the Rayleigh descriptors (colmol; taur of band 24) and, from the laysolfr layer,
//                        sfluxzen;
//   sw_optics_kernel     clouds / aerosols of the general path (not MiMA's configuration).
// Night columns (coszen < 1e-10) are skipped and written as zeros (rad.nomcica:502-510).
// Compiled with -fmad=false: fused multiply-adds appear only where written as fma().
#include "sw_bands.cuh"

namespace rrtmg {

__constant__ unsigned char c_sw_ngb[NGPTSW];

int sw_upload_const(const SwConst &c)
{
    unsigned char ngb[NGPTSW];
    for (int b = 0; b < NBNDSW; ++b)
        for (int i = 0; i < c.band[b].ng; ++i) ngb[c.band[b].g0 + i] = (unsigned char)b;
    if (cudaMemcpyToSymbol(c_sw, &c, sizeof(SwConst)) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(c_sw_ngb, ngb, sizeof ngb) != cudaSuccess) return -1;
    if (sw_column_upload_const(c)) return -1;
    return sw_solver_upload_const(c, ngb);
}

// =====================================================================================================
// prep: per column -- night marker, laytrop (setcoef.f90), and the layer whose binary-species parameter
//       selects the solar source of each band (SW/src/rrtmg_sw_taumol.f90, "laysolfr" logic).
//       With w.f != nullptr (stage capture, test hook) the per-cell setcoef state is also written out.
// =====================================================================================================
// Stage 1, thread <-> (column, layer): the cell's reference-pressure index and "below 100 hPa" flag (setcoef),
// which is all the column logic below needs; stage capture also dumps the full setcoef state here.
__global__ void __launch_bounds__(128) sw_prep_cell_kernel(SwIn in, SwWork w)
{
    const int nc = w.nc, nlay = w.nlay;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nc * nlay) return;
    const int l = (int)(i / nc);
    const int col = (int)(i - (size_t)l * nc);
    // night column: nothing downstream reads it -- except the fused column kernel, whose warps carry the night lanes of a
    // partly sunlit tile along (results discarded) and need valid table indices for them
    if (!w.fused && in.coszen[col] < ZEPZEN) return;
    SwPair p;
    const bool lower = sw_cell(in, col, l, p);
    w.cs_jp[i] = (unsigned char)(p.jp | (lower ? 0x80 : 0));
    if (w.fused) {
        double *f = w.tf + w.tfld(l, col);
#define TF(k) f[(k) * 32]
        TF(SF_COUNT) = __hiloint2double(0, (int)sw_pack(p.jp, p.jt, p.jt1, p.inds, p.indf));
        TF(SF_FAC00) = p.fac00; TF(SF_FAC01) = p.fac01; TF(SF_FAC10) = p.fac10; TF(SF_FAC11) = p.fac11;
        TF(SF_COLH2O) = p.colh2o; TF(SF_COLCO2) = p.colco2; TF(SF_COLO3) = p.colo3;
        TF(SF_COLCH4) = p.colch4; TF(SF_COLO2) = p.colo2; TF(SF_COLMOL) = p.colmol; TF(SF_COLN2O) = p.coln2o;
        TF(SF_SELFFAC) = p.selffac; TF(SF_SELFFRAC) = p.selffrac; TF(SF_FORFAC) = p.forfac; TF(SF_FORFRAC) = p.forfrac;
#undef TF
    }
    if (w.f) {
        const size_t wo = i;
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

// Stage 2, thread <-> column.
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
        const unsigned char v = w.cs_jp[(size_t)l * nc + col];
        if (v & 0x80) laytrop = laytrop + 1;
        jpv[l + 1] = (unsigned char)(v & 0x7f);
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
constexpr int TM_STRIDE = 18;

template <int NG>
struct BandAcc {
    double t[NG];
    const double *__restrict__ tab;
    double *t24;                     // band 24 only: this cell's slot in taur24
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
    // Rayleigh optical depth is not materialised per g-point for the bands where taur(g) = colmol * rayl(g)
    // with one table row rayl per band: the solver evaluates that.  Band 24's coefficient depends on the
    // cell (eta-interpolated below laytrop) and goes to the small side array taur24[col][lay][8].
    __device__ __forceinline__ void rayl1(int off, double wgt)
    {
        if (t24) {
#pragma unroll
            for (int g = 0; g < (NG < 8 ? NG : 8); ++g) t24[g] = wgt * __ldg(tab + off + g);
        }
    }
    __device__ __forceinline__ void rayl2(int o0, double w0, int o1, double w1)
    {
        if (t24) {
#pragma unroll
            for (int g = 0; g < (NG < 8 ? NG : 8); ++g) t24[g] = fma(w1, __ldg(tab + o1 + g), w0 * __ldg(tab + o0 + g));
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

template <int BAND>
__device__ __forceinline__ void sw_band(const SwTables &T, const SwPair &p, bool valid, bool lower, int lay1,
                                        const int *__restrict__ laysolfr, double *slab,
                                        double *__restrict__ taug, double *t24, double *sflx_col,
                                        size_t cell0, size_t colstride, unsigned vmask)
{
    constexpr int NG = sw_ng(BAND);
    const int lane = threadIdx.x & 31;
    const SwBand &B = c_sw.band[BAND];
    const int g0 = B.g0;
    if (valid) {
        BandAcc<NG> pw;
        pw.tab = T.tab + B.base;
        pw.t24 = BAND == 8 ? t24 : nullptr;
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
            *reinterpret_cast<double2 *>(taug + cell0 + (size_t)c * colstride + g0 + 2 * j) = a;
        }
    }
    __syncwarp();
}

// As in the LW kernel: 16 warps per block step through the bands together (instruction-cache reuse);
// work items are linearised (32-column tile, layer) pairs.
constexpr int TM_BLOCK_WARPS = 4;
__global__ void __launch_bounds__(32 * TM_BLOCK_WARPS, 4) sw_taumol_kernel(SwTables T, SwIn in, SwWork w, int g_tm_sync)
{
    extern __shared__ __align__(16) double s_dyn[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nlay = w.nlay, nc = w.nc;
    const int ntile = (nc + 31) / 32;
    const long item = (long)blockIdx.x * TM_BLOCK_WARPS + wid;
    const bool live = item < (long)ntile * nlay;
    const int tile = live ? (int)(item / nlay) : 0;
    const int lay = live ? (int)(item - (long)tile * nlay) : 0;
    const int c0 = tile * 32;
    const int col = c0 + lane;
    bool valid = live && col < nc;
    int laytrop = 0;
    if (valid) {
        laytrop = w.laytrop[col];
        if (laytrop < 0) valid = false;              // night column: the solver never reads its staging
    }
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    SwPair p;
    bool lower = false;
    if (valid) {
        sw_cell(in, col, lay, p);
        lower = (lay + 1) <= laytrop;
    }
    double *slab = s_dyn + (size_t)wid * (32 * TM_STRIDE);
    double *t24 = w.taur24 + ((size_t)(valid ? col : 0) * nlay + lay) * 8;
    if (valid) w.colmol[(size_t)col * nlay + lay] = p.colmol;
    const int *ls = w.laysolfr + (size_t)(valid ? col : 0) * 14;
    double *sflx = w.sfluxzen + (size_t)(valid ? col : 0) * NGPTSW;
    const size_t colstride = (size_t)nlay * NGPTSW;
    const size_t cell0 = ((size_t)c0 * nlay + lay) * NGPTSW;
#define SW_BAND(b) sw_band<b>(T, p, valid, lower, lay + 1, ls, slab, w.taug, t24, sflx, cell0, colstride, vmask); if (((b) & (g_tm_sync - 1)) == g_tm_sync - 1) __syncthreads()
    SW_BAND(0); SW_BAND(1); SW_BAND(2); SW_BAND(3); SW_BAND(4); SW_BAND(5); SW_BAND(6);
    SW_BAND(7); SW_BAND(8); SW_BAND(9); SW_BAND(10); SW_BAND(11); SW_BAND(12); SW_BAND(13);
#undef SW_BAND
}

// Test hook (stage capture): expand the Rayleigh descriptors to taur[col][lay][112] exactly as the solver
// evaluates them.
__global__ void sw_expand_taur_kernel(SwTables T, SwWork w)
{
    const size_t n = (size_t)w.nc * w.nlay * NGPTSW;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int g = (int)(i % NGPTSW);
        const size_t cell = i / NGPTSW;
        const int col = (int)(cell / w.nlay);
        const int band = c_sw_ngb[g];
        const SwBand &B = c_sw.band[band];
        double tr = 0.0;
        if (w.laytrop[col] >= 0) {
            if (band == 8) tr = w.taur24[cell * 8 + g - B.g0];
            else tr = w.colmol[cell] * T.tab[B.base + B.sec[SS_RAYL] * B.rs + g - B.g0];
        }
        w.taur[i] = tr;
    }
}

// General path only: cloud optical properties through cldprop_sw's inflag = 0 branch (delta-M scaling with the forward
// scattering fraction, SW/src/rrtmg_sw_cldprop.f90:120-166), aerosol properties as given (iaer = 10,
// rad.nomcica:633-640), transposed to [col][lay][band][6] so that the g-point lanes of a band read one address.
// Also the `stop 'PARTIAL CLOUD NOT ALLOWED'` test of rad.nomcica:534-539 (flag, checked by the host).
__device__ SwCldConst d_swcld;          // 55 KB: global memory, read through the read-only path
int sw_upload_cld(const SwCldConst &c) { return cudaMemcpyToSymbol(d_swcld, &c, sizeof c) == cudaSuccess ? 0 : -1; }

// cldprop_sw, inflag = 2, one (layer, band) (rrtmg_sw_cldprop.f90:168-345): ice option 1 (Ebert and Curry), 2 (Streamer),
// 3 (Fu), liquid option 1 (Hu and Stamnes), then the delta-scaled combination.  Returns 0 or the number of the Fortran
// `stop`: 1 ICE RADIUS OUT OF BOUNDS, 2 ICE GENERALIZED EFFECTIVE SIZE OUT OF BOUNDS, 3 LIQUID EFFECTIVE RADIUS OUT OF
// BOUNDS, 4 a range check on an interpolated property.
__device__ int sw_cldprop_band(int ib, int iceflag, double ciwp, double clwp, double radice, double radliq,
                               double &tauc, double &omgc, double &asyc)
{
    const SwCldConst &K = d_swcld;
    const double eps = 1.e-06, cldmin = 1.e-20;
    double extcoice = 0., ssacoice = 0., gice = 0., forwice = 0., extcoliq = 0., ssacoliq = 0., gliq = 0., forwliq = 0.;
    if (ciwp == 0.0) {
    } else if (iceflag == 1) {
        if (radice < 13.0 || radice > 130.) return 1;
        const double wavenum2[14] = {3250., 4000., 4650., 5150., 6150., 7700., 8050., 12850., 16000., 22650., 29000., 38000., 50000., 2600.};
        int icx = 4;
        if (wavenum2[ib] > 1.43e04) icx = 0;
        else if (wavenum2[ib] > 7.7e03) icx = 1;
        else if (wavenum2[ib] > 5.3e03) icx = 2;
        else if (wavenum2[ib] > 4.0e03) icx = 3;
        extcoice = K.abari[icx] + K.bbari[icx] / radice;
        ssacoice = 1. - K.cbari[icx] - K.dbari[icx] * radice;
        gice = K.ebari[icx] + K.fbari[icx] * radice;
        if (gice >= 1.0) gice = 1.0 - eps;
        forwice = gice * gice;
        if (extcoice < 0.0 || ssacoice > 1.0 || ssacoice < 0.0 || gice > 1.0 || gice < 0.0) return 4;
    } else if (iceflag == 2) {
        if (radice < 5.0 || radice > 131.0) return 1;
        const double factor = (radice - 2.) / 3.;
        int index = (int)factor;
        if (index == 43) index = 42;
        const double fint = factor - (double)index;
        const int o = (index - 1) + 43 * ib;
        extcoice = K.extice2[o] + fint * (K.extice2[o + 1] - K.extice2[o]);
        ssacoice = K.ssaice2[o] + fint * (K.ssaice2[o + 1] - K.ssaice2[o]);
        gice = K.asyice2[o] + fint * (K.asyice2[o + 1] - K.asyice2[o]);
        forwice = gice * gice;
        if (extcoice < 0.0 || ssacoice > 1.0 || ssacoice < 0.0 || gice > 1.0 || gice < 0.0) return 4;
    } else {
        if (radice < 5.0 || radice > 140.0) return 2;
        const double factor = (radice - 2.) / 3.;
        int index = (int)factor;
        if (index == 46) index = 45;
        const double fint = factor - (double)index;
        const int o = (index - 1) + 46 * ib;
        extcoice = K.extice3[o] + fint * (K.extice3[o + 1] - K.extice3[o]);
        ssacoice = K.ssaice3[o] + fint * (K.ssaice3[o + 1] - K.ssaice3[o]);
        gice = K.asyice3[o] + fint * (K.asyice3[o + 1] - K.asyice3[o]);
        const double fdelta = K.fdlice3[o] + fint * (K.fdlice3[o + 1] - K.fdlice3[o]);
        if (fdelta < 0.0 || fdelta > 1.0) return 4;
        forwice = fdelta + 0.5 / ssacoice;
        if (forwice > gice) forwice = gice;
        if (extcoice < 0.0 || ssacoice > 1.0 || ssacoice < 0.0 || gice > 1.0 || gice < 0.0) return 4;
    }
    if (clwp != 0.0) {
        if (radliq < 2.5 || radliq > 60.) return 3;
        int index = (int)(radliq - 1.5);
        if (index == 0) index = 1;
        if (index == 58) index = 57;
        const double fint = radliq - 1.5 - (double)index;
        const int o = (index - 1) + 58 * ib;
        extcoliq = K.extliq1[o] + fint * (K.extliq1[o + 1] - K.extliq1[o]);
        ssacoliq = K.ssaliq1[o] + fint * (K.ssaliq1[o + 1] - K.ssaliq1[o]);
        if (fint < 0. && ssacoliq > 1.) ssacoliq = K.ssaliq1[o];
        gliq = K.asyliq1[o] + fint * (K.asyliq1[o + 1] - K.asyliq1[o]);
        forwliq = gliq * gliq;
        if (extcoliq < 0.0 || ssacoliq > 1.0 || ssacoliq < 0.0 || gliq > 1.0 || gliq < 0.0) return 4;
    }
    const double tauliqorig = clwp * extcoliq;
    const double tauiceorig = ciwp * extcoice;
    const double ssaliq = ssacoliq * (1.0 - forwliq) / (1.0 - forwliq * ssacoliq);
    const double tauliq = (1.0 - forwliq * ssacoliq) * tauliqorig;
    const double ssaice = ssacoice * (1.0 - forwice) / (1.0 - forwice * ssacoice);
    const double tauice = (1.0 - forwice * ssacoice) * tauiceorig;
    const double scatliq = ssaliq * tauliq;
    double scatice = ssaice * tauice;
    double taucloud = tauliq + tauice;
    if (taucloud == 0.0) taucloud = cldmin;
    if (scatice == 0.0) scatice = cldmin;
    tauc = taucloud;
    omgc = (scatliq + scatice) / taucloud;
    if (iceflag == 3)
        asyc = (1.0 / (scatliq + scatice)) * (scatliq * (gliq - forwliq) / (1.0 - forwliq) + scatice * ((gice - forwice) / (1.0 - forwice)));
    else
        asyc = (scatliq * (gliq - forwliq) / (1.0 - forwliq) + scatice * (gice - forwice) / (1.0 - forwice)) / (scatliq + scatice);
    return 0;
}

__global__ void __launch_bounds__(128) sw_optics_kernel(SwIn in, SwWork w)
{
    const int nc = w.nc, nlay = w.nlay;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nc * nlay) return;
    const int l = (int)(i / nc);
    const int col = (int)(i - (size_t)l * nc);
    const size_t ld = (size_t)in.ld;
    if (in.coszen[col] < ZEPZEN) return;      // night column: skipped before the cloud tests in the reference too (rad.nomcica:497-505)
    const double cldmin = 1.e-20, zepsec = 1.e-06;
    double cf = 0.0;
    double tauctot = 0.0;
    if (in.icld >= 1) {
        cf = in.cldfr[col + (size_t)l * ld];
        if (cf > zepsec && cf < 1.0 - zepsec) atomicOr(w.err, 1);
        if (in.taucld)
            for (int ib = 0; ib < 14; ++ib) tauctot = tauctot + in.taucld[ib + 14 * (col + (size_t)l * ld)];
    }
    const size_t ol = col + (size_t)l * ld;
    const bool wp = in.icld >= 1 && in.inflg == 2;
    const double ciwp = wp ? in.cicewp[ol] : 0.0, clwp = wp ? in.cliqwp[ol] : 0.0;
    const bool cloudy = in.icld >= 1 && cf >= cldmin && ((ciwp + clwp) >= cldmin || tauctot >= cldmin);
    w.clfr[(size_t)col * nlay + l] = cf;
    double *o = w.opt + ((size_t)col * nlay + l) * 14 * 6;
    for (int ib = 0; ib < 14; ++ib) {
        double tauc = 0.0, omgc = 1.0, asyc = 0.0;
        if (cloudy && wp) {
            const int stop = sw_cldprop_band(ib, in.iceflg, ciwp, clwp, in.reice[ol], in.reliq[ol], tauc, omgc, asyc);
            if (stop) atomicMax(w.err + 1, stop);
        } else if (cloudy) {
            const size_t q = ib + 14 * (col + (size_t)l * ld);
            const double taucldorig_a = in.taucld[q];
            const double ffp = in.fsfcld[q];
            const double ffp1 = 1.0 - ffp;
            const double ffpssa = 1.0 - ffp * in.ssacld[q];
            omgc = ffp1 * in.ssacld[q] / ffpssa;
            tauc = ffpssa * taucldorig_a;
            asyc = (in.asmcld[q] - ffp) / (ffp1);
        }
        double taua = 0.0, omga = 1.0, asya = 0.0;
        if (in.iaer == 10) {
            const size_t q = col + ld * (l + (size_t)nlay * ib);
            taua = in.tauaer[q]; asya = in.asmaer[q]; omga = in.ssaaer[q];
        } else if (in.iaer == 6) {            // six ECMWF aerosol types (rad.nomcica:608-640)
            taua = 0.0; asya = 0.0; omga = 0.0;
            for (int ia = 0; ia < 6; ++ia) {
                const double e = in.ecaer[col + ld * (l + (size_t)nlay * ia)];
                taua = taua + c_sw.rsrtaua[ib][ia] * e;
                omga = omga + c_sw.rsrtaua[ib][ia] * e * c_sw.rsrpiza[ib][ia];
                asya = asya + c_sw.rsrtaua[ib][ia] * e * c_sw.rsrpiza[ib][ia] * c_sw.rsrasya[ib][ia];
            }
            if (taua == 0.0) {
                asya = 0.0; omga = 1.0;
            } else {
                if (omga != 0.0) asya = asya / omga;
                omga = omga / taua;
            }
        }
        o[ib * 6 + 0] = tauc; o[ib * 6 + 1] = omgc; o[ib * 6 + 2] = asyc;
        o[ib * 6 + 3] = taua; o[ib * 6 + 4] = omga; o[ib * 6 + 5] = asya;
    }
}

int sw_run_pass(const SwTables &t, const SwIn &in, const SwOut &out, SwWork &w, cudaStream_t s)
{
    int extra = 0;
    // no clouds, no aerosols, no stage capture: taumol and the two-stream solver fused per column (sw_column.cu)
    w.fused = g_tune.sw_fused && !w.opt && !w.f && w.tf;
    if (w.opt) { sw_optics_kernel<<<(unsigned)(((size_t)w.nc * w.nlay + 127) / 128), 128, 0, s>>>(in, w); extra = 1; }
    ktimer_begin(K_SW_PREP, s);
    sw_prep_cell_kernel<<<(unsigned)(((size_t)w.nc * w.nlay + 127) / 128), 128, 0, s>>>(in, w);
    sw_prep_kernel<<<(w.nc + 127) / 128, 128, 0, s>>>(in, w);
    ktimer_end(s);
    if (w.fused) {
        ktimer_begin(K_SW_COLUMN, s);
        const int n = sw_launch_column(t, in, out, w, s);
        ktimer_end(s);
        return 2 + n;
    }
    {
        const long items = (long)((w.nc + 31) / 32) * w.nlay;
        const size_t smem = (size_t)TM_BLOCK_WARPS * 32 * TM_STRIDE * sizeof(double);
        cudaFuncSetAttribute(sw_taumol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ktimer_begin(K_SW_TAUMOL, s);
        sw_taumol_kernel<<<(unsigned)((items + TM_BLOCK_WARPS - 1) / TM_BLOCK_WARPS), 32 * TM_BLOCK_WARPS, smem, s>>>(t, in, w, g_tune.taumol_sync > 0 ? g_tune.taumol_sync : 1);
        ktimer_end(s);
    }
    if (w.taur) sw_expand_taur_kernel<<<1184, 256, 0, s>>>(t, w);
    ktimer_begin(K_SW_SOLVER, s);
    const int nsv = sw_launch_solver(t, in, out, w, s);
    ktimer_end(s);
    return 3 + nsv + extra;
}

} // namespace rrtmg
