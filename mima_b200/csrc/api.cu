// api.cu -- host side of librrtmg_b200.so: coefficient registry, rrtmg_*_ini (g-point reduction, lookup
// tables, device table packing), workspace management, column chunking and the extern "C" ABI declared
// in include/rrtmg_b200.h.  No compute happens on the host; if CUDA is unavailable every entry point
// fails with RRTMG_B200_ERR_CUDA.
//
// Reference for the init path: LW/src/rrtmg_lw_init.f90:28-175 (+ :178-281 lwdatinit, :284-363 lwcmbdat,
// :366-2659 cmbgb1..16), SW/src/rrtmg_sw_init.f90:28-154 (+ :157-241, :244-367, :473-1516).
#include "../../include/rrtmg_b200.h"
#include "rrtmg_dev.cuh"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

using namespace rrtmg;

namespace {

struct HostArr {
    std::vector<int> dims;
    std::vector<double> data;   // column-major
    long size() const { long n = 1; for (int d : dims) n *= d; return n; }
};

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int ensure(size_t n)
    {
        if (n <= bytes) return 0;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        if (cudaMalloc(&p, n) != cudaSuccess) return -1;
        bytes = n;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

struct State {
    std::mutex mu;
    std::map<std::string, HostArr> reg;          // registered original (16-g) arrays
    std::map<std::string, HostArr> reduced;      // reduced arrays in Fortran order (test hook)
    std::string err;
    long launches = 0;
    int chunk = 0;
    int host_chunk = 0;
    int host_slots = 4;       // slots of the host-pointer pipelines (2..MAXSLOT)
    int run_chunk = 0;
    bool capture = false;
    // LW
    bool lw_ready = false;
    LwConst lwc;
    LwTables lwt{};
    DevBuf lw_tab, lw_slices, lw_totplnk, lw_exptfn, lw_work, lw_cap, lw_err;
    bool lw_have_cld = false;
    LwWork lw_last{};
    int lw_last_ncol = 0;
    // SW
    bool sw_ready = false;
    SwConst swc;
    SwTables swt{};
    DevBuf sw_tab, sw_slices, sw_exptbl, sw_work, sw_err;
    bool sw_have_cld = false;
    SwWork sw_last{};
    int sw_last_ncol = 0;
    // option "share_inputs": rrtmg_b200_sw keeps its device copies of the eleven arrays both codes read (play, plev, tlay,
    // tlev, tsfc, h2o, o3, co2, ch4, n2o, o2) as full-batch arrays, and the rrtmg_b200_lw call that follows it with the
    // same host pointers and sizes uses them instead of uploading again (run_rrtmg passes the same arrays to both,
    // rrtm_radiation.f90:686-712 and 722-748).  One shot: any other call forgets the copies.
    bool share_inputs = false;
    struct Shared {
        bool valid = false;
        int ncol = 0, nlay = 0;
        const double *host[11] = {};
        double *dev[11] = {};
        DevBuf buf;
    } shared;
};
State G;

// ---- per-kernel event timing (off by default; bench.py turns it on for the roofline numbers)
struct KTimer {
    bool on = false;
    struct Rec { int id; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    double ms[K_COUNT] = {};
    long n[K_COUNT] = {};
    cudaEvent_t get()
    {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    void collect()
    {
        for (Rec &r : recs) {
            cudaEventSynchronize(r.b);
            float t = 0.f;
            if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms[r.id] += t; n[r.id] += 1; }
            pool.push_back(r.a);
            pool.push_back(r.b);
        }
        recs.clear();
    }
};
KTimer KT;

int fail(int code, const std::string &msg)
{
    G.err = msg;
    return code;
}
#define CUDA_OK(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(RRTMG_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// ------------------------------------------------------------------------------------------------
// g-point reduction.  Each band's 16 original g-points are merged into ngc groups of consecutive
// points (group sizes ngn); absorption-like quantities are averaged with the Gaussian weights wt
// normalised inside the group, Planck fractions and the solar source are summed.
// ------------------------------------------------------------------------------------------------
const double kWt[16] = {0.1527534276, 0.1491729617, 0.1420961469, 0.1316886544, 0.1181945205, 0.1019300893,
                        0.0832767040, 0.0626720116, 0.0424925000, 0.0046269894, 0.0038279891, 0.0030260086,
                        0.0022199750, 0.0014140010, 0.0005330000, 0.0000750000};
const int kLwNgc[16] = {10, 12, 16, 14, 16, 8, 12, 8, 12, 6, 8, 8, 4, 2, 2, 2};
const std::vector<std::vector<int>> kLwNgn = {
    {1, 1, 2, 2, 2, 2, 2, 2, 1, 1}, {1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2},
    {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1}, {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 3},
    {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1}, {2, 2, 2, 2, 2, 2, 2, 2},
    {2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2}, {2, 2, 2, 2, 2, 2, 2, 2},
    {1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2}, {2, 2, 2, 2, 4, 4},
    {1, 1, 2, 2, 2, 2, 3, 3}, {1, 1, 1, 1, 2, 2, 4, 4},
    {3, 3, 4, 6}, {8, 8}, {8, 8}, {4, 12}};
const int kSwNgc[14] = {6, 12, 8, 8, 10, 10, 2, 10, 8, 6, 6, 8, 6, 12};
const std::vector<std::vector<int>> kSwNgn = {
    {2, 2, 2, 2, 4, 4}, {1, 1, 1, 1, 1, 2, 1, 2, 1, 2, 1, 2}, {1, 1, 1, 1, 2, 2, 4, 4}, {1, 1, 1, 1, 2, 2, 4, 4},
    {1, 1, 1, 1, 1, 1, 1, 1, 2, 6}, {1, 1, 1, 1, 1, 1, 1, 1, 2, 6}, {8, 8}, {2, 2, 1, 1, 1, 1, 1, 1, 2, 4},
    {2, 2, 2, 2, 2, 2, 2, 2}, {1, 1, 2, 2, 4, 6}, {1, 1, 2, 2, 4, 6}, {1, 1, 1, 1, 1, 1, 4, 6},
    {1, 1, 2, 2, 4, 6}, {1, 1, 1, 1, 2, 2, 2, 2, 1, 1, 1, 1}};

struct GMap {
    int ngc;
    int first[16], count[16];
    double rw[16];
};
GMap make_gmap(int ngc, const std::vector<int> &ngn)
{
    GMap m;
    m.ngc = ngc;
    int pos = 0;
    for (int k = 0; k < ngc; ++k) {
        m.first[k] = pos;
        m.count[k] = ngn[k];
        double wsum = 0.0;
        for (int i = 0; i < ngn[k]; ++i) wsum = wsum + kWt[pos + i];
        for (int i = 0; i < ngn[k]; ++i) m.rw[pos + i] = (ngc < 16) ? kWt[pos + i] / wsum : 1.0;
        pos += ngn[k];
    }
    return m;
}

enum class GAxis { First, Last };
// Reduce the 16-wide g axis; result keeps the Fortran layout with 16 replaced by ngc.
HostArr reduce_g(const HostArr &a, const GMap &m, GAxis axis, bool weighted)
{
    HostArr r;
    r.dims = a.dims;
    const long total = a.size();
    const long other = total / 16;
    if (axis == GAxis::Last) {
        r.dims.back() = m.ngc;
        r.data.assign((size_t)other * m.ngc, 0.0);
        for (int k = 0; k < m.ngc; ++k)
            for (long e = 0; e < other; ++e) {
                double s = 0.0;
                for (int i = m.first[k]; i < m.first[k] + m.count[k]; ++i) {
                    const double v = a.data[(size_t)i * other + e];
                    s = weighted ? s + v * m.rw[i] : s + v;
                }
                r.data[(size_t)k * other + e] = s;
            }
    } else {
        r.dims.front() = m.ngc;
        r.data.assign((size_t)other * m.ngc, 0.0);
        for (long o = 0; o < other; ++o)
            for (int k = 0; k < m.ngc; ++k) {
                double s = 0.0;
                for (int i = m.first[k]; i < m.first[k] + m.count[k]; ++i) {
                    const double v = a.data[(size_t)o * 16 + i];
                    s = weighted ? s + v * m.rw[i] : s + v;
                }
                r.data[(size_t)o * m.ngc + k] = s;
            }
    }
    return r;
}

const HostArr *find(const std::string &name)
{
    auto it = G.reg.find(name);
    return it == G.reg.end() ? nullptr : &it->second;
}

// Reduce registry array "<pfx>.<oname>" and remember it as "<pfx>.<rname>"; nullptr if absent.
const HostArr *reduce_named(const std::string &pfx, const char *oname, const char *rname, const GMap &m)
{
    const HostArr *a = find(pfx + "." + oname);
    if (!a) return nullptr;
    const std::string on(oname);
    const bool plain = on == "fracrefao" || on == "fracrefbo" || on == "sfluxrefo";
    const bool gfirst = plain || on == "raylao" || a->dims.size() == 1;
    if ((gfirst ? a->dims.front() : a->dims.back()) != 16) return nullptr;
    HostArr r = reduce_g(*a, m, gfirst ? GAxis::First : GAxis::Last, !plain);
    auto &slot = G.reduced[pfx + "." + rname];
    slot = std::move(r);
    return &slot;
}

// Append a reduced (rows, ng) Fortran array (g last) or (ng, rows) (g first) to the band table as
// [row][rs] (rs >= ng, zero padded); returns the first row index.
int pow2_at_least(int n) { int r = 2; while (r < n) r *= 2; return r; }
int append_rows(std::vector<double> &tab, int base, int ng, const HostArr *a, bool gfirst)
{
    const int rs = pow2_at_least(ng);
    const int row0 = (int)((tab.size() - base) / rs);
    if (!a) return -1;
    const long rows = a->size() / ng;
    const size_t o = tab.size();
    tab.resize(o + (size_t)rows * rs, 0.0);
    for (long r = 0; r < rows; ++r)
        for (int ig = 0; ig < ng; ++ig)
            tab[o + (size_t)r * rs + ig] = gfirst ? a->data[(size_t)r * ng + ig] : a->data[(size_t)ig * rows + r];
    return row0;
}
int append_const_row(std::vector<double> &tab, int base, int ng, const double *vals)
{
    const int rs = pow2_at_least(ng);
    const int row0 = (int)((tab.size() - base) / rs);
    for (int ig = 0; ig < rs; ++ig) tab.push_back(ig < ng ? vals[ig] : 0.0);
    return row0;
}

// The per-task slices of the band tables for the column kernels (ColSlices, rrtmg_dev.cuh): task t of band b gets the rows of
// the band's table restricted to its g-points [g0, g0 + n), row stride slice_rs(b), as one contiguous 128-byte aligned piece.
template <class Band>
int build_slices(const std::vector<double> &tab, const Band *bands, int nband, ColTask (*task)(int), int (*slice_rs)(int),
                 DevBuf &buf, ColSlices &out)
{
    std::vector<double> sl;
    out.max_bytes = 0;
    for (int t = 0; t < COL_NTASK; ++t) {
        const ColTask k = task(t);
        const Band &B = bands[k.band];
        const size_t end = k.band + 1 < nband ? (size_t)bands[k.band + 1].base : tab.size();
        const int rows = (int)((end - (size_t)B.base) / B.rs), srs = slice_rs(k.band);
        while (sl.size() % 16) sl.push_back(0.0);
        out.off[t] = (int)sl.size();
        out.bytes[t] = rows * srs * 8;
        if (out.bytes[t] > out.max_bytes) out.max_bytes = out.bytes[t];
        for (int r = 0; r < rows; ++r)
            for (int j = 0; j < srs; ++j)
                sl.push_back(j < k.n && k.g0 + j < B.rs ? tab[(size_t)B.base + (size_t)r * B.rs + k.g0 + j] : 0.0);
    }
    if (buf.ensure(sl.size() * 8)) return -1;
    if (cudaMemcpy(buf.p, sl.data(), sl.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    out.data = (const double *)buf.p;
    return 0;
}

int copy_exact(const char *name, double *dst, long n)
{
    const HostArr *a = find(name);
    if (!a || a->size() != n) return -1;
    std::memcpy(dst, a->data.data(), sizeof(double) * (size_t)n);
    return 0;
}

// ------------------------------------------------------------------------------------------------
int lw_init_impl(double cpdair)
{
    LwConst &c = G.lwc;
    std::memset(&c, 0, sizeof c);
    double pref[59];
    if (copy_exact("lwref.pref", pref, 59) || copy_exact("lwref.preflog", c.preflog, 59) ||
        copy_exact("lwref.tref", c.tref, 59) || copy_exact("lwref.chi_mls", c.chi_mls, 7 * 59))
        return fail(RRTMG_B200_ERR_TABLES, "lwref.{pref,preflog,tref,chi_mls} missing or wrong shape");
    const HostArr *totplnk = find("lwref.totplnk");
    if (!totplnk || totplnk->size() != 181 * 16) return fail(RRTMG_B200_ERR_TABLES, "lwref.totplnk missing");
    auto chi = [&](int m, int j) { return c.chi_mls[(j - 1) * 7 + (m - 1)]; };
    for (int j = 1; j <= 59; ++j) {
        c.rat_h2oco2[j - 1] = chi(1, j) / chi(2, j);
        c.rat_h2oo3[j - 1] = chi(1, j) / chi(3, j);
        c.rat_h2on2o[j - 1] = chi(1, j) / chi(4, j);
        c.rat_h2och4[j - 1] = chi(1, j) / chi(6, j);
        c.rat_n2oco2[j - 1] = chi(4, j) / chi(2, j);
        c.rat_o3co2[j - 1] = chi(3, j) / chi(2, j);
    }
    const double delwave[16] = {340., 150., 130., 70., 120., 160., 100., 100., 210., 90., 320., 280., 170., 130., 220., 650.};
    const double a0[16] = {1.66, 1.55, 1.58, 1.66, 1.54, 1.454, 1.89, 1.33, 1.668, 1.66, 1.66, 1.66, 1.66, 1.66, 1.66, 1.66};
    const double a1[16] = {0.00, 0.25, 0.22, 0.00, 0.13, 0.446, -0.10, 0.40, -0.006, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00};
    const double a2[16] = {0.00, -12.0, -11.7, 0.00, -0.72, -0.243, 0.19, -0.062, 0.414, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00};
    for (int b = 0; b < 16; ++b) { c.delwave[b] = delwave[b]; c.a0[b] = a0[b]; c.a1[b] = a1[b]; c.a2[b] = a2[b]; }
    const double grav = 9.8066, secdy = 8.6400e4;
    c.heatfac = grav * secdy / (cpdair * 1.e2);
    c.oneminus = 1. - 1.e-6;
    c.fluxfac = (2. * std::asin(1.)) * 2.e4;
    c.bpade = 1.0 / 0.278;

    // band-specific reference ratios (taumol.f90, e.g. :485-494): {planck_a, planck_b, m_a, m_b, m_a3}
    auto R = [&](int m1, int j1, int m2) { return chi(m1, j1) / chi(m2, j1); };
    double rr[16][5] = {};
    rr[2][0] = R(1, 9, 2);  rr[2][1] = R(1, 13, 2); rr[2][2] = R(1, 3, 2); rr[2][3] = R(1, 13, 2);
    rr[3][0] = R(1, 11, 2); rr[3][1] = R(3, 13, 2);
    rr[4][0] = R(1, 5, 2);  rr[4][1] = R(3, 43, 2); rr[4][2] = R(1, 7, 2);
    rr[6][0] = R(1, 3, 3);  rr[6][2] = R(1, 3, 3);
    rr[8][0] = R(1, 9, 6);  rr[8][2] = R(1, 3, 6);
    rr[11][0] = R(1, 10, 2);
    rr[12][0] = R(1, 5, 4); rr[12][2] = R(1, 1, 4); rr[12][4] = R(1, 3, 4);
    rr[14][0] = R(4, 1, 2); rr[14][2] = R(4, 1, 2);
    rr[15][0] = R(1, 6, 6);

    std::vector<double> tab;
    int g0 = 0;
    for (int b = 0; b < 16; ++b) {
        LwBand &B = c.band[b];
        char pfx[8];
        std::snprintf(pfx, sizeof pfx, "lw%02d", b + 1);
        const GMap m = make_gmap(kLwNgc[b], kLwNgn[b]);
        const int ng = m.ngc;
        B.ng = ng;
        B.rs = pow2_at_least(ng);
        B.g0 = g0;
        g0 += ng;
        while (tab.size() % 16) tab.push_back(0.0);      // 128-byte aligned band base
        B.base = (int)tab.size();
        for (int k = 0; k < 5; ++k) B.refrat[k] = rr[b][k];
        for (int k = 0; k < LS_COUNT; ++k) B.sec[k] = -1;
        auto add = [&](int sec, const char *o, const char *r) {
            const HostArr *a = reduce_named(pfx, o, r, m);
            if (a) B.sec[sec] = append_rows(tab, B.base, ng, a, std::string(o).rfind("fracref", 0) == 0);
            return a != nullptr;
        };
        if (!add(LS_ABSA, "kao", "absa") || !add(LS_SELF, "selfrefo", "selfref") || !add(LS_FOR, "forrefo", "forref") ||
            !add(LS_FRACA, "fracrefao", "fracrefa"))
            return fail(RRTMG_B200_ERR_TABLES, std::string(pfx) + ": kao/selfrefo/forrefo/fracrefao missing");
        add(LS_ABSB, "kbo", "absb");
        add(LS_FRACB, "fracrefbo", "fracrefb");
        // minor species and cross sections, per band (rrlw_kgNN.f90)
        struct Slot { int band, sec; const char *o, *r; };
        static const Slot slots[] = {
            {1, LS_MA1, "kao_mn2", "ka_mn2"},    {1, LS_MB1, "kbo_mn2", "kb_mn2"},
            {3, LS_MA1, "kao_mn2o", "ka_mn2o"},  {3, LS_MB1, "kbo_mn2o", "kb_mn2o"},
            {5, LS_MA1, "kao_mo3", "ka_mo3"},    {5, LS_X1, "ccl4o", "ccl4"},
            {6, LS_MA1, "kao_mco2", "ka_mco2"},  {6, LS_X1, "cfc11adjo", "cfc11adj"}, {6, LS_X2, "cfc12o", "cfc12"},
            {7, LS_MA1, "kao_mco2", "ka_mco2"},  {7, LS_MB1, "kbo_mco2", "kb_mco2"},
            {8, LS_MA1, "kao_mco2", "ka_mco2"},  {8, LS_MA2, "kao_mo3", "ka_mo3"},    {8, LS_MA3, "kao_mn2o", "ka_mn2o"},
            {8, LS_MB1, "kbo_mco2", "kb_mco2"},  {8, LS_MB2, "kbo_mn2o", "kb_mn2o"},
            {8, LS_X1, "cfc12o", "cfc12"},       {8, LS_X2, "cfc22adjo", "cfc22adj"},
            {9, LS_MA1, "kao_mn2o", "ka_mn2o"},  {9, LS_MB1, "kbo_mn2o", "kb_mn2o"},
            {11, LS_MA1, "kao_mo2", "ka_mo2"},   {11, LS_MB1, "kbo_mo2", "kb_mo2"},
            {13, LS_MA1, "kao_mco2", "ka_mco2"}, {13, LS_MA2, "kao_mco", "ka_mco"},   {13, LS_MB1, "kbo_mo3", "kb_mo3"},
            {15, LS_MA1, "kao_mn2", "ka_mn2"}};
        for (const Slot &s : slots)
            if (s.band == b + 1 && !add(s.sec, s.o, s.r))
                return fail(RRTMG_B200_ERR_TABLES, std::string(pfx) + "." + s.o + " missing");
        // stratospheric g-point scaling of bands 4 and 7 (taumol.f90:1009-1015, :1645-1650)
        if (b == 3) {
            double gsc[14];
            for (int i = 0; i < 14; ++i) gsc[i] = 1.0;
            const double f[7] = {0.92, 0.88, 1.07, 1.1, 0.99, 0.88, 0.943};
            for (int i = 0; i < 7; ++i) gsc[7 + i] = f[i];
            B.sec[LS_GSCALE] = append_const_row(tab, B.base, ng, gsc);
        } else if (b == 6) {
            double gsc[12];
            for (int i = 0; i < 12; ++i) gsc[i] = 1.0;
            const double f[6] = {0.92, 0.88, 1.07, 1.1, 0.99, 0.855};
            for (int i = 0; i < 6; ++i) gsc[5 + i] = f[i];
            B.sec[LS_GSCALE] = append_const_row(tab, B.base, ng, gsc);
        }
    }
    // lookup tables (rrtmg_lw_init.f90:106-123), interleaved {exp_tbl, tfn_tbl}
    std::vector<double> et(2 * (NTBL + 1));
    {
        const double expeps = 1.e-20;
        std::vector<double> tau(NTBL + 1), ex(NTBL + 1), tf(NTBL + 1);
        tau[0] = 0.0; tau[NTBL] = 1.e10; ex[0] = 1.0; ex[NTBL] = expeps; tf[0] = 0.0; tf[NTBL] = 1.0;
        for (int itr = 1; itr <= NTBL - 1; ++itr) {
            const double tfn = (double)itr / (double)NTBL;
            tau[itr] = c.bpade * tfn / (1. - tfn);
            ex[itr] = std::exp(-tau[itr]);
            if (ex[itr] <= expeps) ex[itr] = expeps;
            if (tau[itr] < 0.06) tf[itr] = tau[itr] / 6.;
            else tf[itr] = 1. - 2. * ((1. / tau[itr]) - (ex[itr] / (1. - ex[itr])));
        }
        for (int i = 0; i <= NTBL; ++i) { et[2 * i] = ex[i]; et[2 * i + 1] = tf[i]; }
        HostArr &he = G.reduced["lw.exp_tbl"]; he.dims = {NTBL + 1}; he.data = ex;
        HostArr &ht = G.reduced["lw.tfn_tbl"]; ht.dims = {NTBL + 1}; ht.data = tf;
    }
    const HostArr *totplnkderiv = find("lwref.totplnkderiv");
    if (!totplnkderiv || totplnkderiv->size() != 181 * 16) return fail(RRTMG_B200_ERR_TABLES, "lwref.totplnkderiv missing");
    if (G.lw_tab.ensure(tab.size() * 8) || G.lw_totplnk.ensure(2 * 181 * 16 * 8) || G.lw_exptfn.ensure(et.size() * 8))
        return fail(RRTMG_B200_ERR_CUDA, "cudaMalloc failed for LW tables");
    CUDA_OK(cudaMemcpy(G.lw_tab.p, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(G.lw_totplnk.p, totplnk->data.data(), 181 * 16 * 8, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(G.lw_exptfn.p, et.data(), et.size() * 8, cudaMemcpyHostToDevice));
    G.lwt.tab = (const double *)G.lw_tab.p;
    CUDA_OK(cudaMemcpy((double *)G.lw_totplnk.p + 181 * 16, totplnkderiv->data.data(), 181 * 16 * 8, cudaMemcpyHostToDevice));
    G.lwt.totplnk = (const double *)G.lw_totplnk.p;
    G.lwt.totplnkderiv = (const double *)G.lw_totplnk.p + 181 * 16;
    G.lwt.exptfn = (const double *)G.lw_exptfn.p;
    if (build_slices(tab, c.band, NBNDLW, lw_task, lw_slice_rs, G.lw_slices, G.lwt.sl))
        return fail(RRTMG_B200_ERR_CUDA, "cudaMalloc / copy failed for the LW table slices");
    if (lw_upload_const(c)) return fail(RRTMG_B200_ERR_CUDA, "cudaMemcpyToSymbol(c_lw) failed");
    {   // cloud absorption coefficients of cldprop (lwcldpr); optional: needed for inflglw > 0 only
        static LwCldConst k;
        const HostArr *a1 = find("lwcld.abscld1"), *l0 = find("lwcld.absliq0"), *i0 = find("lwcld.absice0"), *i1 = find("lwcld.absice1"),
                      *i2 = find("lwcld.absice2"), *i3 = find("lwcld.absice3"), *l1 = find("lwcld.absliq1");
        k.have = a1 && l0 && i0 && i1 && i2 && i3 && l1 && i0->size() == 2 && i1->size() == 10 && i2->size() == 43 * 16 &&
                 i3->size() == 46 * 16 && l1->size() == 58 * 16;
        if (k.have) {
            k.abscld1 = a1->data[0]; k.absliq0 = l0->data[0];
            std::memcpy(k.absice0, i0->data.data(), sizeof k.absice0); std::memcpy(k.absice1, i1->data.data(), sizeof k.absice1);
            std::memcpy(k.absice2, i2->data.data(), sizeof k.absice2); std::memcpy(k.absice3, i3->data.data(), sizeof k.absice3);
            std::memcpy(k.absliq1, l1->data.data(), sizeof k.absliq1);
            if (lw_upload_cld(k)) return fail(RRTMG_B200_ERR_CUDA, "cudaMemcpyToSymbol(d_lwcld) failed");
        }
        G.lw_have_cld = k.have != 0;
    }
    G.lw_ready = true;
    return RRTMG_B200_OK;
}

int sw_init_impl(double cpdair)
{
    SwConst &c = G.swc;
    std::memset(&c, 0, sizeof c);
    double pref[59];
    if (copy_exact("swref.pref", pref, 59) || copy_exact("swref.preflog", c.preflog, 59) ||
        copy_exact("swref.tref", c.tref, 59))
        return fail(RRTMG_B200_ERR_TABLES, "swref.{pref,preflog,tref} missing or wrong shape");
    const double grav = 9.8066, secdy = 8.6400e4;
    c.heatfac = grav * secdy / (cpdair * 1.e2);
    c.oneminus = 1.0 - 1.e-06;
    c.bpade = 1.0 / 0.278;

    std::vector<double> tab;
    int g0 = 0;
    for (int b = 0; b < 14; ++b) {
        SwBand &B = c.band[b];
        char pfx[8];
        std::snprintf(pfx, sizeof pfx, "sw%d", b + 16);
        const GMap m = make_gmap(kSwNgc[b], kSwNgn[b]);
        const int ng = m.ngc;
        B.ng = ng;
        B.rs = pow2_at_least(ng);
        B.g0 = g0;
        g0 += ng;
        while (tab.size() % 16) tab.push_back(0.0);      // 128-byte aligned band base
        B.base = (int)tab.size();
        for (int k = 0; k < SS_COUNT; ++k) B.sec[k] = -1;
        auto add = [&](int sec, const char *o, const char *r, bool gfirst) {
            const HostArr *a = reduce_named(pfx, o, r, m);
            if (a) B.sec[sec] = append_rows(tab, B.base, ng, a, gfirst);
            return a;
        };
        add(SS_ABSA, "kao", "absa", false);
        add(SS_ABSB, "kbo", "absb", false);
        add(SS_SELF, "selfrefo", "selfref", false);
        add(SS_FOR, "forrefo", "forref", false);
        const HostArr *sf = add(SS_SFLUX, "sfluxrefo", "sfluxref", true);
        if (!sf) return fail(RRTMG_B200_ERR_TABLES, std::string(pfx) + ".sfluxrefo missing");
        B.nsflux = (int)(sf->size() / ng);
        // Rayleigh: scalar (one row filled with it), per-g row, or band 24's (ng,9) + raylb
        const HostArr *rs = find(std::string(pfx) + ".rayl");
        if (rs) {
            double row[16];
            for (int i = 0; i < ng; ++i) row[i] = rs->data[0];
            B.sec[SS_RAYL] = append_const_row(tab, B.base, ng, row);
            B.nrayl = 1;
        } else if (add(SS_RAYL, "raylo", "rayl", true)) {
            B.nrayl = 1;
        } else if (add(SS_RAYL, "raylao", "rayla", true)) {
            B.nrayl = 9;
            if (!add(SS_RAYLB, "raylbo", "raylb", true)) return fail(RRTMG_B200_ERR_TABLES, std::string(pfx) + ".raylbo missing");
        } else {
            return fail(RRTMG_B200_ERR_TABLES, std::string(pfx) + ": no Rayleigh coefficients");
        }
        if (b + 16 == 20) add(SS_X1, "absch4o", "absch4", true);
        if (b + 16 == 24 || b + 16 == 25) { add(SS_X1, "abso3ao", "abso3a", true); add(SS_X2, "abso3bo", "abso3b", true); }
        if (b + 16 == 29) { add(SS_X1, "absh2oo", "absh2o", true); add(SS_X2, "absco2o", "absco2", true); }
    }
    // ECMWF aerosol optical properties (swaerpr), (nbndsw, naerec) column-major
    {
        const HostArr *ta = find("swaer.rsrtaua"), *pa = find("swaer.rsrpiza"), *ga = find("swaer.rsrasya");
        c.have_aer = ta && pa && ga && ta->size() == 84 && pa->size() == 84 && ga->size() == 84;
        for (int ib = 0; ib < 14; ++ib)
            for (int ia = 0; ia < 6; ++ia) {
                c.rsrtaua[ib][ia] = c.have_aer ? ta->data[ib + 14 * ia] : 0.0;
                c.rsrpiza[ib][ia] = c.have_aer ? pa->data[ib + 14 * ia] : 0.0;
                c.rsrasya[ib][ia] = c.have_aer ? ga->data[ib + 14 * ia] : 0.0;
            }
    }
    // exp_tbl (rrtmg_sw_init.f90:96-105), interleaved with its reciprocal
    std::vector<double> et(2 * (NTBL + 1));
    {
        const double expeps = 1.e-20;
        std::vector<double> ex(NTBL + 1);
        ex[0] = 1.0; ex[NTBL] = expeps;
        for (int itr = 1; itr <= NTBL - 1; ++itr) {
            const double tfn = (double)itr / (double)NTBL;
            const double tau = c.bpade * tfn / (1. - tfn);
            ex[itr] = std::exp(-tau);
            if (ex[itr] <= expeps) ex[itr] = expeps;
        }
        for (int i = 0; i <= NTBL; ++i) { et[2 * i] = ex[i]; et[2 * i + 1] = 1. / ex[i]; }
        HostArr &he = G.reduced["sw.exp_tbl"]; he.dims = {NTBL + 1}; he.data = ex;
    }
    if (G.sw_tab.ensure(tab.size() * 8) || G.sw_exptbl.ensure(et.size() * 8))
        return fail(RRTMG_B200_ERR_CUDA, "cudaMalloc failed for SW tables");
    CUDA_OK(cudaMemcpy(G.sw_tab.p, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(G.sw_exptbl.p, et.data(), et.size() * 8, cudaMemcpyHostToDevice));
    G.swt.tab = (const double *)G.sw_tab.p;
    G.swt.exptbl = (const double *)G.sw_exptbl.p;
    if (build_slices(tab, c.band, NBNDSW, sw_task, sw_slice_rs, G.sw_slices, G.swt.sl))
        return fail(RRTMG_B200_ERR_CUDA, "cudaMalloc / copy failed for the SW table slices");
    if (sw_upload_const(c)) return fail(RRTMG_B200_ERR_CUDA, "cudaMemcpyToSymbol(c_sw) failed");
    {   // cloud optical properties of cldprop_sw (swcldpr); optional: needed for inflgsw = 2 only
        static SwCldConst k;
        struct { const char *nm; double *dst; size_t n; } t[16] = {
            {"swcld.extliq1", k.extliq1, 58 * 14}, {"swcld.ssaliq1", k.ssaliq1, 58 * 14}, {"swcld.asyliq1", k.asyliq1, 58 * 14},
            {"swcld.extice2", k.extice2, 43 * 14}, {"swcld.ssaice2", k.ssaice2, 43 * 14}, {"swcld.asyice2", k.asyice2, 43 * 14},
            {"swcld.extice3", k.extice3, 46 * 14}, {"swcld.ssaice3", k.ssaice3, 46 * 14}, {"swcld.asyice3", k.asyice3, 46 * 14},
            {"swcld.fdlice3", k.fdlice3, 46 * 14}, {"swcld.abari", k.abari, 5}, {"swcld.bbari", k.bbari, 5}, {"swcld.cbari", k.cbari, 5},
            {"swcld.dbari", k.dbari, 5}, {"swcld.ebari", k.ebari, 5}, {"swcld.fbari", k.fbari, 5}};
        k.have = 1;
        for (auto &e : t) {
            const HostArr *a = find(e.nm);
            if (!a || (size_t)a->size() != e.n) { k.have = 0; break; }
            std::memcpy(e.dst, a->data.data(), e.n * sizeof(double));
        }
        if (k.have && sw_upload_cld(k)) return fail(RRTMG_B200_ERR_CUDA, "cudaMemcpyToSymbol(d_swcld) failed");
        G.sw_have_cld = k.have != 0;
    }
    G.sw_ready = true;
    return RRTMG_B200_OK;
}

// ------------------------------------------------------------------------------------------------
// workspace carving
struct Carver {
    char *p;
    size_t off = 0;
    explicit Carver(void *base) : p((char *)base) {}
    template <class T> T *take(size_t n)
    {
        off = (off + 255) & ~(size_t)255;
        T *r = p ? (T *)(p + off) : nullptr;
        off += n * sizeof(T);
        return r;
    }
};
size_t lw_carve(LwWork &w, void *base, int nc, int nlay, bool fields, bool cloud = false)
{
    Carver c(base);
    w.nc = nc; w.nlay = nlay;
    const size_t np = (size_t)nc * nlay;
    w.laytrop = c.take<int>(nc);
    // per-cell setcoef state: what the tasks of the fused clear-sky kernel read, and what the stage-capture test hook returns
    (void)fields;
    w.idx = c.take<uint32_t>(np);
    w.f = c.take<double>((size_t)((nc + 31) & ~31) * nlay * LF_SLOTS);      // [field][lay][col], or tile-major in the fused path
    w.idrv = 0;
    w.dplankbnd = c.take<double>((size_t)nc * 16);
    w.cs_coldry = c.take<double>(np);
    w.cs_wkl1 = c.take<double>(np);
    w.cs_lower = c.take<unsigned char>(np);
    w.secdiff = c.take<double>((size_t)nc * 16);
    w.planklay = c.take<double>(np * 16);
    w.planklev = c.take<double>((size_t)nc * (nlay + 1) * 16);
    w.plankbnd = c.take<double>((size_t)nc * 16);
    // staging: [col][lay][140] twice for the staged kernels; the fused kernel sees the same storage as one field of pairs
    // over whole 32-column tiles
    w.ncp = (nc + 31) & ~31;
    const size_t npp = (size_t)w.ncp * nlay;
    w.colst = c.take<double>(2 * npp * NGPTLW);
    w.taug = w.colst;
    w.fracs = w.colst + npp * NGPTLW;
    w.part = c.take<double>((size_t)LW_NTASK * 2 * (nlay + 1) * w.ncp);
    w.taucloud = cloud ? c.take<double>(np * 16) : nullptr;
    w.ncbands = cloud ? c.take<int>(nc) : nullptr;
    w.err = cloud ? (int *)G.lw_err.p : nullptr;
    return c.off + 256;
}
size_t sw_carve(SwWork &w, void *base, int nc, int nlay, bool fields, bool general = false)
{
    Carver c(base);
    w.nc = nc; w.nlay = nlay;
    const size_t np = (size_t)nc * nlay;
    w.laytrop = c.take<int>(nc);
    w.laysolfr = c.take<int>((size_t)nc * 14);
    w.cs_jp = c.take<unsigned char>(np);
    w.idx = fields ? c.take<uint32_t>(np) : nullptr;
    w.f = fields ? c.take<double>(np * SF_COUNT) : nullptr;
    w.taug = c.take<double>(np * NGPTSW);
    w.colmol = c.take<double>(np);
    w.taur24 = c.take<double>(np * 8);
    w.taur = fields ? c.take<double>(np * NGPTSW) : nullptr;      // expanded from rdesc (test hook)
    w.sfluxzen = c.take<double>((size_t)nc * NGPTSW);
    w.part = c.take<double>((size_t)nc * (NGPTSW / 16) * 2 * (nlay + 1));
#ifdef RRTMG_B200_DEV_VARIANTS
    w.stack = c.take<double>((size_t)SW_STACK_SLOTS * (nlay + 1) * 96);
#else
    w.stack = nullptr;
#endif
    w.ncp = (nc + 31) & ~31;
    if (!general) {             // fused clear-sky kernel: tile-major setcoef state, scratch field, task sums
        w.tf = c.take<double>((size_t)w.ncp * nlay * SF_SLOTS);
        w.colst = c.take<double>((size_t)w.ncp * nlay * SW_NSLOT);
        w.cpart = c.take<double>((size_t)SW_NTASK * 2 * (nlay + 1) * w.ncp);
    }
    w.opt = general ? c.take<double>(np * 14 * 6) : nullptr;
    w.clfr = general ? c.take<double>(np) : nullptr;
    w.err = general ? (int *)G.sw_err.p : nullptr;
    return c.off + 256;
}

int pick_chunk(int ncol)
{
    // columns per pass; T170L60 step with the column kernels: 16384 28.6, 32768 25.5, 65536 24.3, 131072 23.8 ms
    int ch = G.chunk > 0 ? G.chunk : 131072;
    return ch < ncol ? ch : ncol;
}
// The workspace of a pass is 59 GB at 131072 columns x 60 layers (the fields of the fused and of the staged kernels).  MiMA's shipped job size is 32 ranks (exp/nci_runscript.sh:7),
// i.e. four ranks per GPU on one 8-GPU box: the pass is halved until its workspace fits into what the device has free
// (beyond what this buffer already holds), so several ranks can share a GPU without an allocation failure.
// cudaMemGetInfo is a slow, synchronising driver call (measured: 2 ms per call, 4.4 ms per LW+SW step): it is asked only
// when the workspace has to grow, and the answer is remembered per (requested pass, layers, kind).
struct FitCache { int req = 0, nlay = 0, kind = -1, fit = 0; };
template <class Carve>
int fit_chunk(int chunk, int nlay, int kind, const DevBuf &have, FitCache &fc, Carve bytes_of)
{
    if (fc.req == chunk && fc.nlay == nlay && fc.kind == kind) return fc.fit;
    int fit = chunk;
    if (bytes_of(chunk) > have.bytes) {
        size_t freeb = 0, total = 0;
        if (cudaMemGetInfo(&freeb, &total) == cudaSuccess)
            while (fit > 1024 && bytes_of(fit) > have.bytes + (size_t)(0.9 * (double)freeb)) fit = (fit + 1) / 2;
    }
    fc.req = chunk; fc.nlay = nlay; fc.kind = kind; fc.fit = fit;
    return fit;
}
FitCache g_fit_lw, g_fit_sw;

// ------------------------------------------------------------------------------------------------
struct LwOpt {                // the optional cloud arguments of rrtmg_lw
    int inflglw = 0;
    const double *cldfr = nullptr, *taucld = nullptr;
    int iceflglw = 0, liqflglw = 0;
    const double *cicewp = nullptr, *cliqwp = nullptr, *reice = nullptr, *reliq = nullptr;
};
int lw_validate(int ncol, int nlay, int *icld, int idrv, const LwOpt &o = LwOpt())
{
    if (!G.lw_ready) return fail(RRTMG_B200_ERR_NOT_INITIALIZED, "rrtmg_b200_lw_init has not been called");
    if (ncol < 0 || nlay < 1 || nlay > MAXLAY) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "ncol/nlay out of range (1 <= nlay <= 128)");
    if (icld && (*icld < 0 || *icld > 3)) *icld = 2;     // LW rad.nomcica:437
    if (icld && *icld != 0) {
        if (o.inflglw < 0 || o.inflglw > 2) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "rrtmg_lw: inflglw must be 0, 1 or 2");
        if (!o.cldfr || !o.taucld) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "rrtmg_lw: icld > 0 needs cldfr and taucld");
        if (o.inflglw >= 1 && (!o.cicewp || !o.cliqwp)) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "rrtmg_lw: inflglw > 0 needs cicewp and cliqwp");
        if (o.inflglw == 2 && (o.iceflglw < 0 || o.iceflglw > 3 || o.liqflglw < 0 || o.liqflglw > 1))
            return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "rrtmg_lw: iceflglw must be 0..3 and liqflglw 0..1");
        if (o.inflglw == 2 && (!o.reice || !o.reliq)) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "rrtmg_lw: inflglw = 2 needs reice and reliq");
        if (o.inflglw >= 1 && !G.lw_have_cld) return fail(RRTMG_B200_ERR_TABLES, "rrtmg_lw: inflglw > 0 needs the lwcld.* tables");
    }
    if (idrv != 0 && idrv != 1) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "rrtmg_lw: idrv must be 0 or 1");
    return RRTMG_B200_OK;
}
// the optional cloud / aerosol arguments of rrtmg_sw (host or device pointers, as the entry point's other arrays)
struct SwOpt {
    int inflgsw = 0;
    const double *cldfr = nullptr, *taucld = nullptr, *ssacld = nullptr, *asmcld = nullptr, *fsfcld = nullptr;
    const double *tauaer = nullptr, *ssaaer = nullptr, *asmaer = nullptr;
    const double *ecaer = nullptr;
    int iceflgsw = 0, liqflgsw = 0;
    const double *cicewp = nullptr, *cliqwp = nullptr, *reice = nullptr, *reliq = nullptr;
};
int sw_validate(int ncol, int nlay, int *icld, int *iaer, const SwOpt &o = SwOpt())
{
    if (!G.sw_ready) return fail(RRTMG_B200_ERR_NOT_INITIALIZED, "rrtmg_b200_sw_init has not been called");
    if (ncol < 0 || nlay < 1 || nlay > MAXLAY) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "ncol/nlay out of range (1 <= nlay <= 128)");
    if (icld && (*icld < 0 || *icld > 3)) *icld = 2;                           // SW rad.nomcica:468
    if (iaer && *iaer != 0 && *iaer != 6 && *iaer != 10) *iaer = 0;            // SW rad.nomcica:473
    if (icld && *icld != 0) {
        if (o.inflgsw != 0 && o.inflgsw != 2)
            return fail(RRTMG_B200_ERR_UNSUPPORTED, "rrtmg_sw: inflgsw must be 0 (optical properties given) or 2 (water paths and radii); cldprop_sw has no other branch");
        if (!o.cldfr) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "rrtmg_sw: icld > 0 needs cldfr");
        if (o.inflgsw == 0 && (!o.taucld || !o.ssacld || !o.asmcld || !o.fsfcld))
            return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "rrtmg_sw: icld > 0 with inflgsw = 0 needs taucld, ssacld, asmcld, fsfcld");
        if (o.inflgsw == 2) {
            if (!o.cicewp || !o.cliqwp || !o.reice || !o.reliq)
                return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "rrtmg_sw: inflgsw = 2 needs cicewp, cliqwp, reice, reliq");
            if (o.iceflgsw < 1 || o.iceflgsw > 3 || o.liqflgsw != 1)
                return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "rrtmg_sw: inflgsw = 2 needs iceflgsw in 1..3 and liqflgsw = 1 (the options cldprop_sw defines)");
            if (!G.sw_have_cld) return fail(RRTMG_B200_ERR_TABLES, "rrtmg_sw: inflgsw = 2 needs the swcld.* tables");
        }
    }
    if (iaer && *iaer == 6) {
        if (!G.swc.have_aer) return fail(RRTMG_B200_ERR_TABLES, "rrtmg_sw: iaer = 6 needs the swaer.rsrtaua/rsrpiza/rsrasya tables");
        if (!o.ecaer) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "rrtmg_sw: iaer = 6 needs ecaer");
    }
    if (iaer && *iaer == 10 && (!o.tauaer || !o.ssaaer || !o.asmaer))
        return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "rrtmg_sw: iaer = 10 needs tauaer, ssaaer, asmaer");
    return RRTMG_B200_OK;
}

// One device pass over nc columns whose interface arrays start at in/out (leading dimensions in in.ld/out.ld).
// `work` must already hold lw_carve(., nc_max, nlay, fields) bytes.
int lw_chunk(const LwIn &in, const LwOut &out, int nc, int nlay, void *work, bool fields, cudaStream_t st, bool last)
{
    LwWork w;
    lw_carve(w, work, nc, nlay, fields, in.icld >= 1);
    w.idrv = out.duflx_dt ? 1 : 0;
    double *cap = nullptr;
    if (fields) {
        if (G.lw_cap.ensure(2 * (size_t)nc * nlay * NGPTLW * 8)) return fail(RRTMG_B200_ERR_CUDA, "cudaMalloc failed (capture)");
        cap = (double *)G.lw_cap.p;
    }
    G.launches += lw_run_pass(G.lwt, in, out, w, st, cap);
    CUDA_OK(cudaGetLastError());
    if (last) { G.lw_last = w; G.lw_last_ncol = fields ? nc : 0; }
    return RRTMG_B200_OK;
}
int sw_chunk(const SwIn &in, const SwOut &out, int nc, int nlay, void *work, bool fields, cudaStream_t st, bool last)
{
    SwWork w;
    sw_carve(w, work, nc, nlay, fields, in.icld >= 1 || in.iaer != 0);
    CUDA_OK(cudaMemsetAsync(w.sfluxzen, 0, (size_t)nc * NGPTSW * 8, st));
    G.launches += sw_run_pass(G.swt, in, out, w, st);
    CUDA_OK(cudaGetLastError());
    if (last) { G.sw_last = w; G.sw_last_ncol = fields ? nc : 0; }
    return RRTMG_B200_OK;
}

// general SW path: the partial-cloud flag (SW rad.nomcica:534-539) lives in one device word, cleared before the first
// pass and read back (one stream synchronisation) after the last
int sw_err_begin(bool general)
{
    if (!general) return RRTMG_B200_OK;
    if (G.sw_err.ensure(256)) return fail(RRTMG_B200_ERR_CUDA, "cudaMalloc failed");
    CUDA_OK(cudaMemset(G.sw_err.p, 0, 8));
    return RRTMG_B200_OK;
}
int sw_err_end(bool general)
{
    if (!general) return RRTMG_B200_OK;
    int flag[2] = {0, 0};
    CUDA_OK(cudaMemcpy(flag, G.sw_err.p, 8, cudaMemcpyDeviceToHost));
    if (flag[0] & 1) return fail(RRTMG_B200_ERR_PARTIAL_CLOUD, "rrtmg_sw: PARTIAL CLOUD NOT ALLOWED (0 < cldfr < 1 with icld > 0)");
    static const char *msg[5] = {"", "ICE RADIUS OUT OF BOUNDS", "ICE GENERALIZED EFFECTIVE SIZE OUT OF BOUNDS",
                                 "LIQUID EFFECTIVE RADIUS OUT OF BOUNDS", "an interpolated cloud property is out of range"};
    if (flag[1] >= 1 && flag[1] <= 4) return fail(RRTMG_B200_ERR_CLOUD_INPUT, std::string("rrtmg_sw cldprop_sw: ") + msg[flag[1]]);
    return RRTMG_B200_OK;
}

// cloudy LW: cldprop's Fortran `stop`s (radii out of range) come back through one device word
int lw_err_begin(bool cloud)
{
    if (!cloud) return RRTMG_B200_OK;
    if (G.lw_err.ensure(256)) return fail(RRTMG_B200_ERR_CUDA, "cudaMalloc failed");
    CUDA_OK(cudaMemset(G.lw_err.p, 0, 4));
    return RRTMG_B200_OK;
}
int lw_err_end(bool cloud)
{
    if (!cloud) return RRTMG_B200_OK;
    int flag = 0;
    CUDA_OK(cudaMemcpy(&flag, G.lw_err.p, 4, cudaMemcpyDeviceToHost));
    static const char *msg[5] = {"", "ICE RADIUS TOO SMALL", "ICE RADIUS OUT OF BOUNDS", "ICE GENERALIZED EFFECTIVE SIZE OUT OF BOUNDS",
                                 "LIQUID EFFECTIVE RADIUS OUT OF BOUNDS"};
    if (flag >= 1 && flag <= 4) return fail(RRTMG_B200_ERR_CLOUD_INPUT, std::string("rrtmg_lw cldprop: ") + msg[flag]);
    return RRTMG_B200_OK;
}
void lw_set_optional(LwIn &in, const int *icld, const LwOpt &o)
{
    if (!icld || *icld < 1) return;
    in.icld = *icld; in.cldfr = o.cldfr; in.taucld = o.taucld;
    in.inflg = o.inflglw; in.iceflg = o.iceflglw; in.liqflg = o.liqflglw;
    if (o.inflglw >= 1) { in.cicewp = o.cicewp; in.cliqwp = o.cliqwp; }
    if (o.inflglw == 2) { in.reice = o.reice; in.reliq = o.reliq; }
}

int lw_device_impl(int ncol, int nlay, int *icld, int idrv, const LwIn &in0, const LwOut &out0, cudaStream_t st, DevBuf *work = nullptr,
                   const LwOpt &opt = LwOpt())
{
    DevBuf &wk = work ? *work : G.lw_work;
    if (const int rc = lw_validate(ncol, nlay, icld, idrv, opt)) return rc;
    if (idrv == 1 && (!out0.duflx_dt || !out0.duflxc_dt)) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "rrtmg_lw: idrv = 1 needs duflx_dt and duflxc_dt");
    if (ncol == 0) return RRTMG_B200_OK;
    int chunk = pick_chunk(ncol);
    LwWork w;
    const bool cloudy = in0.icld >= 1;
    chunk = fit_chunk(chunk, nlay, (cloudy ? 1 : 0) | (G.capture ? 2 : 0) | (work ? 4 : 0), wk, g_fit_lw,
                      [&](int c) { LwWork t; return lw_carve(t, nullptr, c, nlay, G.capture && ncol <= c, cloudy); });
    const bool fields = G.capture && ncol <= chunk;
    if (const int rc = lw_err_begin(cloudy)) return rc;
    if (wk.ensure(lw_carve(w, nullptr, chunk, nlay, fields, cloudy))) return fail(RRTMG_B200_ERR_CUDA, "cudaMalloc failed for the LW workspace");
    for (int c0 = 0; c0 < ncol; c0 += chunk) {
        const int nc = (ncol - c0 < chunk) ? ncol - c0 : chunk;
        LwIn in = in0;
        LwOut out = out0;
#define OFF(p) if (in.p) in.p += c0
        OFF(play); OFF(plev); OFF(tlay); OFF(tlev); OFF(tsfc); OFF(h2o); OFF(o3); OFF(co2); OFF(ch4); OFF(n2o);
        OFF(o2); OFF(cfc11); OFF(cfc12); OFF(cfc22); OFF(ccl4); OFF(emis); OFF(tauaer); OFF(cldfr);
        OFF(cicewp); OFF(cliqwp); OFF(reice); OFF(reliq);
#undef OFF
        if (in.taucld) in.taucld += (size_t)16 * c0;
        out.uflx += c0; out.dflx += c0; out.hr += c0; out.uflxc += c0; out.dflxc += c0; out.hrc += c0;
        if (out.duflx_dt) { out.duflx_dt += c0; out.duflxc_dt += c0; }
        if (const int rc = lw_chunk(in, out, nc, nlay, wk.p, fields, st, c0 + nc >= ncol)) return rc;
    }
    if (cloudy) CUDA_OK(cudaStreamSynchronize(st));
    return lw_err_end(cloudy);
}

int sw_device_impl(int ncol, int nlay, int *icld, int *iaer, const SwIn &in0, const SwOut &out0, cudaStream_t st, DevBuf *work = nullptr,
                   const SwOpt &opt = SwOpt())
{
    DevBuf &wk = work ? *work : G.sw_work;
    if (const int rc = sw_validate(ncol, nlay, icld, iaer, opt)) return rc;
    if (ncol == 0) return RRTMG_B200_OK;
    int chunk = pick_chunk(ncol);
    SwWork w;
    const bool general = in0.icld >= 1 || in0.iaer != 0;
    chunk = fit_chunk(chunk, nlay, (general ? 1 : 0) | (G.capture ? 2 : 0) | (work ? 4 : 0), wk, g_fit_sw,
                      [&](int c) { SwWork t; return sw_carve(t, nullptr, c, nlay, G.capture && ncol <= c, general); });
    const bool fields = G.capture && ncol <= chunk;
    if (const int rc = sw_err_begin(general)) return rc;
    if (wk.ensure(sw_carve(w, nullptr, chunk, nlay, fields, general))) return fail(RRTMG_B200_ERR_CUDA, "cudaMalloc failed for the SW workspace");
    for (int c0 = 0; c0 < ncol; c0 += chunk) {
        const int nc = (ncol - c0 < chunk) ? ncol - c0 : chunk;
        SwIn in = in0;
        SwOut out = out0;
#define OFF(p) if (in.p) in.p += c0
        OFF(play); OFF(plev); OFF(tlay); OFF(tlev); OFF(tsfc); OFF(h2o); OFF(o3); OFF(co2); OFF(ch4); OFF(n2o);
        OFF(o2); OFF(asdir); OFF(asdif); OFF(aldir); OFF(aldif); OFF(coszen);
        OFF(cldfr); OFF(tauaer); OFF(ssaaer); OFF(asmaer); OFF(ecaer); OFF(cicewp); OFF(cliqwp); OFF(reice); OFF(reliq);
#undef OFF
#define OFF14(p) if (in.p) in.p += (size_t)14 * c0
        OFF14(taucld); OFF14(ssacld); OFF14(asmcld); OFF14(fsfcld);
#undef OFF14
        out.uflx += c0; out.dflx += c0; out.hr += c0; out.uflxc += c0; out.dflxc += c0; out.hrc += c0;
        if (const int rc = sw_chunk(in, out, nc, nlay, wk.p, fields, st, c0 + nc >= ncol)) return rc;
    }
    if (general) CUDA_OK(cudaStreamSynchronize(st));
    return sw_err_end(general);
}

void sw_set_optional(SwIn &in, const int *icld, const int *iaer, const SwOpt &o)
{
    in.icld = icld ? *icld : 0;
    in.iaer = iaer ? *iaer : 0;
    if (in.icld >= 1) {
        in.cldfr = o.cldfr; in.inflg = o.inflgsw;
        if (o.inflgsw == 0) { in.taucld = o.taucld; in.ssacld = o.ssacld; in.asmcld = o.asmcld; in.fsfcld = o.fsfcld; }
        else { in.iceflg = o.iceflgsw; in.liqflg = o.liqflgsw; in.cicewp = o.cicewp; in.cliqwp = o.cliqwp; in.reice = o.reice; in.reliq = o.reliq; }
    }
    if (in.iaer == 10) { in.tauaer = o.tauaer; in.ssaaer = o.ssaaer; in.asmaer = o.asmaer; }
    if (in.iaer == 6) in.ecaer = o.ecaer;
}

// adjflux (SW rad.nomcica:953-972, earth_sun :734-758)
double sw_adjflux(double adjes, int dyofyr, double scon)
{
    double adjflx = adjes;
    if (dyofyr > 0) {
        const double pi = 2. * std::asin(1.);
        const double gamma = 2. * pi * (dyofyr - 1) / 365.;
        adjflx = 1.000110 + .034221 * std::cos(gamma) + .001289 * std::sin(gamma) + .000719 * std::cos(2. * gamma) +
                 .000077 * std::sin(2. * gamma);
    }
    const double solvar = scon / 1.36822e+03;
    return adjflx * solvar;
}

// Host-pointer ABI: the batch is cut into column chunks that flow through a pipeline of `host_slots` slots (one stream and
// one set of device buffers each), so the H2D copies of the next chunks and the D2H copies of the previous ones overlap the
// kernels of chunk i.  A slot's copy-in, kernels and copy-out are in stream order, so with two slots the copy-in of chunk
// i+2 waits for the copy-out of chunk i: a chunk's period is max(kernels, (copy-in + kernels + copy-out) / 2), and the
// copy-in of a chunk (PCIe) takes about as long as its SW kernels.  With n slots the period is max(copy-in, kernels, copy-out,
// sum / n); T170L60 end to end (shim-style / all outputs / run_rrtmg, ms per step): 2 slots 29.8 / 39.7 / 28.5, 3 slots 25.0 /
// 29.5 / 26.4, 4 slots (default) 24.2 / 26.6 / 24.0, against 21.8 ms of kernels (profiles/r02ar_sweep_host_slots.txt).  A column chunk of a column-major (ncol, rows) array is a 2-D copy with source
// pitch ncol*8.
constexpr int MAXSLOT = 6;
struct Pipe {
    cudaStream_t st[MAXSLOT] = {};
    DevBuf in[MAXSLOT], out[MAXSLOT], work[MAXSLOT];
    int ready()
    {
        for (int i = 0; i < MAXSLOT; ++i)
            if (!st[i] && cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking) != cudaSuccess) return -1;
        return 0;
    }
    // slot i could not be allocated: give its partial buffers back, clear the allocation error, keep slots 0..i-1
    int shrink(int i)
    {
        in[i].release(); out[i].release(); work[i].release();
        (void)cudaGetLastError();
        return i;
    }
    void release()
    {
        for (int i = 0; i < MAXSLOT; ++i) {
            in[i].release(); out[i].release(); work[i].release();
            if (st[i]) { cudaStreamDestroy(st[i]); st[i] = nullptr; }
        }
    }
};
Pipe P_lw, P_sw;
// error exit of a pipelined host call: copies of the other block may still be in flight to or from the caller's buffers
int drain(Pipe &P, int rc)
{
    for (int i = 0; i < MAXSLOT; ++i)
        if (P.st[i]) cudaStreamSynchronize(P.st[i]);
    return rc;
}

struct Slot {                 // bump allocator over one slot's device buffer + the chunk being copied
    char *base;
    size_t off;
    int c0, nc, ncol;
    cudaStream_t st;
    bool ok;
    double *take(size_t rows)
    {
        double *d = (double *)(base + off);
        off += (((size_t)nc * rows * 8 + 255) & ~(size_t)255);
        return d;
    }
    const double *up(const double *h, size_t rows)          // host (ncol, rows) columns [c0, c0+nc) -> device (nc, rows)
    {
        if (!h) return nullptr;
        double *d = take(rows);
        const cudaError_t e = (nc == ncol)
            ? cudaMemcpyAsync(d, h, (size_t)nc * rows * 8, cudaMemcpyHostToDevice, st)
            : cudaMemcpy2DAsync(d, (size_t)nc * 8, h + c0, (size_t)ncol * 8, (size_t)nc * 8, rows, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) ok = false;
        return d;
    }
    // host (ncol, rows) columns [c0, c0+nc) -> the same columns of a full-batch device array (ncol, rows); returns the
    // pointer to column c0 (leading dimension ncol)
    const double *up_full(const double *h, double *dfull, size_t rows)
    {
        if (!h) return nullptr;
        const cudaError_t e = cudaMemcpy2DAsync(dfull + c0, (size_t)ncol * 8, h + c0, (size_t)ncol * 8, (size_t)nc * 8, rows, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) ok = false;
        return dfull + c0;
    }
    const double *up_banded(const double *h, size_t nb, size_t rows)   // host (nb, ncol, rows) -> device (nb, nc, rows)
    {
        if (!h) return nullptr;
        double *d = take(nb * rows);
        const cudaError_t e = (nc == ncol)
            ? cudaMemcpyAsync(d, h, (size_t)nc * nb * rows * 8, cudaMemcpyHostToDevice, st)
            : cudaMemcpy2DAsync(d, (size_t)nc * nb * 8, h + nb * c0, (size_t)ncol * nb * 8, (size_t)nc * nb * 8, rows, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) ok = false;
        return d;
    }
    void down(double *h, const double *d, size_t rows)          // h == NULL: output not wanted
    {
        if (!h) return;
        const cudaError_t e = (nc == ncol)
            ? cudaMemcpyAsync(h, d, (size_t)nc * rows * 8, cudaMemcpyDeviceToHost, st)
            : cudaMemcpy2DAsync(h + c0, (size_t)ncol * 8, d, (size_t)nc * 8, (size_t)nc * 8, rows, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) ok = false;
    }
};

int host_slots(int nblocks)
{
    const int n = G.host_slots < 2 ? 2 : (G.host_slots > MAXSLOT ? MAXSLOT : G.host_slots);
    return nblocks < n ? (nblocks < 1 ? 1 : nblocks) : n;
}

int host_chunk(int ncol)
{
    if (G.capture) return ncol;                        // stage dumps need the whole batch in one pass
    // measured e2e at T170L60 (131072 columns): 4096 53.6, 8192 48.0, 16384 47.3, 32768 50.0 ms.  A batch of fewer than four
    // such blocks (a rank's share in strong scaling: 16384 columns at 8 GPUs) is cut into about four, so that copies and
    // kernels still overlap: 8 GPUs x 16384 columns 9.2 ms in one block, 7.4 ms in blocks of 4096 (profiles/r02_summary.md)
    int hc = G.host_chunk > 0 ? G.host_chunk : 16384;
    if (G.host_chunk <= 0 && ncol < 4 * hc) {
        hc = ((ncol + 3) / 4 + 1023) / 1024 * 1024;
        if (hc < 4096) hc = 4096;
    }
    if (G.chunk > 0 && G.chunk < hc) hc = G.chunk;     // option "chunk" bounds every device pass
    return hc < ncol ? hc : ncol;
}

} // namespace

namespace rrtmg {
Tuning g_tune = {0, 0, 0, 4, 2, 4, {0, 3, 0, 0, 0, 0, 0, 0}, 1, 1, 0};   // taumol_sync: one block barrier per 4 bands (measured 1: 9.72, 2: 9.53, 4: 9.46, 16: 9.67 ms LW)
void ktimer_begin(int id, cudaStream_t s)
{
    if (!KT.on) return;
    KTimer::Rec r{id, KT.get(), KT.get()};
    cudaEventRecord(r.a, s);
    KT.recs.push_back(r);
}
void ktimer_end(cudaStream_t s)
{
    if (!KT.on) return;
    cudaEventRecord(KT.recs.back().b, s);
}
} // namespace rrtmg

// ------------------------------------------------------------------------------------------------
// radiation driver (run_rrtmg on the device)
namespace {
struct DrvSlot {                     // one stage of the host-pointer pipeline (and slot 0: the device-pointer entry)
    DevBuf in, out, host_in, host_out, lw_work, sw_work;
    cudaStream_t st = nullptr;
    cudaStream_t side = nullptr;         // LW runs here, next to SW on the main stream (fork/join by events)
    cudaEvent_t fork = nullptr, join = nullptr;
    int side_ready()
    {
        if (!side && cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) != cudaSuccess) return -1;
        if (!fork && cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) != cudaSuccess) return -1;
        if (!join && cudaEventCreateWithFlags(&join, cudaEventDisableTiming) != cudaSuccess) return -1;
        return 0;
    }
};
struct DrvState {
    DrvSlot slot[MAXSLOT];
    DevBuf gas, misc;
    size_t gas_n = 0;
    double gas_val[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    bool gas_set = false;
} D;

struct ZenithHost { ZenithArgs a; int dyofyr; int sec_l, day_l; };

// scalar part of compute_zenith (astro.f90:95-127) and Time_loc of run_rrtmg (rrtm_radiation.f90:550-558)
ZenithHost zenith_scalars(const rrtmg_b200_rad_config &c, int seconds, int days, int dt)
{
    const double PI = 3.14159265358979323846;
    ZenithHost z;
    const double deg2rad = PI / 180.;
    const int daysperyear = c.days_per_year;
    const double radpersec = 2 * PI / 86400.;
    const double radperday = 2 * PI / daysperyear;
    z.a.radpersec = radpersec;
    z.a.radsec = seconds * radpersec;
    z.a.dt_pi = dt * radpersec;
    z.a.dt = dt;
    int d = days - (int)(c.equinox_day * daysperyear);
    int dy = d % daysperyear;
    if (dy < 0) dy += daysperyear;                    // modulo
    z.dyofyr = dy;
    const double radday = dy * radperday;
    z.a.dec_sin = std::sin(c.obliq * deg2rad) * std::sin(radday);
    const double dec = std::asin(z.a.dec_sin);
    z.a.dec_cos = std::cos(dec);
    z.a.dec_tan = std::tan(dec);
    return z;
}
void local_time(const rrtmg_b200_rad_config &c, int seconds, int days, int &sec_l, int &day_l)
{
    sec_l = seconds; day_l = days;
    if (c.solday > 0) { day_l = c.solday; return; }
    if (c.slowdown_rad != 1.0) {
        long long tot = (long long)days * 86400 + seconds;
        tot = (long long)((double)tot * c.slowdown_rad);          // int(seconds*slowdown_rad)
        sec_l = (int)(tot % 86400);
        day_l = (int)(tot / 86400);
    }
}

int run_rrtmg_device_impl(const rrtmg_b200_rad_config &c, int si, int sj, int sk, int seconds, int days,
                          const double *lat, const double *lon, const double *p_full, const double *p_half,
                          const double *albedo, const double *q, const double *t, const double *t_surf,
                          const double *z_full, const double *z_half, const double *t_half_in, const double *o3f,
                          double *tdt, double *coszen, double *flux_sw, double *flux_lw, double *tdt_rad,
                          double *tdt_sw, double *tdt_lw, double *olr, double *isr, double *t_half_out,
                          cudaStream_t st, DrvSlot &S, bool own_work, int top_flag = -1)
{
    if (!G.lw_ready || !G.sw_ready) return fail(RRTMG_B200_ERR_NOT_INITIALIZED, "run_rrtmg: rrtmg_b200_lw_init / sw_init have not been called");
    if (si < 1 || sj < 1 || sk < 2 || sk > MAXLAY) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "run_rrtmg: grid extents out of range");
    if (c.lonstep < 1 || si % c.lonstep != 0) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "run_rrtmg: lonstep must divide the number of longitudes");
    if (c.days_per_year < 1) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "run_rrtmg: days_per_year must be positive");
    // astro.f90:99-104: `use_dyofyr is TRUE but the calendar year does not have 365 days. STOPPING` (FATAL)
    if (c.use_dyofyr && c.days_per_year != 365)
        return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "run_rrtmg: use_dyofyr needs a 365-day calendar (astro.f90:99-104)");
    // rrtm_radiation_init resolves dt_rad_avg < 0 to dt_rad (:336-340); dt_rad is not part of this interface, so the
    // caller must pass the resolved value
    if (c.do_rad_time_avg && c.dt_rad_avg < 0)
        return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "run_rrtmg: dt_rad_avg < 0 must be resolved to dt_rad by the caller (rrtm_radiation.f90:336-340)");
    if (!lat || !lon || !p_full || !p_half || !albedo || !q || !t || !t_surf || !coszen)
        return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "run_rrtmg: required array is NULL");
    if (!t_half_in && (!z_full || !z_half)) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "run_rrtmg: need t_half or z_full + z_half");
    RadGeom g{si, sj, sk, c.lonstep, si / c.lonstep, (si / c.lonstep) * sj};
    const size_t np = (size_t)si * sj, L = sk, V = sk + 1, nc = g.ncols;
    auto pad = [](size_t n) { return (n * 8 + 255) & ~(size_t)255; };
    // packed RRTMG inputs + t_half + zonal-mean scratch
    const size_t in_bytes = 4 * pad(nc * L) + 2 * pad(nc * V) + 3 * pad(nc) + pad(np * V) + pad((size_t)sj * sk + 2 * sj) + pad(sj * L) + 256;
    const size_t out_bytes = 4 * pad(nc * V) + 2 * pad(nc * L) + 4 * pad(nc * V) + 2 * pad(nc * L);
    if (S.in.ensure(in_bytes) || S.out.ensure(out_bytes)) return fail(RRTMG_B200_ERR_CUDA, "cudaMalloc failed (run_rrtmg buffers)");
    Carver ci(S.in.p), co(S.out.p);
    double *pfull = ci.take<double>(nc * L), *tfull = ci.take<double>(nc * L), *h2o = ci.take<double>(nc * L), *o3 = ci.take<double>(nc * L);
    double *phalf = ci.take<double>(nc * V), *thalf = ci.take<double>(nc * V);
    double *cosz_rr = ci.take<double>(nc), *albedo_rr = ci.take<double>(nc), *tsrf = ci.take<double>(nc);
    double *t_half_buf = ci.take<double>(np * V);
    double *zm_buf = ci.take<double>((size_t)sj * sk + 2 * sj);
    double *qzm_buf = ci.take<double>((size_t)sj * L);
    int *flag = ci.take<int>(1);
    double *swuflx = co.take<double>(nc * V), *swdflx = co.take<double>(nc * V), *swuflxc = co.take<double>(nc * V), *swdflxc = co.take<double>(nc * V);
    double *swhr = co.take<double>(nc * L), *swhrc = co.take<double>(nc * L);
    double *uflx = co.take<double>(nc * V), *dflx = co.take<double>(nc * V), *uflxc = co.take<double>(nc * V), *dflxc = co.take<double>(nc * V);
    double *hr = co.take<double>(nc * L), *hrc = co.take<double>(nc * L);
    // constant gas arrays: co2, and the seven secondary gases when requested (RR/rrtm_radiation.f90:360, 679-748)
    const double gv[8] = {c.co2ppmv * 1.e-6, c.ch4_val, c.n2o_val, c.o2_val, c.cfc11_val, c.cfc12_val, c.cfc22_val, c.ccl4_val};
    const int ngas = c.include_secondary_gases ? 8 : 1;
    // (a prefix of a constant array is the same constant array: the capacity only grows, and the arrays are shared by
    // the pipeline stages; the fill is ordered before every consumer by a device-wide sync, done only when values change)
    bool refill = !D.gas_set || D.gas_n < nc * L;
    for (int i = 0; i < 8; ++i) refill = refill || D.gas_val[i] != gv[i];
    const size_t gcap = refill ? (nc * L > D.gas_n ? nc * L : D.gas_n) : D.gas_n;
    if (refill) {
        CUDA_OK(cudaDeviceSynchronize());
        if (D.gas.ensure(8 * pad(gcap))) return fail(RRTMG_B200_ERR_CUDA, "cudaMalloc failed (gas arrays)");
    }
    Carver cg(D.gas.p);
    double *gas[8];
    for (int i = 0; i < 8; ++i) gas[i] = cg.take<double>(gcap);
    if (refill) {
        for (int i = 0; i < 8; ++i) { G.launches += drv_fill(gas[i], gcap, gv[i], st); D.gas_val[i] = gv[i]; }
        CUDA_OK(cudaStreamSynchronize(st));
        D.gas_n = gcap; D.gas_set = true;
    }
    // Time_loc, zenith angle (also an output)
    int sec_l, day_l;
    local_time(c, seconds, days, sec_l, day_l);
    const int dt = c.do_rad_time_avg ? c.dt_rad_avg : 0;
    ZenithHost z = zenith_scalars(c, sec_l, day_l, dt);
    G.launches += drv_zenith((int)np, lat, lon, coszen, z.a, st);
    const int dyofyr = c.use_dyofyr ? z.dyofyr : 0;                                    // :592
    const double *t_half = t_half_in;
    if (!t_half) {
        double *dst = t_half_out ? t_half_out : t_half_buf;
        G.launches += drv_interp_temp((int)np, sk, z_full, z_half, t_surf, t, dst, st);
        t_half = dst;
    } else if (t_half_out && t_half_out != t_half_in) {
        CUDA_OK(cudaMemcpyAsync(t_half_out, t_half_in, np * V * 8, cudaMemcpyDeviceToDevice, st));
    }
    PackArgs pa{};
    pa.p_full = p_full; pa.p_half = p_half; pa.t = t; pa.t_half = t_half; pa.q = q; pa.o3f = o3f; pa.coszen = coszen;
    pa.albedo = albedo; pa.t_surf = t_surf; pa.lat = lat; pa.qzm = nullptr; pa.top_flag = nullptr;
    pa.pfull = pfull; pa.phalf = phalf; pa.tfull = tfull; pa.thalf = thalf; pa.h2o = h2o; pa.o3 = o3;
    pa.cosz_rr = cosz_rr; pa.albedo_rr = albedo_rr; pa.tsrf = tsrf;
    pa.qmin = c.h2o_lower_limit; pa.tmin = c.temp_lower_limit; pa.tmax = c.temp_upper_limit;
    pa.scale_ozone = c.scale_ozone; pa.o3_val = c.o3_val;
    pa.do_fixed_water = c.do_fixed_water; pa.fixed_water = c.fixed_water; pa.fixed_water_pres = c.fixed_water_pres;
    pa.fixed_water_lat = c.fixed_water_lat;
    G.launches += drv_pack(g, pa, c.do_zm_tracers ? qzm_buf : nullptr, flag, st, top_flag);
    // the two RRTMG calls with MiMA's fixed switches (icld = iaer = idrv = 0, emis = 1, tauaer = 0)
    int icld = 0, iaer = 0;
    SwIn sin{(int)nc, pfull, phalf, tfull, thalf, tsrf, h2o, o3, gas[0],
             ngas > 1 ? gas[1] : nullptr, ngas > 1 ? gas[2] : nullptr, ngas > 1 ? gas[3] : nullptr,
             albedo_rr, albedo_rr, albedo_rr, albedo_rr, cosz_rr, sw_adjflux(c.solrad, dyofyr, c.solr_cnst)};
    SwOut sout{(int)nc, swuflx, swdflx, swhr, swuflxc, swdflxc, swhrc};
    // SW on the caller's stream, LW on a side stream: the FP64-bound SW solver and the L1-bound LW kernels overlap
    // at the kernel boundaries (41.9 vs 43.2 ms per T170L60 step)
    if (S.side_ready()) return fail(RRTMG_B200_ERR_CUDA, "cudaStreamCreate failed (run_rrtmg side stream)");
    CUDA_OK(cudaEventRecord(S.fork, st));
    CUDA_OK(cudaStreamWaitEvent(S.side, S.fork, 0));
    if (const int rc = sw_device_impl((int)nc, sk, &icld, &iaer, sin, sout, st, own_work ? &S.sw_work : nullptr)) return rc;
    LwIn lin{(int)nc, pfull, phalf, tfull, thalf, tsrf, h2o, o3, gas[0],
             ngas > 1 ? gas[1] : nullptr, ngas > 1 ? gas[2] : nullptr, ngas > 1 ? gas[3] : nullptr,
             ngas > 1 ? gas[4] : nullptr, ngas > 1 ? gas[5] : nullptr, ngas > 1 ? gas[6] : nullptr, ngas > 1 ? gas[7] : nullptr,
             nullptr, nullptr};
    LwOut lout{(int)nc, uflx, dflx, hr, uflxc, dflxc, hrc};
    if (const int rc = lw_device_impl((int)nc, sk, &icld, 0, lin, lout, S.side, own_work ? &S.lw_work : nullptr)) return rc;
    CUDA_OK(cudaEventRecord(S.join, S.side));
    CUDA_OK(cudaStreamWaitEvent(st, S.join, 0));
    UnpackArgs ua{swhr, swuflx, swdflx, hr, uflx, dflx, tdt, tdt_rad, tdt_sw, tdt_lw, flux_sw, flux_lw, olr, isr};
    G.launches += drv_unpack(g, ua, c.do_zm_rad ? zm_buf : nullptr, st);
    CUDA_OK(cudaGetLastError());
    return RRTMG_B200_OK;
}
} // namespace

// =================================================================================================
extern "C" {

/* Accumulated device time [ms] and launch count per kernel since the last reset (needs option
 * "kernel_timing" = 1).  Order: lw_prep, lw_taumol, lw_rtrn, sw_prep, sw_taumol, sw_solver, lw_column, sw_column (arrays
 * of 8; lw_column / sw_column = the fused clear-sky kernels with their flux kernels, which replace taumol and the solver
 * when they apply). */
int rrtmg_b200_kernel_times(double *ms, long *launches, int reset)
{
    KT.collect();
    for (int i = 0; i < K_COUNT; ++i) {
        if (ms) ms[i] = KT.ms[i];
        if (launches) launches[i] = KT.n[i];
        if (reset) { KT.ms[i] = 0.0; KT.n[i] = 0; }
    }
    return RRTMG_B200_OK;
}

const char *rrtmg_b200_last_error(void) { return G.err.c_str(); }
long rrtmg_b200_launch_count(void) { return G.launches; }

int rrtmg_b200_set_device(int local_rank)
{
    int n = 0;
    CUDA_OK(cudaGetDeviceCount(&n));
    if (n <= 0) return fail(RRTMG_B200_ERR_CUDA, "no CUDA device");
    CUDA_OK(cudaSetDevice(((local_rank % n) + n) % n));
    return RRTMG_B200_OK;
}

int rrtmg_b200_set_table(const char *name, const double *data, int ndim, const int *dims)
{
    if (!name || !data || ndim < 1 || ndim > 4 || !dims) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "set_table: bad argument");
    std::lock_guard<std::mutex> lk(G.mu);
    HostArr a;
    a.dims.assign(dims, dims + ndim);
    for (int d : a.dims) if (d < 1) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "set_table: non-positive dimension");
    a.data.assign(data, data + a.size());
    G.reg[name] = std::move(a);
    return RRTMG_B200_OK;
}

int rrtmg_b200_load_tables(const char *path)
{
    FILE *f = path ? std::fopen(path, "rb") : nullptr;
    if (!f) return fail(RRTMG_B200_ERR_TABLES, std::string("cannot open ") + (path ? path : "(null)"));
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<unsigned char> raw((size_t)sz);
    const bool ok = std::fread(raw.data(), 1, (size_t)sz, f) == (size_t)sz;
    std::fclose(f);
    if (!ok || sz < 12 || std::memcmp(raw.data(), "RRTMGTB1", 8) != 0) return fail(RRTMG_B200_ERR_TABLES, "not an RRTMGTB1 blob");
    uint32_t n;
    std::memcpy(&n, raw.data() + 8, 4);
    const size_t recsz = 60;
    if ((size_t)sz < 12 + n * recsz) return fail(RRTMG_B200_ERR_TABLES, "truncated blob");
    const unsigned char *data = raw.data() + 12 + n * recsz;
    for (uint32_t i = 0; i < n; ++i) {
        const unsigned char *p = raw.data() + 12 + i * recsz;
        char name[33];
        std::memcpy(name, p, 32);
        name[32] = 0;
        uint32_t nd, d[4];
        uint64_t off;
        std::memcpy(&nd, p + 32, 4);
        std::memcpy(d, p + 36, 16);
        std::memcpy(&off, p + 52, 8);
        if (nd < 1 || nd > 4) return fail(RRTMG_B200_ERR_TABLES, "bad record in blob");
        int dims[4];
        size_t cnt = 1;
        for (uint32_t k = 0; k < nd; ++k) { dims[k] = (int)d[k]; cnt *= d[k]; }
        if (data + (off + cnt) * 8 > raw.data() + sz) return fail(RRTMG_B200_ERR_TABLES, "blob record out of range");
        std::vector<double> tmp(cnt);
        std::memcpy(tmp.data(), data + off * 8, cnt * 8);
        const int rc = rrtmg_b200_set_table(name, tmp.data(), (int)nd, dims);
        if (rc) return rc;
    }
    return RRTMG_B200_OK;
}

int rrtmg_b200_lw_init(double cpdair)
{
    std::lock_guard<std::mutex> lk(G.mu);
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return fail(RRTMG_B200_ERR_CUDA, "no CUDA device: this library has no CPU path");
    return lw_init_impl(cpdair);
}
int rrtmg_b200_sw_init(double cpdair)
{
    std::lock_guard<std::mutex> lk(G.mu);
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return fail(RRTMG_B200_ERR_CUDA, "no CUDA device: this library has no CPU path");
    return sw_init_impl(cpdair);
}

int rrtmg_b200_finalize(void)
{
    std::lock_guard<std::mutex> lk(G.mu);
    for (DevBuf *b : {&G.lw_slices, &G.sw_slices, &G.lw_tab, &G.lw_totplnk, &G.lw_exptfn, &G.lw_work, &G.lw_cap, &G.lw_err, &G.sw_tab, &G.sw_exptbl, &G.sw_work, &G.sw_err})
        b->release();
    P_lw.release();
    P_sw.release();
    for (DrvSlot &S : D.slot) {
        for (DevBuf *b : {&S.in, &S.out, &S.host_in, &S.host_out, &S.lw_work, &S.sw_work}) b->release();
        if (S.st) { cudaStreamDestroy(S.st); S.st = nullptr; }
        if (S.side) { cudaStreamDestroy(S.side); S.side = nullptr; }
        if (S.fork) { cudaEventDestroy(S.fork); S.fork = nullptr; }
        if (S.join) { cudaEventDestroy(S.join); S.join = nullptr; }
    }
    D.gas.release(); D.misc.release();
    G.shared.buf.release(); G.shared.valid = false;
    g_fit_lw = FitCache(); g_fit_sw = FitCache();
    D.gas_set = false; D.gas_n = 0;
    G.lw_ready = G.sw_ready = false;
    G.lw_last_ncol = G.sw_last_ncol = 0;
    G.reduced.clear();
    G.reg.clear();            // registered coefficient arrays: a new init starts from what is registered after this call
    return RRTMG_B200_OK;
}

long rrtmg_b200_get_table(const char *name, double *out, long capacity)
{
    auto it = G.reduced.find(name ? name : "");
    if (it == G.reduced.end()) return -1;
    const long n = it->second.size();
    if (out && capacity >= n) std::memcpy(out, it->second.data.data(), (size_t)n * 8);
    return n;
}

int rrtmg_b200_set_chunk(int ncol_per_pass)
{
    if (ncol_per_pass < 0) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "chunk must be >= 0");
    G.chunk = ncol_per_pass;
    return RRTMG_B200_OK;
}

int rrtmg_b200_lw_device(int ncol, int nlay, int *icld, int idrv,
                         const double *play, const double *plev, const double *tlay, const double *tlev,
                         const double *tsfc, const double *h2ovmr, const double *o3vmr, const double *co2vmr,
                         const double *ch4vmr, const double *n2ovmr, const double *o2vmr,
                         const double *cfc11vmr, const double *cfc12vmr, const double *cfc22vmr,
                         const double *ccl4vmr, const double *emis,
                         int inflglw, int iceflglw, int liqflglw, const double *cldfr, const double *taucld,
                         const double *cicewp, const double *cliqwp, const double *reice, const double *reliq,
                         const double *tauaer,
                         double *uflx, double *dflx, double *hr, double *uflxc, double *dflxc, double *hrc,
                         double *duflx_dt, double *duflxc_dt, void *stream)
{
    if (!play || !plev || !tlay || !tlev || !tsfc || !h2ovmr || !o3vmr || !co2vmr || !uflx || !dflx || !hr ||
        !uflxc || !dflxc || !hrc)
        return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "rrtmg_lw: required array is NULL");
    LwIn in{ncol, play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr,
            cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr, emis, tauaer};
    const LwOpt opt{inflglw, cldfr, taucld, iceflglw, liqflglw, cicewp, cliqwp, reice, reliq};
    if (const int rc = lw_validate(ncol, nlay, icld, idrv, opt)) return rc;       // also normalises *icld
    lw_set_optional(in, icld, opt);
    LwOut out{ncol, uflx, dflx, hr, uflxc, dflxc, hrc};
    if (idrv == 1) { out.duflx_dt = duflx_dt; out.duflxc_dt = duflxc_dt; }
    return lw_device_impl(ncol, nlay, icld, idrv, in, out, (cudaStream_t)stream, nullptr, opt);
}

int rrtmg_b200_lw(int ncol, int nlay, int *icld, int idrv,
                  const double *play, const double *plev, const double *tlay, const double *tlev,
                  const double *tsfc, const double *h2ovmr, const double *o3vmr, const double *co2vmr,
                  const double *ch4vmr, const double *n2ovmr, const double *o2vmr,
                  const double *cfc11vmr, const double *cfc12vmr, const double *cfc22vmr,
                  const double *ccl4vmr, const double *emis,
                  int inflglw, int iceflglw, int liqflglw, const double *cldfr,
                  const double *taucld, const double *cicewp, const double *cliqwp,
                  const double *reice, const double *reliq, const double *tauaer,
                  double *uflx, double *dflx, double *hr, double *uflxc, double *dflxc, double *hrc,
                  double *duflx_dt, double *duflxc_dt)
{
    const LwOpt opt{inflglw, cldfr, taucld, iceflglw, liqflglw, cicewp, cliqwp, reice, reliq};
    // uflxc, dflxc, hrc (and duflxc_dt) may be NULL: the clear-sky result is then not copied back (MiMA never reads it)
    if (!play || !plev || !tlay || !tlev || !tsfc || !h2ovmr || !o3vmr || !co2vmr || !uflx || !dflx || !hr)
        return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "rrtmg_lw: required array is NULL");
    if (const int rc = lw_validate(ncol, nlay, icld, idrv, opt)) return rc;
    if (idrv == 1 && !duflx_dt) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "rrtmg_lw: idrv = 1 needs duflx_dt");
    if (ncol == 0) return RRTMG_B200_OK;
    const bool cloud = icld && *icld >= 1;
    if (const int rc = lw_err_begin(cloud)) return rc;
    if (P_lw.ready()) return fail(RRTMG_B200_ERR_CUDA, "cudaStreamCreate failed");
    const int hc = host_chunk(ncol);
    const bool fields = G.capture;
    const size_t L = nlay, V = nlay + 1;
    const size_t in_bytes = (size_t)hc * (13 * L + 2 * V + 1 + 16 + 16 * L + (cloud ? 21 * L : 0)) * 8 + 40 * 256;
    const size_t out_bytes = (size_t)hc * (6 * V + 2 * L) * 8 + 10 * 256;
    LwWork wsz;
    const size_t work_bytes = lw_carve(wsz, nullptr, hc, nlay, fields, cloud);
    int nslot = host_slots((ncol + hc - 1) / hc);
    for (int i = 0; i < nslot; ++i)
        if (P_lw.in[i].ensure(in_bytes) || P_lw.out[i].ensure(out_bytes) || P_lw.work[i].ensure(work_bytes)) {
            if (i < 2) return fail(RRTMG_B200_ERR_CUDA, "cudaMalloc failed for the LW pipeline buffers");
            nslot = P_lw.shrink(i);                     // no room for a deeper pipeline: run with the slots that fit
        }
    // inputs left on the device by the rrtmg_b200_sw call just before (option share_inputs)?
    const double *hp[11] = {play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr};
    bool shared = G.share_inputs && G.shared.valid && G.shared.ncol == ncol && G.shared.nlay == nlay;
    for (int k = 0; shared && k < 11; ++k) shared = G.shared.host[k] == hp[k];
    // the interface arrays of a block share one leading dimension: with LW-only array inputs (uploaded per block) present
    // the shared full-batch copies are not used
    if (cfc11vmr || cfc12vmr || cfc22vmr || ccl4vmr || emis || tauaer || cloud || fields) shared = false;
    G.shared.valid = false;
    int idx = 0;
    for (int c0 = 0; c0 < ncol; c0 += hc, ++idx) {
        const int nc = (ncol - c0 < hc) ? ncol - c0 : hc;
        const int slot = idx % nslot;
        cudaStream_t st = P_lw.st[slot];
        Slot a{(char *)P_lw.in[slot].p, 0, c0, nc, ncol, st, true};
        auto in11 = [&](int k, const double *h, size_t rows) { return shared ? (h ? G.shared.dev[k] + c0 : nullptr) : a.up(h, rows); };
        LwIn in{shared ? ncol : nc, in11(0, play, L), in11(1, plev, V), in11(2, tlay, L), in11(3, tlev, V), in11(4, tsfc, 1),
                in11(5, h2ovmr, L), in11(6, o3vmr, L), in11(7, co2vmr, L), in11(8, ch4vmr, L), in11(9, n2ovmr, L), in11(10, o2vmr, L),
                a.up(cfc11vmr, L), a.up(cfc12vmr, L), a.up(cfc22vmr, L), a.up(ccl4vmr, L), a.up(emis, 16), a.up(tauaer, 16 * L)};
        if (cloud) {
            LwOpt dopt{inflglw, a.up(cldfr, L), a.up_banded(taucld, 16, L), iceflglw, liqflglw,
                       inflglw >= 1 ? a.up(cicewp, L) : nullptr, inflglw >= 1 ? a.up(cliqwp, L) : nullptr,
                       inflglw == 2 ? a.up(reice, L) : nullptr, inflglw == 2 ? a.up(reliq, L) : nullptr};
            lw_set_optional(in, icld, dopt);
        }
        if (!a.ok) return drain(P_lw, fail(RRTMG_B200_ERR_CUDA, "H2D copy failed (LW)"));
        Slot o{(char *)P_lw.out[slot].p, 0, c0, nc, ncol, st, true};
        LwOut out{nc, o.take(V), o.take(V), o.take(L), o.take(V), o.take(V), o.take(L)};
        if (idrv == 1) { out.duflx_dt = o.take(V); out.duflxc_dt = o.take(V); }
        if (const int rc = lw_chunk(in, out, nc, nlay, P_lw.work[slot].p, fields, st, c0 + nc >= ncol)) return drain(P_lw, rc);
        if (idrv == 1) { o.down(duflx_dt, out.duflx_dt, V); o.down(duflxc_dt, out.duflxc_dt, V); }
        o.down(uflx, out.uflx, V); o.down(dflx, out.dflx, V); o.down(hr, out.hr, L);
        o.down(uflxc, out.uflxc, V); o.down(dflxc, out.dflxc, V); o.down(hrc, out.hrc, L);      // skipped when NULL
        if (!o.ok) return drain(P_lw, fail(RRTMG_B200_ERR_CUDA, "D2H copy failed (LW)"));
    }
    for (int i = 0; i < nslot; ++i) CUDA_OK(cudaStreamSynchronize(P_lw.st[i]));
    return lw_err_end(cloud);
}

int rrtmg_b200_sw_device(int ncol, int nlay, int *icld, int *iaer,
                         const double *play, const double *plev, const double *tlay, const double *tlev,
                         const double *tsfc, const double *h2ovmr, const double *o3vmr, const double *co2vmr,
                         const double *ch4vmr, const double *n2ovmr, const double *o2vmr,
                         const double *asdir, const double *asdif, const double *aldir, const double *aldif,
                         const double *coszen, double adjes, int dyofyr, double scon,
                         int inflgsw, int iceflgsw, int liqflgsw, const double *cldfr,
                         const double *taucld, const double *ssacld, const double *asmcld, const double *fsfcld,
                         const double *cicewp, const double *cliqwp, const double *reice, const double *reliq,
                         const double *tauaer, const double *ssaaer, const double *asmaer, const double *ecaer,
                         double *swuflx, double *swdflx, double *swhr, double *swuflxc, double *swdflxc,
                         double *swhrc, void *stream)
{
    if (!play || !plev || !tlay || !tlev || !tsfc || !h2ovmr || !o3vmr || !co2vmr || !asdir || !asdif || !aldir ||
        !aldif || !coszen || !swuflx || !swdflx || !swhr || !swuflxc || !swdflxc || !swhrc)
        return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "rrtmg_sw: required array is NULL");
    SwOpt opt{inflgsw, cldfr, taucld, ssacld, asmcld, fsfcld, tauaer, ssaaer, asmaer, ecaer, iceflgsw, liqflgsw, cicewp, cliqwp, reice, reliq};
    if (const int rc = sw_validate(ncol, nlay, icld, iaer, opt)) return rc;       // also normalises *icld, *iaer
    SwIn in{ncol, play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr,
            asdir, asdif, aldir, aldif, coszen, sw_adjflux(adjes, dyofyr, scon)};
    sw_set_optional(in, icld, iaer, opt);
    SwOut out{ncol, swuflx, swdflx, swhr, swuflxc, swdflxc, swhrc};
    return sw_device_impl(ncol, nlay, icld, iaer, in, out, (cudaStream_t)stream, nullptr, opt);
}

int rrtmg_b200_sw(int ncol, int nlay, int *icld, int *iaer,
                  const double *play, const double *plev, const double *tlay, const double *tlev,
                  const double *tsfc, const double *h2ovmr, const double *o3vmr, const double *co2vmr,
                  const double *ch4vmr, const double *n2ovmr, const double *o2vmr,
                  const double *asdir, const double *asdif, const double *aldir, const double *aldif,
                  const double *coszen, double adjes, int dyofyr, double scon,
                  int inflgsw, int iceflgsw, int liqflgsw, const double *cldfr,
                  const double *taucld, const double *ssacld, const double *asmcld, const double *fsfcld,
                  const double *cicewp, const double *cliqwp, const double *reice, const double *reliq,
                  const double *tauaer, const double *ssaaer, const double *asmaer, const double *ecaer,
                  double *swuflx, double *swdflx, double *swhr, double *swuflxc, double *swdflxc, double *swhrc)
{
    const SwOpt opt{inflgsw, cldfr, taucld, ssacld, asmcld, fsfcld, tauaer, ssaaer, asmaer, ecaer, iceflgsw, liqflgsw, cicewp, cliqwp, reice, reliq};
    // swuflxc, swdflxc, swhrc may be NULL: the clear-sky result is then not copied back
    if (!play || !plev || !tlay || !tlev || !tsfc || !h2ovmr || !o3vmr || !co2vmr || !asdir || !asdif || !aldir ||
        !aldif || !coszen || !swuflx || !swdflx || !swhr)
        return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "rrtmg_sw: required array is NULL");
    if (const int rc = sw_validate(ncol, nlay, icld, iaer, opt)) return rc;
    if (ncol == 0) return RRTMG_B200_OK;
    if (P_sw.ready()) return fail(RRTMG_B200_ERR_CUDA, "cudaStreamCreate failed");
    const bool cloud = icld && *icld >= 1, aer = iaer && *iaer == 10, aer6 = iaer && *iaer == 6, general = cloud || aer || aer6;
    if (const int rc = sw_err_begin(general)) return rc;
    const int hc = host_chunk(ncol);
    const bool fields = G.capture;
    const size_t L = nlay, V = nlay + 1;
    const size_t in_bytes = (size_t)hc * (9 * L + 2 * V + 6 + (cloud ? 57 * L : 0) + (aer ? 42 * L : 0) + (aer6 ? 6 * L : 0)) * 8 + 56 * 256;
    const size_t out_bytes = (size_t)hc * (4 * V + 2 * L) * 8 + 8 * 256;
    SwWork wsz;
    const size_t work_bytes = sw_carve(wsz, nullptr, hc, nlay, fields, general);
    int nslot = host_slots((ncol + hc - 1) / hc);
    for (int i = 0; i < nslot; ++i)
        if (P_sw.in[i].ensure(in_bytes) || P_sw.out[i].ensure(out_bytes) || P_sw.work[i].ensure(work_bytes)) {
            if (i < 2) return fail(RRTMG_B200_ERR_CUDA, "cudaMalloc failed for the SW pipeline buffers");
            nslot = P_sw.shrink(i);                     // no room for a deeper pipeline: run with the slots that fit
        }
    const double adjflux = sw_adjflux(adjes, dyofyr, scon);
    // option share_inputs: the eleven arrays rrtmg_lw reads too are uploaded into full-batch device arrays and kept
    const bool share = G.share_inputs && !general && !fields;
    G.shared.valid = false;
    if (share) {
        const size_t rows[11] = {L, V, L, V, 1, L, L, L, L, L, L};
        const double *hp[11] = {play, plev, tlay, tlev, tsfc, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr};
        size_t total = 0;
        for (int k = 0; k < 11; ++k) total += hp[k] ? (((size_t)ncol * rows[k] * 8 + 255) & ~(size_t)255) : 0;
        if (G.shared.buf.ensure(total + 256)) return fail(RRTMG_B200_ERR_CUDA, "cudaMalloc failed for the shared inputs");
        size_t off = 0;
        for (int k = 0; k < 11; ++k) {
            G.shared.host[k] = hp[k];
            G.shared.dev[k] = hp[k] ? (double *)((char *)G.shared.buf.p + off) : nullptr;
            off += hp[k] ? (((size_t)ncol * rows[k] * 8 + 255) & ~(size_t)255) : 0;
        }
        G.shared.ncol = ncol; G.shared.nlay = nlay;
    }
    int idx = 0;
    for (int c0 = 0; c0 < ncol; c0 += hc, ++idx) {
        const int nc = (ncol - c0 < hc) ? ncol - c0 : hc;
        const int slot = idx % nslot;
        cudaStream_t st = P_sw.st[slot];
        Slot a{(char *)P_sw.in[slot].p, 0, c0, nc, ncol, st, true};
        auto in11 = [&](int k, const double *h, size_t rows) { return share ? a.up_full(h, G.shared.dev[k], rows) : a.up(h, rows); };
        const double *d_play = in11(0, play, L), *d_plev = in11(1, plev, V), *d_tlay = in11(2, tlay, L), *d_tlev = in11(3, tlev, V);
        const double *d_tsfc = in11(4, tsfc, 1), *d_h2o = in11(5, h2ovmr, L), *d_o3 = in11(6, o3vmr, L), *d_co2 = in11(7, co2vmr, L);
        const double *d_ch4 = in11(8, ch4vmr, L), *d_n2o = in11(9, n2ovmr, L), *d_o2 = in11(10, o2vmr, L);
        // MiMA passes the same albedo array four times (rrtm_radiation.f90:690): upload once per distinct pointer
        const double *d_asdir = a.up(asdir, 1);
        const double *d_asdif = asdif == asdir ? d_asdir : a.up(asdif, 1);
        const double *d_aldir = aldir == asdir ? d_asdir : a.up(aldir, 1);
        const double *d_aldif = aldif == asdir ? d_asdir : (aldif == aldir ? d_aldir : a.up(aldif, 1));
        const double *d_cosz = a.up(coszen, 1);
        SwOpt dopt;
        if (cloud) {
            dopt.cldfr = a.up(cldfr, L);
            dopt.inflgsw = inflgsw; dopt.iceflgsw = iceflgsw; dopt.liqflgsw = liqflgsw;
            if (inflgsw == 0) {
                dopt.taucld = a.up_banded(taucld, 14, L); dopt.ssacld = a.up_banded(ssacld, 14, L);
                dopt.asmcld = a.up_banded(asmcld, 14, L); dopt.fsfcld = a.up_banded(fsfcld, 14, L);
            } else {
                dopt.cicewp = a.up(cicewp, L); dopt.cliqwp = a.up(cliqwp, L); dopt.reice = a.up(reice, L); dopt.reliq = a.up(reliq, L);
            }
        }
        if (aer) { dopt.tauaer = a.up(tauaer, 14 * L); dopt.ssaaer = a.up(ssaaer, 14 * L); dopt.asmaer = a.up(asmaer, 14 * L); }
        if (aer6) dopt.ecaer = a.up(ecaer, 6 * L);
        if (!a.ok) return drain(P_sw, fail(RRTMG_B200_ERR_CUDA, "H2D copy failed (SW)"));
        SwIn in{share ? ncol : nc, d_play, d_plev, d_tlay, d_tlev, d_tsfc, d_h2o, d_o3, d_co2, d_ch4, d_n2o, d_o2,
                d_asdir, d_asdif, d_aldir, d_aldif, d_cosz, adjflux};
        sw_set_optional(in, icld, iaer, dopt);
        Slot o{(char *)P_sw.out[slot].p, 0, c0, nc, ncol, st, true};
        SwOut out{nc, o.take(V), o.take(V), o.take(L), o.take(V), o.take(V), o.take(L)};
        if (const int rc = sw_chunk(in, out, nc, nlay, P_sw.work[slot].p, fields, st, c0 + nc >= ncol)) return drain(P_sw, rc);
        o.down(swuflx, out.uflx, V); o.down(swdflx, out.dflx, V); o.down(swhr, out.hr, L);
        o.down(swuflxc, out.uflxc, V); o.down(swdflxc, out.dflxc, V); o.down(swhrc, out.hrc, L);
        if (!o.ok) return drain(P_sw, fail(RRTMG_B200_ERR_CUDA, "D2H copy failed (SW)"));
    }
    for (int i = 0; i < nslot; ++i) CUDA_OK(cudaStreamSynchronize(P_sw.st[i]));
    G.shared.valid = share;
    return sw_err_end(general);
}

// ---- stage dumps ---------------------------------------------------------------------------------
static long dump_field(const void *dev, size_t count, bool is_int, std::vector<double> &host)
{
    host.resize(count);
    if (is_int) {
        std::vector<int> tmp(count);
        if (cudaMemcpy(tmp.data(), dev, count * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        for (size_t i = 0; i < count; ++i) host[i] = tmp[i];
    } else if (cudaMemcpy(host.data(), dev, count * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (long)count;
}

long rrtmg_b200_get_stage(const char *which, double *out, long capacity)
{
    if (!which) return -1;
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    const std::string k(which);
    const bool lw = k.rfind("lw.", 0) == 0, sw = k.rfind("sw.", 0) == 0;
    if (!lw && !sw) return -1;
    const std::string f = k.substr(3);
    const int nc = lw ? G.lw_last_ncol : G.sw_last_ncol;
    if (nc <= 0) return -1;
    const int nlay = lw ? G.lw_last.nlay : G.sw_last.nlay;
    std::vector<double> h;
    long n = -1;
    // [lay][col] fields are already (ncol,nlay) column-major
    static const char *lwf[LF_COUNT] = {"fac00", "fac01", "fac10", "fac11", "colh2o", "colco2", "colo3", "coln2o", "colco",
                                        "colch4", "colo2", "colbrd", "selffac", "selffrac", "forfac", "forfrac", "minorfrac",
                                        "scaleminor", "scaleminorn2", "coldry", "pavel", "wx1", "wx2", "wx3", "wx4"};
    static const char *swf[SF_COUNT] = {"fac00", "fac01", "fac10", "fac11", "colh2o", "colco2", "colo3", "colch4", "colo2",
                                        "colmol", "coln2o", "selffac", "selffrac", "forfac", "forfrac"};
    const size_t np = (size_t)nc * nlay;
    const bool have_fields = lw ? G.lw_last.f != nullptr : G.sw_last.f != nullptr;
    auto transposed = [&](const double *dev, int inner, int mid) -> long {
        // device [col][mid][inner] -> host (ncol, mid, inner) column-major
        std::vector<double> t;
        if (dump_field(dev, (size_t)nc * mid * inner, false, t) < 0) return -1;
        h.resize(t.size());
        for (int c = 0; c < nc; ++c)
            for (int m = 0; m < mid; ++m)
                for (int i = 0; i < inner; ++i)
                    h[((size_t)i * mid + m) * nc + c] = t[((size_t)c * mid + m) * inner + i];
        return (long)h.size();
    };
    if (f == "laytrop") n = dump_field(lw ? (void *)G.lw_last.laytrop : (void *)G.sw_last.laytrop, nc, true, h);
    else if (f == "jp" || f == "jt" || f == "jt1" || f == "indself" || f == "indfor" || f == "indminor") {
        if (!have_fields) return -1;      // needs option capture_stages
        std::vector<uint32_t> tmp(np);
        if (cudaMemcpy(tmp.data(), lw ? G.lw_last.idx : G.sw_last.idx, np * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        h.resize(np);
        for (size_t i = 0; i < np; ++i) {
            const LwIdx ix = lw_unpack(tmp[i]);   // SW packing is a prefix of the LW packing
            h[i] = f == "jp" ? ix.jp : f == "jt" ? ix.jt : f == "jt1" ? ix.jt1 : f == "indself" ? ix.inds : f == "indfor" ? ix.indf : ix.indm;
        }
        n = (long)np;
    } else if (lw && f == "planklay") n = transposed(G.lw_last.planklay, 16, nlay);
    else if (lw && f == "planklev") n = transposed(G.lw_last.planklev, 16, nlay + 1);
    else if (lw && f == "plankbnd") n = transposed(G.lw_last.plankbnd, 16, 1);
    else if (lw && f == "secdiff") n = transposed(G.lw_last.secdiff, 16, 1);
    else if (lw && (f == "taug" || f == "fracs")) {
        if (!G.lw_cap.p) return -1;
        const double *base = (const double *)G.lw_cap.p + (f == "fracs" ? np * NGPTLW : 0);
        n = transposed(base, NGPTLW, nlay);
    } else if (sw && f == "taug") n = transposed(G.sw_last.taug, NGPTSW, nlay);
    else if (sw && f == "taur") { if (!G.sw_last.taur) return -1; n = transposed(G.sw_last.taur, NGPTSW, nlay); }
    else if (sw && f == "sfluxzen") n = transposed(G.sw_last.sfluxzen, NGPTSW, 1);
    else if (sw && f == "laysolfr") {
        std::vector<double> t;
        if (dump_field(G.sw_last.laysolfr, (size_t)nc * 14, true, t) < 0) return -1;
        h.resize(t.size());
        for (int c = 0; c < nc; ++c)
            for (int b = 0; b < 14; ++b) h[(size_t)b * nc + c] = t[(size_t)c * 14 + b];
        n = (long)h.size();
    } else {
        const int nf = lw ? (int)LF_COUNT : (int)SF_COUNT;
        if (!have_fields) return -1;
        for (int i = 0; i < nf; ++i)
            if (f == (lw ? lwf[i] : swf[i])) n = dump_field(lw ? G.lw_last.fld(i) : G.sw_last.fld(i), np, false, h);
    }
    if (n < 0) return -1;
    if (out && capacity >= n) std::memcpy(out, h.data(), (size_t)n * 8);
    return n;
}

void rrtmg_b200_rad_config_default(rrtmg_b200_rad_config *c)
{
    if (!c) return;
    std::memset(c, 0, sizeof *c);
    c->do_rad_time_avg = 1; c->dt_rad_avg = 86400; c->lonstep = 1; c->days_per_year = 360;
    c->scale_ozone = 1.0;
    c->h2o_lower_limit = 2.e-7; c->temp_lower_limit = 100.; c->temp_upper_limit = 370.;
    c->co2ppmv = 300.;
    c->fixed_water = 2.e-06; c->fixed_water_pres = 100.e02; c->fixed_water_lat = 90.;
    c->slowdown_rad = 1.0;
    c->obliq = 23.439; c->solr_cnst = 1368.22; c->solrad = 1.0; c->equinox_day = 0.25;
}

int rrtmg_b200_compute_zenith(const rrtmg_b200_rad_config *cfg, int seconds, int days, int dt, int n,
                              const double *lat, const double *lon, double *cosz, int *dyofyr)
{
    if (!cfg || !lat || !lon || !cosz || n < 0) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "compute_zenith: bad argument");
    if (cfg->days_per_year < 1) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "compute_zenith: days_per_year must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(RRTMG_B200_ERR_CUDA, "no CUDA device: this library has no CPU path");
    const ZenithHost z = zenith_scalars(*cfg, seconds, days, dt);
    if (dyofyr) *dyofyr = z.dyofyr;
    if (n == 0) return RRTMG_B200_OK;
    if (D.misc.ensure((size_t)n * 24)) return fail(RRTMG_B200_ERR_CUDA, "cudaMalloc failed (compute_zenith)");
    double *d = (double *)D.misc.p;
    CUDA_OK(cudaMemcpy(d, lat, (size_t)n * 8, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(d + n, lon, (size_t)n * 8, cudaMemcpyHostToDevice));
    G.launches += drv_zenith(n, d, d + n, d + 2 * (size_t)n, z.a, nullptr);
    CUDA_OK(cudaMemcpy(cosz, d + 2 * (size_t)n, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return RRTMG_B200_OK;
}

int rrtmg_b200_interp_temp(int si, int sj, int sk, const double *z_full, const double *z_half,
                           const double *t_surf_rad, const double *t, double *t_half)
{
    if (si < 1 || sj < 1 || sk < 2 || !z_full || !z_half || !t_surf_rad || !t || !t_half)
        return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "interp_temp: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(RRTMG_B200_ERR_CUDA, "no CUDA device: this library has no CPU path");
    const size_t np = (size_t)si * sj, L = sk, V = sk + 1;
    if (D.misc.ensure((2 * np * L + 2 * np * V + np) * 8)) return fail(RRTMG_B200_ERR_CUDA, "cudaMalloc failed (interp_temp)");
    double *zf = (double *)D.misc.p, *zh = zf + np * L, *tt = zh + np * V, *ts = tt + np * L, *th = ts + np;
    CUDA_OK(cudaMemcpy(zf, z_full, np * L * 8, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(zh, z_half, np * V * 8, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(tt, t, np * L * 8, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(ts, t_surf_rad, np * 8, cudaMemcpyHostToDevice));
    G.launches += drv_interp_temp((int)np, sk, zf, zh, ts, tt, th, nullptr);
    CUDA_OK(cudaMemcpy(t_half, th, np * V * 8, cudaMemcpyDeviceToHost));
    return RRTMG_B200_OK;
}

int rrtmg_b200_run_rrtmg_device(const rrtmg_b200_rad_config *cfg, int si, int sj, int sk, int seconds, int days,
                                const double *lat, const double *lon, const double *p_full, const double *p_half,
                                const double *albedo, const double *q, const double *t, const double *t_surf_rad,
                                const double *z_full, const double *z_half, const double *t_half_in, const double *o3f,
                                double *tdt, double *coszen, double *flux_sw, double *flux_lw,
                                double *tdt_rad, double *tdt_sw, double *tdt_lw, double *olr, double *isr,
                                double *t_half_out, void *stream)
{
    if (!cfg) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "run_rrtmg: config is NULL");
    return run_rrtmg_device_impl(*cfg, si, sj, sk, seconds, days, lat, lon, p_full, p_half, albedo, q, t, t_surf_rad,
                                 z_full, z_half, t_half_in, o3f, tdt, coszen, flux_sw, flux_lw, tdt_rad, tdt_sw, tdt_lw,
                                 olr, isr, t_half_out, (cudaStream_t)stream, D.slot[0], false);
}

int rrtmg_b200_run_rrtmg(const rrtmg_b200_rad_config *cfg, int si, int sj, int sk, int seconds, int days,
                         const double *lat, const double *lon, const double *p_full, const double *p_half,
                         const double *albedo, const double *q, const double *t, const double *t_surf_rad,
                         const double *z_full, const double *z_half, const double *t_half_in, const double *o3f,
                         double *tdt, double *coszen, double *flux_sw, double *flux_lw,
                         double *tdt_rad, double *tdt_sw, double *tdt_lw, double *olr, double *isr,
                         double *t_half_out)
{
    if (!cfg) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "run_rrtmg: config is NULL");
    if (si < 1 || sj < 1 || sk < 2) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "run_rrtmg: grid extents out of range");
    if (!lat || !lon || !p_full || !p_half || !albedo || !q || !t || !t_surf_rad || !coszen)
        return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "run_rrtmg: required array is NULL");
    if (cfg->lonstep < 1 || si % cfg->lonstep != 0) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "run_rrtmg: lonstep must divide the number of longitudes");
    if (!t_half_in && (!z_full || !z_half)) return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "run_rrtmg: need t_half or z_full + z_half");
    // Blocks of latitude rows flow through a two-stage pipeline (two streams, two buffer sets): the H2D copy of block
    // i+1 and the D2H copy of block i-1 overlap the kernels of block i.  Everything in run_rrtmg is local to a
    // latitude row (the zonal means too), and a row block of an (si, sj, n) field is n runs of si*rows doubles.
    const size_t np_all = (size_t)si * sj, L = sk, V = sk + 1;
    int target = G.run_chunk > 0 ? G.run_chunk : 16384;           // RRTMG columns per block (e2e at T170L60: 8192 44.7, 16384 44.3, 32768 45.2 ms)
    {
        const long nrr = (long)(si / cfg->lonstep) * sj;           // RRTMG columns of the call: small batches in about four blocks
        if (G.run_chunk <= 0 && nrr < 4L * target) {
            target = (int)(((nrr + 3) / 4 + 1023) / 1024 * 1024);
            if (target < 4096) target = 4096;
        }
    }
    if (G.chunk > 0 && G.chunk < target) target = G.chunk;
    int rows = (int)((long)target * cfg->lonstep / si);
    if (rows < 1) rows = 1;
    if (rows > sj) rows = sj;
    const int nslot = host_slots((sj + rows - 1) / rows);
    auto pad = [](size_t n) { return (n * 8 + 255) & ~(size_t)255; };
    const size_t npb = (size_t)si * rows;
    const size_t hin = 5 * pad(npb) + 6 * pad(npb * L) + 3 * pad(npb * V);
    const size_t hout = 5 * pad(npb) + 4 * pad(npb * L) + pad(npb * V);
    for (int i = 0; i < nslot; ++i) {
        DrvSlot &S = D.slot[i];
        if (!S.st && cudaStreamCreateWithFlags(&S.st, cudaStreamNonBlocking) != cudaSuccess) return fail(RRTMG_B200_ERR_CUDA, "cudaStreamCreate failed");
        if (S.host_in.ensure(hin) || S.host_out.ensure(hout)) return fail(RRTMG_B200_ERR_CUDA, "cudaMalloc failed (run_rrtmg host staging)");
    }
    // run_rrtmg replaces the top interface pressure of ALL the rank's columns when the smallest one is not positive
    // (rrtm_radiation.f90:655-656): decided here over the whole field, not per row block
    int top_flag = 0;
    for (int j = 0; j < sj && !top_flag; ++j)
        for (int ii = 0; ii < si; ii += cfg->lonstep)
            if (p_half[(size_t)ii + (size_t)si * j] * 0.01 <= 0.0) { top_flag = 1; break; }
    int idx = 0;
    for (int j0 = 0; j0 < sj; j0 += rows, ++idx) {
        const int nr = (sj - j0 < rows) ? sj - j0 : rows;
        DrvSlot &S = D.slot[idx % nslot];
        cudaStream_t st = S.st;
        // a block is "columns" [j0*si, (j0+nr)*si) of a column-major (si*sj, n) array
        Slot a{(char *)S.host_in.p, 0, j0 * si, nr * si, (int)np_all, st, true};
        const double *d_lat = a.up(lat, 1), *d_lon = a.up(lon, 1), *d_alb = a.up(albedo, 1), *d_ts = a.up(t_surf_rad, 1);
        const double *d_pf = a.up(p_full, L), *d_ph = a.up(p_half, V), *d_q = a.up(q, L), *d_t = a.up(t, L);
        const double *d_zf = t_half_in ? nullptr : a.up(z_full, L), *d_zh = t_half_in ? nullptr : a.up(z_half, V);
        const double *d_th = a.up(t_half_in, V), *d_o3 = a.up(o3f, L);
        double *d_tdt = tdt ? const_cast<double *>(a.up(tdt, L)) : nullptr;
        if (!a.ok) return fail(RRTMG_B200_ERR_CUDA, "H2D copy failed (run_rrtmg)");
        Slot o{(char *)S.host_out.p, 0, j0 * si, nr * si, (int)np_all, st, true};
        double *d_cz = o.take(1);
        double *d_fsw = flux_sw ? o.take(1) : nullptr, *d_flw = flux_lw ? o.take(1) : nullptr;
        double *d_olr = olr ? o.take(1) : nullptr, *d_isr = isr ? o.take(1) : nullptr;
        double *d_trad = tdt_rad ? o.take(L) : nullptr, *d_tsw = tdt_sw ? o.take(L) : nullptr, *d_tlw = tdt_lw ? o.take(L) : nullptr;
        double *d_tho = t_half_out ? o.take(V) : nullptr;
        if (const int rc = run_rrtmg_device_impl(*cfg, si, nr, sk, seconds, days, d_lat, d_lon, d_pf, d_ph, d_alb, d_q, d_t, d_ts,
                                                 d_zf, d_zh, d_th, d_o3, d_tdt, d_cz, d_fsw, d_flw, d_trad, d_tsw, d_tlw,
                                                 d_olr, d_isr, d_tho, st, S, true, top_flag)) {
            for (int i = 0; i < nslot; ++i) cudaStreamSynchronize(D.slot[i].st);
            return rc;
        }
        if (tdt) o.down(tdt, d_tdt, L);
        o.down(coszen, d_cz, 1);
        if (flux_sw) o.down(flux_sw, d_fsw, 1);
        if (flux_lw) o.down(flux_lw, d_flw, 1);
        if (olr) o.down(olr, d_olr, 1);
        if (isr) o.down(isr, d_isr, 1);
        if (tdt_rad) o.down(tdt_rad, d_trad, L);
        if (tdt_sw) o.down(tdt_sw, d_tsw, L);
        if (tdt_lw) o.down(tdt_lw, d_tlw, L);
        if (t_half_out) o.down(t_half_out, d_tho, V);
        if (!o.ok) return fail(RRTMG_B200_ERR_CUDA, "D2H copy failed (run_rrtmg)");
    }
    for (int i = 0; i < nslot; ++i) CUDA_OK(cudaStreamSynchronize(D.slot[i].st));
    return RRTMG_B200_OK;
}

int rrtmg_b200_tables_info(int *lw_synthetic, int *lw_ready, int *sw_ready)
{
    // the packaged LW k-distribution is a synthetic stand-in (the reference checkout has no rrtmg_lw_k_g.f90); its blob
    // carries the marker "lwmeta.synthetic", real coefficient sets do not
    if (lw_synthetic) *lw_synthetic = find("lwmeta.synthetic") != nullptr;
    if (lw_ready) *lw_ready = G.lw_ready ? 1 : 0;
    if (sw_ready) *sw_ready = G.sw_ready ? 1 : 0;
    return RRTMG_B200_OK;
}

int rrtmg_b200_set_option(const char *key, long value)
{
    const std::string k(key ? key : "");
    if (k == "chunk") return rrtmg_b200_set_chunk((int)value);
    if (k == "host_chunk") { G.host_chunk = (int)value; return RRTMG_B200_OK; }
    if (k == "host_slots" && value >= 2 && value <= MAXSLOT) { G.host_slots = (int)value; return RRTMG_B200_OK; }
    if (k == "run_chunk") { G.run_chunk = (int)value; return RRTMG_B200_OK; }
    if (k == "capture_stages") { G.capture = value != 0; return RRTMG_B200_OK; }
    if (k == "share_inputs") { G.share_inputs = value != 0; G.shared.valid = false; return RRTMG_B200_OK; }
    if (k == "lw_rtrn_pad_kb") { g_tune.lw_rtrn_pad_kb = (int)value; return RRTMG_B200_OK; }
    if (k == "taumol_sync") { g_tune.taumol_sync = (int)value; return RRTMG_B200_OK; }
    if (k == "lw_fused") { g_tune.lw_fused = value != 0; return RRTMG_B200_OK; }
    if (k == "sw_fused") { g_tune.sw_fused = value != 0; return RRTMG_B200_OK; }
    if (k == "col_warps" && (value == 0 || value == 8 || value == 16)) { g_tune.col_warps = (int)value; return RRTMG_B200_OK; }
#ifdef RRTMG_B200_DEV_VARIANTS
    if (k == "dev_variants") return RRTMG_B200_OK;
    if (k == "lw_rtrn_variant") { g_tune.lw_rtrn_variant = (int)value; return RRTMG_B200_OK; }
    if (k == "sw_solver_variant") { g_tune.sw_solver_variant = (int)value; return RRTMG_B200_OK; }
    if (k == "sw_solver_store") { g_tune.sw_solver_store = (int)value; return RRTMG_B200_OK; }
#else
    // the earlier kernel forms (SW solver variants 0-3, direct-load lw_rtrn) exist in development builds only
    // (RRTMG_B200_DEV_VARIANTS=1 python -m mima_b200.build --force)
    if (k == "lw_rtrn_variant" && value >= 2) { g_tune.lw_rtrn_variant = (int)value; return RRTMG_B200_OK; }
    if (k == "sw_solver_variant" && value == 4) { g_tune.sw_solver_variant = (int)value; return RRTMG_B200_OK; }
#endif
    if (k == "sw_solver_pad_kb") { g_tune.sw_solver_pad_kb = (int)value; return RRTMG_B200_OK; }
    if (k.size() == 2 && k[0] == 'x' && k[1] >= '0' && k[1] <= '7') { g_tune.x[k[1] - '0'] = (int)value; return RRTMG_B200_OK; }
    if (k == "kernel_timing") { KT.collect(); KT.on = value != 0; return RRTMG_B200_OK; }
    return fail(RRTMG_B200_ERR_BAD_ARGUMENT, "unknown option " + k);
}

} // extern "C"
