// test_fused_kernel_emu.cpp -- runs the SOURCE of the fused streaming kernels (csrc/kernels_fused.cu: Pass A k_down,
// Pass B k_up, their launch/geometry code included) on the CPU, lane by lane (tests/cpp/emu/host_emulation.h, model F:
// 32 fibers per warp, warp shuffles as collectives), and compares every output with the CPU oracle's operators
// (oracle/pmg_oracle.c) BIT FOR BIT:
//   * whole levels: xb = S^nu(x), coarse f = R(f - A xb); x = S^nu(xb + P e) and the residual-norm partials;
//     all sweep counts, weighted / plain Jacobi, x == 0 form, both prolongations, every tuning variant, several
//     chunk geometries (the SM count the geometry code sees is varied);
//   * row slabs as the multi-GPU path runs them: 2 and 3 "ranks" in one process, halo rows read in place from the
//     neighbour's arrays through HaloPeers (flags pre-set, the epoch publication checked), the interior / boundary
//     split, the rows Pass A / Pass B finish beyond the slab, and that nothing outside the permitted rows is written.
// Test infrastructure only (no GPU needed); the GPU suite checks the compiled kernels.
#include <cmath>
#include <cstring>
#include <random>
#include <vector>

#include "../../parallel-geometric-multigrid-for-poisson-problem_b200/csrc/kernels_fused.cu"
#include "../../oracle/oracle.h"

namespace pmg {
void count_launch(int) {}
unsigned long long launches_so_far() { return 0; }
JacobiCoef jacobi_coef(double h, double omega)
{
    JacobiCoef c;
    c.h2 = h * h;
    c.omega = omega;
    c.om1 = 1.0 - omega;
    c.w4 = 0.25 * omega;
    c.weighted = (omega != 1.0);
    return c;
}
}  // namespace pmg

using namespace pmg;

static int g_bad = 0;
static std::mt19937_64 g_rng(12345);
static double rnd() { return std::uniform_real_distribution<double>(-1.0, 1.0)(g_rng); }

static void check(bool ok, const char *what, int n, int a = 0, int b = 0)
{
    if (!ok) {
        std::printf("  MISMATCH: %s (n=%d, %d, %d)\n", what, n, a, b);
        ++g_bad;
    }
}

// a dense n x n field
struct Dense {
    int n;
    std::vector<double> v;
    explicit Dense(int n_) : n(n_), v((size_t)n_ * n_, 0.0) {}
    double &at(int y, int x) { return v[(size_t)y * n + x]; }
    void randomize(bool ring)
    {
        for (int y = 0; y < n; ++y)
            for (int x = 0; x < n; ++x) at(y, x) = (ring || (y > 0 && y < n - 1 && x > 0 && x < n - 1)) ? rnd() : 0.0;
    }
};

// the solver's padded layout of `rows` rows of an n-column level: logical (0,0) at base[PADY * pitch + PADX]
struct Padded {
    int n, rows, pitch;
    std::vector<double> base;
    Padded(int n_, int rows_) : n(n_), rows(rows_), pitch(level_pitch(n_)), base((size_t)level_pitch(n_) * (rows_ + 2 * PADY), 0.0) {}
    double *p() { return base.data() + (size_t)PADY * pitch + PADX; }
    double &at(int y, int x) { return p()[(ptrdiff_t)y * pitch + x]; }
    // rows [a, b) of the dense field d (global row offset y0) into local rows [a - y0, b - y0)
    void load(Dense &d, int y0, int a, int b)
    {
        for (int y = a; y < b; ++y)
            for (int x = 0; x < n; ++x) at(y - y0, x) = d.at(y, x);
    }
    void fill_rows(int a, int b, double val)  // logical columns of local rows [a, b)
    {
        for (int y = a; y < b; ++y)
            for (int x = 0; x < n; ++x) at(y, x) = val;
    }
    // everything outside local rows [a, b) x columns [0, n) still equals `other` (same shape)
    bool untouched_outside(const Padded &other, int a, int b) const
    {
        for (int r = -PADY; r < rows + PADY; ++r)
            for (int c = -PADX; c < pitch - PADX; ++c) {
                if (r >= a && r < b && c >= 0 && c < n) continue;
                size_t i = (size_t)(r + PADY) * pitch + PADX + c;
                if (std::memcmp(&base[i], &other.base[i], 8) != 0) return false;
            }
        return true;
    }
};

static bool same_bits(double a, double b) { return std::memcmp(&a, &b, 8) == 0; }

static void partition(int n, int ranks, int r, int &y0, int &y1)  // pmg_partition_rows
{
    long pairs = (n - 1) / 2;
    long a = pairs * r / ranks, b = pairs * (r + 1) / ranks;
    y0 = (int)(2 * a);
    y1 = (r == ranks - 1) ? n : (int)(2 * b);
}

struct Expected {
    Dense xb, cf, xnew;
    double norm2;
    Expected(int n) : xb(n), cf((n - 1) / 2 + 1), xnew(n), norm2(0.0) {}
};

// the oracle's operator sequence for one level visit
static Expected expected_visit(Dense &x, Dense &f, Dense &e, double h, double omega, int nu1, int nu2, int prolong, bool x_is_zero)
{
    const int n = x.n, nc = (n - 1) / 2 + 1;
    Expected E(n);
    E.xb = x;
    if (x_is_zero) std::fill(E.xb.v.begin(), E.xb.v.end(), 0.0);
    orc_jacobi(E.xb.v.data(), f.v.data(), n, n, h, omega, nu1 - 1, 0.0, nullptr);
    Dense r(n);
    orc_residual(r.v.data(), E.xb.v.data(), f.v.data(), n, n, h);
    orc_restrict_fw(r.v.data(), E.cf.v.data(), n, nc);
    E.xnew = E.xb;
    orc_prolong_add(E.xnew.v.data(), e.v.data(), n, nc, prolong);
    orc_jacobi(E.xnew.v.data(), f.v.data(), n, n, h, omega, nu2 - 1, 0.0, nullptr);
    Dense r2(n);
    orc_residual(r2.v.data(), E.xnew.v.data(), f.v.data(), n, n, h);
    double nn = orc_norm(r2.v.data(), (long)n * n);
    E.norm2 = nn * nn;
    return E;
}

// ---- cross-cycle pass: Pass B of cycle k + Pass A of cycle k+1 in one launch (k_cross) -------------------------------
static void cross_level(int n, double omega, int prolong, int sms, int minb)
{
    emu_num_sms = sms;
    fused_set_cross_minb(minb);
    const int nc = (n - 1) / 2 + 1;
    const double h = 1.0 / (n - 1);
    Dense xb(n), f(n), e(nc);
    xb.randomize(true);
    f.randomize(true);
    e.randomize(false);
    // what the two separate passes produce
    Dense xk = xb;
    orc_prolong_add(xk.v.data(), e.v.data(), n, nc, prolong);
    orc_jacobi(xk.v.data(), f.v.data(), n, n, h, omega, 1, 0.0, nullptr);
    Dense r(n);
    orc_residual(r.v.data(), xk.v.data(), f.v.data(), n, n, h);
    const double nn = orc_norm(r.v.data(), (long)n * n);
    Dense xb2 = xk;
    orc_jacobi(xb2.v.data(), f.v.data(), n, n, h, omega, 1, 0.0, nullptr);
    Dense r2(n), cf(nc);
    orc_residual(r2.v.data(), xb2.v.data(), f.v.data(), n, n, h);
    orc_restrict_fw(r2.v.data(), cf.v.data(), n, nc);
    // the kernel
    Padded pin(n, n), pout(n, n), pxk(n, n), pf(n, n), pe(nc, nc), pcf(nc, nc);
    pin.load(xb, 0, 0, n);
    pf.load(f, 0, 0, n);
    pe.load(e, 0, 0, nc);
    pout.fill_rows(0, n, std::nan(""));
    pxk.fill_rows(0, n, std::nan(""));
    Padded pin0 = pin, pf0 = pf, pe0 = pe, pcf0 = pcf;
    FusedLevel lv{};
    lv.x = (minb == 2) ? nullptr : pxk.p();  // nullptr: x_k is not written (one-GPU solve)
    lv.xb = pin.p();
    lv.f = pf.p();
    lv.n = n;
    lv.pitch = pin.pitch;
    lv.h = h;
    std::vector<double> partials((size_t)fused_max_partials(n), 0.0);
    int np = 0;
    launch_fused_cross(lv, pout.p(), pe.p(), pcf.p(), pe.pitch, omega, prolong, partials.data(), &np, nullptr, nullptr);
    bool ok_x = true, ok_c = true, ok_k = true;
    for (int y = 0; y < n; ++y)
        for (int x = 0; x < n; ++x) {
            ok_x = ok_x && same_bits(pout.at(y, x), xb2.at(y, x));
            ok_k = ok_k && (minb == 2 ? std::isnan(pxk.at(y, x)) : same_bits(pxk.at(y, x), xk.at(y, x)));
        }
    check(ok_k, "cross: x_k = S^2(xb + P e)", n, sms, minb);
    for (int y = 0; y < nc; ++y)
        for (int x = 0; x < nc; ++x) ok_c = ok_c && same_bits(pcf.at(y, x), cf.at(y, x));
    double sum = 0.0;
    for (int i = 0; i < np; ++i) sum += partials[(size_t)i];
    check(ok_x, "cross: xb' = S^2(S^2(xb + P e))", n, sms, minb);
    check(ok_c, "cross: coarse f = R(f - A xb')", n, sms, minb);
    check(std::fabs(sum - nn * nn) <= 1e-12 * nn * nn, "cross: residual norm of x_k", n, sms, minb);
    check(pin.untouched_outside(pin0, 0, 0) && pf.untouched_outside(pf0, 0, 0) && pe.untouched_outside(pe0, 0, 0),
          "cross: inputs untouched", n, sms, minb);
    check(pcf.untouched_outside(pcf0, 1, nc - 1), "cross: coarse ring and padding untouched", n, sms, minb);
    std::printf("cross-cycle pass n=%d omega=%.3f prolong=%d sms=%d minb=%d: %s\n", n, omega, prolong, sms, minb,
                (ok_x && ok_c) ? "bit-identical" : "MISMATCH");
}

// ---- Pass A in its prolong-in form: xb = S^2(P e), coarse f = R(f - A xb), P e never written (nested iteration) --------
static void prolong_down_level(int n, double omega, int prolong, int sms)
{
    emu_num_sms = sms;
    const int nc = (n - 1) / 2 + 1;
    const double h = 1.0 / (n - 1);
    Dense f(n), e(nc);
    f.randomize(true);
    e.randomize(false);
    Dense x0(n);  // zeroed fine grid (MultiGrid.hpp:161)
    orc_prolong_add(x0.v.data(), e.v.data(), n, nc, prolong);
    Dense xb = x0;
    orc_jacobi(xb.v.data(), f.v.data(), n, n, h, omega, 1, 0.0, nullptr);
    Dense r(n), cf(nc);
    orc_residual(r.v.data(), xb.v.data(), f.v.data(), n, n, h);
    orc_restrict_fw(r.v.data(), cf.v.data(), n, nc);
    Padded px(n, n), pxb(n, n), pf(n, n), pe(nc, nc), pcf(nc, nc);
    px.fill_rows(0, n, std::nan(""));  // the iterate array must not be read
    pf.load(f, 0, 0, n);
    pe.load(e, 0, 0, nc);
    Padded pf0 = pf, pe0 = pe, pcf0 = pcf, pxb0 = pxb;
    FusedLevel lv{};
    lv.x = px.p();
    lv.xb = pxb.p();
    lv.f = pf.p();
    lv.n = n;
    lv.pitch = px.pitch;
    lv.h = h;
    launch_fused_down_prolong(lv, pe.p(), pe.pitch, pcf.p(), pcf.pitch, omega, prolong, nullptr, nullptr);
    bool ok_x = true, ok_c = true;
    for (int y = 0; y < n; ++y)
        for (int x = 0; x < n; ++x) ok_x = ok_x && same_bits(pxb.at(y, x), xb.at(y, x));
    for (int y = 0; y < nc; ++y)
        for (int x = 0; x < nc; ++x) ok_c = ok_c && same_bits(pcf.at(y, x), cf.at(y, x));
    check(ok_x, "prolong-in Pass A: xb = S^2(P e)", n, sms, prolong);
    check(ok_c, "prolong-in Pass A: coarse f = R(f - A xb)", n, sms, prolong);
    check(pf.untouched_outside(pf0, 0, 0) && pe.untouched_outside(pe0, 0, 0), "prolong-in Pass A: inputs untouched", n, sms, prolong);
    check(pxb.untouched_outside(pxb0, 0, n), "prolong-in Pass A wrote xb outside the level", n, sms, prolong);
    check(pcf.untouched_outside(pcf0, 1, nc - 1), "prolong-in Pass A: coarse ring and padding untouched", n, sms, prolong);
    std::printf("prolong-in Pass A n=%d omega=%.3f prolong=%d sms=%d: %s\n", n, omega, prolong, sms,
                (ok_x && ok_c) ? "bit-identical" : "MISMATCH");
}

// ---- whole level (one GPU) ---------------------------------------------------------------------------------------
static void whole_level(int n, double omega, int nu1, int nu2, int prolong, bool x_is_zero, int variant, int sms)
{
    emu_num_sms = sms;
    fused_set_variant(variant);
    const int nc = (n - 1) / 2 + 1;
    const double h = 1.0 / (n - 1);
    Dense x(n), f(n), e(nc);
    x.randomize(true);
    f.randomize(true);
    e.randomize(false);
    Expected E = expected_visit(x, f, e, h, omega, nu1, nu2, prolong, x_is_zero);
    Padded px(n, n), pxb(n, n), pf(n, n), pcf(nc, nc), pe(nc, nc);
    px.load(x, 0, 0, n);
    if (x_is_zero) px.fill_rows(0, n, std::nan(""));  // must not be read
    pf.load(f, 0, 0, n);
    pe.load(e, 0, 0, nc);
    Padded pxb0 = pxb, pcf0 = pcf;
    FusedLevel lv{};
    lv.x = px.p();
    lv.xb = pxb.p();
    lv.f = pf.p();
    lv.n = n;
    lv.pitch = px.pitch;
    lv.h = h;
    launch_fused_down(lv, pcf.p(), pcf.pitch, nu1, omega, x_is_zero, nullptr);
    bool ok = true;
    for (int y = 0; y < n && ok; ++y)
        for (int xx = 0; xx < n; ++xx)
            if (!same_bits(pxb.at(y, xx), x_is_zero && (y == 0 || y == n - 1 || xx == 0 || xx == n - 1) ? 0.0 : E.xb.at(y, xx))) {
                ok = false;
                break;
            }
    check(ok, "Pass A: xb != S^nu1(x)", n, nu1, variant);
    ok = true;
    for (int y = 0; y < nc && ok; ++y)
        for (int xx = 0; xx < nc; ++xx)
            if (!same_bits(pcf.at(y, xx), E.cf.at(y, xx))) ok = false;
    check(ok, "Pass A: coarse f != R(f - A xb)", n, nu1, variant);
    check(pxb.untouched_outside(pxb0, 0, n), "Pass A wrote xb outside the level", n);
    check(pcf.untouched_outside(pcf0, 0, nc), "Pass A wrote the coarse array outside the level", n);
    // Pass B from the expected xb (ring of xb as Pass A left it)
    Padded pxn(n, n);
    Padded pxn0 = pxn;
    lv.x = pxn.p();
    std::vector<double> partials((size_t)fused_max_partials(n), 0.0);
    int np = 0;
    launch_fused_up(lv, pe.p(), pe.pitch, nu2, omega, prolong, partials.data(), &np, nullptr);
    ok = true;
    for (int y = 0; y < n && ok; ++y)
        for (int xx = 0; xx < n; ++xx) {
            double want = E.xnew.at(y, xx);
            if (x_is_zero && (y == 0 || y == n - 1 || xx == 0 || xx == n - 1)) want = 0.0;
            if (!same_bits(pxn.at(y, xx), want)) ok = false;
        }
    check(ok, "Pass B: x != S^nu2(xb + P e)", n, nu2, variant);
    check(pxn.untouched_outside(pxn0, 0, n), "Pass B wrote x outside the level", n);
    double s2 = 0.0;
    for (int i = 0; i < np; ++i) s2 += partials[i];
    if (!x_is_zero)  // (with the zeroed ring the oracle's residual next to the ring differs; the iterate check covers it)
        check(std::fabs(s2 - E.norm2) <= 1e-12 * E.norm2, "Pass B: residual norm", n, nu2, variant);
    // the same Pass B without the norm must give the same iterate
    Padded pxm(n, n);
    lv.x = pxm.p();
    launch_fused_up(lv, pe.p(), pe.pitch, nu2, omega, prolong, nullptr, nullptr, nullptr);
    check(pxm.base == pxn.base, "Pass B with / without norm differ", n, nu2, variant);
    std::printf("whole level n=%d omega=%.3f nu=(%d,%d) prolong=%d x0=%d variant=%d sms=%d: %s\n", n, omega, nu1, nu2, prolong,
                (int)x_is_zero, variant, sms, g_bad ? "see above" : "ok");
    std::fflush(stdout);
}

// plain smoothing pass (coarse_f == nullptr): the FMG sweeps and pmg_smooth
static void smoothing_pass(int n, double omega, int nu, int sms)
{
    emu_num_sms = sms;
    fused_set_variant(-1);
    const double h = 1.0 / (n - 1);
    Dense x(n), f(n);
    x.randomize(true);
    f.randomize(true);
    Dense want = x;
    orc_jacobi(want.v.data(), f.v.data(), n, n, h, omega, nu - 1, 0.0, nullptr);
    Padded px(n, n), pxb(n, n), pf(n, n);
    px.load(x, 0, 0, n);
    pf.load(f, 0, 0, n);
    FusedLevel lv{};
    lv.x = px.p();
    lv.xb = pxb.p();
    lv.f = pf.p();
    lv.n = n;
    lv.pitch = px.pitch;
    lv.h = h;
    launch_fused_down(lv, nullptr, 0, nu, omega, false, nullptr);
    bool ok = true;
    for (int y = 0; y < n; ++y)
        for (int xx = 0; xx < n; ++xx) ok = ok && same_bits(pxb.at(y, xx), want.at(y, xx));
    check(ok, "smoothing pass", n, nu);
    std::printf("smoothing pass n=%d nu=%d sms=%d: %s\n", n, nu, sms, ok ? "ok" : "MISMATCH");
}

// ---- row slabs: `ranks` ranks in one process, halo rows read in place through HaloPeers ------------------------------
struct Rank {
    int y0, y1, ny, yc0, yc1, nyc;
    Padded x, xb, f, cf, e;
    int inbox[2];  // {from_up, from_dn}
    Rank(int n, int nc, int y0_, int y1_, int yc0_, int yc1_)
        : y0(y0_), y1(y1_), ny(y1_ - y0_), yc0(yc0_), yc1(yc1_), nyc(yc1_ - yc0_), x(n, y1_ - y0_), xb(n, y1_ - y0_),
          f(n, y1_ - y0_), cf(nc, yc1_ - yc0_), e(nc, yc1_ - yc0_), inbox{0, 0} {}
};

static void slab_visit(int n, int ranks, bool first_visit, bool split, int prolong, int sms, bool level0, bool prologue = false)
{
    emu_num_sms = sms;
    fused_set_variant(-1);
    fused_set_halo_prologue(prologue ? 1 : 0);  // halo rows copied by a prologue of Pass A instead of streamed in place
    const double omega = 2.0 / 3.0;
    const int nc = (n - 1) / 2 + 1, nu1 = 2, nu2 = 2, epoch = 7;
    const double h = 1.0 / (n - 1);
    Dense x(n), f(n), e(nc);
    x.randomize(level0);  // non-zero Dirichlet ring only on the finest level
    f.randomize(false);
    e.randomize(false);
    Expected E = expected_visit(x, f, e, h, omega, nu1, nu2, prolong, first_visit);
    std::vector<Rank> R;
    for (int r = 0; r < ranks; ++r) {
        int y0, y1, c0, c1;
        partition(n, ranks, r, y0, y1);
        c0 = y0 / 2;  // the solver only accepts partitions that nest: coarse row jc lives with fine row 2 jc
        c1 = (y1 == n) ? nc : y1 / 2;
        R.emplace_back(n, nc, y0, y1, c0, c1);
    }
    const double poison = std::nan("");
    for (int r = 0; r < ranks; ++r) {
        Rank &K = R[r];
        const bool up = r > 0, dn = r < ranks - 1;
        // owned rows; halo rows poisoned where the kernel must read the neighbour instead
        K.x.load(x, K.y0, K.y0, K.y1);
        K.f.load(f, K.y0, K.y0, K.y1);
        if (first_visit) K.x.fill_rows(0, K.ny, poison);  // x == 0 is known: never read
        if (up) K.x.fill_rows(-PADY, 0, poison);
        if (dn) K.x.fill_rows(K.ny, K.ny + PADY, poison);
        if (level0) {  // the finest level's f halo is local (exchanged once in pmg_set_rhs)
            K.f.load(f, K.y0, std::max(0, K.y0 - PADY), K.y0);
            K.f.load(f, K.y0, K.y1, std::min(n, K.y1 + PADY));
        } else {
            if (up) K.f.fill_rows(-PADY, 0, poison);
            if (dn) K.f.fill_rows(K.ny, K.ny + PADY, poison);
        }
        // coarse correction: rows [yc0 - 4, yc1 + 4) (what the child level / the agglomerated level provides)
        K.e.fill_rows(-PADY, K.nyc + PADY, poison);
        for (int xx = -PADX; xx < K.e.pitch - PADX; ++xx)  // padding columns stay zero
            for (int y = -PADY; y < K.nyc + PADY; ++y)
                if (xx < 0 || xx >= nc) K.e.at(y, xx) = 0.0;
        K.e.load(e, K.yc0, std::max(0, K.yc0 - 4), std::min(nc, K.yc1 + 4));
        if (K.yc0 - 4 < 0) K.e.fill_rows(-PADY, 0, 0.0);               // rows beyond the grid: zero padding
        if (K.yc1 + 4 > nc) K.e.fill_rows(K.nyc, K.nyc + PADY, 0.0);
        K.inbox[0] = K.inbox[1] = epoch;  // the neighbours have published
    }
    std::vector<int> outbox(2 * ranks, 0);  // what each rank publishes: [r][to_up, to_dn]
    for (int r = 0; r < ranks; ++r) {
        Rank &K = R[r];
        const bool up = r > 0, dn = r < ranks - 1;
        FusedLevel lv{};
        lv.x = K.x.p();
        lv.xb = K.xb.p();
        lv.f = K.f.p();
        lv.n = n;
        lv.pitch = K.x.pitch;
        lv.h = h;
        lv.ny = K.ny;
        lv.yoff = K.y0;
        HaloPeers hp{};
        if (!first_visit) {
            hp.x_up = up ? R[r - 1].x.p() + (ptrdiff_t)R[r - 1].ny * K.x.pitch : nullptr;  // row -1 == neighbour's last row
            hp.x_dn = dn ? R[r + 1].x.p() : nullptr;
            hp.x_keep = K.x.p();
        }
        if (!level0) {
            hp.f_up = up ? R[r - 1].f.p() + (ptrdiff_t)R[r - 1].ny * K.f.pitch : nullptr;
            hp.f_dn = dn ? R[r + 1].f.p() : nullptr;
            hp.f_keep = K.f.p();
        }
        hp.flag_up = up ? &K.inbox[0] : nullptr;
        hp.flag_dn = dn ? &K.inbox[1] : nullptr;
        hp.pub_up = up ? &outbox[2 * r] : nullptr;
        hp.pub_dn = dn ? &outbox[2 * r + 1] : nullptr;
        hp.epoch = epoch;
        static const int base_part = 5;  // graph replay form: part of the epoch comes from device memory
        if (prologue || split) {
            hp.epoch = epoch - base_part;
            hp.epoch_base = &base_part;
        }
        int err = 0;
        hp.err = &err;
        Padded xb0 = K.xb, cf0 = K.cf;
        if (split) {  // cycle_dist: boundary strips [-6, 8), [ny - 8, ny + 6) with the peers, interior without
            lv.hp = hp;
            if (up) {
                lv.span_lo = -6;
                lv.span_hi = PADY;
                launch_fused_down(lv, K.cf.p(), K.cf.pitch, nu1, omega, first_visit, nullptr);
                lv.hp.pub_up = lv.hp.pub_dn = nullptr;
            }
            if (dn) {
                lv.span_lo = K.ny - PADY;
                lv.span_hi = K.ny + 6;
                launch_fused_down(lv, K.cf.p(), K.cf.pitch, nu1, omega, first_visit, nullptr);
            }
            lv.hp = HaloPeers{};
            lv.span_lo = up ? PADY : 0;
            lv.span_hi = dn ? K.ny - PADY : K.ny;
            launch_fused_down(lv, K.cf.p(), K.cf.pitch, nu1, omega, first_visit, nullptr);
        } else {
            lv.hp = hp;
            lv.span_lo = up ? -6 : 0;
            lv.span_hi = dn ? K.ny + 6 : K.ny;
            launch_fused_down(lv, K.cf.p(), K.cf.pitch, nu1, omega, first_visit, nullptr);
        }
        check(err == 0, "flag wait failed", n, r);
        check((!up || outbox[2 * r] == epoch) && (!dn || outbox[2 * r + 1] == epoch), "epoch not published", n, r);
        // xb on rows [-6, ny + 6) (clipped to the grid), coarse f on the owned coarse rows
        const int a = up ? -6 : 0, b = dn ? K.ny + 6 : K.ny;
        bool ok = true;
        for (int y = a; y < b; ++y)
            for (int xx = 0; xx < n; ++xx) ok = ok && same_bits(K.xb.at(y, xx), E.xb.at(y + K.y0, xx));
        check(ok, "slab Pass A: xb", n, r, ranks);
        check(K.xb.untouched_outside(xb0, a, b), "slab Pass A wrote xb outside [-6, ny + 6)", n, r);
        ok = true;
        for (int y = 0; y < K.nyc; ++y)
            for (int xx = 0; xx < nc; ++xx) ok = ok && same_bits(K.cf.at(y, xx), E.cf.at(y + K.yc0, xx));
        check(ok, "slab Pass A: coarse f", n, r, ranks);
        check(K.cf.untouched_outside(cf0, 0, K.nyc), "slab Pass A wrote coarse f outside the owned rows", n, r);
        if (prologue && !first_visit) {  // the prologue stored the neighbours' x rows in the local halo rows
            ok = true;
            for (int y = (up ? -PADY : 0); y < (dn ? K.ny + PADY : K.ny); ++y)
                if (y < 0 || y >= K.ny)
                    for (int xx = 0; xx < n; ++xx) ok = ok && same_bits(K.x.at(y, xx), x.at(y + K.y0, xx));
            check(ok, "slab Pass A (prologue): x halo rows", n, r, ranks);
        }
        if (!level0) {  // the fetched f halo rows were kept locally for Pass B
            ok = true;
            for (int y = (up ? -PADY + 2 : 0); y < (dn ? K.ny + PADY - 2 : K.ny); ++y)
                for (int xx = 1; xx < n - 1; ++xx) ok = ok && same_bits(K.f.at(y, xx), f.at(y + K.y0, xx));
            check(ok, "slab Pass A: f halo rows kept for Pass B", n, r, ranks);
        }
    }
    // Pass B: rows [-ext, ny + ext), ext = 0 on the finest level, 4 below it
    double s2 = 0.0;
    for (int r = 0; r < ranks; ++r) {
        Rank &K = R[r];
        const bool up = r > 0, dn = r < ranks - 1;
        FusedLevel lv{};
        Padded xn(n, K.ny);
        Padded xn0 = xn;
        lv.x = xn.p();
        lv.xb = K.xb.p();
        lv.f = K.f.p();
        lv.n = n;
        lv.pitch = K.x.pitch;
        lv.h = h;
        lv.ny = K.ny;
        lv.yoff = K.y0;
        const int ext = level0 ? 0 : 4;
        lv.ext_lo = up ? ext : 0;
        lv.ext_hi = dn ? ext : 0;
        std::vector<double> partials((size_t)fused_max_partials(n), 0.0);
        int np = 0;
        launch_fused_up(lv, K.e.p(), K.e.pitch, nu2, omega, prolong, level0 ? partials.data() : nullptr, &np, nullptr);
        const int a = -lv.ext_lo, b = K.ny + lv.ext_hi;
        bool ok = true;
        for (int y = a; y < b; ++y)
            for (int xx = 0; xx < n; ++xx) ok = ok && same_bits(xn.at(y, xx), E.xnew.at(y + K.y0, xx));
        check(ok, "slab Pass B: x", n, r, ranks);
        check(xn.untouched_outside(xn0, a, b), "slab Pass B wrote x outside its rows", n, r);
        for (int i = 0; i < np; ++i) s2 += partials[i];
    }
    if (level0 && !first_visit) check(std::fabs(s2 - E.norm2) <= 1e-12 * E.norm2, "slab Pass B: rank-summed residual norm", n, ranks);
    std::printf("slabs n=%d ranks=%d first_visit=%d split=%d prolong=%d level0=%d sms=%d prologue=%d: %s\n", n, ranks,
                (int)first_visit, (int)split, prolong, (int)level0, sms, (int)prologue, g_bad ? "see above" : "ok");
    fused_set_halo_prologue(0);
    std::fflush(stdout);
}

// ---- cross-cycle pass on row slabs: the input's halo rows come from the neighbours' arrays through the halo prologue ----
static void slab_cross(int n, int ranks, int prolong, int sms, int minb = 3)
{
    emu_num_sms = sms;
    fused_set_cross_minb(minb);
    const double omega = 2.0 / 3.0;
    const int nc = (n - 1) / 2 + 1, epoch = 11;
    const double h = 1.0 / (n - 1);
    Dense xb(n), f(n), e(nc);
    xb.randomize(true);
    f.randomize(false);
    e.randomize(false);
    Dense xk = xb;
    orc_prolong_add(xk.v.data(), e.v.data(), n, nc, prolong);
    orc_jacobi(xk.v.data(), f.v.data(), n, n, h, omega, 1, 0.0, nullptr);
    Dense r(n);
    orc_residual(r.v.data(), xk.v.data(), f.v.data(), n, n, h);
    const double nn = orc_norm(r.v.data(), (long)n * n);
    Dense xb2 = xk;
    orc_jacobi(xb2.v.data(), f.v.data(), n, n, h, omega, 1, 0.0, nullptr);
    Dense r2(n), cf(nc);
    orc_residual(r2.v.data(), xb2.v.data(), f.v.data(), n, n, h);
    orc_restrict_fw(r2.v.data(), cf.v.data(), n, nc);
    std::vector<Rank> R;
    for (int q = 0; q < ranks; ++q) {
        int y0, y1;
        partition(n, ranks, q, y0, y1);
        R.emplace_back(n, nc, y0, y1, y0 / 2, (y1 == n) ? nc : y1 / 2);
    }
    const double poison = std::nan("");
    for (int q = 0; q < ranks; ++q) {
        Rank &K = R[q];
        const bool up = q > 0, dn = q < ranks - 1;
        K.xb.load(xb, K.y0, K.y0, K.y1);  // the input array: owned rows; halo rows must come from the neighbours
        if (up) K.xb.fill_rows(-PADY, 0, poison);
        if (dn) K.xb.fill_rows(K.ny, K.ny + PADY, poison);
        K.f.load(f, K.y0, std::max(0, K.y0 - PADY), std::min(n, K.y1 + PADY));  // level 0: the f halo is local
        K.e.fill_rows(-PADY, K.nyc + PADY, poison);
        for (int xx = -PADX; xx < K.e.pitch - PADX; ++xx)
            for (int y = -PADY; y < K.nyc + PADY; ++y)
                if (xx < 0 || xx >= nc) K.e.at(y, xx) = 0.0;
        K.e.load(e, K.yc0, std::max(0, K.yc0 - 4), std::min(nc, K.yc1 + 4));
        if (K.yc0 - 4 < 0) K.e.fill_rows(-PADY, 0, 0.0);
        if (K.yc1 + 4 > nc) K.e.fill_rows(K.nyc, K.nyc + PADY, 0.0);
        K.inbox[0] = K.inbox[1] = epoch;
    }
    std::vector<int> outbox(2 * ranks, 0);
    double s2 = 0.0;
    for (int q = 0; q < ranks; ++q) {
        Rank &K = R[q];
        const bool up = q > 0, dn = q < ranks - 1;
        Padded out(n, K.ny), out0 = out;
        FusedLevel lv{};
        lv.x = K.x.p();  // x_k goes here
        lv.xb = K.xb.p();
        lv.f = K.f.p();
        lv.n = n;
        lv.pitch = K.x.pitch;
        lv.h = h;
        lv.ny = K.ny;
        lv.yoff = K.y0;
        HaloPeers hp{};
        hp.x_up = up ? R[q - 1].xb.p() + (ptrdiff_t)R[q - 1].ny * K.x.pitch : nullptr;
        hp.x_dn = dn ? R[q + 1].xb.p() : nullptr;
        hp.x_keep = K.xb.p();
        hp.flag_up = up ? &K.inbox[0] : nullptr;
        hp.flag_dn = dn ? &K.inbox[1] : nullptr;
        hp.pub_up = up ? &outbox[2 * q] : nullptr;
        hp.pub_dn = dn ? &outbox[2 * q + 1] : nullptr;
        hp.epoch = epoch;
        int err = 0;
        hp.err = &err;
        lv.hp = hp;
        Padded xk0 = K.x, cf0 = K.cf;
        std::vector<double> partials((size_t)fused_max_partials(n), 0.0);
        int np = 0;
        launch_fused_cross(lv, out.p(), K.e.p(), K.cf.p(), K.cf.pitch, omega, prolong, partials.data(), &np, nullptr, nullptr);
        check(err == 0, "cross slab: flag wait failed", n, q);
        check((!up || outbox[2 * q] == epoch) && (!dn || outbox[2 * q + 1] == epoch), "cross slab: epoch not published", n, q);
        bool ok_b = true, ok_k = true, ok_c = true;
        for (int y = 0; y < K.ny; ++y)
            for (int xx = 0; xx < n; ++xx) {
                ok_b = ok_b && same_bits(out.at(y, xx), xb2.at(y + K.y0, xx));
                ok_k = ok_k && same_bits(K.x.at(y, xx), xk.at(y + K.y0, xx));
            }
        for (int y = 0; y < K.nyc; ++y)
            for (int xx = 0; xx < nc; ++xx) ok_c = ok_c && same_bits(K.cf.at(y, xx), cf.at(y + K.yc0, xx));
        check(ok_b, "cross slab: xb'", n, q, ranks);
        check(ok_k, "cross slab: x_k", n, q, ranks);
        check(ok_c, "cross slab: coarse f", n, q, ranks);
        check(out.untouched_outside(out0, 0, K.ny) && K.x.untouched_outside(xk0, 0, K.ny), "cross slab wrote outside the owned rows", n, q);
        check(K.cf.untouched_outside(cf0, 0, K.nyc), "cross slab wrote coarse f outside the owned rows", n, q);
        for (int i = 0; i < np; ++i) s2 += partials[(size_t)i];
    }
    check(std::fabs(s2 - nn * nn) <= 1e-12 * nn * nn, "cross slab: residual norm summed over the ranks", n, ranks);
    std::printf("cross-cycle pass on %d slabs n=%d prolong=%d sms=%d: done\n", ranks, n, prolong, sms);
}

int main(int argc, char **argv)
{
    const bool full = argc > 1 && std::strcmp(argv[1], "full") == 0;
    const double w = 2.0 / 3.0;
    // whole levels: the headline V(2,2) flavours on every variant, then the other sweep counts
    for (int v = 0; v < fused_num_variants(); ++v) whole_level(65, w, 2, 2, ORC_PROLONG_REFERENCE, false, v, 148);
    for (int v = 5; v < fused_num_variants(); ++v) whole_level(257, w, 2, 2, ORC_PROLONG_FULL, v == 6, v, 3);  // 128-column strips
    whole_level(65, w, 2, 2, ORC_PROLONG_REFERENCE, true, -1, 148);
    whole_level(33, 1.0, 2, 2, ORC_PROLONG_FULL, false, -1, 148);
    whole_level(129, w, 2, 2, ORC_PROLONG_REFERENCE, false, -1, 2);  // few SMs: 14-row chunks, chunk overlap exercised
    whole_level(17, w, 1, 3, ORC_PROLONG_FULL, false, -1, 148);
    whole_level(33, 0.8, 3, 1, ORC_PROLONG_REFERENCE, true, -1, 1);
    whole_level(33, w, 4, 4, ORC_PROLONG_REFERENCE, false, -1, 148);
    for (int nu = 1; nu <= 4; ++nu) smoothing_pass(65, w, nu, nu == 3 ? 1 : 148);
    // cross-cycle pass (Pass B of cycle k + Pass A of cycle k+1): strips of 52 owned columns, several chunk geometries
    prolong_down_level(65, w, ORC_PROLONG_REFERENCE, 148);
    prolong_down_level(129, w, ORC_PROLONG_FULL, 2);
    prolong_down_level(257, 1.0, ORC_PROLONG_REFERENCE, 5);
    cross_level(65, w, ORC_PROLONG_REFERENCE, 148, 3);
    cross_level(129, w, ORC_PROLONG_FULL, 148, 4);
    cross_level(129, 1.0, ORC_PROLONG_REFERENCE, 2, 2);   // few SMs: tall chunks, chunk overlap exercised
    cross_level(257, w, ORC_PROLONG_REFERENCE, 1, 3);
    // 4 columns per lane (strips of 128 columns own 112), prefetch depths 2 / 3 / 4
    cross_level(257, w, ORC_PROLONG_REFERENCE, 148, 2);
    cross_level(129, w, ORC_PROLONG_FULL, 2, 7);
    cross_level(513, w, ORC_PROLONG_REFERENCE, 4, 8);
    if (full) {
        cross_level(1025, w, ORC_PROLONG_REFERENCE, 148, 2);   // ten strips of 112 columns, the B200's wave
        prolong_down_level(513, w, ORC_PROLONG_FULL, 148);
    }
    slab_cross(129, 2, ORC_PROLONG_REFERENCE, 148);
    slab_cross(257, 3, ORC_PROLONG_FULL, 2);
    slab_cross(257, 4, ORC_PROLONG_REFERENCE, 148);
    slab_cross(257, 3, ORC_PROLONG_REFERENCE, 148, 2);
    slab_cross(513, 2, ORC_PROLONG_FULL, 3, 7);
    // row slabs
    slab_visit(129, 2, false, false, ORC_PROLONG_REFERENCE, 148, true);   // finest level, iterate exchanged
    slab_visit(129, 2, true, false, ORC_PROLONG_REFERENCE, 148, false);   // coarse level, first visit: f exchanged
    slab_visit(129, 3, false, false, ORC_PROLONG_FULL, 2, false);         // W re-visit on a coarse level, middle rank
    slab_visit(257, 2, false, true, ORC_PROLONG_REFERENCE, 148, true);    // interior / boundary split
    // the same with the halo prologue (opt-in flavour of Pass A)
    slab_visit(129, 2, false, false, ORC_PROLONG_REFERENCE, 148, true, true);
    slab_visit(129, 2, true, false, ORC_PROLONG_REFERENCE, 148, false, true);
    slab_visit(129, 3, false, false, ORC_PROLONG_FULL, 2, false, true);
    slab_visit(257, 4, false, true, ORC_PROLONG_REFERENCE, 148, true, true);
    if (full) {
        slab_visit(257, 4, true, true, ORC_PROLONG_REFERENCE, 4, false);
        slab_visit(257, 3, false, true, ORC_PROLONG_FULL, 148, true);
        whole_level(129, 1.0, 2, 2, ORC_PROLONG_REFERENCE, true, 4, 3);
    }
    std::printf(g_bad ? "FAILED (%d)\n" : "all bit-identical\n", g_bad);
    return g_bad ? 1 : 0;
}
