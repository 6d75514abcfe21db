// kernels_small.cu -- the levels n <= 65 of a V- or W-cycle inside ONE CTA, second generation.
//
// Same operators, point for point, as MultigridSolver::v_cycle / w_cycle (2_part_MG/MultiGrid.hpp:57-136) restricted
// to small N -- results are bit-identical to the per-operator kernels -- but organised for LATENCY, because at these
// sizes nothing else matters: a W-cycle (gamma = 2) at N = 16385 visits the 5x5 grid 4096 times per cycle, and in a
// V-cycle on 8 GPUs this kernel is on every rank's critical path.
//
//   * every level gets a thread group sized to its interior (63x63 -> 1024 threads, 31x31 -> 1024, 15x15 -> 256,
//     7x7 -> 64, 3x3 -> one warp): one point per thread per stage wherever it fits, so a stage is one dependent
//     chain (shared load -> 6-8 fp64 ops -> shared store) instead of a strided loop;
//   * a level's group synchronises on ITS OWN named barrier (bar.sync id = level + 1 with the group's thread count;
//     __syncwarp for a one-warp group).  Warps outside the group run ahead through the (uniform) cycle state machine
//     and park at the barrier that opens the next operator they take part in, so the deep levels -- visited gamma^k
//     times -- never pay for a 1024-thread barrier;
//   * the per-level table lives in shared memory and all per-thread state (ping-pong parities, visit counters) in
//     registers: no dynamically indexed local arrays, hence no local-memory traffic on the dependent chain;
//   * sweeps write only interior points: both ping-pong buffers carry the same ring (loaded once on the top level,
//     identically zero on the error levels below it).
#include "pmg_internal.h"

namespace pmg {

namespace {

constexpr int VS_MAXL = 8;

struct VsTable {
    int n[VS_MAXL];        // points per side
    int off[VS_MAXL];      // offset (doubles) of the level's three n*n arrays: ping, pong, f
    int shift[VS_MAXL];    // log2 of the columns of the level's thread grid (>= n - 2)
    int rows[VS_MAXL];     // rows of the thread grid
    int threads[VS_MAXL];  // cols * rows
    int warps[VS_MAXL];    // warps in the level's group
    double h2[VS_MAXL], inv_h2[VS_MAXL];
    int nl;
};

__host__ __device__ inline void vs_grid(int n, int max_threads, int &shift, int &rows)
{
    int m = n - 2 < 1 ? 1 : n - 2;
    shift = 0;
    while ((1 << shift) < m) ++shift;
    int cols = 1 << shift;
    rows = cols;
    if (rows > m) rows = m;
    if (cols * rows > max_threads) rows = max_threads / cols;
    if (rows < 1) rows = 1;
}

__device__ __forceinline__ void vs_sync(int k, int warps)
{
    if (warps == 1)
        __syncwarp();
    else
#ifdef PMG_HOST_EMULATION
        emu_bar_sync(k + 1, warps * 32);
#else
        asm volatile("bar.sync %0, %1;" ::"r"(k + 1), "r"(warps * 32) : "memory");
#endif
}

#ifdef PMG_HOST_EMULATION
alignas(16) static double g_vs2_smem[200 * 1024 / 8];
#else
extern __shared__ __align__(16) double g_vs2_smem[];
#endif

// `sweeps` weighted-Jacobi sweeps (Smoother.hpp:61-70) of a level's interior by its thread group; returns with the
// result in `cur` (pointers swapped in registers), every sweep closed by the group's barrier
template <bool WEIGHTED>
__device__ __forceinline__ void vs_sweeps(double *&cur, double *&oth, const double *f, int n, int tx, int ty, int rows,
                                          bool active, const JacobiCoef &c, int sweeps, int k, int warps)
{
    for (int s = 0; s < sweeps; ++s) {
        if (active)
            for (int y = 1 + ty; y <= n - 2; y += rows) {
                const int i = y * n + 1 + tx;
                double acc = dadd(dadd(dadd(dadd(dmul(c.h2, f[i]), cur[i - 1]), cur[i + 1]), cur[i - n]), cur[i + n]);
                oth[i] = WEIGHTED ? dadd(dmul(c.om1, cur[i]), dmul(c.w4, acc)) : dmul(0.25, acc);
            }
        vs_sync(k, warps);
        double *t = cur;
        cur = oth;
        oth = t;
    }
}

template <bool WEIGHTED>
__global__ void __launch_bounds__(1024)
    k_vcycle_small2(double *__restrict__ xg, const double *__restrict__ fg, int n0, int pitch_x, int pitch_f,
                    int n_coarse, double h0, double omega, int nu1, int nu2, int coarse_sweeps, int lo, int x_is_zero,
                    int gamma, const int *__restrict__ done)
{
    __shared__ VsTable T;
    pdl_prologue();
    if (done != nullptr && *done) return;
    const int t = threadIdx.x;
    if (t == 0) {
        int m = n0, nl = 0, off = 0;
        double hh = h0;
        for (;;) {
            T.n[nl] = m;
            T.off[nl] = off;
            vs_grid(m, (int)blockDim.x, T.shift[nl], T.rows[nl]);
            T.threads[nl] = (1 << T.shift[nl]) * T.rows[nl];
            T.warps[nl] = (T.threads[nl] + 31) / 32;
            T.h2[nl] = hh * hh;
            T.inv_h2[nl] = 1.0 / (hh * hh);
            off += 3 * m * m;
            ++nl;
            if (m <= n_coarse || m <= 3 || nl == VS_MAXL) break;
            m = (m - 1) / 2 + 1;
            hh = 2 * hh;  // MultiGrid.hpp:83
        }
        T.nl = nl;
    }
    __syncthreads();
    const int nl = T.nl;
    // zero everything once: the rings of all levels below the top one stay zero for the whole kernel
    {
        const int total = T.off[nl - 1] + 3 * T.n[nl - 1] * T.n[nl - 1];
        for (int i = t; i < total; i += blockDim.x) g_vs2_smem[i] = 0.0;
    }
    __syncthreads();
    // top level: f and the iterate (ring included, mirrored into both ping-pong buffers) from global memory
    {
        const int P = 1 << T.shift[0], R = T.rows[0];
        double *a = g_vs2_smem, *b = a + n0 * n0, *fs = b + n0 * n0;
        if (t < T.threads[0])
            for (int y = t >> T.shift[0]; y < n0; y += R)
                for (int x = t & (P - 1); x < n0; x += P) {
                    const int i = y * n0 + x;
                    fs[i] = fg[(size_t)y * pitch_f + x];
                    const double v = x_is_zero ? 0.0 : xg[(size_t)y * pitch_x + x];
                    a[i] = v;
                    b[i] = v;
                }
    }
    __syncthreads();

    JacobiCoef c;
    c.omega = omega;
    c.om1 = 1.0 - omega;
    c.w4 = 0.25 * omega;
    c.weighted = WEIGHTED ? 1 : 0;
    c.h2 = 0.0;
    const int warp = t >> 5;
    unsigned par = 0;             // bit k: level k's current iterate is in its second buffer
    unsigned long long vis = 0;   // 8 bits per level: coarse visits made during the current visit of level k

    // one operator of the cycle on level k, executed by that level's thread group
    auto level_view = [&](int k, double *&cur, double *&oth, double *&f, int &n, int &tx, int &ty, int &rows,
                          bool &active, int &warps) {
        n = T.n[k];
        double *base = g_vs2_smem + T.off[k];
        const int nn = n * n;
        const bool second = (par >> k) & 1u;
        cur = base + (second ? nn : 0);
        oth = base + (second ? 0 : nn);
        f = base + 2 * nn;
        const int sh = T.shift[k];
        tx = t & ((1 << sh) - 1);
        ty = t >> sh;
        rows = T.rows[k];
        active = (t < T.threads[k]) && (tx < n - 2);
        warps = T.warps[k];
        c.h2 = T.h2[k];
    };

    if (nl == 1) {
        double *cur, *oth, *f;
        int n, tx, ty, rows, warps;
        bool active;
        level_view(0, cur, oth, f, n, tx, ty, rows, active, warps);
        if (warp < warps) vs_sweeps<WEIGHTED>(cur, oth, f, n, tx, ty, rows, active, c, coarse_sweeps, 0, warps);
        par ^= (unsigned)(coarse_sweeps & 1);
    } else {
        int k = 0;
        bool down = true;
        for (;;) {
            double *cur, *oth, *f;
            int n, tx, ty, rows, warps;
            bool active;
            if (down) {
                level_view(k, cur, oth, f, n, tx, ty, rows, active, warps);
                if (k == nl - 1) {  // coarsest level (MultiGrid.hpp:59-63)
                    if (warp < warps) vs_sweeps<WEIGHTED>(cur, oth, f, n, tx, ty, rows, active, c, coarse_sweeps, k, warps);
                    par ^= (unsigned)(coarse_sweeps & 1) << k;
                    down = false;
                    --k;
                } else {
                    // pre-smooth, residual, restriction into level k+1, whose iterate is zeroed (MultiGrid.hpp:66-82)
                    if (warp < warps) {
                        vs_sweeps<WEIGHTED>(cur, oth, f, n, tx, ty, rows, active, c, nu1, k, warps);
                        const double inv_h2 = T.inv_h2[k];
                        double *r = oth;
                        if (active)
                            for (int y = 1 + ty; y <= n - 2; y += rows) {
                                const int i = y * n + 1 + tx;
                                r[i] = residual_point(inv_h2, f[i], cur[i], cur[i - 1], cur[i + 1], cur[i - n], cur[i + n]);
                            }
                        vs_sync(k, warps);
                        const int mc = T.n[k + 1];
                        double *cb = g_vs2_smem + T.off[k + 1];
                        double *xc = cb + (((par >> (k + 1)) & 1u) ? mc * mc : 0);
                        double *fc = cb + 2 * mc * mc;
                        if (t < T.threads[k] && tx < mc - 2)
                            for (int jc = 1 + ty; jc <= mc - 2; jc += rows) {
                                const int ic = 1 + tx;
                                const double *q = r + (2 * jc) * n + 2 * ic;
                                fc[jc * mc + ic] = restrict_point(q[0], q[1], q[-1], q[n], q[-n], q[-n - 1], q[-n + 1],
                                                                  q[n - 1], q[n + 1]);
                                xc[jc * mc + ic] = 0.0;
                            }
                        vs_sync(k, warps);
                    }
                    par ^= (unsigned)(nu1 & 1) << k;
                    vis &= ~(0xffull << (8 * k));
                    ++k;
                }
            } else {  // one visit of level k+1 has finished
                const unsigned done_visits = (unsigned)((vis >> (8 * k)) & 0xffull) + 1u;
                vis = (vis & ~(0xffull << (8 * k))) | ((unsigned long long)done_visits << (8 * k));
                if ((int)done_visits < gamma) {
                    ++k;
                    down = true;
                } else {
                    // prolongation-and-add from level k+1, post-smooth (MultiGrid.hpp:86-89)
                    level_view(k, cur, oth, f, n, tx, ty, rows, active, warps);
                    if (warp < warps) {
                        // warps of this group that sat out the coarser levels wait here for the ones that did not
                        if (warps > T.warps[k + 1]) vs_sync(k, warps);
                        const int mc = T.n[k + 1];
                        const double *cb = g_vs2_smem + T.off[k + 1];
                        const double *e = cb + (((par >> (k + 1)) & 1u) ? mc * mc : 0);
                        if (active && 1 + tx >= lo)
                            for (int y = 1 + ty; y <= n - 2; y += rows) {
                                if (y < lo) continue;
                                const int x = 1 + tx;
                                const int i = y * n + x;
                                const double *q = e + (y >> 1) * mc + (x >> 1);
                                double v;
                                if ((y & 1) == 0)
                                    v = ((x & 1) == 0) ? q[0] : dmul(0.5, dadd(q[0], q[1]));
                                else
                                    v = ((x & 1) == 0) ? dmul(0.5, dadd(q[0], q[mc]))
                                                       : dmul(0.25, dadd(dadd(dadd(q[0], q[1]), q[mc]), q[mc + 1]));
                                cur[i] = dadd(cur[i], v);
                            }
                        vs_sync(k, warps);
                        vs_sweeps<WEIGHTED>(cur, oth, f, n, tx, ty, rows, active, c, nu2, k, warps);
                    }
                    par ^= (unsigned)(nu2 & 1) << k;
                    if (k == 0) break;
                    --k;
                }
            }
        }
    }
    __syncthreads();
    {
        const int P = 1 << T.shift[0], R = T.rows[0];
        const double *res = g_vs2_smem + ((par & 1u) ? n0 * n0 : 0);
        if (t < T.threads[0])
            for (int y = t >> T.shift[0]; y < n0; y += R)
                for (int x = t & (P - 1); x < n0; x += P) xg[(size_t)y * pitch_x + x] = res[y * n0 + x];
    }
}

size_t vs2_smem_bytes(int n0, int n_coarse)
{
    size_t d = 0;
    for (int m = n0, k = 0; k < VS_MAXL; ++k) {
        d += 3 * (size_t)m * m;
        if (m <= n_coarse || m <= 3) break;
        m = (m - 1) / 2 + 1;
    }
    return d * sizeof(double);
}

}  // namespace

#ifndef PMG_HOST_EMULATION
bool vcycle_small_v2_supported(int gamma) { return gamma >= 1 && gamma <= 255; }

void launch_vcycle_small_v2(double *x, const double *f, int n0, int pitch_x, int pitch_f, int n_coarse, double h0,
                            double omega, int nu1, int nu2, int coarse_sweeps, int prolong_mode, bool x_is_zero,
                            int gamma, cudaStream_t st, const int *done)
{
    const size_t smem = vs2_smem_bytes(n0, n_coarse);
    {  // per-device attribute: once per device ordinal, not once per process
        static unsigned long long mask = 0;
        int dev = 0;
        cudaGetDevice(&dev);
        if (!((mask >> (dev & 63)) & 1ull)) {
            cudaFuncSetAttribute(k_vcycle_small2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            cudaFuncSetAttribute(k_vcycle_small2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            mask |= 1ull << (dev & 63);
        }
    }
    int shift = 0, rows = 1;
    vs_grid(n0, 1024, shift, rows);
    int threads = (((1 << shift) * rows + 31) / 32) * 32;
    if (threads > 1024) threads = 1024;
    const int lo = prolong_mode == PMG_PROLONG_FULL ? 1 : 2;
    if (omega != 1.0)
        launch_pdl(k_vcycle_small2<true>, dim3(1), dim3(threads), smem, st, x, f, n0, pitch_x, pitch_f, n_coarse, h0, omega, nu1,
                   nu2, coarse_sweeps, lo, x_is_zero ? 1 : 0, gamma, done);
    else
        launch_pdl(k_vcycle_small2<false>, dim3(1), dim3(threads), smem, st, x, f, n0, pitch_x, pitch_f, n_coarse, h0, omega, nu1,
                   nu2, coarse_sweeps, lo, x_is_zero ? 1 : 0, gamma, done);
    count_launch();
}

#endif  // PMG_HOST_EMULATION

}  // namespace pmg
