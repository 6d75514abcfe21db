// kernels_coarse.cu -- the coarse end of a V- or W-cycle on chip: third generation of the small-level kernel.
//
// Same operators, point for point and in the reference's evaluation order, as MultigridSolver::v_cycle / w_cycle
// (2_part_MG/MultiGrid.hpp:57-136) -- results are bit-identical to the per-operator kernels -- organised around what
// the B200 measures for a dependent step at these sizes (tools/micro/latency.cu, profiles/r2_latency_micro.log):
//
//      dependent fp64 add / mul                      8 cycles
//      STS -> __syncwarp -> LDS of a neighbour      42 cycles        (one warp)
//      STS -> bar.sync(256 threads) -> LDS          43 cycles per barrier,  91 with 1024 threads
//      one weighted-Jacobi point through smem       90 cycles per sweep    (one warp)
//
// The second generation (kernels_small.cu) spent ~545 cycles per barrier stage: every stage re-derived its indices from
// a run-time level table, re-loaded f, and walked a generic row loop.  Here the level size N is a TEMPLATE parameter:
//   * thread mapping, strides and shared-memory offsets are compile-time constants -- a sweep is 4 LDS, 7 dependent
//     fp64 operations, one STS and the group's barrier;
//   * h^2 f and the centre value stay in registers across the sweeps of a stage;
//   * the cycle recursion is the template recursion visit<N> -> visit<(N-1)/2+1> (gamma calls), no state machine;
//   * the CTA has 8 warps (256 threads: a 32-warp CTA caps every thread at 64 registers, and the deep levels then run
//     at half speed; the big levels are bound by shared-memory bandwidth and the fp64 pipe, not by thread count):
//     65: 16 points per thread, 33: 4, 17: one; 9 and 5: ONE warp, __syncwarp only.  A group synchronises on its own
//     named barrier; warps outside a group park on the parent's join barrier while the deep levels -- visited gamma^k
//     times -- run.
// A W(gamma = 2) visit of the levels 65..5 drops from 77 us to about 20 us, a V visit from 27 us to about 7 us.
#include "pmg_internal.h"

#ifndef PMG_HOST_EMULATION
#include <cooperative_groups.h>
#endif

namespace pmg {

namespace {

// Shared memory and the three thread-block-cluster primitives the cluster kernel uses.  Under PMG_HOST_EMULATION
// (tests/cpp/emu: the kernel SOURCE run on the CPU, test infrastructure only) they act on the emulated CTAs instead.
#ifdef PMG_HOST_EMULATION
#define g_co_smem (emu_cta_smem_doubles())
__device__ __forceinline__ int co_cluster_rank() { return (int)blockIdx.x; }
__device__ __forceinline__ double *co_map_rank(double *p, int rank) { return emu_map_shared_rank(p, rank); }
__device__ __forceinline__ void co_cluster_sync() { emu_cluster_sync(); }
#else
extern __shared__ __align__(16) double g_co_smem[];
__device__ __forceinline__ int co_cluster_rank() { return (int)cooperative_groups::this_cluster().block_rank(); }
__device__ __forceinline__ double *co_map_rank(double *p, int rank)
{
    return cooperative_groups::this_cluster().map_shared_rank(p, (unsigned)rank);
}
// barrier.cluster.arrive.release + wait.acquire: DSMEM stores issued before it are visible to every CTA after it
__device__ __forceinline__ void co_cluster_sync() { cooperative_groups::this_cluster().sync(); }
#endif

__host__ __device__ constexpr int co_log2(int v) { return v <= 1 ? 0 : 1 + co_log2(v >> 1); }

// three N x N arrays per level (ping, pong, f), small levels first so that a kernel whose top level is small needs
// only the front of the layout: offsets in doubles
__host__ __device__ constexpr int co_off(int n) { return n <= 3 ? 0 : co_off((n - 1) / 2 + 1) + 3 * ((n - 1) / 2 + 1) * ((n - 1) / 2 + 1); }
__host__ __device__ constexpr int co_total(int n) { return co_off(n) + 3 * n * n; }

template <int N>
struct Geo {
    static constexpr int M = N - 2;  // interior points per side
    static constexpr int TX = (N >= 65) ? 64 : (N >= 33) ? 32 : (N >= 17) ? 16 : (N >= 9) ? 8 : (N >= 5) ? 4 : 1;
    static constexpr int WARPS = (N >= 17) ? 8 : 1;
    static constexpr int THREADS = WARPS * 32;
    static constexpr int TY = THREADS / TX;             // rows of the thread grid
    static constexpr int PTS = (M + TY - 1) / TY;       // points per thread (row y = 1 + ty + p * TY)
    static constexpr int BAR = co_log2(N - 1);          // named barrier of the level's group (1..6)
    static_assert(TX >= M && N <= 65 && N >= 3, "level size");
};

template <int N>
__device__ __forceinline__ void group_sync()
{
    if (Geo<N>::WARPS == 1) {
        __syncwarp();
    } else {
#ifdef PMG_HOST_EMULATION
        emu_bar_sync(Geo<N>::BAR, Geo<N>::THREADS);
#else
        asm volatile("bar.sync %0, %1;" ::"n"(Geo<N>::BAR), "n"(Geo<N>::THREADS) : "memory");
#endif
    }
}

struct CoCtx {
    double *sm;        // start of the level layout (co_off)
    double h0;         // spacing of the top level n0
    int n0;
    int n_coarse, nu1, nu2, coarse_sweeps, lo, gamma;
    JacobiCoef coef;   // h2 is set per level
    unsigned par;      // bit Geo<N>::BAR: level N's current iterate is in its second buffer
    const double *htab;  // shared memory: [2k] = h^2, [2k+1] = 1/h^2 of level 2^k + 1 (co_fill_htab)
    bool prof_cta0;    // this thread's CTA is the one that reports the profile
    long long tsum;    // (thread 0 only) cycles attributed to finished visits so far: per-level profile, see co_prof
};

// Per-level cycle profile of the last launch (thread 0 of CTA 0): own cycles of all visits of level 2^k + 1 in slot k,
// slot 0 = the whole kernel.  Costs two clock reads per visit; read back through pmg_coarse_profile (tools/small_kernel_probe.py).
#ifdef PMG_HOST_EMULATION
inline long long co_clock() { return 0; }
static long long g_co_prof[16];
#else
__device__ __forceinline__ long long co_clock() { return clock64(); }
__device__ long long g_co_prof[16];
#endif
struct CoProf {
    long long t0, s0;
    __device__ __forceinline__ void begin(const CoCtx &c)
    {
        t0 = co_clock();
        s0 = c.tsum;
    }
    __device__ __forceinline__ void end(CoCtx &c, int slot, bool mine)
    {
        const long long own = (co_clock() - t0) - (c.tsum - s0);
        c.tsum += own;
        if (mine) g_co_prof[slot] += own;
    }
};

// h^2 and 1/h^2 of every level, computed ONCE per launch with the expressions the per-operator kernels use
// (h doubles per level, MultiGrid.hpp:83 -- exact; 1.0 / (h * h) as in launch_residual): an fp64 division costs several
// hundred cycles, and a W-cycle visits the deep levels hundreds of times per launch
constexpr int CO_HTAB = 20;
__device__ __forceinline__ void co_fill_htab(double *tab, int t, double h0, int n0)
{
    if (t >= 1 && t <= 9) {
        const int n = (1 << t) + 1;
        if (n <= n0) {
            const double h = h0 * (double)((n0 - 1) / (n - 1));
            tab[2 * t] = h * h;
            tab[2 * t + 1] = 1.0 / (h * h);
        }
    }
}

// bilinear prolongation weight pattern (MultiGrid.hpp:208-226) without divergent branches: all four coarse neighbours are
// loaded (they exist for every interior fine point), the three candidate sums are formed in the reference's order and the
// one that applies to the point's parity is selected -- the selected value is bit-identical to the branchy form
__device__ __forceinline__ double prolong_add(double fine, const double *q, int nc, int x, int y)
{
    const double q0 = q[0], q1 = q[1], qn = q[nc], qn1 = q[nc + 1];
    const double sx = dadd(q0, q1);
    const double sy = dadd(q0, qn);
    const double sxy = dadd(dadd(sx, qn), qn1);
    const bool ox = (x & 1) != 0, oy = (y & 1) != 0;
    const double sum = oy ? (ox ? sxy : sy) : (ox ? sx : q0);
    const double w = oy ? (ox ? 0.25 : 0.5) : (ox ? 0.5 : 1.0);
    return dfma_pow2(w, sum, fine);  // the weight is a power of two: one rounding, that of fine + w * sum
}

// this thread's points of level N: index into an N x N array, or -1
template <int N>
struct Pts {
    int idx[Geo<N>::PTS];
    bool act[Geo<N>::PTS];
    __device__ __forceinline__ void init(int t)
    {
        const int tx = t & (Geo<N>::TX - 1), ty = t / Geo<N>::TX;
#pragma unroll
        for (int p = 0; p < Geo<N>::PTS; ++p) {
            const int y = 1 + ty + p * Geo<N>::TY;
            act[p] = (tx < Geo<N>::M) && (y <= Geo<N>::M);
            idx[p] = act[p] ? y * N + 1 + tx : N + 1;  // inactive threads address a valid interior point and never store
        }
    }
};

// `count` weighted-Jacobi sweeps (Smoother.hpp:61-70) by the level's group; result in `cur` (pointers swapped)
template <int N, bool WEIGHTED>
__device__ __forceinline__ void co_sweeps(double *&cur, double *&oth, const double *f, const JacobiCoef &c, int count,
                                          const Pts<N> &P)
{
    double hf[Geo<N>::PTS], xc[Geo<N>::PTS];
#pragma unroll
    for (int p = 0; p < Geo<N>::PTS; ++p) {
        hf[p] = dmul(c.h2, f[P.idx[p]]);
        xc[p] = cur[P.idx[p]];
    }
    for (int s = 0; s < count; ++s) {
#pragma unroll
        for (int p = 0; p < Geo<N>::PTS; ++p) {
            const int i = P.idx[p];
            double acc = dadd(dadd(dadd(dadd(hf[p], cur[i - 1]), cur[i + 1]), cur[i - N]), cur[i + N]);
            double v = WEIGHTED ? dadd(dmul(c.om1, xc[p]), dmul(c.w4, acc)) : dmul(0.25, acc);
            if (P.act[p]) oth[i] = v;
            xc[p] = v;
        }
        group_sync<N>();
        double *tmp = cur;
        cur = oth;
        oth = tmp;
    }
}

// One visit of level N by its thread group (t < Geo<N>::THREADS): MultiGrid.hpp:57-94 (V) / 96-136 (W).
template <int N, bool WEIGHTED>
__device__ void co_visit(CoCtx &c, const int t)
{
    using G = Geo<N>;
    constexpr int NN = N * N;
    double *base = c.sm + co_off(N);
    const bool second = (c.par >> G::BAR) & 1u;
    double *cur = base + (second ? NN : 0), *oth = base + (second ? 0 : NN);
    const double *f = base + 2 * NN;
    c.coef.h2 = c.htab[2 * G::BAR];
    Pts<N> P;
    P.init(t);
    CoProf prof;
    prof.begin(c);
    const bool prof_mine = (t == 0) && (c.prof_cta0);
    if (N <= 3 || N <= c.n_coarse) {  // coarsest level (MultiGrid.hpp:59-63)
        co_sweeps<N, WEIGHTED>(cur, oth, f, c.coef, c.coarse_sweeps, P);
        c.par ^= (unsigned)(c.coarse_sweeps & 1) << G::BAR;
        prof.end(c, G::BAR, prof_mine);
        return;
    }
    if constexpr (N > 3) {
        constexpr int NC = (N - 1) / 2 + 1;
        using GC = Geo<NC>;
        // pre-smooth, residual (MultiGrid.hpp:66-71); r lives in the spare buffer
        co_sweeps<N, WEIGHTED>(cur, oth, f, c.coef, c.nu1, P);
        c.par ^= (unsigned)(c.nu1 & 1) << G::BAR;
        {
            const double inv_h2 = c.htab[2 * G::BAR + 1];
#pragma unroll
            for (int p = 0; p < G::PTS; ++p) {
                const int i = P.idx[p];
                double r = residual_point(inv_h2, f[i], cur[i], cur[i - 1], cur[i + 1], cur[i - N], cur[i + N]);
                if (P.act[p]) oth[i] = r;
            }
            group_sync<N>();
        }
        const double h_own2 = c.coef.h2;
        if (t < GC::THREADS) {
            // restriction into level NC, whose iterate is zeroed (MultiGrid.hpp:74-82), by the coarse group
            double *cb = c.sm + co_off(NC);
            double *xc = cb + (((c.par >> GC::BAR) & 1u) ? NC * NC : 0);
            double *fc = cb + 2 * NC * NC;
            Pts<NC> Q;
            Q.init(t);
#pragma unroll
            for (int p = 0; p < GC::PTS; ++p) {
                const int ic = Q.idx[p] % NC, jc = Q.idx[p] / NC;
                const double *q = oth + (2 * jc) * N + 2 * ic;
                double v = restrict_point(q[0], q[1], q[-1], q[N], q[-N], q[-N - 1], q[-N + 1], q[N - 1], q[N + 1]);
                if (Q.act[p]) {
                    fc[Q.idx[p]] = v;
                    xc[Q.idx[p]] = 0.0;
                }
            }
            group_sync<NC>();
            for (int v = 0; v < c.gamma; ++v) co_visit<NC, WEIGHTED>(c, t);  // :81-83 / :121-125
        } else {
            // warps that sit the coarser levels out still need level NC's buffer parity for the prolongation below:
            // every visit flips it by the same, known amount
            const unsigned per_visit = (unsigned)(((NC <= 3 || NC <= c.n_coarse) ? c.coarse_sweeps : c.nu1 + c.nu2) & 1);
            c.par ^= (per_visit & (unsigned)(c.gamma & 1)) << GC::BAR;
        }
        if (G::WARPS > GC::WARPS) group_sync<N>();  // join: the warps that sat out the coarser levels wait here
        c.coef.h2 = h_own2;
        // prolongation-and-add (MultiGrid.hpp:86, :208-226), post-smooth (:89)
        {
            const double *cb = c.sm + co_off(NC);
            const double *e = cb + (((c.par >> GC::BAR) & 1u) ? NC * NC : 0);
#pragma unroll
            for (int p = 0; p < G::PTS; ++p) {
                const int i = P.idx[p];
                const int y = i / N, x = i - y * N;
                const double v = prolong_add(cur[i], e + (y >> 1) * NC + (x >> 1), NC, x, y);
                if (P.act[p] && y >= c.lo && x >= c.lo) cur[i] = v;
            }
            group_sync<N>();
        }
        co_sweeps<N, WEIGHTED>(cur, oth, f, c.coef, c.nu2, P);
        c.par ^= (unsigned)(c.nu2 & 1) << G::BAR;
    }
    prof.end(c, G::BAR, prof_mine);
}

// The levels n0 (<= 65) ... n_coarse of one V- (gamma = 1) or W-cycle in a single CTA.  x (whole array incl. ring)
// receives the result; x_is_zero: start from 0 instead of reading x.
template <int N0, bool WEIGHTED>
__global__ void __launch_bounds__(Geo<N0>::THREADS)
    k_coarse_local(double *__restrict__ xg, const double *__restrict__ fg, int pitch_x, int pitch_f, int n_coarse, double h0,
                   double omega, int nu1, int nu2, int coarse_sweeps, int lo, int x_is_zero, int gamma,
                   const int *__restrict__ done)
{
    __shared__ double htab[CO_HTAB];
    pdl_prologue();
    if (done != nullptr && *done) return;
    const int t = threadIdx.x;
    constexpr int NN = N0 * N0;
    co_fill_htab(htab, t, h0, N0);
    // zero the levels below the top one: their rings stay zero for the whole kernel
    for (int i = t; i < co_off(N0); i += Geo<N0>::THREADS) g_co_smem[i] = 0.0;
    // top level: f and the iterate (ring included, mirrored into both buffers) from global memory
    {
        double *a = g_co_smem + co_off(N0), *b = a + NN, *fs = b + NN;
        for (int i = t; i < NN; i += Geo<N0>::THREADS) {
            const int y = i / N0, x = i - y * N0;
            fs[i] = fg[(size_t)y * pitch_f + x];
            const double v = x_is_zero ? 0.0 : xg[(size_t)y * pitch_x + x];
            a[i] = v;
            b[i] = v;
        }
    }
    group_sync<N0>();
    CoCtx c;
    c.sm = g_co_smem;
    c.h0 = h0;
    c.n0 = N0;
    c.n_coarse = n_coarse;
    c.nu1 = nu1;
    c.nu2 = nu2;
    c.coarse_sweeps = coarse_sweeps;
    c.lo = lo;
    c.gamma = gamma;
    c.coef.omega = omega;
    c.coef.om1 = 1.0 - omega;
    c.coef.w4 = 0.25 * omega;
    c.coef.weighted = WEIGHTED ? 1 : 0;
    c.coef.h2 = 0.0;
    c.par = 0u;
    c.htab = htab;
    c.tsum = 0;
    c.prof_cta0 = true;
    const long long k0 = co_clock();
    if (t == 0)
        for (int i = 0; i < 16; ++i) g_co_prof[i] = 0;
    co_visit<N0, WEIGHTED>(c, t);
    {
        const double *res = g_co_smem + co_off(N0) + (((c.par >> Geo<N0>::BAR) & 1u) ? NN : 0);
        for (int i = t; i < NN; i += Geo<N0>::THREADS) {
            const int y = i / N0, x = i - y * N0;
            xg[(size_t)y * pitch_x + x] = res[i];
        }
    }
    if (t == 0) g_co_prof[0] = co_clock() - k0;
}

template <int N0>
void coarse_local_launch(double *x, const double *f, int pitch_x, int pitch_f, int n_coarse, double h0, double omega,
                         int nu1, int nu2, int coarse_sweeps, int lo, bool x_is_zero, int gamma, cudaStream_t st,
                         const int *done)
{
    const size_t smem = (size_t)co_total(N0) * sizeof(double);
#ifdef PMG_HOST_EMULATION
    (void)smem;
    (void)st;
    emu_launch_threads(Geo<N0>::THREADS, [&] {
        if (omega != 1.0)
            k_coarse_local<N0, true>(x, f, pitch_x, pitch_f, n_coarse, h0, omega, nu1, nu2, coarse_sweeps, lo, x_is_zero ? 1 : 0,
                                     gamma, done);
        else
            k_coarse_local<N0, false>(x, f, pitch_x, pitch_f, n_coarse, h0, omega, nu1, nu2, coarse_sweeps, lo, x_is_zero ? 1 : 0,
                                      gamma, done);
    });
#else
    {  // per-device attribute: once per (kernel, device ordinal)
        static unsigned long long mask = 0;
        int dev = 0;
        cudaGetDevice(&dev);
        if (!((mask >> (dev & 63)) & 1ull)) {
            cudaFuncSetAttribute(k_coarse_local<N0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(k_coarse_local<N0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            mask |= 1ull << (dev & 63);
        }
    }
    if (omega != 1.0)
        launch_pdl(k_coarse_local<N0, true>, dim3(1), dim3(Geo<N0>::THREADS), smem, st, x, f, pitch_x, pitch_f, n_coarse, h0,
                   omega, nu1, nu2, coarse_sweeps, lo, x_is_zero ? 1 : 0, gamma, done);
    else
        launch_pdl(k_coarse_local<N0, false>, dim3(1), dim3(Geo<N0>::THREADS), smem, st, x, f, pitch_x, pitch_f, n_coarse, h0,
                   omega, nu1, nu2, coarse_sweeps, lo, x_is_zero ? 1 : 0, gamma, done);
#endif
}


// =====================================================================================================================
// The cluster kernel: the levels NTOP (257 or 129) ... n_coarse of one V- or W-cycle in ONE launch of a 16-CTA
// thread-block cluster, the levels >= 33 distributed over the CTAs' shared memories by rows (DSMEM).
//
//   * 256 threads per CTA; level N >= 65: CTA c owns rows [c R, (c+1) R), R = (N-1)/16 (the last CTA also keeps the ring row), stored with
//     one halo row above and below.  A sweep computes the own rows, STORES its first / last row straight into the
//     neighbours' halo rows (st.shared::cluster through map_shared_rank) and closes with the cluster barrier --
//     5 cluster barriers per level visit (nu1 + nu2 sweeps and the residual); restriction and prolongation are local:
//     coarse row jc lives in the CTA that owns fine row 2 jc, and the prolongation recomputes the fine halo rows from
//     the coarse halo rows instead of exchanging them (same arithmetic, hence the same bits as the owner's);
//   * levels <= 33 live in CTA 0 and run the single-CTA code above (co_visit<33>): the restriction 65 -> 33 writes
//     into CTA 0's shared memory, the prolongation 33 -> 65 reads it back from there; the other CTAs wait on the
//     cluster barrier meanwhile.
// 257^2 x 3 arrays x 8 B = 1.6 MB does not fit one SM (227 KB) but fits sixteen (164 KB each); shared-memory bandwidth
// and fp64 throughput -- what bounds a 65^2 or 33^2 sweep inside ONE SM -- scale with the CTAs.  It replaces the
// streaming passes of the levels 257 and 129 (5-6 us each, latency-bound: 4 launches) and the single-CTA kernel.
// =====================================================================================================================
constexpr int CL = 16;        // CTAs per cluster (non-portable size: cudaFuncAttributeNonPortableClusterSizeAllowed)
constexpr int CL_THREADS = 256;  // 8 warps: a cluster barrier costs 400 cycles instead of 600 with 32 warps, and no
                                  // thread is capped at 64 registers (spills are refetched from L2 after every cluster
                                  // barrier, which invalidates L1)
constexpr int DIST_MIN = 65;  // smallest distributed level; below it CTA 0 works alone (a 33^2 visit costs 5 300 cycles in
                              // one CTA and 8 400 distributed: 7 cluster barriers at ~1 000 cycles each)

template <int N>
struct DGeo {
    static constexpr int R = (N - 1) / CL;              // own rows per CTA
    static constexpr int TX = N - 1;                    // columns of the thread grid (power of two >= N - 2)
    static constexpr int TY = CL_THREADS / TX;
    static constexpr int PTS = (R + TY - 1) / TY;       // own points per thread (local row ty + p * TY)
    static constexpr int ROWS = R + 2;                  // halo above, own rows, halo below (= ring row in the last CTA)
    static constexpr int ELEMS = ROWS * N;
    static constexpr int BIT = co_log2(N - 1);          // parity bit in CoCtx::par
    static_assert(N >= DIST_MIN && (N - 1) % (2 * CL) == 0 && TX * TY == CL_THREADS, "distributed level size");
};

// per-CTA layout: the local levels (<= 17) first, then the distributed ones from small to large (offsets in doubles)
__host__ __device__ constexpr int d_off(int n) { return n <= DIST_MIN ? co_total((DIST_MIN - 1) / 2 + 1) : d_off((n - 1) / 2 + 1) + 3 * ((((n - 1) / 2 + 1) - 1) / CL + 2) * ((n - 1) / 2 + 1); }
__host__ __device__ constexpr int d_total(int n) { return d_off(n) + 3 * ((n - 1) / CL + 2) * n; }

template <int N>
struct DPts {
    int idx[DGeo<N>::PTS];
    bool act[DGeo<N>::PTS], push_up[DGeo<N>::PTS], push_dn[DGeo<N>::PTS];
    int tx;
    __device__ __forceinline__ void init(int t, int rank)
    {
        using G = DGeo<N>;
        tx = t & (G::TX - 1);
        const int ty = t / G::TX;
#pragma unroll
        for (int p = 0; p < G::PTS; ++p) {
            const int jr = ty + p * G::TY;       // local own row
            const int y = rank * G::R + jr;      // global row
            act[p] = (tx < N - 2) && (jr < G::R) && (y >= 1) && (y <= N - 2);
            idx[p] = act[p] ? (jr + 1) * N + 1 + tx : N + 1;
            push_up[p] = act[p] && jr == 0 && rank > 0;            // my first row is the upper neighbour's lower halo
            push_dn[p] = act[p] && jr == G::R - 1 && rank < CL - 1;  // my last row is the lower neighbour's upper halo
        }
    }
};

template <int N, bool WEIGHTED>
__device__ __forceinline__ void dist_sweeps(double *&cur, double *&oth, const double *f, const JacobiCoef &c, int count,
                                            const DPts<N> &P, int rank)
{
    using G = DGeo<N>;
    double hf[G::PTS], xc[G::PTS];
#pragma unroll
    for (int p = 0; p < G::PTS; ++p) {
        hf[p] = dmul(c.h2, f[P.idx[p]]);
        xc[p] = cur[P.idx[p]];
    }
    for (int s = 0; s < count; ++s) {
#pragma unroll
        for (int p = 0; p < G::PTS; ++p) {
            const int i = P.idx[p];
            double acc = dadd(dadd(dadd(dadd(hf[p], cur[i - 1]), cur[i + 1]), cur[i - N]), cur[i + N]);
            double v = WEIGHTED ? dadd(dmul(c.om1, xc[p]), dmul(c.w4, acc)) : dmul(0.25, acc);
            if (P.act[p]) oth[i] = v;
            if (P.push_up[p]) *co_map_rank(oth + (G::R + 1) * N + 1 + P.tx, rank - 1) = v;
            if (P.push_dn[p]) *co_map_rank(oth + 1 + P.tx, rank + 1) = v;
            xc[p] = v;
        }
        co_cluster_sync();
        double *tmp = cur;
        cur = oth;
        oth = tmp;
    }
}

// One visit of the distributed level N by the whole cluster (every thread of every CTA runs this).
template <int N, bool WEIGHTED>
__device__ void dist_visit(CoCtx &c, const int t, const int rank)
{
    using G = DGeo<N>;
    constexpr int NC = (N - 1) / 2 + 1;
    double *base = c.sm + d_off(N);
    const bool second = (c.par >> G::BIT) & 1u;
    double *cur = base + (second ? G::ELEMS : 0), *oth = base + (second ? 0 : G::ELEMS);
    const double *f = base + 2 * G::ELEMS;
    c.coef.h2 = c.htab[2 * G::BIT];
    const double h_own2 = c.coef.h2;
    DPts<N> P;
    P.init(t, rank);
    CoProf prof;
    prof.begin(c);
    // pre-smooth (MultiGrid.hpp:66)
    dist_sweeps<N, WEIGHTED>(cur, oth, f, c.coef, c.nu1, P, rank);
    c.par ^= (unsigned)(c.nu1 & 1) << G::BIT;
    // residual into the spare buffer (:69-71); the lower neighbour's restriction reads my last row
    {
        const double inv_h2 = c.htab[2 * G::BIT + 1];
#pragma unroll
        for (int p = 0; p < G::PTS; ++p) {
            const int i = P.idx[p];
            double r = residual_point(inv_h2, f[i], cur[i], cur[i - 1], cur[i + 1], cur[i - N], cur[i + N]);
            if (P.act[p]) oth[i] = r;
            if (P.push_dn[p]) *co_map_rank(oth + 1 + P.tx, rank + 1) = r;
        }
        co_cluster_sync();
    }
    if constexpr (NC >= DIST_MIN) {
        // restriction into the distributed level NC (coarse row jc lives with fine row 2 jc); its iterate, halo rows
        // included, is zeroed (:74-82).  Everything is local: a CTA barrier is enough.
        using GC = DGeo<NC>;
        double *cb = c.sm + d_off(NC);
        double *xcc = cb + (((c.par >> GC::BIT) & 1u) ? GC::ELEMS : 0);
        double *fc = cb + 2 * GC::ELEMS;
        const int txc = t & (GC::TX - 1), tyc = t / GC::TX;
#pragma unroll
        for (int p = 0; p < GC::PTS; ++p) {
            const int jrc = tyc + p * GC::TY;
            const int yc = rank * GC::R + jrc;
            if (txc < NC - 2 && jrc < GC::R && yc >= 1 && yc <= NC - 2) {
                const double *q = oth + (2 * jrc + 1) * N + 2 * (1 + txc);
                fc[(jrc + 1) * NC + 1 + txc] =
                    restrict_point(q[0], q[1], q[-1], q[N], q[-N], q[-N - 1], q[-N + 1], q[N - 1], q[N + 1]);
                xcc[(jrc + 1) * NC + 1 + txc] = 0.0;
            }
        }
        for (int x = t; x < NC; x += CL_THREADS) {
            xcc[x] = 0.0;
            xcc[(GC::R + 1) * NC + x] = 0.0;
        }
        __syncthreads();
        for (int v = 0; v < c.gamma; ++v) dist_visit<NC, WEIGHTED>(c, t, rank);  // :81-83 / :121-125
    } else {
        // restriction into the first LOCAL level NC, which lives in CTA 0: this CTA's R/2 coarse rows (fine rows 2 jc),
        // written straight into CTA 0's shared memory
        using GC = Geo<NC>;
        constexpr int RC = G::R / 2;
        static_assert(NC == (DIST_MIN - 1) / 2 + 1 && RC >= 1 && (NC - 2) * RC <= CL_THREADS, "the first local level");
        double *cb0 = co_map_rank(c.sm + co_off(NC), 0);
        double *xcc = cb0 + (((c.par >> GC::BAR) & 1u) ? NC * NC : 0);
        double *fc = cb0 + 2 * NC * NC;
        {
            const int jrc = t / (NC - 2), txc = t - jrc * (NC - 2);  // (coarse row of mine, coarse column - 1)
            const int yc = rank * RC + jrc;
            if (jrc < RC && yc >= 1 && yc <= NC - 2) {
                const double *q = oth + (2 * jrc + 1) * N + 2 * (1 + txc);
                fc[yc * NC + 1 + txc] =
                    restrict_point(q[0], q[1], q[-1], q[N], q[-N], q[-N - 1], q[-N + 1], q[N - 1], q[N + 1]);
                xcc[yc * NC + 1 + txc] = 0.0;
            }
        }
        co_cluster_sync();
        if (rank == 0 && t < GC::THREADS) {
            for (int v = 0; v < c.gamma; ++v) co_visit<NC, WEIGHTED>(c, t);
        } else {
            // everybody else still needs level NC's buffer parity for the prolongation: the same flips, known in advance
            const unsigned per_visit = (unsigned)(((NC <= 3 || NC <= c.n_coarse) ? c.coarse_sweeps : c.nu1 + c.nu2) & 1);
            c.par ^= (per_visit & (unsigned)(c.gamma & 1)) << GC::BAR;
        }
        co_cluster_sync();
    }
    c.coef.h2 = h_own2;
    // prolongation-and-add (:86, :208-226) on the own rows AND the two halo rows (recomputed, not exchanged)
    {
        const double *e;
        if constexpr (NC >= DIST_MIN) {
            using GC = DGeo<NC>;
            e = c.sm + d_off(NC) + (((c.par >> GC::BIT) & 1u) ? GC::ELEMS : 0);
        } else {
            e = co_map_rank(c.sm + co_off(NC) + (((c.par >> Geo<NC>::BAR) & 1u) ? NC * NC : 0), 0);
        }
        const int x = 1 + P.tx;
        for (int j = t / G::TX; j < G::ROWS; j += G::TY) {
            const int y = rank * G::R + j - 1;  // global row of stored row j
            if (P.tx < N - 2 && y >= c.lo && y <= N - 2 && x >= c.lo) {
                const double *q;
                if constexpr (NC >= DIST_MIN)
                    q = e + (((j - 1) >> 1) + 1) * NC + (x >> 1);  // stored coarse row of global coarse row y >> 1
                else
                    q = e + (y >> 1) * NC + (x >> 1);
                cur[j * N + x] = prolong_add(cur[j * N + x], q, NC, x, y);
            }
        }
        __syncthreads();
    }
    // post-smooth (:89)
    dist_sweeps<N, WEIGHTED>(cur, oth, f, c.coef, c.nu2, P, rank);
    c.par ^= (unsigned)(c.nu2 & 1) << G::BIT;
    prof.end(c, G::BIT, t == 0 && rank == 0);
}

template <int NTOP, bool WEIGHTED>
__global__ void __launch_bounds__(CL_THREADS, 1)
    k_coarse_cluster(double *__restrict__ xg, const double *__restrict__ fg, int pitch_x, int pitch_f, int n_coarse, double h0,
                     double omega, int nu1, int nu2, int coarse_sweeps, int lo, int x_is_zero, int gamma,
                     const int *__restrict__ done)
{
    using G = DGeo<NTOP>;
    __shared__ double htab[CO_HTAB];
    pdl_prologue();
    if (done != nullptr && *done) return;  // uniform over the cluster: nobody reaches a cluster barrier
    const int t = threadIdx.x;
    const int rank = co_cluster_rank();
    double *sm = g_co_smem;
    co_fill_htab(htab, t, h0, NTOP);
    // zero everything once: rings and halo rows of the levels below the top one stay zero where nothing writes them
    for (int i = t; i < d_total(NTOP); i += CL_THREADS) sm[i] = 0.0;
    __syncthreads();
    // top level: my rows with their halo rows, f and the iterate (mirrored into both buffers), from global memory
    {
        double *a = sm + d_off(NTOP), *b = a + G::ELEMS, *fs = b + G::ELEMS;
        for (int i = t; i < G::ELEMS; i += CL_THREADS) {
            const int j = i / NTOP, x = i - j * NTOP;
            const int y = rank * G::R + j - 1;
            if (y >= 0 && y <= NTOP - 1) {
                fs[i] = fg[(size_t)y * pitch_f + x];
                const double v = x_is_zero ? 0.0 : xg[(size_t)y * pitch_x + x];
                a[i] = v;
                b[i] = v;
            }
        }
    }
    co_cluster_sync();  // also: nobody pushes into a neighbour that is still zeroing its shared memory
    CoCtx c;
    c.sm = sm;
    c.h0 = h0;
    c.n0 = NTOP;
    c.n_coarse = n_coarse;
    c.nu1 = nu1;
    c.nu2 = nu2;
    c.coarse_sweeps = coarse_sweeps;
    c.lo = lo;
    c.gamma = gamma;
    c.coef.omega = omega;
    c.coef.om1 = 1.0 - omega;
    c.coef.w4 = 0.25 * omega;
    c.coef.weighted = WEIGHTED ? 1 : 0;
    c.coef.h2 = 0.0;
    c.par = 0u;
    c.htab = htab;
    c.tsum = 0;
    c.prof_cta0 = (rank == 0);
    const long long k0 = co_clock();
    if (t == 0 && rank == 0)
        for (int i = 0; i < 16; ++i) g_co_prof[i] = 0;
    dist_visit<NTOP, WEIGHTED>(c, t, rank);
    {
        const double *res = sm + d_off(NTOP) + (((c.par >> G::BIT) & 1u) ? G::ELEMS : 0);
        for (int i = t; i < G::R * NTOP; i += CL_THREADS) {
            const int jr = i / NTOP, x = i - jr * NTOP;
            xg[(size_t)(rank * G::R + jr) * pitch_x + x] = res[(jr + 1) * NTOP + x];
        }
        if (rank == CL - 1)  // the ring row travels with the last CTA
            for (int x = t; x < NTOP; x += CL_THREADS) xg[(size_t)(NTOP - 1) * pitch_x + x] = res[(G::R + 1) * NTOP + x];
    }
    if (t == 0 && rank == 0) g_co_prof[0] = co_clock() - k0;
}

template <int NTOP>
bool coarse_cluster_launch(double *x, const double *f, int pitch_x, int pitch_f, int n_coarse, double h0, double omega, int nu1,
                           int nu2, int coarse_sweeps, int lo, bool x_is_zero, int gamma, cudaStream_t st, const int *done,
                           bool probe_only = false)
{
    const size_t smem = (size_t)d_total(NTOP) * sizeof(double);
#ifdef PMG_HOST_EMULATION
    if (probe_only) return true;
    (void)smem;
    (void)st;
    emu_launch_cluster(CL, CL_THREADS, [&] {
        if (omega != 1.0)
            k_coarse_cluster<NTOP, true>(x, f, pitch_x, pitch_f, n_coarse, h0, omega, nu1, nu2, coarse_sweeps, lo,
                                         x_is_zero ? 1 : 0, gamma, done);
        else
            k_coarse_cluster<NTOP, false>(x, f, pitch_x, pitch_f, n_coarse, h0, omega, nu1, nu2, coarse_sweeps, lo,
                                          x_is_zero ? 1 : 0, gamma, done);
    });
    return true;
#else
    static int usable[64] = {0};  // per device ordinal: 0 unknown, 1 yes, -1 no (a 16-CTA cluster cannot be scheduled)
    int dev = 0;
    cudaGetDevice(&dev);
    int &ok = usable[dev & 63];
    auto kt = k_coarse_cluster<NTOP, true>;
    auto kf = k_coarse_cluster<NTOP, false>;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL);
    cfg.blockDim = dim3(CL_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    if (ok == 0) {
        bool good = cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
                    cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
                    cudaFuncSetAttribute(kt, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                    cudaFuncSetAttribute(kf, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
        int n_clusters = 0;
        if (good) {
            cudaLaunchConfig_t q = cfg;
            q.numAttrs = 1;
            good = cudaOccupancyMaxActiveClusters(&n_clusters, kt, &q) == cudaSuccess && n_clusters >= 1;
        }
        cudaGetLastError();
        ok = good ? 1 : -1;
    }
    if (ok < 0) return false;
    if (probe_only) return true;
    cudaError_t e;
    if (omega != 1.0)
        e = cudaLaunchKernelEx(&cfg, kt, x, f, pitch_x, pitch_f, n_coarse, h0, omega, nu1, nu2, coarse_sweeps, lo,
                               x_is_zero ? 1 : 0, gamma, done);
    else
        e = cudaLaunchKernelEx(&cfg, kf, x, f, pitch_x, pitch_f, n_coarse, h0, omega, nu1, nu2, coarse_sweeps, lo,
                               x_is_zero ? 1 : 0, gamma, done);
    return e == cudaSuccess;
#endif
}

}  // namespace

bool coarse_cluster_top(int n) { return n == 257 || n == 129; }

void coarse_profile_read(long long out[16])
{
#ifndef PMG_HOST_EMULATION
    cudaMemcpyFromSymbol(out, g_co_prof, 16 * sizeof(long long));
#else
    for (int i = 0; i < 16; ++i) out[i] = g_co_prof[i];
#endif
}

// sets the kernel attributes of the current device and asks the occupancy calculator whether one 16-CTA cluster with
// this much shared memory can be resident; call outside stream capture (pmg_create does)
bool coarse_cluster_available(int n0)
{
    if (n0 == 257) return coarse_cluster_launch<257>(nullptr, nullptr, 0, 0, 5, 1.0, 1.0, 1, 1, 1, 2, true, 1, nullptr, nullptr, true);
    if (n0 == 129) return coarse_cluster_launch<129>(nullptr, nullptr, 0, 0, 5, 1.0, 1.0, 1, 1, 1, 2, true, 1, nullptr, nullptr, true);
    return false;
}

bool launch_coarse_cluster(double *x, const double *f, int n0, int pitch_x, int pitch_f, int n_coarse, double h0, double omega,
                           int nu1, int nu2, int coarse_sweeps, int prolong_mode, bool x_is_zero, int gamma, cudaStream_t st,
                           const int *done)
{
    const int lo = prolong_mode == PMG_PROLONG_FULL ? 1 : 2;
    if (n_coarse > (DIST_MIN - 1) / 2 + 1 || gamma < 1) return false;  // the local part starts at level 33
    bool ok = false;
    if (n0 == 257)
        ok = coarse_cluster_launch<257>(x, f, pitch_x, pitch_f, n_coarse, h0, omega, nu1, nu2, coarse_sweeps, lo, x_is_zero, gamma, st, done);
    else if (n0 == 129)
        ok = coarse_cluster_launch<129>(x, f, pitch_x, pitch_f, n_coarse, h0, omega, nu1, nu2, coarse_sweeps, lo, x_is_zero, gamma, st, done);
    if (ok) count_launch();
    return ok;
}

bool coarse_local_supported(int n0, int gamma) { return n0 >= 3 && n0 <= VSMALL_TOP && gamma >= 1; }

void launch_coarse_local(double *x, const double *f, int n0, int pitch_x, int pitch_f, int n_coarse, double h0, double omega,
                         int nu1, int nu2, int coarse_sweeps, int prolong_mode, bool x_is_zero, int gamma, cudaStream_t st,
                         const int *done)
{
    const int lo = prolong_mode == PMG_PROLONG_FULL ? 1 : 2;
    switch (n0) {
        case 65: coarse_local_launch<65>(x, f, pitch_x, pitch_f, n_coarse, h0, omega, nu1, nu2, coarse_sweeps, lo, x_is_zero, gamma, st, done); break;
        case 33: coarse_local_launch<33>(x, f, pitch_x, pitch_f, n_coarse, h0, omega, nu1, nu2, coarse_sweeps, lo, x_is_zero, gamma, st, done); break;
        case 17: coarse_local_launch<17>(x, f, pitch_x, pitch_f, n_coarse, h0, omega, nu1, nu2, coarse_sweeps, lo, x_is_zero, gamma, st, done); break;
        case 9: coarse_local_launch<9>(x, f, pitch_x, pitch_f, n_coarse, h0, omega, nu1, nu2, coarse_sweeps, lo, x_is_zero, gamma, st, done); break;
        case 5: coarse_local_launch<5>(x, f, pitch_x, pitch_f, n_coarse, h0, omega, nu1, nu2, coarse_sweeps, lo, x_is_zero, gamma, st, done); break;
        case 3: coarse_local_launch<3>(x, f, pitch_x, pitch_f, n_coarse, h0, omega, nu1, nu2, coarse_sweeps, lo, x_is_zero, gamma, st, done); break;
        default: return;
    }
    count_launch();
}

}  // namespace pmg
