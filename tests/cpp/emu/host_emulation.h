// host_emulation.h -- TEST INFRASTRUCTURE: lets the source of a single-CTA CUDA kernel (csrc/kernels_small.cu)
// be compiled by g++ and executed by one OS thread per CUDA thread, so that its barrier protocol (named barriers with
// sub-CTA thread groups, warps that run ahead through the cycle state machine) and its arithmetic can be checked
// against the CPU oracle in the `-m "not gpu"` suite.  Included by csrc/pmg_internal.h only under
// -DPMG_HOST_EMULATION; never part of libpmg.so.
//
// Emulated: threadIdx / blockDim (thread_local), __shared__ (plain statics shared by the threads),
// __syncthreads / __syncwarp / bar.sync id,count (counting barriers; a barrier whose arrival count never
// completes shows up as a time-out instead of a hang), __dadd_rn / __dsub_rn / __dmul_rn (plain IEEE operations;
// the harness is built with -ffp-contract=off so nothing is fused).
#pragma once
#include <chrono>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(x) alignas(x)
#define __shared__ static

struct EmuDim3 {
    unsigned x = 1, y = 1, z = 1;
};
inline thread_local EmuDim3 threadIdx;
inline EmuDim3 blockDim;
inline EmuDim3 gridDim;
typedef void *cudaStream_t;

inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }

// Model C (below) runs every CUDA thread of a whole thread-block cluster as a fiber inside ONE OS thread; a fiber that
// has to wait at a barrier yields to the scheduler through this hook
struct EmuBarrier;
inline bool g_emu_fiber_mode = false;
inline void emu_fiber_block(EmuBarrier *b, unsigned long gen, const char *what, int id, int count);

// counting barrier; every waiter of generation g leaves once `count` threads have arrived
struct EmuBarrier {
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0;
    unsigned long generation = 0;
    void wait(int count, const char *what, int id)
    {
        if (g_emu_fiber_mode) {  // single OS thread: no locking, block by yielding
            const unsigned long g = generation;
            if (++arrived == count) {
                arrived = 0;
                ++generation;
                return;
            }
            if (arrived > count) {
                std::fprintf(stderr, "emulation: %s %d over-subscribed (%d > %d)\n", what, id, arrived, count);
                std::abort();
            }
            emu_fiber_block(this, g, what, id, count);
            return;
        }
        std::unique_lock<std::mutex> lk(mu);
        const unsigned long g = generation;
        if (++arrived == count) {
            arrived = 0;
            ++generation;
            cv.notify_all();
            return;
        }
        if (arrived > count) {
            std::fprintf(stderr, "emulation: %s %d over-subscribed (%d > %d)\n", what, id, arrived, count);
            std::abort();
        }
        if (!cv.wait_for(lk, std::chrono::seconds(60), [&] { return generation != g; })) {
            std::fprintf(stderr, "emulation: DEADLOCK at %s %d (%d of %d threads arrived)\n", what, id, arrived, count);
            std::abort();
        }
    }
};

constexpr int EMU_MAX_CTAS = 16;  // CTAs of one cluster (Model C); the thread model uses CTA 0 only
inline thread_local EmuDim3 blockIdx;
inline EmuBarrier g_emu_named[EMU_MAX_CTAS][16];  // id 0 = __syncthreads
inline EmuBarrier g_emu_warp_bar[EMU_MAX_CTAS][32];
inline EmuBarrier g_emu_cluster_bar;
inline int g_emu_cluster_ctas = 1;
inline unsigned emu_cta() { return g_emu_fiber_mode ? blockIdx.x : 0u; }

inline void __syncthreads() { g_emu_named[emu_cta()][0].wait((int)blockDim.x, "__syncthreads", 0); }
inline void __syncwarp()
{
    g_emu_warp_bar[emu_cta()][threadIdx.x >> 5].wait(32, "__syncwarp of warp", (int)(threadIdx.x >> 5));
}
inline void emu_bar_sync(int id, int count)
{
    if (id < 1 || id > 15 || count % 32 != 0 || count > (int)blockDim.x) {
        std::fprintf(stderr, "emulation: bad bar.sync %d, %d\n", id, count);
        std::abort();
    }
    g_emu_named[emu_cta()][id].wait(count, "bar.sync", id);
}
inline void emu_cluster_sync()
{
    g_emu_cluster_bar.wait(g_emu_cluster_ctas * (int)blockDim.x, "cluster barrier", 0);
}

// dynamic shared memory of the running CTA as doubles (kernels_coarse.cu) and its DSMEM view of another CTA
constexpr size_t EMU_CTA_SMEM_DOUBLES = 228 * 1024 / 8;
inline double *g_emu_cta_smem[EMU_MAX_CTAS] = {nullptr};
inline double *emu_cta_smem_doubles()
{
    double *&p = g_emu_cta_smem[emu_cta()];
    if (!p) p = new double[EMU_CTA_SMEM_DOUBLES];
    return p;
}
inline double *emu_map_shared_rank(double *p, int rank)
{
    double *mine = emu_cta_smem_doubles();
    if (rank < 0 || rank >= g_emu_cluster_ctas || p < mine || p >= mine + EMU_CTA_SMEM_DOUBLES) {
        std::fprintf(stderr, "emulation: bad map_shared_rank (rank %d)\n", rank);
        std::abort();
    }
    double *&other = g_emu_cta_smem[rank];
    if (!other) other = new double[EMU_CTA_SMEM_DOUBLES];
    return other + (p - mine);
}

// one "kernel launch" of a single CTA in the thread model: one OS thread per CUDA thread
#include <functional>
#include <thread>
#include <vector>
inline void emu_launch_threads(int threads, const std::function<void()> &body)
{
    blockDim.x = (unsigned)threads;
    (void)emu_cta_smem_doubles();  // allocate CTA 0's shared memory before the threads race for it
    std::vector<std::thread> pool;
    pool.reserve((size_t)threads);
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&body, t] {
            threadIdx.x = (unsigned)t;
            body();
        });
    for (auto &th : pool) th.join();
}

// =====================================================================================================================
// Model F ("fibers"): warp-collective kernels without CTA barriers (csrc/kernels_fused.cu).  The 32 lanes of a warp are
// 32 ucontext fibers inside ONE OS thread, resumed round-robin; a lane yields at every warp collective.  In round r each
// lane first completes collective #r-1 (all 32 deposits are in place, double-buffered) and then runs to collective #r.
// Warps and CTAs are executed one after the other -- legal for kernels whose warps are independent of each other.
// A collective reached by only part of a warp aborts ("divergent collective").
// =====================================================================================================================
#include <ucontext.h>

#include <cstring>
#include <functional>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct double2 {
    double x, y;
};
inline double2 make_double2(double x, double y) { return double2{x, y}; }
template <typename T>
inline T __ldg(const T *p) { return *p; }
template <typename T>
inline T __ldcg(const T *p) { return *p; }
using std::max;
using std::min;

constexpr size_t EMU_SMEM_BYTES = 228 * 1024;
alignas(16) inline thread_local unsigned char emu_smem[EMU_SMEM_BYTES];  // the running CTA's shared memory
inline unsigned emu_smem_addr(const void *p) { return (unsigned)((const unsigned char *)p - emu_smem); }
#define __cvta_generic_to_shared(p) emu_smem_addr(p)

struct EmuWarp {
    static constexpr size_t STACK = 512 * 1024;
    ucontext_t sched, lane[32];
    std::vector<unsigned char> stacks;
    bool finished[32];
    int cur = 0;
    unsigned long long slot[2][32];
    unsigned long phase[32];
    std::function<void()> body;
    unsigned warp_in_cta = 0;
    EmuWarp() : stacks(32 * STACK) {}
};
inline thread_local EmuWarp *g_emu_warp = nullptr;

inline void emu_lane_entry()
{
    EmuWarp *w = g_emu_warp;
    w->body();
    w->finished[w->cur] = true;
    swapcontext(&w->lane[w->cur], &w->sched);
}

// run `body` once per lane of warp `warp_in_cta` of the current CTA
inline void emu_run_warp(EmuWarp &w, unsigned warp_in_cta, const std::function<void()> &body)
{
    w.body = body;
    w.warp_in_cta = warp_in_cta;
    g_emu_warp = &w;
    for (int l = 0; l < 32; ++l) {
        w.finished[l] = false;
        w.phase[l] = 0;
        getcontext(&w.lane[l]);
        w.lane[l].uc_stack.ss_sp = w.stacks.data() + (size_t)l * EmuWarp::STACK;
        w.lane[l].uc_stack.ss_size = EmuWarp::STACK;
        w.lane[l].uc_link = &w.sched;
        makecontext(&w.lane[l], (void (*)())emu_lane_entry, 0);
    }
    for (;;) {
        int alive = 0, done = 0;
        for (int l = 0; l < 32; ++l) {
            if (w.finished[l]) {
                ++done;
                continue;
            }
            w.cur = l;
            threadIdx.x = warp_in_cta * 32 + (unsigned)l;
            swapcontext(&w.sched, &w.lane[l]);
            if (!w.finished[l]) ++alive;
        }
        if (alive == 0) break;
        int fin = 0;
        for (int l = 0; l < 32; ++l) fin += w.finished[l] ? 1 : 0;
        if (fin != 0) {
            std::fprintf(stderr, "emulation: divergent collective (%d lanes returned, %d wait in a shuffle)\n", fin, alive);
            std::abort();
        }
    }
}

// deposit, yield, then read lane `src` (own value when src is outside the warp)
template <typename T>
inline T emu_collective(T v, int src)
{
    static_assert(sizeof(T) <= 8, "8-byte slots");
    EmuWarp *w = g_emu_warp;
    const int l = w->cur;
    const unsigned long k = w->phase[l]++;
    unsigned long long raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    w->slot[k & 1][l] = raw;
    swapcontext(&w->lane[l], &w->sched);
    if (src < 0 || src > 31) return v;
    raw = w->slot[k & 1][src];
    T out;
    std::memcpy(&out, &raw, sizeof(T));
    return out;
}
template <typename T>
inline T __shfl_sync(unsigned, T v, int src) { return emu_collective(v, src); }
template <typename T>
inline T __shfl_up_sync(unsigned, T v, int d) { return emu_collective(v, g_emu_warp->cur - d); }
template <typename T>
inline T __shfl_down_sync(unsigned, T v, int d) { return emu_collective(v, g_emu_warp->cur + d); }
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_collective(v, g_emu_warp->cur ^ m); }

// one "kernel launch": CTAs and their warps one after the other, each CTA starting from poisoned shared memory
inline void emu_launch_warps(dim3 grid, dim3 block, const std::function<void()> &body)
{
    static thread_local EmuWarp warp;
    blockDim.x = block.x;
    gridDim.x = grid.x;
    for (unsigned b = 0; b < grid.x; ++b) {
        blockIdx.x = b;
        std::memset(emu_smem, 0xff, EMU_SMEM_BYTES);  // NaN pattern: a slot read before it was written shows up
        for (unsigned wi = 0; wi < block.x / 32; ++wi) emu_run_warp(warp, wi, body);
    }
}

inline int emu_num_sms = 148;  // what the geometry code sees as the SM count (148 unless a test wants fewer, larger chunks)


// =====================================================================================================================
// Model C ("cluster"): a whole thread-block cluster -- n_cta CTAs x `threads` CUDA threads -- as fibers inside ONE OS
// thread (csrc/kernels_coarse.cu, k_coarse_cluster).  Every barrier (bar.sync id,count per CTA, __syncwarp per warp, the
// cluster barrier) is a counting barrier; a fiber that has to wait yields to the scheduler, which resumes it once the
// barrier's generation has moved.  A pass over all fibers in which nobody can run is reported as DEADLOCK with the
// barrier each CTA's first blocked thread sits at.  Execution is sequential and deterministic, so data races are NOT
// detected here (compute-sanitizer's racecheck on the GPU does that); what is checked is the barrier protocol -- every
// thread of every CTA arrives at every barrier it is counted in -- the DSMEM addressing and the arithmetic.
// =====================================================================================================================
#include <sys/mman.h>

struct EmuFiberC {
    ucontext_t ctx;
    unsigned cta = 0, tid = 0;
    bool finished = false;
    EmuBarrier *waiting_on = nullptr;
    unsigned long wait_gen = 0;
    const char *wait_what = "";
    int wait_id = 0, wait_count = 0;
};
struct EmuClusterRun {
    std::vector<EmuFiberC> fibers;
    ucontext_t sched;
    int cur = -1;
    std::function<void()> body;
};
inline EmuClusterRun *g_emu_cluster_run = nullptr;

inline void emu_fiber_block(EmuBarrier *b, unsigned long gen, const char *what, int id, int count)
{
    EmuClusterRun *r = g_emu_cluster_run;
    EmuFiberC &f = r->fibers[(size_t)r->cur];
    f.waiting_on = b;
    f.wait_gen = gen;
    f.wait_what = what;
    f.wait_id = id;
    f.wait_count = count;
    swapcontext(&f.ctx, &r->sched);
    f.waiting_on = nullptr;
}

inline void emu_fiber_entry_c()
{
    EmuClusterRun *r = g_emu_cluster_run;
    r->body();
    EmuFiberC &f = r->fibers[(size_t)r->cur];
    f.finished = true;
    swapcontext(&f.ctx, &r->sched);
}

inline void emu_launch_cluster(int n_cta, int threads, const std::function<void()> &body)
{
    if (n_cta > EMU_MAX_CTAS) std::abort();
    constexpr size_t STACK = 64 * 1024;
    const size_t n = (size_t)n_cta * (size_t)threads;
    unsigned char *stacks = (unsigned char *)mmap(nullptr, n * STACK, PROT_READ | PROT_WRITE,
                                                  MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (stacks == (unsigned char *)MAP_FAILED) {
        std::perror("emulation: mmap of fiber stacks");
        std::abort();
    }
    EmuClusterRun run;
    run.fibers.resize(n);
    run.body = body;
    g_emu_cluster_run = &run;
    g_emu_fiber_mode = true;
    g_emu_cluster_ctas = n_cta;
    blockDim.x = (unsigned)threads;
    gridDim.x = (unsigned)n_cta;
    for (int c = 0; c < n_cta; ++c) {  // poisoned shared memory: a slot read before it was written shows up as NaN
        blockIdx.x = (unsigned)c;
        std::memset(emu_cta_smem_doubles(), 0xff, EMU_CTA_SMEM_DOUBLES * sizeof(double));
    }
    for (size_t i = 0; i < n; ++i) {
        EmuFiberC &f = run.fibers[i];
        f.cta = (unsigned)(i / (size_t)threads);
        f.tid = (unsigned)(i % (size_t)threads);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = stacks + i * STACK;
        f.ctx.uc_stack.ss_size = STACK;
        f.ctx.uc_link = &run.sched;
        makecontext(&f.ctx, (void (*)())emu_fiber_entry_c, 0);
    }
    for (;;) {
        size_t ran = 0, alive = 0;
        for (size_t i = 0; i < n; ++i) {
            EmuFiberC &f = run.fibers[i];
            if (f.finished) continue;
            ++alive;
            if (f.waiting_on && f.waiting_on->generation == f.wait_gen) continue;  // still blocked
            run.cur = (int)i;
            blockIdx.x = f.cta;
            threadIdx.x = f.tid;
            swapcontext(&run.sched, &f.ctx);
            ++ran;
        }
        if (alive == 0) break;
        if (ran == 0) {
            std::fprintf(stderr, "emulation: DEADLOCK in the cluster, %zu threads blocked:\n", alive);
            for (int c = 0; c < n_cta; ++c)
                for (size_t i = (size_t)c * threads; i < (size_t)(c + 1) * threads; ++i)
                    if (!run.fibers[i].finished) {
                        const EmuFiberC &f = run.fibers[i];
                        std::fprintf(stderr, "  cta %u thread %u waits at %s %d (%d of %d arrived)\n", f.cta, f.tid, f.wait_what,
                                     f.wait_id, f.waiting_on ? f.waiting_on->arrived : -1, f.wait_count);
                        break;
                    }
            std::abort();
        }
    }
    g_emu_fiber_mode = false;
    g_emu_cluster_ctas = 1;
    g_emu_cluster_run = nullptr;
    munmap(stacks, n * STACK);
}
