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

// counting barrier; every waiter of generation g leaves once `count` threads have arrived
struct EmuBarrier {
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0;
    unsigned long generation = 0;
    void wait(int count, const char *what, int id)
    {
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

inline EmuBarrier g_emu_named[16];  // id 0 = __syncthreads
inline EmuBarrier g_emu_warp_bar[32];

inline void __syncthreads() { g_emu_named[0].wait((int)blockDim.x, "__syncthreads", 0); }
inline void __syncwarp() { g_emu_warp_bar[threadIdx.x >> 5].wait(32, "__syncwarp of warp", (int)(threadIdx.x >> 5)); }
inline void emu_bar_sync(int id, int count)
{
    if (id < 1 || id > 15 || count % 32 != 0 || count > (int)blockDim.x) {
        std::fprintf(stderr, "emulation: bad bar.sync %d, %d\n", id, count);
        std::abort();
    }
    g_emu_named[id].wait(count, "bar.sync", id);
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

inline thread_local EmuDim3 blockIdx;
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
