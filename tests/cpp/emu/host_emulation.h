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
extern thread_local EmuDim3 threadIdx;
extern EmuDim3 blockDim;
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

extern EmuBarrier g_emu_named[16];  // id 0 = __syncthreads
extern EmuBarrier g_emu_warp[32];

inline void __syncthreads() { g_emu_named[0].wait((int)blockDim.x, "__syncthreads", 0); }
inline void __syncwarp() { g_emu_warp[threadIdx.x >> 5].wait(32, "__syncwarp of warp", (int)(threadIdx.x >> 5)); }
inline void emu_bar_sync(int id, int count)
{
    if (id < 1 || id > 15 || count % 32 != 0 || count > (int)blockDim.x) {
        std::fprintf(stderr, "emulation: bad bar.sync %d, %d\n", id, count);
        std::abort();
    }
    g_emu_named[id].wait(count, "bar.sync", id);
}
