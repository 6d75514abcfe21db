// test_small_kernel_emu.cpp -- runs the SOURCE of the single-CTA small-level kernel (csrc/kernels_small.cu,
// k_vcycle_small2) on CPU threads, one OS thread per CUDA thread (tests/cpp/emu/host_emulation.h), and compares the
// result bit for bit with the CPU oracle's V / W cycle (oracle/pmg_oracle.c).  Checks, without a GPU, the
// per-level thread groups, the named-barrier protocol (a missing or surplus arrival aborts as DEADLOCK /
// over-subscribed), the parity / visit-counter state machine and the arithmetic order.
//   g++ -std=c++17 -O1 -ffp-contract=off -pthread -DPMG_HOST_EMULATION -Itests/cpp/emu -I<pkg>/csrc -Iinclude
//       tests/cpp/test_small_kernel_emu.cpp -Loracle -loracle
#include <cmath>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#include "../../parallel-geometric-multigrid-for-poisson-problem_b200/csrc/kernels_small.cu"
#include "../../oracle/oracle.h"


static int run_case(int n0, int gamma, double omega, int nu1, int nu2, int lo, bool x_is_zero, unsigned seed)
{
    const int n_coarse = 5, coarse_sweeps = 11;
    const double h0 = 4.0 / 1024.0;  // any level spacing; exact in binary like the solver's 2^-k
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> u(-1.0, 1.0);
    std::vector<double> x((size_t)n0 * n0, 0.0), f((size_t)n0 * n0, 0.0);
    for (int y = 1; y < n0 - 1; ++y)
        for (int i = 1; i < n0 - 1; ++i) {
            f[(size_t)y * n0 + i] = u(rng);
            if (!x_is_zero) x[(size_t)y * n0 + i] = u(rng);
        }
    std::vector<double> want = x;
    if (x_is_zero) std::fill(want.begin(), want.end(), 0.0);
    orc_cycle(want.data(), f.data(), n0, h0, gamma > 1 ? ORC_CYCLE_W : ORC_CYCLE_V, omega, 0.0, gamma, nu1 - 1, nu2 - 1,
              lo == 1 ? ORC_PROLONG_FULL : ORC_PROLONG_REFERENCE);
    if (x_is_zero)  // the kernel must not read x: poison it
        for (double &v : x) v = std::nan("");
    int shift = 0, rows = 1;
    pmg::vs_grid(n0, 1024, shift, rows);
    int threads = (((1 << shift) * rows + 31) / 32) * 32;
    if (threads > 1024) threads = 1024;
    blockDim.x = (unsigned)threads;
    std::vector<std::thread> pool;
    pool.reserve(threads);
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t] {
            threadIdx.x = (unsigned)t;
            if (omega != 1.0)
                pmg::k_vcycle_small2<true>(x.data(), f.data(), n0, n0, n0, n_coarse, h0, omega, nu1, nu2, coarse_sweeps, lo,
                                           x_is_zero ? 1 : 0, gamma, nullptr);
            else
                pmg::k_vcycle_small2<false>(x.data(), f.data(), n0, n0, n0, n_coarse, h0, omega, nu1, nu2, coarse_sweeps,
                                            lo, x_is_zero ? 1 : 0, gamma, nullptr);
        });
    for (auto &th : pool) th.join();
    const bool same = std::memcmp(x.data(), want.data(), x.size() * sizeof(double)) == 0;
    std::printf("n0=%d gamma=%d omega=%.3f nu=(%d,%d) lo=%d x_is_zero=%d threads=%d: %s\n", n0, gamma, omega, nu1, nu2, lo,
                (int)x_is_zero, threads, same ? "bit-identical" : "MISMATCH");
    std::fflush(stdout);
    return same ? 0 : 1;
}

int main(int argc, char **argv)
{
    const bool full = argc > 1 && std::strcmp(argv[1], "full") == 0;
    int bad = 0;
    const double w23 = 2.0 / 3.0;
    // every top size the solver can hand to the kernel, V and W, both prolongations, weighted and plain Jacobi
    bad += run_case(65, 1, w23, 2, 2, 2, true, 1);
    bad += run_case(65, 2, w23, 2, 2, 2, false, 2);
    bad += run_case(33, 1, 1.0, 2, 2, 1, true, 3);
    bad += run_case(33, 3, w23, 1, 3, 2, true, 4);
    bad += run_case(17, 2, 0.8, 2, 1, 1, false, 5);
    bad += run_case(9, 2, w23, 2, 2, 2, true, 6);
    bad += run_case(5, 1, w23, 2, 2, 2, false, 7);
    if (full) {
        bad += run_case(65, 3, 1.0, 3, 4, 1, true, 8);
        bad += run_case(17, 1, w23, 4, 4, 2, true, 9);
        bad += run_case(9, 3, 1.0, 1, 1, 1, false, 10);
    }
    std::printf(bad ? "FAILED (%d)\n" : "all bit-identical\n", bad);
    return bad ? 1 : 0;
}
