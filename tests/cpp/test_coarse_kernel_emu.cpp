// test_coarse_kernel_emu.cpp -- runs the SOURCE of the third-generation small-level kernel (csrc/kernels_coarse.cu,
// k_coarse_local<N0>) on CPU threads, one OS thread per CUDA thread (tests/cpp/emu/host_emulation.h), and compares the
// result bit for bit with the CPU oracle's V / W cycle (oracle/pmg_oracle.c).  Checks, without a GPU, the template
// recursion over the level sizes, the per-level thread groups and their named barriers (a missing or surplus arrival
// aborts as DEADLOCK / over-subscribed), the buffer-parity bookkeeping of the warps that sit levels out, and the
// arithmetic order.
//   g++ -std=c++17 -O1 -ffp-contract=off -pthread -DPMG_HOST_EMULATION -Itests/cpp/emu -I<pkg>/csrc -Iinclude
//       tests/cpp/test_coarse_kernel_emu.cpp -Loracle -loracle
#include <cmath>
#include <cstring>
#include <random>
#include <vector>

#include "../../parallel-geometric-multigrid-for-poisson-problem_b200/csrc/kernels_coarse.cu"
#include "../../oracle/oracle.h"

namespace pmg {
void count_launch(int) {}
}  // namespace pmg

static int run_case(int n0, int gamma, double omega, int nu1, int nu2, int lo, bool x_is_zero, unsigned seed, int n_coarse = 5,
                    int coarse_sweeps = 11, bool cluster = false)
{
    const double h0 = 4.0 / 1024.0;  // any level spacing; exact in binary like the solver's 2^-k
    const int pitch = n0 + 7;        // the kernel must honour the pitch
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> u(-1.0, 1.0);
    std::vector<double> x((size_t)n0 * n0, 0.0), f((size_t)n0 * n0, 0.0);
    for (int y = 0; y < n0; ++y)
        for (int i = 0; i < n0; ++i) {
            const bool ring = y == 0 || i == 0 || y == n0 - 1 || i == n0 - 1;
            if (!ring) f[(size_t)y * n0 + i] = u(rng);
            if (!x_is_zero) x[(size_t)y * n0 + i] = u(rng);  // non-zero Dirichlet ring on the top level
        }
    std::vector<double> want = x;
    if (x_is_zero) std::fill(want.begin(), want.end(), 0.0);
    if (n_coarse != 5 || coarse_sweeps != 11) {
        std::printf("oracle is fixed to n_coarse = 5 / 11 sweeps\n");
        return 1;
    }
    orc_cycle(want.data(), f.data(), n0, h0, gamma > 1 ? ORC_CYCLE_W : ORC_CYCLE_V, omega, 0.0, gamma, nu1 - 1, nu2 - 1,
              lo == 1 ? ORC_PROLONG_FULL : ORC_PROLONG_REFERENCE);
    std::vector<double> xp((size_t)n0 * pitch, std::nan("")), fp((size_t)n0 * pitch, std::nan(""));
    for (int y = 0; y < n0; ++y)
        for (int i = 0; i < n0; ++i) {
            fp[(size_t)y * pitch + i] = f[(size_t)y * n0 + i];
            if (!x_is_zero) xp[(size_t)y * pitch + i] = x[(size_t)y * n0 + i];  // x_is_zero: poisoned, must not be read
        }
    if (cluster) {
        if (!pmg::launch_coarse_cluster(xp.data(), fp.data(), n0, pitch, pitch, n_coarse, h0, omega, nu1, nu2, coarse_sweeps,
                                        lo == 1 ? PMG_PROLONG_FULL : PMG_PROLONG_REFERENCE, x_is_zero, gamma, nullptr, nullptr)) {
            std::printf("cluster kernel refused n0=%d\n", n0);
            return 1;
        }
    } else {
        pmg::launch_coarse_local(xp.data(), fp.data(), n0, pitch, pitch, n_coarse, h0, omega, nu1, nu2, coarse_sweeps,
                                 lo == 1 ? PMG_PROLONG_FULL : PMG_PROLONG_REFERENCE, x_is_zero, gamma, nullptr, nullptr);
    }
    bool same = true;
    for (int y = 0; y < n0 && same; ++y)
        same = std::memcmp(&xp[(size_t)y * pitch], &want[(size_t)y * n0], n0 * sizeof(double)) == 0;
    if (!same) {
        int shown = 0;
        for (int y = 0; y < n0 && shown < 6; ++y)
            for (int i = 0; i < n0 && shown < 6; ++i)
                if (std::memcmp(&xp[(size_t)y * pitch + i], &want[(size_t)y * n0 + i], 8) != 0) {
                    std::printf("  first mismatches: (row %d, col %d) got %.17g want %.17g\n", y, i, xp[(size_t)y * pitch + i],
                                want[(size_t)y * n0 + i]);
                    ++shown;
                }
    }
    std::printf("%s n0=%d gamma=%d omega=%.3f nu=(%d,%d) lo=%d x_is_zero=%d: %s\n", cluster ? "cluster" : "local  ", n0, gamma,
                omega, nu1, nu2, lo, (int)x_is_zero, same ? "bit-identical" : "MISMATCH");
    std::fflush(stdout);
    return same ? 0 : 1;
}

int main(int argc, char **argv)
{
    const bool full = argc > 1 && std::strcmp(argv[1], "full") == 0;
    int bad = 0;
    const double w23 = 2.0 / 3.0;
    // every top size the solver can hand to the kernel, V and W, both prolongations, weighted and plain Jacobi
    bad += run_case(5, 1, w23, 2, 2, 2, false, 7);
    bad += run_case(9, 2, w23, 2, 2, 2, true, 6);
    bad += run_case(17, 2, 0.8, 2, 1, 1, false, 5);
    bad += run_case(33, 1, 1.0, 2, 2, 1, true, 3);
    bad += run_case(33, 3, w23, 1, 3, 2, true, 4);
    bad += run_case(65, 1, w23, 2, 2, 2, true, 1);
    bad += run_case(65, 2, w23, 2, 2, 2, false, 2);
    // the 16-CTA cluster kernel (levels >= 33 distributed over the CTAs' shared memories): both top sizes, V and W
    bad += run_case(129, 1, w23, 2, 2, 2, true, 21, 5, 11, true);
    bad += run_case(129, 2, w23, 2, 2, 2, false, 22, 5, 11, true);
    bad += run_case(257, 1, w23, 2, 2, 2, true, 23, 5, 11, true);
    if (full) {
        bad += run_case(257, 2, 1.0, 2, 2, 1, false, 24, 5, 11, true);
        bad += run_case(129, 3, 0.8, 1, 2, 1, true, 25, 5, 11, true);
        bad += run_case(65, 3, 1.0, 3, 4, 1, true, 8);
        bad += run_case(17, 1, w23, 4, 4, 2, true, 9);
        bad += run_case(9, 3, 1.0, 1, 1, 1, false, 10);
        bad += run_case(33, 2, w23, 3, 2, 2, false, 11);
    }
    std::printf(bad ? "FAILED (%d)\n" : "all bit-identical\n", bad);
    return bad ? 1 : 0;
}
