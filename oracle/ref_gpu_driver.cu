/*
 * ref_gpu_driver.cu -- times the reference's OWN CUDA multigrid (3_part_parallel/Parallel_Mg.cu, unmodified,
 * #included from /root/reference where it lies) on this GPU.  TEST / BASELINE INFRASTRUCTURE ONLY: built by
 * oracle/Makefile into oracle/_ref/ref_gpu_exec (git-ignored), run only by bench.py as the
 * "reference CUDA build" baseline.  Protocol = ParallelTestRunner::run_v_cycle
 * (3_part_parallel/ParallelTestRunner.cu:152-186): managed phi / f initialised on the host, fixed number of
 * ParallelMultiGridSolver::v_cycle calls timed with std::chrono around each call (+ device sync).
 * Notes: the reference leaks ~16*N*N bytes of managed memory per cycle (Parallel_Mg.cu:38-54), its Jacobi
 * kernel updates in place (racy), it has no omega and no convergence test -- this is a time-per-cycle baseline.
 * usage: ref_gpu_exec N cycles [prefetch=1]
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "Parallel_Mg.cu"

int main(int argc, char **argv)
{
    int N = argc > 1 ? atoi(argv[1]) : 4097;
    int cycles = argc > 2 ? atoi(argv[2]) : 3;
    int prefetch = argc > 3 ? atoi(argv[3]) : 1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        printf("{\"error\": \"no CUDA device\"}\n");
        return 2;
    }
    size_t L = (size_t)N * N;
    double h = a / (N - 1);
    double *phi, *f;
    cudaMallocManaged(&phi, L * sizeof(double));
    cudaMallocManaged(&f, L * sizeof(double));
    DynamicGridUtils::initialize_zeros(phi, (int)L);
    DynamicGridUtils::compute_rhs(f, N, N, h);
    if (prefetch) {
        cudaMemPrefetchAsync(phi, L * sizeof(double), 0);
        cudaMemPrefetchAsync(f, L * sizeof(double), 0);
    }
    cudaDeviceSynchronize();
    ParallelMultiGridSolver mg(3);
    printf("{\"n\": %d, \"prefetch\": %d, \"block\": %d, \"cycle_ms\": [", N, prefetch, num_thread);
    for (int it = 0; it < cycles; ++it) {
        auto t0 = std::chrono::high_resolution_clock::now();
        mg.v_cycle(phi, f, N, h);
        cudaDeviceSynchronize();
        auto t1 = std::chrono::high_resolution_clock::now();
        printf("%s%.4f", it ? ", " : "", std::chrono::duration<double, std::milli>(t1 - t0).count());
    }
    cudaError_t e = cudaGetLastError();
    printf("], \"cuda_error\": \"%s\"}\n", cudaGetErrorString(e));
    return 0;
}
