// cluster_sync.cu -- cost of a thread-block-cluster barrier and of a DSMEM halo push on B200 (cycles, clock64):
// what bounds one stage of a level distributed over the shared memories of a cluster (csrc/kernels_coarse.cu).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_sync cluster_sync.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

template <int MODE>
__global__ void k(double *out, long long *cyc, int iters)
{
    __shared__ double sm[2048];
    cg::cluster_group cl = cg::this_cluster();
    const int t = threadIdx.x, rank = (int)cl.block_rank(), n = (int)cl.num_blocks();
    sm[t] = t;
    sm[t + 1024] = 0;
    cl.sync();
    double x = t;
    double *up = cl.map_shared_rank(sm, (rank + n - 1) % n);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
            cl.sync();
        } else if (MODE == 1) {  // push one row into the neighbour, barrier, read it back locally
            if (t < 256) up[1024 + t] = x;
            cl.sync();
            x = __dadd_rn(sm[1024 + ((t + 1) & 255)], 1.0);
        } else if (MODE == 2) {  // split barrier: arrive, independent work, wait
            if (t < 256) up[1024 + t] = x;
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
#pragma unroll
            for (int u = 0; u < 8; ++u) x = __dadd_rn(x, 1.0);
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
            x = __dadd_rn(sm[1024 + ((t + 1) & 255)], x);
        }
    }
    long long t1 = clock64();
    if (t == 0 && rank == 0) cyc[0] = t1 - t0;
    out[rank * blockDim.x + t] = x;
}

template <int MODE>
void run(const char *name, int ctas, int threads)
{
    double *out;
    long long *cyc, h = 0;
    cudaMalloc(&out, 16 * 1024 * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    const int iters = 200;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(threads);
    cudaLaunchAttribute a[1];
    a[0].id = cudaLaunchAttributeClusterDimension;
    a[0].val.clusterDim.x = ctas;
    a[0].val.clusterDim.y = 1;
    a[0].val.clusterDim.z = 1;
    cfg.attrs = a;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k<MODE>, out, cyc, iters);
    cudaLaunchKernelEx(&cfg, k<MODE>, out, cyc, iters);
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%-58s ctas=%2d threads=%4d  %8.1f cycles per step  (%s)\n", name, ctas, threads, (double)h / iters, cudaGetErrorString(e));
    fflush(stdout);
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    for (int ctas : {2, 4, 8, 16})
        for (int threads : {32, 256, 1024}) run<0>("cluster.sync", ctas, threads);
    for (int ctas : {8, 16}) {
        run<1>("DSMEM row push + cluster.sync + local read", ctas, 1024);
        run<2>("DSMEM row push + arrive / 8 DADD / wait + local read", ctas, 1024);
    }
    return 0;
}
