// latency.cu -- dependent-issue latencies that bound the small multigrid levels on B200 (cycles, %clock64).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency latency.cu ; run: ./latency
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double *out, long long *cyc, double a, double b, int iters)
{
    __shared__ double sm[2048];
    const int t = threadIdx.x;
    sm[t] = a + t;
    sm[t + 1024] = b;
    __syncthreads();
    double x = a + t * 1e-9;
    long long t0 = clock64();
    if (MODE == 0) {  // dependent DADD
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 16; ++u) x = __dadd_rn(x, b);
        }
    } else if (MODE == 1) {  // dependent DMUL
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 16; ++u) x = __dmul_rn(x, b);
        }
    } else if (MODE == 2) {  // dependent DFMA
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 16; ++u) x = __fma_rn(x, b, a);
        }
    } else if (MODE == 3) {  // STS -> syncwarp -> LDS neighbour -> DADD (one warp smem round trip)
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                sm[t] = x;
                __syncwarp();
                x = __dadd_rn(sm[(t + 1) & 31], b);
                __syncwarp();
            }
        }
    } else if (MODE == 4) {  // shuffle of a double + DADD
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 16; ++u) x = __dadd_rn(__shfl_down_sync(0xffffffffu, x, 1), b);
        }
    } else if (MODE == 5) {  // STS -> bar.sync (all threads of the block) -> LDS -> DADD
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                sm[t] = x;
                __syncthreads();
                x = __dadd_rn(sm[(t + 1) & (blockDim.x - 1)], b);
                __syncthreads();
            }
        }
    } else if (MODE == 6) {  // a full weighted-Jacobi point: 5 LDS, 9 fp64 ops, STS, syncwarp
        double hf = b * 0.001;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const double *c = sm + ((u & 1) ? 1024 : 0);
                double *o = sm + ((u & 1) ? 0 : 1024);
                int p = 40 + (t & 31);
                double acc = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(hf, c[p - 1]), c[p + 1]), c[p - 32]), c[p + 32]);
                double jac = __dmul_rn(0.25, acc);
                o[p] = __dadd_rn(__dmul_rn(a, c[p]), __dmul_rn(b, jac));
                __syncwarp();
            }
        }
        x = sm[40 + t];
    }
    long long t1 = clock64();
    if (t == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + t] = x;
}

template <int MODE>
void run(const char *name, int threads)
{
    double *out;
    long long *cyc, h = 0;
    cudaMalloc(&out, 1024 * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    const int iters = 64;
    k<MODE><<<1, threads>>>(out, cyc, 1.0000001, 0.9999999, iters);
    k<MODE><<<1, threads>>>(out, cyc, 1.0000001, 0.9999999, iters);
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%-52s threads=%4d  %7.1f cycles per step  (%s)\n", name, threads, (double)h / (iters * 16), cudaGetErrorString(e));
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    run<0>("dependent DADD", 32);
    run<1>("dependent DMUL", 32);
    run<2>("dependent DFMA", 32);
    run<0>("dependent DADD, 8 warps", 256);
    run<0>("dependent DADD, 32 warps", 1024);
    run<3>("STS.64 + syncwarp + LDS.64 neighbour + DADD", 32);
    run<4>("SHFL(double) + DADD", 32);
    run<5>("STS + bar.sync + LDS + DADD + bar.sync", 64);
    run<5>("STS + bar.sync + LDS + DADD + bar.sync", 256);
    run<5>("STS + bar.sync + LDS + DADD + bar.sync", 1024);
    run<6>("weighted-Jacobi point via smem + syncwarp", 32);
    return 0;
}
