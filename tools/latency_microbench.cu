// tools/latency_microbench.cu -- dependent-issue latencies on B200 (one warp per SM sub-partition,
// one dependent chain per thread): DADD, DMUL, DFMA, the sign-test + 64-bit select used by the
// soil ladders, FP64 compare + select, shared-memory load.  Development aid.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency_microbench tools/latency_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void chain_kernel(int iters, double *out, long long *cycles, double a, double b)
{
    __shared__ double sh[256];
    sh[threadIdx.x] = a;
    sh[threadIdx.x + 128] = b;
    __syncthreads();
    double x = 1.0 + threadIdx.x * 1e-3, y = 0.5;
    int idx = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (MODE == 0) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(b));
            if (MODE == 1) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(a));
            if (MODE == 2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(a), "d"(b));
            if (MODE == 3) {   // t = x - b ; x = (sign clear) ? t : y      (ladder link: DADD + ISETP + 2 SEL)
                double t;
                asm volatile("sub.rn.f64 %0, %1, %2;" : "=d"(t) : "d"(x), "d"(b));
                x = (__double2hiint(t) >= 0) ? t : y;
            }
            if (MODE == 4) {   // same with an FP64 compare
                double t;
                asm volatile("sub.rn.f64 %0, %1, %2;" : "=d"(t) : "d"(x), "d"(b));
                x = (t >= 0.0) ? t : y;
            }
            if (MODE == 5) {   // dependent shared-memory load
                idx = (int)sh[idx & 255] & 127;
            }
            if (MODE == 6) {   // DADD -> DADD -> sign select (V2 fill link)
                double w, t;
                asm volatile("sub.rn.f64 %0, %1, %2;" : "=d"(w) : "d"(y), "d"(x));
                asm volatile("sub.rn.f64 %0, %1, %2;" : "=d"(t) : "d"(a), "d"(w));
                x = (__double2hiint(t) >= 0) ? 0.0 : t;
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + idx;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
void run(const char *name, double *out, long long *cyc, int links)
{
    const int iters = 4096;
    chain_kernel<MODE><<<1, 128>>>(iters, out, cyc, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    chain_kernel<MODE><<<1, 128>>>(iters, out, cyc, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost);
    printf("%-52s %7.2f cycles per link\n", name, (double)c / (iters * 16.0) / links);
}

int main()
{
    double *out;
    long long *cyc;
    cudaMalloc(&out, sizeof(double) * 128);
    cudaMalloc(&cyc, sizeof(long long));
    run<0>("DADD dependent", out, cyc, 1);
    run<1>("DMUL dependent", out, cyc, 1);
    run<2>("DFMA dependent", out, cyc, 1);
    run<3>("DADD + sign test (ISETP) + 64-bit select", out, cyc, 1);
    run<4>("DADD + DSETP + 64-bit select", out, cyc, 1);
    run<5>("LDS.64 + F2I + LOP dependent", out, cyc, 1);
    run<6>("DADD + DADD + sign select (fill link)", out, cyc, 1);
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
