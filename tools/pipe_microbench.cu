// tools/pipe_microbench.cu -- what bounds a mixed FP64 instruction stream on B200?
//
// Measures, per SM sub-partition, the throughput of register-resident instruction mixes:
// pure DFMA / DADD / DMUL / DSETP, and DFMA interleaved with k independent ALU (LOP3/FSEL) or
// FMA-pipe (IMAD/FFMA) instructions per DFMA.  If a DFMA only holds the FP64 pipe for two
// cycles, time(DFMA + 1 other) == time(DFMA); if it also costs two issue slots, the times add.
// Development aid for DESIGN.md's roofline discussion; not part of the product library.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_microbench tools/pipe_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8

template <int MODE>
__global__ void __launch_bounds__(256) mix_kernel(long long iters, double *out, double a, double b, int sel)
{
    double x[CHAINS];
    unsigned u[CHAINS];
    float f[CHAINS];
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) {
        x[j] = 1.0 + j + threadIdx.x * 1e-3;
        u[j] = threadIdx.x * 7 + j;
        f[j] = 0.5f + j;
    }
    int pred_acc = 0;
    for (long long i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) {
            if (MODE == 0 || MODE >= 10) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[j]) : "d"(a), "d"(b));
            if (MODE == 1) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[j]) : "d"(b));
            if (MODE == 2) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x[j]) : "d"(a));
            if (MODE == 3) {
                int p;
                asm volatile("{ .reg .pred q; setp.lt.f64 q, %1, %2; selp.s32 %0, 1, 0, q; }" : "=r"(p) : "d"(x[j]), "d"(b));
                pred_acc += p;
            }
            // extra ALU-pipe work per DFMA
            if (MODE == 10 || MODE == 12) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[j]) : "r"(sel), "r"(sel + 1));
            if (MODE == 12) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[j]) : "r"(sel + 2), "r"(sel + 3));
            // extra FMA-pipe work per DFMA
            if (MODE == 11 || MODE == 13) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[j]) : "f"(0.999f), "f"(0.001f));
            if (MODE == 13) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[j]) : "r"(sel), "r"(sel + 1));
            // 64-bit select (2 x SEL) per DFMA, as the ladders do
            if (MODE == 14) {
                asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; selp.b32 %0, %0, %1, q; }" : "+r"(u[j]) : "r"(sel), "r"(sel));
                asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; selp.f32 %0, %0, %1, q; }" : "+f"(f[j]) : "f"(0.5f), "r"(sel));
            }
        }
    }
    double s = pred_acc;
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) s += x[j] + u[j] + f[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int sms, double *out)
{
    const long long iters = 1 << 14;
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        mix_kernel<MODE><<<blocks, threads>>>(iters, out, 0.9999999, 1e-7, rep);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double n = (double)blocks * threads * CHAINS * iters;
    printf("%-44s %8.3f ms  %7.3f T primary-op/s\n", name, best, n / (best * 1e-3) / 1e12);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double *out;
    cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 8 * 256);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>("DFMA", p.multiProcessorCount, out);
    run<1>("DADD", p.multiProcessorCount, out);
    run<2>("DMUL", p.multiProcessorCount, out);
    run<3>("DSETP (+selp+iadd)", p.multiProcessorCount, out);
    run<10>("DFMA + 1 LOP3", p.multiProcessorCount, out);
    run<12>("DFMA + 2 LOP3", p.multiProcessorCount, out);
    run<11>("DFMA + 1 FFMA", p.multiProcessorCount, out);
    run<13>("DFMA + 1 FFMA + 1 LOP3", p.multiProcessorCount, out);
    run<14>("DFMA + 64-bit select (2 setp + 2 selp)", p.multiProcessorCount, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
