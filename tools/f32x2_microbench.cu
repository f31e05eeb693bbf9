// tools/f32x2_microbench.cu -- is the packed binary32 path (fma/add/mul.rn.f32x2, sm_100+) worth
// a two-members-per-thread FP32 kernel?  Measures scalar FFMA against FFMA2 alone and mixed with
// ALU-pipe work, per SM sub-partition.  Development aid, not part of the product library.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_microbench tools/f32x2_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8

template <int MODE>
__global__ void __launch_bounds__(256) mix_kernel(long long iters, float *out, float a, float b, int sel)
{
    float f[CHAINS];
    unsigned long long v[CHAINS];
    unsigned u[CHAINS];
    float g[CHAINS];
    unsigned long long w[CHAINS];
    unsigned long long a2, b2;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(b2) : "f"(b));
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) {
        f[j] = 0.5f + j + threadIdx.x * 1e-3f;
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(v[j]) : "f"(f[j]));
        u[j] = threadIdx.x * 7 + j;
        g[j] = 1.0f - 1e-7f * (j + 1) + sel * 1e-9f;
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(w[j]) : "f"(g[j]));
    }
    for (long long i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) {
            if (MODE == 0 || MODE == 10) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[j]) : "f"(a), "f"(b));
            if (MODE == 1 || MODE == 11) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[j]) : "l"(a2), "l"(b2));
            if (MODE == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[j]) : "l"(b2));
            if (MODE == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v[j]) : "l"(a2));
            if (MODE == 10 || MODE == 11) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[j]) : "r"(sel), "r"(sel + 1));
            // three DISTINCT register operands per instruction (no operand-reuse cache hits), as real code has:
            // chain j multiplies by chain j+1's value and adds chain j+2's
            if (MODE == 20) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[j]) : "f"(g[(j + 1) % CHAINS]), "f"(g[(j + 3) % CHAINS]));
            if (MODE == 21) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(v[j]) : "l"(w[(j + 1) % CHAINS]), "l"(w[(j + 3) % CHAINS]));
            if (MODE == 22) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(g[(j + 1) % CHAINS]));
            if (MODE == 23) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(g[(j + 1) % CHAINS]));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) {
        float lo, hi;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[j]));
        s += f[j] + lo + hi + u[j] + g[j];
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(w[j]));
        s += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int sms, float *out, int flops_per_op)
{
    const long long iters = 1 << 14;
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        mix_kernel<MODE><<<blocks, threads>>>(iters, out, 0.9999999f, 1e-7f, rep);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double n = (double)blocks * threads * CHAINS * iters;
    printf("%-36s %8.3f ms  %7.3f T instr/s  %7.3f T lane-op/s\n", name, best, n / (best * 1e-3) / 1e12,
           n * flops_per_op / (best * 1e-3) / 1e12);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    float *out;
    cudaMalloc(&out, sizeof(float) * p.multiProcessorCount * 8 * 256);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>("FFMA", p.multiProcessorCount, out, 1);
    run<1>("FFMA2 (fma.rn.f32x2)", p.multiProcessorCount, out, 2);
    run<2>("FADD2 (add.rn.f32x2)", p.multiProcessorCount, out, 2);
    run<3>("FMUL2 (mul.rn.f32x2)", p.multiProcessorCount, out, 2);
    run<10>("FFMA + 1 LOP3", p.multiProcessorCount, out, 1);
    run<11>("FFMA2 + 1 LOP3", p.multiProcessorCount, out, 2);
    run<20>("FFMA, 3 distinct register operands", p.multiProcessorCount, out, 1);
    run<21>("FFMA2, 3 distinct register operands", p.multiProcessorCount, out, 2);
    run<22>("FADD, 2 distinct register operands", p.multiProcessorCount, out, 1);
    run<23>("FMUL, 2 distinct register operands", p.multiProcessorCount, out, 1);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
