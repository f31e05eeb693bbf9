// tools/operand_microbench.cu -- does the FP64 pipe of B200 sustain one warp instruction per two
// cycles per sub-partition when every operand is a DIFFERENT register pair?
//
// The peak probe of the library (and tools/pipe_microbench.cu) runs x = fma(x, a, b) with a and b
// shared by all chains: one register operand per instruction plus two that sit in the operand
// reuse cache or a constant bank.  The step kernel's instructions read two or three distinct
// 64-bit register pairs each (DFMA ly, pw, ly ...).  This measures DFMA / DADD / DMUL with 1, 2
// and 3 distinct register operands, at 8 and at 5 warps per sub-partition (the step kernel runs
// at 5), and a "leak pass" shaped mix (6 x DFMA a,-b,a over distinct pairs + a 5-add chain).
// Development aid for DESIGN.md's roofline discussion; not part of the product library.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o operand_microbench tools/operand_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CH 8

template <int MODE>
__global__ void __launch_bounds__(256) k(long long iters, double *out, double a, double b)
{
    double x[CH], y[CH], z[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        x[j] = 1.0 + j + threadIdx.x * 1e-3;
        y[j] = 0.999 - j * 1e-4 + threadIdx.x * 1e-9;
        z[j] = 1e-7 * (j + 1) + threadIdx.x * 1e-12;
    }
    for (long long i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            if (MODE == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[j]) : "d"(a), "d"(b));          // 1 register operand
            if (MODE == 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[j]) : "d"(y[j]), "d"(b));       // 2
            if (MODE == 2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[j]) : "d"(y[j]), "d"(z[j]));    // 3 distinct
            if (MODE == 3) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(x[j]) : "d"(y[j]), "d"(z[j]));    // 3 distinct, accumulate form
            if (MODE == 4) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[j]) : "d"(b));                      // DADD 1
            if (MODE == 5) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[j]) : "d"(y[j]));                   // DADD 2 distinct
            if (MODE == 6) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x[j]) : "d"(y[j]));                   // DMUL 2 distinct
            if (MODE == 7) {   // leak-pass shape: x = fma(-x, y, x) (two distinct pairs, x read twice)
                double nx;
                asm volatile("neg.f64 %0, %1;" : "=d"(nx) : "d"(x[j]));
                asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(x[j]) : "d"(nx), "d"(y[j]));
            }
        }
        if (MODE == 8) {   // 6 leak FMAs + a dependent 5-addition total, like one pass of the wet hour
            double t;
#pragma unroll
            for (int j = 0; j < 6; ++j) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(x[j]) : "d"(y[j]), "d"(z[j]));
            asm volatile("add.rn.f64 %0, %1, %2;" : "=d"(t) : "d"(x[0]), "d"(x[1]));
#pragma unroll
            for (int j = 2; j < 6; ++j) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(t) : "d"(x[j]));
            asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[6]) : "d"(t));
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < CH; ++j) s += x[j] + y[j] + z[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int sms, double *out, int ctas_per_sm, int threads, int ops_per_iter)
{
    const long long iters = 1 << 14;
    const int blocks = sms * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, threads>>>(iters, out, 0.9999999, 1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double n = (double)blocks * threads * ops_per_iter * iters;
    const double per_smsp_cycle = n / 32.0 / (best * 1e-3) / (sms * 4.0) / 1.965e9;
    printf("%-58s %2d warps/SMSP %8.3f ms %7.3f T op/s  %.3f warp-instr/cycle/SMSP\n", name, ctas_per_sm * threads / 128, best,
           n / (best * 1e-3) / 1e12, per_smsp_cycle);
}

#define BOTH(MODE, NAME, OPS)                 \
    run<MODE>(NAME, sms, out, 8, 128, OPS);   \
    run<MODE>(NAME, sms, out, 5, 128, OPS);   \
    run<MODE>(NAME, sms, out, 4, 256, OPS);

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double *out;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
    printf("%s, %d SMs\n", p.name, sms);
    BOTH(0, "DFMA x = x*a + b            (1 register operand)", CH)
    BOTH(1, "DFMA x = x*y[j] + b         (2 distinct pairs)", CH)
    BOTH(2, "DFMA x = x*y[j] + z[j]      (3 distinct pairs)", CH)
    BOTH(3, "DFMA x = y[j]*z[j] + x      (3 distinct pairs)", CH)
    BOTH(4, "DADD x = x + b              (1 register operand)", CH)
    BOTH(5, "DADD x = x + y[j]           (2 distinct pairs)", CH)
    BOTH(6, "DMUL x = x * y[j]           (2 distinct pairs)", CH)
    BOTH(7, "DFMA x = (-x)*y[j] + x      (leak update)", CH)
    BOTH(8, "6 leak DFMA + 5-add total + 1 add (12 FP64 per iter)", 12)
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
