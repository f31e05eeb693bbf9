// Latin Hypercube sampler on the device -- the construction of
// smartpy/montecarlo/lhs.py:133-167 ((stratum + U) / N through the inverse uniform CDF, one
// stratum per row and column) with the host permutation replaced by a keyed bijection, so that
// ANY row range of the sample can be produced independently: no permutation array, no sort, no
// host round trip, and the shards of a multi-GPU run are self-generating while the sample stays
// stratified over the WHOLE run, not per GPU.
//
// Per column p and row i (all 64-bit integer arithmetic wraps):
//   mix(z)      : z ^= z >> 30; z *= 0xBF58476D1CE4E5B9; z ^= z >> 27; z *= 0x94D049BB133111EB; z ^= z >> 31
//   K_p         = mix(seed + 0x9E3779B97F4A7C15 * (p + 1))
//   stratum     = cycle-walked 6-round balanced Feistel network on 2h bits (4^h >= N, h >= 1):
//                 (L, R) <- (R, L ^ (mix(K_p ^ (R + 0xD6E8FEB86659FD93 * (round + 1))) & (2^h - 1))),
//                 repeated until the value is < N  (a permutation of 0..N-1)
//   jitter      = (mix(K_p ^ mix(i + 0x632BE59BD9B4E019)) >> 11) * 2^-53            in [0, 1)
//   value       = ((stratum + jitter) / N) * (hi_p - lo_p) + lo_p                    (IEEE, no FMA)
// tests/test_gpu_sampler.py restates this in numpy and checks the kernel bit for bit.
#include <cuda_runtime.h>
#include <cstdint>
#include <string>

#include "smart_b200.h"

int smart_internal_fail(int code, const char *msg);     // smart_kernels.cu
void smart_internal_count(int n);                       // smart_kernels.cu (smart_launch_count)

namespace {

constexpr int kMaxParams = SMART_LHS_MAX_PARAMS;

struct Bounds {
    double lo[kMaxParams];
    double width[kMaxParams];
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

__global__ void __launch_bounds__(256)
lhs_rows_kernel(unsigned long long seed, long long n_total, long long row_first, long long n_rows, int n_params,
                int half_bits, Bounds b, double *__restrict__ out)
{
    const long long cells = n_rows * n_params;
    const unsigned long long half_mask = (1ull << half_bits) - 1ull;
    const double n_as_double = (double)n_total;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (long long)gridDim.x * blockDim.x) {
        const long long r = c / n_params;
        const int p = (int)(c - r * n_params);
        const unsigned long long i = (unsigned long long)(row_first + r);
        const unsigned long long key = mix64(seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(p + 1));
        unsigned long long x = i;
        do {
            unsigned long long left = x >> half_bits, right = x & half_mask;
#pragma unroll
            for (int round = 0; round < 6; ++round) {
                const unsigned long long f = mix64(key ^ (right + 0xD6E8FEB86659FD93ull * (unsigned long long)(round + 1))) & half_mask;
                const unsigned long long t = left ^ f;
                left = right;
                right = t;
            }
            x = (left << half_bits) | right;
        } while (x >= (unsigned long long)n_total);
        const double jitter = (double)(mix64(key ^ mix64(i + 0x632BE59BD9B4E019ull)) >> 11) * 0x1.0p-53;
        const double q = __ddiv_rn(__dadd_rn((double)x, jitter), n_as_double);
        out[c] = __dadd_rn(__dmul_rn(q, b.width[p]), b.lo[p]);          // coalesced: cells are row-major
    }
}

}  // namespace

extern "C" int smart_lhs_rows(uint64_t seed, int64_t n_total, int64_t row_first, int64_t n_rows, int32_t n_params,
                              const double *bounds, double *out, void *stream)
{
    if (n_total < 1 || row_first < 0 || n_rows < 0 || row_first + n_rows > n_total || n_params < 1 ||
        n_params > kMaxParams || !bounds || (n_rows > 0 && !out))
        return smart_internal_fail(SMART_ERR_BAD_ARG, "smart_lhs_rows: bad argument (rows must lie in 0..n_total, "
                                                      "1..SMART_LHS_MAX_PARAMS columns)");
    if (n_rows == 0) return SMART_OK;
    Bounds b;
    for (int p = 0; p < kMaxParams; ++p) {
        b.lo[p] = p < n_params ? bounds[2 * p] : 0.0;
        b.width[p] = p < n_params ? bounds[2 * p + 1] - bounds[2 * p] : 0.0;
    }
    int half_bits = 1;
    while (half_bits < 31 && (1ll << (2 * half_bits)) < n_total) ++half_bits;
    if ((1ll << (2 * half_bits)) < n_total)
        return smart_internal_fail(SMART_ERR_BAD_ARG, "smart_lhs_rows: sample larger than 2^62 rows");
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long blocks = (n_rows * n_params + 255) / 256;
    if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
    lhs_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(seed, n_total, row_first, n_rows, n_params,
                                                                        half_bits, b, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return smart_internal_fail(SMART_ERR_CUDA, (std::string("smart_lhs_rows: ") + cudaGetErrorString(e)).c_str());
    smart_internal_count(1);
    return SMART_OK;
}
