// Conditioning of a scored sample on the device: the selections the
// reference makes with numpy masks and a full argsort after reading the sample database back
// (smartpy/montecarlo/glue.py:222-289, smartpy/montecarlo/best.py:221-287), done on the score
// table the batch kernel left in HBM.
//
//   smart_condition_rows : AND of the (column, kind, values) conditions -> ascending row indices
//                          (three passes: per-block counts, scan of the counts, ordered scatter)
//   smart_best_rows      : rows of the k best values of one column among the rows that pass the
//                          constraints, ascending with the best last.  Radix select of the k-th
//                          largest (key, row) pair -- 8 byte-wide passes over the order-preserving
//                          64-bit image of the score, then 4 over the row index to break ties the
//                          way a stable ascending sort followed by [-k:] does -- a gather of the k
//                          winners and a bitonic sort of those k only.  Nothing is sorted in full.
//
// Everything is stream-ordered, allocation-free (caller's workspace) and HBM-bound: 12 passes
// over 8 N bytes of keys.  sm_100a only.
#include <cuda_runtime.h>
#include <cstdint>
#include <string>

#include "smart_b200.h"
#include "smart_step.cuh"

int smart_internal_fail(int code, const char *msg);     // smart_kernels.cu (sets smart_last_error)
void smart_internal_count(int n);                       // smart_kernels.cu (smart_launch_count)

namespace {

#define SEL_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return smart_internal_fail(SMART_ERR_CUDA,                                              \
                                       (std::string(#expr) + ": " + cudaGetErrorString(_e)).c_str()); \
    } while (0)

constexpr int kMaxConds = SMART_MAX_CONDITIONS;
constexpr int kRowBlock = 1024;          // rows per CTA in the mask passes (one row per thread)
constexpr int kSortChunk = 2048;         // (key, row) pairs sorted in shared memory by one CTA
constexpr unsigned kPadRow = 0xFFFFFFFFu;

struct CondSet {
    int n;
    int col[kMaxConds];
    int kind[kMaxConds];
    double lo[kMaxConds];
    double hi[kMaxConds];
};

// glue.py:246-286 -- comparisons with NaN are false, as with numpy
__device__ __forceinline__ bool row_passes(const double *__restrict__ row, const CondSet &c)
{
    bool ok = true;
#pragma unroll 1
    for (int i = 0; i < c.n; ++i) {
        const double x = row[c.col[i]];
        bool sel;
        switch (c.kind[i]) {
        case SMART_COND_EQUAL: sel = (x == c.lo[i]); break;
        case SMART_COND_MIN: sel = (x >= c.lo[i]); break;
        case SMART_COND_MAX: sel = (x <= c.lo[i]); break;
        case SMART_COND_INSIDE: sel = (x >= c.lo[i]) && (x <= c.hi[i]); break;
        default: sel = (x <= c.lo[i]) && (x >= c.hi[i]); break;      // 'outside', the reference's literal rule
        }
        ok = ok && sel;
    }
    return ok;
}

// order-preserving image of a double: a < b  <=>  key(a) < key(b); NaN sorts last (largest) as in
// numpy's argsort; -0.0 == +0.0; 0 is never produced, it marks the rows the constraints removed
__device__ __forceinline__ unsigned long long ordered_key(double x)
{
    if (x != x) return ~0ull;
    const unsigned long long u = (unsigned long long)__double_as_longlong(x + 0.0);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

struct SelectState {
    unsigned long long hist[256];
    unsigned long long key_prefix;    // digits of the k-th largest key fixed so far
    unsigned long long k_rem;         // rank still to resolve among the rows matching the prefix
    unsigned long long kept;          // rows that pass the constraints
    unsigned long long k;             // requested number of rows
    unsigned int row_prefix;          // digits of the tie-breaking row index fixed so far
    unsigned int n_out;               // gather cursor
};

// ------------------------------------------------------------------ mask + ordered compaction
__global__ void __launch_bounds__(kRowBlock)
count_rows_kernel(const double *__restrict__ scores, long long n, int ld, CondSet c, unsigned int *blk_count)
{
    const long long i = (long long)blockIdx.x * kRowBlock + threadIdx.x;
    const bool keep = i < n && row_passes(scores + i * ld, c);
    const int cnt = __syncthreads_count(keep);
    if (threadIdx.x == 0) blk_count[blockIdx.x] = (unsigned)cnt;
}

// exclusive scan of the per-block counts by ONE CTA (the list is N/1024 long); total -> *count_out
__global__ void __launch_bounds__(1024)
scan_counts_kernel(const unsigned int *__restrict__ blk_count, unsigned long long *blk_offset, long long n_blk,
                   long long *count_out)
{
    __shared__ unsigned long long warp_sum[32];
    __shared__ unsigned long long carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < n_blk; base += 1024) {
        const long long i = base + threadIdx.x;
        const unsigned long long v = i < n_blk ? blk_count[i] : 0;
        unsigned long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) warp_sum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = warp_sum[lane], winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long o = __shfl_up_sync(0xffffffffu, winc, d);
                if (lane >= d) winc += o;
            }
            warp_sum[lane] = winc - w;          // exclusive over the warps
        }
        __syncthreads();
        const unsigned long long excl = carry + warp_sum[warp] + inc - v;
        if (i < n_blk) blk_offset[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count_out = (long long)carry;
}

__global__ void __launch_bounds__(kRowBlock)
scatter_rows_kernel(const double *__restrict__ scores, long long n, int ld, CondSet c,
                    const unsigned long long *__restrict__ blk_offset, long long *rows_out)
{
    __shared__ int warp_cnt[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long i = (long long)blockIdx.x * kRowBlock + threadIdx.x;
    const bool keep = i < n && row_passes(scores + i * ld, c);
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[warp] = __popc(ballot);
    __syncthreads();
    if (warp == 0) {
        const int w = warp_cnt[lane];
        int winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += o;
        }
        warp_cnt[lane] = winc - w;
    }
    __syncthreads();
    if (keep) rows_out[blk_offset[blockIdx.x] + warp_cnt[warp] + __popc(ballot & ((1u << lane) - 1u))] = i;
}

// ------------------------------------------------------------------ radix select of the k best
__global__ void __launch_bounds__(256)
build_keys_kernel(const double *__restrict__ scores, long long n, int ld, int target, CondSet c,
                  unsigned long long *__restrict__ keys, SelectState *st)
{
    unsigned int kept = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double *row = scores + i * ld;
        const bool keep = row_passes(row, c);
        keys[i] = keep ? ordered_key(row[target]) : 0ull;
        kept += keep;
    }
    kept = __reduce_add_sync(0xffffffffu, kept);
    if ((threadIdx.x & 31) == 0 && kept) atomicAdd(&st->kept, (unsigned long long)kept);
}

__global__ void init_state_kernel(SelectState *st, unsigned long long k)
{
    st->hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        st->key_prefix = 0;
        st->k_rem = k;
        st->kept = 0;
        st->k = k;
        st->row_prefix = 0;
        st->n_out = 0;
    }
}

// pass 0..7: digit `pass` (from the top) of the keys that agree with the prefix on the digits above;
// pass 8..11: same on the row index, among the rows whose key equals the resolved key
__global__ void __launch_bounds__(256)
histogram_kernel(const unsigned long long *__restrict__ keys, long long n, int pass, SelectState *st)
{
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long kp = st->key_prefix;
    const unsigned int rp = st->row_prefix;
    const bool on_key = pass < 8;
    const int shift = on_key ? 56 - 8 * pass : 24 - 8 * (pass - 8);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long key = keys[i];
        if (on_key) {
            const bool match = shift == 56 || (key >> (shift + 8)) == (kp >> (shift + 8));
            if (match) atomicAdd(&h[(unsigned)(key >> shift) & 255u], 1u);
        } else if (key == kp) {
            const unsigned int r = (unsigned int)i;
            const bool match = shift == 24 || (r >> (shift + 8)) == (rp >> (shift + 8));
            if (match) atomicAdd(&h[(r >> shift) & 255u], 1u);
        }
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

// pick the bin that holds the k_rem-th largest, fix that digit, clear the histogram
__global__ void __launch_bounds__(256)
pick_digit_kernel(int pass, SelectState *st)
{
    __shared__ unsigned long long h[256];
    h[threadIdx.x] = st->hist[threadIdx.x];
    st->hist[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long above = 0, k_rem = st->k_rem;
        int b = 255;
        for (; b > 0; --b) {
            if (above + h[b] >= k_rem) break;
            above += h[b];
        }
        st->k_rem = k_rem - above;
        if (pass < 8) st->key_prefix |= (unsigned long long)b << (56 - 8 * pass);
        else st->row_prefix |= (unsigned int)b << (24 - 8 * (pass - 8));
    }
}

__global__ void __launch_bounds__(256)
gather_best_kernel(const unsigned long long *__restrict__ keys, long long n, SelectState *st,
                   unsigned long long *__restrict__ cand_key, unsigned int *__restrict__ cand_row, long long cap)
{
    const unsigned long long kp = st->key_prefix;
    const unsigned int rp = st->row_prefix;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long key = keys[i];
        if (key > kp || (key == kp && (unsigned int)i >= rp)) {
            const unsigned int slot = atomicAdd(&st->n_out, 1u);
            if (slot < cap) {
                cand_key[slot] = key;
                cand_row[slot] = (unsigned int)i;
            }
        }
    }
}

__global__ void pad_candidates_kernel(unsigned long long *cand_key, unsigned int *cand_row, long long k, long long padded)
{
    const long long i = k + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < padded) {
        cand_key[i] = ~0ull;
        cand_row[i] = kPadRow;          // above every real (key, row) pair: stays at the tail
    }
}

__device__ __forceinline__ bool pair_greater(unsigned long long ka, unsigned int ra, unsigned long long kb, unsigned int rb)
{
    return ka > kb || (ka == kb && ra > rb);
}

// bitonic network, ascending on (key, row).  Stages kk_first..kk_last, steps j <= chunk/2, in
// shared memory; a stage wider than the chunk has had its j >= chunk steps done in global memory
__global__ void __launch_bounds__(kSortChunk / 2)
bitonic_shared_kernel(unsigned long long *cand_key, unsigned int *cand_row, int chunk, long long kk_first, long long kk_last)
{
    __shared__ unsigned long long sk[kSortChunk];
    __shared__ unsigned int sr[kSortChunk];
    const long long base = (long long)blockIdx.x * chunk;
    for (int t = threadIdx.x; t < chunk; t += blockDim.x) {
        sk[t] = cand_key[base + t];
        sr[t] = cand_row[base + t];
    }
    __syncthreads();
    for (long long kk = kk_first; kk <= kk_last; kk <<= 1) {
        long long j0 = kk >> 1;
        if (j0 > chunk / 2) j0 = chunk / 2;
        for (int j = (int)j0; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < chunk / 2; t += blockDim.x) {
                const int a = 2 * t - (t & (j - 1));          // lower element of comparator t at distance j
                const int b = a + j;
                const bool asc = (((base + a) & kk) == 0);
                const unsigned long long ka = sk[a], kb = sk[b];
                const unsigned int ra = sr[a], rb = sr[b];
                if (pair_greater(ka, ra, kb, rb) == asc) {
                    sk[a] = kb; sk[b] = ka;
                    sr[a] = rb; sr[b] = ra;
                }
            }
            __syncthreads();
        }
    }
    for (int t = threadIdx.x; t < chunk; t += blockDim.x) {
        cand_key[base + t] = sk[t];
        cand_row[base + t] = sr[t];
    }
}

__global__ void __launch_bounds__(256)
bitonic_global_kernel(unsigned long long *cand_key, unsigned int *cand_row, long long half, long long j, long long kk)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= half) return;
    const long long a = 2 * t - (t & (j - 1));
    const long long b = a + j;
    const bool asc = ((a & kk) == 0);
    const unsigned long long ka = cand_key[a], kb = cand_key[b];
    const unsigned int ra = cand_row[a], rb = cand_row[b];
    if (pair_greater(ka, ra, kb, rb) == asc) {
        cand_key[a] = kb; cand_key[b] = ka;
        cand_row[a] = rb; cand_row[b] = ra;
    }
}

__global__ void emit_rows_kernel(const unsigned int *__restrict__ cand_row, long long k, const SelectState *st,
                                 long long *rows_out, long long *kept_out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) rows_out[i] = (long long)cand_row[i];
    if (i == 0 && kept_out) *kept_out = (long long)st->kept;
}

// ascending bitonic sort of `padded` (a power of two) (key, row) pairs: 2048-pair chunks in shared
// memory, global steps only for distances >= 2048
void sort_pairs(unsigned long long *cand_key, unsigned int *cand_row, long long padded, cudaStream_t s)
{
    const int chunk = (int)(padded < kSortChunk ? padded : kSortChunk);
    if (padded <= 1) return;
    const unsigned n_chunks = (unsigned)(padded / chunk);
    const int threads = chunk / 2 < 32 ? 32 : chunk / 2;
    int launches = 1;
    bitonic_shared_kernel<<<n_chunks, threads, 0, s>>>(cand_key, cand_row, chunk, 2, chunk);
    for (long long kk = 2ll * chunk; kk <= padded; kk <<= 1) {
        for (long long j = kk >> 1; j >= chunk; j >>= 1) {
            bitonic_global_kernel<<<(unsigned)((padded / 2 + 255) / 256), 256, 0, s>>>(cand_key, cand_row, padded / 2, j, kk);
            ++launches;
        }
        bitonic_shared_kernel<<<n_chunks, threads, 0, s>>>(cand_key, cand_row, chunk, kk, kk);
        ++launches;
    }
    smart_internal_count(launches);
}

// ------------------------------------------------------------------ grouping of the members of a launch
constexpr int kOrderPad = 128;            // the largest CTA of the step kernel: groups start on its multiples
constexpr double kOrderTSlices = 64.0;    // slices of the T range inside which members are ordered by S * Z

struct OrderState {
    unsigned long long t_min, t_max, x_min, x_max;   // ordered_key images (monotonic in the value)
    unsigned int n_fast;
};

__device__ __forceinline__ double key_to_double(unsigned long long k)
{
    const unsigned long long u = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)u);
}

__global__ void order_init_kernel(OrderState *st)
{
    st->t_min = st->x_min = ~0ull;
    st->t_max = st->x_max = 0ull;
    st->n_fast = 0;
}

// range of T and of S * Z over the batch, number of members inside the fast form's domain
__global__ void __launch_bounds__(256)
order_range_kernel(const double *__restrict__ params, long long n, double dt, OrderState *st)
{
    unsigned long long tmin = ~0ull, tmax = 0ull, xmin = ~0ull, xmax = 0ull;
    unsigned int n_fast = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double par[SMART_N_PARAMS];
#pragma unroll
        for (int k = 0; k < SMART_N_PARAMS; ++k) par[k] = params[i * SMART_N_PARAMS + k];
        const unsigned long long t = ordered_key(par[0]), x = ordered_key(par[4] * par[5]);
        tmin = t < tmin ? t : tmin;
        tmax = t > tmax ? t : tmax;
        xmin = x < xmin ? x : xmin;
        xmax = x > xmax ? x : xmax;
        n_fast += smart::fast_form_ok(par, dt) ? 1u : 0u;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, tmin, off), b = __shfl_xor_sync(0xffffffffu, tmax, off);
        const unsigned long long c = __shfl_xor_sync(0xffffffffu, xmin, off), d = __shfl_xor_sync(0xffffffffu, xmax, off);
        tmin = a < tmin ? a : tmin;
        tmax = b > tmax ? b : tmax;
        xmin = c < xmin ? c : xmin;
        xmax = d > xmax ? d : xmax;
    }
    n_fast = __reduce_add_sync(0xffffffffu, n_fast);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&st->t_min, tmin);
        atomicMax(&st->t_max, tmax);
        atomicMin(&st->x_min, xmin);
        atomicMax(&st->x_max, xmax);
        if (n_fast) atomicAdd(&st->n_fast, n_fast);
    }
}

// key = [needs the branch-faithful form] * 2 * slices + slice of T + S * Z scaled into [0, 1)
__global__ void __launch_bounds__(256)
order_keys_kernel(const double *__restrict__ params, long long n, long long padded, double dt, const OrderState *st,
                  unsigned long long *__restrict__ cand_key, unsigned int *__restrict__ cand_row)
{
    const double tmin = key_to_double(st->t_min), tmax = key_to_double(st->t_max);
    const double xmin = key_to_double(st->x_min), xmax = key_to_double(st->x_max);
    const double tw = tmax - tmin + 1e-300, xw = xmax - xmin + 1e-300;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < padded; i += (long long)gridDim.x * blockDim.x) {
        if (i >= n) {
            cand_key[i] = ~0ull;
            cand_row[i] = kPadRow;
            continue;
        }
        double par[SMART_N_PARAMS];
#pragma unroll
        for (int k = 0; k < SMART_N_PARAMS; ++k) par[k] = params[i * SMART_N_PARAMS + k];
        double slice = floor((par[0] - tmin) / tw * kOrderTSlices);
        slice = slice < kOrderTSlices - 1.0 ? slice : kOrderTSlices - 1.0;
        if (!(slice >= 0.0)) slice = 0.0;                       // NaN parameters: anywhere, but somewhere
        double frac = (par[4] * par[5] - xmin) / xw * 0.999;
        if (!(frac >= 0.0)) frac = 0.0;
        const double group = smart::fast_form_ok(par, dt) ? 0.0 : 2.0 * kOrderTSlices;
        cand_key[i] = ordered_key(group + slice + frac);
        cand_row[i] = (unsigned int)i;
    }
}

// sorted rows -> slots: fast members at [0, F), idle slots up to the next multiple of kOrderPad,
// then the members of the branch-faithful form, idle slots to the end.
// (Tried and dropped: dealing the CTAs' worth of members out through a golden-ratio bijection so
// that every SM draws from the whole sorted order.  A C2-sized launch is one wave, every
// sub-partition keeps its 5 or 6 warps for the whole run and the kernel ends with the slowest one;
// warps of an LHS batch differ by +-4.5 % in cost, and a random draw of 5-6 of them is WORSE
// balanced (max/mean 1.20) than neighbours in the sorted order (1.15); measured 28.3 vs 27.7 ms.)
__global__ void __launch_bounds__(256)
order_emit_kernel(const unsigned int *__restrict__ cand_row, long long n, long long slots, const OrderState *st,
                  long long *__restrict__ order_out)
{
    const long long n_fast = st->n_fast;
    const long long general_at = (n_fast + kOrderPad - 1) / kOrderPad * kOrderPad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += (long long)gridDim.x * blockDim.x) {
        long long src = -1;
        if (i < n_fast) src = i;
        else if (i >= general_at && i - general_at < n - n_fast) src = n_fast + (i - general_at);
        order_out[i] = src < 0 ? -1 : (long long)cand_row[src];
    }
}

// ------------------------------------------------------------------ block-constant forcing
__global__ void __launch_bounds__(256)
fold_blocks_kernel(const double *__restrict__ rows, long long n_blocks, int C, int k, double *__restrict__ out, int *flag)
{
    bool same = true;
    const long long total = n_blocks * C;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (long long)gridDim.x * blockDim.x) {
        const long long b = j / C;
        const int c = (int)(j - b * C);
        const double *p = rows + (b * k) * (long long)C + c;
        const unsigned long long first = (unsigned long long)__double_as_longlong(p[0]);
        for (int h = 1; h < k; ++h)     // bit equality: what an equal split of one value produces
            same = same && ((unsigned long long)__double_as_longlong(p[(long long)h * C]) == first);
        out[j] = p[0];
    }
    if (!__all_sync(0xffffffffu, same) && (threadIdx.x & 31) == 0) atomicExch(flag, 0);
}

__global__ void set_flag_kernel(int *flag, int v) { *flag = v; }

long long pow2_at_least(long long k)
{
    long long p = 1;
    while (p < k) p <<= 1;
    return p;
}

size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

int make_conds(const smart_condition *conds, int32_t n_conds, int32_t ld, CondSet &c)
{
    if (n_conds < 0 || n_conds > kMaxConds || (n_conds > 0 && !conds))
        return smart_internal_fail(SMART_ERR_BAD_ARG, "smart_condition: at most SMART_MAX_CONDITIONS conditions");
    c.n = n_conds;
    for (int i = 0; i < kMaxConds; ++i) {
        c.col[i] = 0; c.kind[i] = 0; c.lo[i] = 0.0; c.hi[i] = 0.0;
    }
    for (int i = 0; i < n_conds; ++i) {
        if (conds[i].column < 0 || conds[i].column >= ld)
            return smart_internal_fail(SMART_ERR_BAD_ARG, "smart_condition: column outside the score table");
        if (conds[i].kind < SMART_COND_EQUAL || conds[i].kind > SMART_COND_OUTSIDE)
            return smart_internal_fail(SMART_ERR_BAD_ARG, "smart_condition: unknown kind of condition");
        c.col[i] = conds[i].column;
        c.kind[i] = conds[i].kind;
        c.lo[i] = conds[i].lo;
        c.hi[i] = conds[i].hi;
    }
    return SMART_OK;
}

int grid_for(long long n, int threads)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long want = (n + threads - 1) / threads;
    const long long cap = (long long)sms * 8;          // grid-stride loops: 8 resident CTAs per SM
    if (want > cap) want = cap;
    return (int)(want < 1 ? 1 : want);
}

}  // namespace

extern "C" {

size_t smart_condition_workspace_bytes(int64_t n_rows, int64_t k)
{
    if (n_rows < 0) n_rows = 0;
    if (k < 0) k = 0;
    const size_t n_blk = (size_t)((n_rows + kRowBlock - 1) / kRowBlock);
    const size_t rows = align256(n_blk * sizeof(unsigned int)) + align256(n_blk * sizeof(unsigned long long));
    const size_t padded = (size_t)pow2_at_least(k > 0 ? k : 1);
    const size_t best = align256(sizeof(SelectState)) + align256((size_t)n_rows * 8) + align256(padded * 8) +
                        align256(padded * 4);
    return rows > best ? rows : best;
}

int smart_condition_rows(const double *scores, int64_t n_rows, int32_t ld, const smart_condition *conds,
                         int32_t n_conds, int64_t *rows_out, int64_t *count_out, void *workspace, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    CondSet c;
    if (int rc = make_conds(conds, n_conds, ld, c)) return rc;
    if (n_rows < 0 || n_rows >= 0xFFFFFFFFll || ld < 1 || !count_out || (n_rows > 0 && (!scores || !rows_out || !workspace)))
        return smart_internal_fail(SMART_ERR_BAD_ARG, "smart_condition_rows: bad argument");
    if (n_rows == 0) {
        SEL_CUDA(cudaMemsetAsync(count_out, 0, sizeof(int64_t), s));
        return SMART_OK;
    }
    const long long n_blk = (n_rows + kRowBlock - 1) / kRowBlock;
    char *w = (char *)workspace;
    unsigned int *blk_count = (unsigned int *)w;
    unsigned long long *blk_offset = (unsigned long long *)(w + align256((size_t)n_blk * sizeof(unsigned int)));
    count_rows_kernel<<<(unsigned)n_blk, kRowBlock, 0, s>>>(scores, n_rows, ld, c, blk_count);
    scan_counts_kernel<<<1, 1024, 0, s>>>(blk_count, blk_offset, n_blk, (long long *)count_out);
    scatter_rows_kernel<<<(unsigned)n_blk, kRowBlock, 0, s>>>(scores, n_rows, ld, c, blk_offset, (long long *)rows_out);
    SEL_CUDA(cudaGetLastError());
    smart_internal_count(3);
    return SMART_OK;
}

int smart_best_rows(const double *scores, int64_t n_rows, int32_t ld, int32_t target_column,
                    const smart_condition *conds, int32_t n_conds, int64_t k, int64_t *rows_out,
                    int64_t *kept_out, void *workspace, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    CondSet c;
    if (int rc = make_conds(conds, n_conds, ld, c)) return rc;
    if (n_rows < 1 || n_rows >= 0xFFFFFFFFll || ld < 1 || target_column < 0 || target_column >= ld || !scores ||
        !rows_out || !workspace)
        return smart_internal_fail(SMART_ERR_BAD_ARG, "smart_best_rows: bad argument");
    if (k < 1 || k > n_rows)
        return smart_internal_fail(SMART_ERR_BAD_ARG, "smart_best_rows: k must be in 1..n_rows");

    const long long padded = pow2_at_least(k);
    char *w = (char *)workspace;
    SelectState *st = (SelectState *)w;
    w += align256(sizeof(SelectState));
    unsigned long long *keys = (unsigned long long *)w;
    w += align256((size_t)n_rows * 8);
    unsigned long long *cand_key = (unsigned long long *)w;
    w += align256((size_t)padded * 8);
    unsigned int *cand_row = (unsigned int *)w;

    const int grid = grid_for(n_rows, 256);
    init_state_kernel<<<1, 256, 0, s>>>(st, (unsigned long long)k);
    build_keys_kernel<<<grid, 256, 0, s>>>(scores, n_rows, ld, target_column, c, keys, st);
    for (int pass = 0; pass < 12; ++pass) {
        histogram_kernel<<<grid, 256, 0, s>>>(keys, n_rows, pass, st);
        pick_digit_kernel<<<1, 256, 0, s>>>(pass, st);
    }
    gather_best_kernel<<<grid, 256, 0, s>>>(keys, n_rows, st, cand_key, cand_row, k);
    if (padded > k)
        pad_candidates_kernel<<<(unsigned)((padded - k + 255) / 256), 256, 0, s>>>(cand_key, cand_row, k, padded);
    smart_internal_count(2 + 24 + 1 + (padded > k ? 1 : 0) + 1);      // init, keys, 12 x (histogram, pick), gather, pad, emit

    sort_pairs(cand_key, cand_row, padded, s);
    emit_rows_kernel<<<(unsigned)((k + 255) / 256), 256, 0, s>>>(cand_row, k, st, (long long *)rows_out, (long long *)kept_out);
    SEL_CUDA(cudaGetLastError());
    return SMART_OK;
}

int64_t smart_member_order_len(int64_t n_members)
{
    if (n_members < 0) n_members = 0;
    // the fast group rounded up to a CTA boundary + the other group: at most one padding run
    return (n_members + kOrderPad - 1) / kOrderPad * kOrderPad + kOrderPad;
}

size_t smart_member_order_workspace_bytes(int64_t n_members)
{
    const size_t padded = (size_t)pow2_at_least(n_members > 0 ? n_members : 1);
    return align256(sizeof(OrderState)) + align256(padded * 8) + align256(padded * 4);
}

int smart_member_order(const double *params, int64_t n_members, double dt_sec, int64_t *order_out, void *workspace,
                       void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (!params || n_members < 1 || n_members >= 0xFFFFFFFFll || !(dt_sec > 0.0) || !order_out || !workspace)
        return smart_internal_fail(SMART_ERR_BAD_ARG, "smart_member_order: bad argument");
    const long long padded = pow2_at_least(n_members);
    char *w = (char *)workspace;
    OrderState *st = (OrderState *)w;
    w += align256(sizeof(OrderState));
    unsigned long long *cand_key = (unsigned long long *)w;
    w += align256((size_t)padded * 8);
    unsigned int *cand_row = (unsigned int *)w;
    const long long slots = smart_member_order_len(n_members);
    order_init_kernel<<<1, 1, 0, s>>>(st);
    order_range_kernel<<<grid_for(n_members, 256), 256, 0, s>>>(params, n_members, dt_sec, st);
    order_keys_kernel<<<grid_for(padded, 256), 256, 0, s>>>(params, n_members, padded, dt_sec, st, cand_key, cand_row);
    smart_internal_count(3);
    sort_pairs(cand_key, cand_row, padded, s);
    order_emit_kernel<<<grid_for(slots, 256), 256, 0, s>>>(cand_row, n_members, slots, st, (long long *)order_out);
    smart_internal_count(1);
    SEL_CUDA(cudaGetLastError());
    return SMART_OK;
}

int smart_fold_blocks(const double *rows, int64_t n_rows, int32_t n_catchments, int32_t k, double *out,
                      int32_t *flag_out, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (!rows || !out || !flag_out || n_rows < 1 || n_catchments < 1 || k < 1 || n_rows % k != 0)
        return smart_internal_fail(SMART_ERR_BAD_ARG, "smart_fold_blocks: bad argument (n_rows must be a multiple of k)");
    const long long n_blocks = n_rows / k;
    set_flag_kernel<<<1, 1, 0, s>>>(flag_out, 1);
    fold_blocks_kernel<<<grid_for(n_blocks * n_catchments, 256), 256, 0, s>>>(rows, n_blocks, n_catchments, k, out, flag_out);
    smart_internal_count(2);
    SEL_CUDA(cudaGetLastError());
    return SMART_OK;
}

}  // extern "C"
