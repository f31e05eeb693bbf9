// smart_kernels.cu -- B200 (sm_100a) kernels and the C ABI of include/smart_b200.h.
//
// One thread = one member (parameter set x catchment).  The member's twelve stores live in
// registers for the whole run; the thread walks the warm-up and then the main period
// sequentially (the recurrence of smartpy/structure.py:181-187 is nonlinear in the state, so
// time cannot be scanned).  Forcing is staged chunk by chunk into shared memory -- with 1-D
// TMA bulk copies (cp.async.bulk + mbarrier, double buffered) for a single catchment, with
// coalesced [t][catchment] tile loads for multi-catchment batches -- and broadcast to every
// member of the CTA.  Reporting (structure.py:188-195) and the objective functions
// (smartpy/montecarlo/montecarlo.py:193-209) are fused into the same pass: discharge goes
// out as coalesced [t_report][member] stores, or not at all when only scores are wanted.
#include "smart_b200.h"
#include "smart_step.cuh"

#include <cuda_runtime.h>
#include <math_constants.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

// Translation units.  A member's binary64 result must not depend on WHICH instantiation of a
// kernel advanced it (register budget, CTA size -- chosen from the batch size): left to itself
// ptxas contracts a * b + c differently from one instantiation to the next.  This file is therefore
// compiled with -fmad=false: the only FMAs are the ones written in the source, and the host
// emulation of smart_step.cuh reproduces the GPU bit for bit.  The binary32-state kernels owe the
// reference a tolerance, not bits, and keep the compiler's contraction (9 % faster): they are
// compiled as a second translation unit, smart_kernels_f32.cu, which defines SMART_TU_F32 and
// includes this file -- it then emits smart_batch_run_f32 and nothing else.
#ifdef SMART_TU_F32
int smart_internal_fail(int code, const char *msg);     // the error channel lives in the binary64 unit
#endif
void smart_internal_count(int n);                       // launch counter, shared by every unit
void *smart_internal_side_stream();                     // per-device side stream (binary64 unit)

namespace {

using namespace smart;

constexpr int kChunkSingle = 512;  // forcing steps per smem stage, single catchment
constexpr int kAccSlots = 8;       // per-thread binary64 accumulators parked in smem
constexpr int kConstSlots = 7;     // per-thread R-typed constants of the fast step parked in smem
constexpr int kBlockSlots = 7;     // per-thread binary64 constants of the dry-block closed form
constexpr int kSmemHeader = 128;   // two mbarriers, padded (the relay ticket of the CTA sits at byte 32)
constexpr int kRelaySlots = 26;    // doubles a member parks between two segments of a relay (run_timeline)
#ifndef SMART_STEP_UNROLL
#define SMART_STEP_UNROLL 1
#endif
constexpr int kStepUnroll = SMART_STEP_UNROLL;   // unroll factor of the per-step time loop
constexpr bool kSubNested = false;               // fill ladder of the hours that report (block-sub mode): one branch for the lower five layers
#ifndef SMART_FAST_REGS_F64
#define SMART_FAST_REGS_F64 96     // register budget of the fast FP64 kernel (sweep 80..104 in profiles/): 20 warps per SM, no spills
#endif
#ifndef SMART_FAST_REGS_F64_LEAN
#define SMART_FAST_REGS_F64_LEAN 80   // lean budget: 25 warps per SM, a few bytes of spills
#endif
#ifndef SMART_SLOW_REGS
#define SMART_SLOW_REGS 128         // branch-faithful kernels (general, fluxes)
#endif
#ifndef SMART_MULTI_REGS_F64
#define SMART_MULTI_REGS_F64 96     // one-warp CTAs of multi-catchment batches (sweep 80..128 in profiles/: 96-128 within 2 %)
#endif
#ifndef SMART_LEAN_RATE
#define SMART_LEAN_RATE 0.94        // throughput of the lean fast kernel relative to the roomy one at steady state
#endif
constexpr double kLeanRate = SMART_LEAN_RATE;
#ifndef SMART_FAST_REGS_F32
#define SMART_FAST_REGS_F32 72     // fast FP32 kernel: 70 registers used, no spills, 28 warps per SM (sweep 64..96)
#endif

#ifndef SMART_TU_F32
thread_local std::string g_err;

int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}
#else
int fail(int code, const std::string &msg) { return smart_internal_fail(code, msg.c_str()); }
#endif

#define SMART_CUDA(expr)                                                                      \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return fail(SMART_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));  \
    } while (0)

// kernels launched by this library since it was loaded (smart_launch_count(): bench.py reports
// the launches of its timed region from this counter, not from a model of the code)
#ifndef SMART_TU_F32
std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
#else
void count_launches(int n) { smart_internal_count(n); }
#endif

// The stream the branch-faithful kernel runs on beside the fast one (one per device, created on
// first use, highest priority so that its few CTAs are placed as soon as a slot frees up).  The
// fork/join events are shared by every caller: the mutex keeps one call's record/wait pairs together.
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    std::mutex mu;
};
#ifndef SMART_TU_F32
SideStream *side_stream_impl()
{
    static std::mutex create_mu;
    static SideStream *per_device[64] = {nullptr};
    static const bool off = getenv("SMART_B200_NO_SIDE_STREAM") != nullptr;
    if (off) return nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> hold(create_mu);
    if (per_device[dev] == nullptr) {
        SideStream *s = new SideStream;
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&s->stream, cudaStreamNonBlocking, hi) != cudaSuccess ||
            cudaEventCreateWithFlags(&s->fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s->join, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            delete s;
            return nullptr;
        }
        per_device[dev] = s;
    }
    return per_device[dev];
}
SideStream *side_stream() { return side_stream_impl(); }
#else
SideStream *side_stream() { return static_cast<SideStream *>(smart_internal_side_stream()); }
#endif

struct KArgs {
    const double *params, *rain, *peva, *area, *obs, *obs_stats, *initial_state;
    void *discharge;
    double *scores, *gw, *last_state;
    double *blk_best_score;
    long long *blk_best_index;
    const long long *order;      // optional grouping: thread i advances member order[i] (< 0: idle thread)
    long long n_threads;         // threads of the launch that may carry a member (N, or the length of order)
    long long N, T, W, ld_q, ld_s, ld_g;
    long long q_stride_bytes;    // ld_q * sizeof(state type): one report row of the discharge output
    int C, mpc, gap, report_type;
    // multi-catchment thread layout (see member_of_thread): whole warps of one catchment first, the
    // remainders of all catchments packed behind them; tile geometry of the remainder region
    int full, rem;               // warps / leftover lanes per catchment: mpc = 32 * full + rem
    int kc_rem, chunk_rem;
    long long n_full_threads;    // 32 * full * C
    int chunk, kc, use_tma, force_general;
    int has_extra, best_col, best_sign, first_report;
    int rep, mode;               // steps per forcing row (1 = one row per step); kModeStep / kModeBlock / kModeBlockSub
    // relay (see smart_batch_kernel): the timeline in n_seg segments of seg_chunks forcing stages, one CTA
    // per (segment, group of members); n_seg <= 1: one CTA walks the whole timeline of its group
    int n_seg, seg_chunks, n_groups;
    unsigned *relay_ticket;      // order in which the CTAs of this launch came to life
    int *relay_progress;         // [n_groups] segments of the group that are done
    double *relay_state;         // [n_groups][kRelaySlots][BLOCK] state parked between segments
    double dt, aar_ro, split[5], gw_constraint;
};

// ------------------------------------------------------------------ PTX helpers (TMA 1-D bulk + mbarrier)
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Ampere-style asynchronous 8-byte copies (SASS LDGSTS) for the [t][catchment] tiles, whose rows
// are too short (kc * 8 bytes) for the 16-byte granules of cp.async.bulk
__device__ __forceinline__ void cp_async_8(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    // try_wait suspends in hardware up to a time limit; the bound only turns a lost copy
    // into a trap instead of a hung GPU.
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 26)) __trap();
}

// doubles per forcing stage: chunk rows + one padding row (the fast step reads one step ahead),
// rounded to 128 bytes so that every stage is a valid cp.async.bulk destination
__host__ __device__ __forceinline__ int stage_doubles(int chunk, int kc) { return ((chunk + 1) * kc + 15) & ~15; }

// ------------------------------------------------------------------ shared-memory layout of one CTA
//   [0, 128)                       two mbarriers (TMA stages), relay ticket, padded
//   double acc[kAccSlots][BLOCK]   per-thread binary64 accumulators touched once per report step
//   R      kconst[kConstSlots][BLOCK]  per-thread constants of the fast step (see smart_step_fast)
//   double td[BLOCK]               parameter T in binary64 (wet/dry predicate)
//   double kblock[kBlockSlots][BLOCK]  dry-block constants (block mode only, see smart_block_fast)
//   double rain[2][chunk * kc]     forcing stages
//   double peva[2][chunk * kc]
enum : int { kVariantFast = 0, kVariantGeneral = 1, kVariantFluxes = 2 };
// How the time loop reads the forcing (host: mode_of):
//   kModeStep     one forcing row per step;
//   kModeBlock    one row per block of a.rep steps AND one report per block ('summary', gap == rep):
//                 wet blocks hour by hour, dry blocks in closed form (smart_block_fast);
//   kModeBlockSub one row per block, reports inside the block (gap divides rep, 'summary' or 'raw'):
//                 no per-step forcing loads or predicates, dry blocks with the soil in closed form
//                 and the stores walked hour by hour (every hour's outflow is reported or summed).
enum : int { kModeStep = 0, kModeBlock = 1, kModeBlockSub = 2 };

// The per-thread columns come FIRST, at offsets known at compile time (the stages, whose size
// depends on the launch, follow them): an access is then `[tid * 8 + constant]` instead of an
// address rebuilt from the stage size at every use.
template <typename R, int BLOCK, int kMode>
struct Smem {
    uint64_t *full;
    double *rain, *peva, *acc;
    R *kconst;
    double *td, *kblock;
    __device__ __forceinline__ Smem(unsigned char *raw, int tile)   // tile = stage_doubles(chunk, kc)
    {
        full = reinterpret_cast<uint64_t *>(raw);
        acc = reinterpret_cast<double *>(raw + kSmemHeader);
        kconst = reinterpret_cast<R *>(acc + kAccSlots * BLOCK);
        td = reinterpret_cast<double *>(kconst + (kConstSlots + 1) * BLOCK);   // +1: keeps 8-byte alignment for R = float
        kblock = td + BLOCK;
        rain = kblock + (kMode == kModeBlock ? kBlockSlots : 0) * BLOCK;       // (128-byte aligned: BLOCK >= 32)
        peva = rain + 2 * tile;
    }
};

// ------------------------------------------------------------------ the time loop
// kMode != kModeStep: the forcing arrays hold one row per block of a.rep steps (constant forcing
// inside the block, the reference's daily -> hourly disaggregation), run and warm-up lengths are
// whole numbers of blocks.  kModeBlock: one row = one reporting step, the fast form advances a
// whole block at a time (smart_block_fast).  kModeBlockSub: reports fall inside the block.
template <typename R, int kVariant, int BLOCK, bool kSingle, int kMode>
__device__ __forceinline__ bool run_timeline(const KArgs &a, MemberState<R> &s, const MemberPar<R> &p,
                                             const FastPar<R> &fp_, const Smem<R, BLOCK, kMode> &sm, long long m, bool active,
                                             int c, int col, int c_base, int kc_cta, int chunk, double area,
                                             double &gw_out, StepOut<R> &o, int seg, double *park)
{
    constexpr bool kFast = kVariant == kVariantFast;
    constexpr bool kWide = sizeof(R) == 8;    // binary64 state: run-long sums stay in registers
    constexpr bool kDaily = kMode != kModeStep;
    const int kc = kSingle ? 1 : kc_cta;
    const int tile = stage_doubles(chunk, kc);
    const int tid = threadIdx.x;
    double &A = sm.acc[0 * BLOCK + tid];      // sum (s - ebar)
    double &B = sm.acc[1 * BLOCK + tid];      // sum (s - ebar)^2
    double &Cc = sm.acc[2 * BLOCK + tid];     // sum (s - ebar)(e - ebar)
    double &E = sm.acc[3 * BLOCK + tid];      // sum (s - e)^2
    double &GN = sm.acc[4 * BLOCK + tid];     // sum (Q_sgw + Q_dgw)
    double &GD = sm.acc[5 * BLOCK + tid];     // sum of the five pathway flows
    double &RIV0 = sm.acc[6 * BLOCK + tid];   // river store at the start of the main run
    double &SCALE = sm.acc[7 * BLOCK + tid];  // mm per step -> m3/s (and / report_gap for 'summary')
    const R *kconst = sm.kconst + tid;

    // rows of the forcing arrays: one per step, or one per block of a.rep steps
    const int rep = kDaily ? a.rep : 1;
    const long long rowsW = a.W / rep, rowsT = a.T / rep;
    const int nWc = static_cast<int>((rowsW + chunk - 1) / chunk);
    const int nTot = nWc + static_cast<int>((rowsT + chunk - 1) / chunk);
    // with one simulation step per reporting step 'raw' and 'summary' coincide (structure.py:190-195)
    const bool summary = kMode == kModeBlock || a.report_type == SMART_REPORT_SUMMARY || a.gap == 1;

    auto chunk_span = [&](int ci, long long &t0, int &n) {
        if (ci < nWc) {
            t0 = static_cast<long long>(ci) * chunk;
            n = static_cast<int>(min(static_cast<long long>(chunk), rowsW - t0));
        } else {
            t0 = static_cast<long long>(ci - nWc) * chunk;
            n = static_cast<int>(min(static_cast<long long>(chunk), rowsT - t0));
        }
    };
    // TMA producer (one thread): even element counts go through cp.async.bulk (16-byte
    // granules); an odd tail element is placed with a plain store BEFORE the releasing arrive.
    auto tma_issue = [&](int ci, int b) {
        long long t0;
        int n;
        chunk_span(ci, t0, n);
        double *dr = sm.rain + b * tile, *dp = sm.peva + b * tile;
        const int n_even = n & ~1;
        if (n & 1) {
            dr[n - 1] = a.rain[t0 + n - 1];
            dp[n - 1] = a.peva[t0 + n - 1];
        }
        mbar_arrive_expect_tx(&sm.full[b], 2u * n_even * 8u);
        if (n_even) {
            tma_bulk_g2s(dr, a.rain + t0, n_even * 8u, &sm.full[b]);
            tma_bulk_g2s(dp, a.peva + t0, n_even * 8u, &sm.full[b]);
        }
    };
    // generic producer (every thread): rows t0..t0+n, columns c_base..c_base+kc of [rows][C], as
    // asynchronous copies so that the next stage loads while the current one is consumed
    auto async_issue = [&](int ci, int b) {
        long long t0;
        int n;
        chunk_span(ci, t0, n);
        double *dr = sm.rain + b * tile, *dp = sm.peva + b * tile;
        for (int idx = tid; idx < n * kc; idx += BLOCK) {
            const int row = idx / kc, cc = idx - row * kc;
            const int cg = c_base + cc;
            if (cg < a.C) {
                const long long g = (t0 + row) * static_cast<long long>(a.C) + cg;
                cp_async_8(dr + idx, a.rain + g);
                cp_async_8(dp + idx, a.peva + g);
            } else {
                dr[idx] = 0.0;
                dp[idx] = 0.0;
            }
        }
        cp_async_commit();
    };

    int countdown = 0x7fffffff;   // never fires during the warm-up
    R acc = R(0), agw = R(0), aall = R(0);
    FastCarry<R> carry;
    carry.tot = carry.part = R(0);
    carry.valid = false;
    o.q_riv = o.q_gw = o.q_all = R(0);
    o.aeva = o.q_ove = o.q_dra = o.q_int = o.q_sgw = o.q_dgw = R(0);

    // one reporting step (structure.py:188-195 + montecarlo.py:193-203), sval in m3/s.  When reports
    // fall inside a block (every hour for C4a) the output cursor advances by one row per report
    // instead of being rebuilt from the report index (a 64-bit multiply-add per store); elsewhere a
    // report is rare and the index costs one register instead of a pointer.
    // The cursor is a byte address advanced by a stride the host put in the launch arguments (an add
    // with a constant-bank operand: no per-report reload and shift of ld_q), and whether there is an
    // output at all is one launch-wide predicate on the store: a tail thread that shadows a member
    // writes that member's own values to that member's own slots a second time instead of carrying a
    // null pointer to test at every report.
    constexpr bool kCursor = kMode == kModeBlockSub;
    const bool has_q = kCursor && a.discharge != nullptr;
    char *q_out = has_q ? reinterpret_cast<char *>(static_cast<R *>(a.discharge) + m) : nullptr;
    int r = 0;

    // Relay: this CTA walks stages [ci_begin, ci_end) of the timeline.  Everything that lives across
    // a stage boundary -- the twelve stores, the running sums in registers and in shared memory, the
    // carried soil total, the report countdown and index -- is parked in `park` (one column per
    // thread, read and written past L1: the previous segment ran on another SM) by the CTA of the
    // previous segment and picked up here; the numbers and the order of operations on them are
    // exactly those of one uninterrupted walk.
    const bool relay = a.n_seg > 1;
    const int ci_begin = relay ? seg * a.seg_chunks : 0;
    const int ci_end = relay ? min(nTot, ci_begin + a.seg_chunks) : nTot;
    if (relay && seg > 0) {
        const double *pk = park;
#pragma unroll
        for (int k = 0; k < 6; ++k) s.ly[k] = static_cast<R>(__ldcg(pk + k * BLOCK));
        s.ove = static_cast<R>(__ldcg(pk + 6 * BLOCK));
        s.dra = static_cast<R>(__ldcg(pk + 7 * BLOCK));
        s.itf = static_cast<R>(__ldcg(pk + 8 * BLOCK));
        s.sgw = static_cast<R>(__ldcg(pk + 9 * BLOCK));
        s.dgw = static_cast<R>(__ldcg(pk + 10 * BLOCK));
        s.riv = static_cast<R>(__ldcg(pk + 11 * BLOCK));
        acc = static_cast<R>(__ldcg(pk + 12 * BLOCK));
        agw = static_cast<R>(__ldcg(pk + 13 * BLOCK));
        aall = static_cast<R>(__ldcg(pk + 14 * BLOCK));
        carry.tot = static_cast<R>(__ldcg(pk + 15 * BLOCK));
#pragma unroll
        for (int k = 0; k < kAccSlots; ++k) sm.acc[k * BLOCK + tid] = __ldcg(pk + (16 + k) * BLOCK);
        const double packed = __ldcg(pk + 24 * BLOCK);
        countdown = __double2hiint(packed);
        r = __double2loint(packed);
        carry.valid = __ldcg(pk + 25 * BLOCK) != 0.0;
        if (sizeof(R) == 8) carry.part = soil_lower(s);   // (what it held: the same five values in the same order)
        if (has_q) q_out += static_cast<long long>(r) * a.q_stride_bytes;
    }
    // the pieces of one report: binary32 sums folded into binary64, the value stored, the value scored
    auto fold_sums = [&]() {
        if (!kWide) {
            GN += static_cast<double>(agw);
            GD += static_cast<double>(aall);
            agw = aall = R(0);
        }
    };
    auto store = [&](R sval) {
        if (kCursor) {
            if (has_q) *reinterpret_cast<R *>(q_out) = sval;
            q_out += a.q_stride_bytes;
        } else if (a.discharge != nullptr && active) {
            static_cast<R *>(a.discharge)[static_cast<long long>(r) * a.ld_q + m] = sval;
        }
    };
    auto score = [&](R sval) {
        const double e = __ldg(&a.obs[static_cast<long long>(r) * a.C + c]);
        if (e == e) {                            // montecarlo.py:195-196 NaN mask
            const double ebar = a.obs_stats[c * SMART_OBS_STATS + 2];
            const double ds = static_cast<double>(sval) - ebar;
            const double de = e - ebar;
            const double df = ds - de;
            A += ds;
            B = fma(ds, ds, B);
            Cc = fma(ds, de, Cc);
            E = fma(df, df, E);
        }
    };
    auto report = [&](R sval) {
        fold_sums();
        store(sval);
        if (a.obs != nullptr) score(sval);
        ++r;
    };

    if (a.use_tma) {
        if (tid == 0) tma_issue(ci_begin, 0);
    } else {
        async_issue(ci_begin, 0);
    }

    for (int ci = ci_begin, stage = 0; ci < ci_end; ++ci, ++stage) {
        const int b = stage & 1;
        long long t0;
        int n;
        chunk_span(ci, t0, n);
        if (a.use_tma) {
            if (tid == 0 && ci + 1 < ci_end) tma_issue(ci + 1, b ^ 1);
            mbar_wait(&sm.full[b], static_cast<uint32_t>((stage >> 1) & 1));
        } else {
            if (ci + 1 < ci_end) {
                async_issue(ci + 1, b ^ 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
        }
        if (ci == nWc) {   // the main run starts here (structure.py:143-146)
            countdown = a.first_report;
            r = 0;
            acc = agw = aall = R(0);
            A = B = Cc = E = GN = GD = 0.0;
            RIV0 = static_cast<double>(s.riv);
            SCALE = area / (1e3 * a.dt) / (summary ? static_cast<double>(a.gap) : 1.0);
        }
        const double *fr = sm.rain + b * tile + col;
        const double *fp = sm.peva + b * tile + col;
        const double *tdp = sm.td + tid;      // T in binary64 (both modes), parked in shared memory
        if (kMode == kModeBlock) {
            const bool in_main = ci >= nWc;
            for (int i = 0; i < n; ++i) {
                const double ex_d = __dsub_rn(__dmul_rn(fr[0], *tdp), fp[0]);   // structure.py:353-355
                if constexpr (kFast) {
                    smart_block_fast<R, BLOCK, kSingle>(s, fp_, kconst, sm.kblock + tid, carry, ex_d, rep, acc, agw);
                } else {
                    for (int h = 0; h < rep; ++h) {
                        smart_step<R, true, kVariant == kVariantFluxes>(s, p, fr[0], fp[0], o);
                        aall += o.q_all;
                        acc += o.q_riv;
                        agw += o.q_gw;
                    }
                }
                fr += kc;
                fp += kc;
                if (in_main) {
                    const R sval = acc * static_cast<R>(SCALE);
                    if (kFast) GD += static_cast<double>(acc);
                    acc = R(0);
                    report(sval);
                }
            }
        } else if (kMode == kModeBlockSub) {
            // reports inside the block: every hour hands its outflows to the reporting code below
            const R scale = static_cast<R>(SCALE);           // (not yet set during the warm-up: unused there)
            const bool in_main = ci >= nWc;
            auto after_hour = [&](R q_riv, R q_gw, R q_all) {
                acc += q_riv;
                if (summary) {
                    agw += q_gw;
                    if (!kFast) aall += q_all;
                }
                if (--countdown == 0) {
                    countdown = a.gap;
                    const R sval = (summary ? acc : q_riv) * scale;             // structure.py:190 | :193
                    if (summary) {
                        if (kFast) {
                            if (kWide) aall += acc;                             // fast form: aall = sum of Q_out
                            else GD += static_cast<double>(acc);
                        }
                    } else {
                        agw += q_gw;
                        aall += q_all;
                    }
                    acc = R(0);
                    report(sval);
                }
            };
            // gap == 1: every step is a report ('raw' and 'summary' coincide) -- no countdown, no
            // per-gap sum; the fast form keeps the sum of Q_out for the groundwater share
            // (the warm-up / main-run and scored / unscored cases are separate copies of the hour loop:
            // as run-time tests they cost ~10 instructions in every hour of C4a)
            auto rows = [&](auto hourly_tag, auto main_tag, auto scored_tag) {
                constexpr bool kHourly = decltype(hourly_tag)::value;
                constexpr bool kMain = decltype(main_tag)::value, kScored = decltype(scored_tag)::value;
                auto every_hour = [&](R q_riv, R q_gw, R q_all) {
                    agw += q_gw;
                    if (kFast) {
                        if (kWide) aall += q_riv;
                        else acc += q_riv;
                    } else {
                        aall += q_all;
                    }
                    if constexpr (kMain) {
                        const R sval = q_riv * scale;
                        fold_sums();
                        store(sval);
                        if constexpr (kScored) score(sval);
                        ++r;
                    }
                };
                auto done = [&](R q_riv, R q_gw, R q_all) {
                    if constexpr (kHourly) every_hour(q_riv, q_gw, q_all);
                    else after_hour(q_riv, q_gw, q_all);
                };
                for (int i = 0; i < n; ++i) {
                    const double rain_i = fr[0], peva_i = fp[0];
                    fr += kc;
                    fp += kc;
                    if constexpr (kFast) {
                        const double ex_d = __dsub_rn(__dmul_rn(rain_i, *tdp), peva_i);   // structure.py:353-355
                        const BlockPar<R> bp = block_par<R, BLOCK>(kconst);
                        const R r_rk = kconst[6 * BLOCK];
                        if (ex_d >= 0.0) {
                            const R ex = static_cast<R>(ex_d);
                            const R hex = fp_.Hz * ex;
                            const unsigned mask = __activemask();
                            if (!carry.valid) carry_form(s, carry);
#pragma unroll 2
                            for (int h = 0; h < rep; ++h) {
                                const R q_riv = s.riv * r_rk;
                                R q_gw, q_in;
                                fast_wet_hour<R, BLOCK, kSubNested>(s, fp_, kconst, bp, carry, ex, hex, mask, q_gw, q_in);
                                done(q_riv, q_gw, q_in);
                            }
                        } else {
                            dry_block_soil<R>(s, kconst[0], fp_.z, ex_d, rep);
                            carry.valid = false;
#pragma unroll 2
                            for (int h = 0; h < rep; ++h) {
                                const R q_riv = s.riv * r_rk;
                                R q_gw, q_in;
                                fast_dry_hour<R, BLOCK>(s, fp_, kconst, bp, q_gw, q_in);
                                done(q_riv, q_gw, q_in);
                            }
                        }
                    } else {
                        for (int h = 0; h < rep; ++h) {
                            smart_step<R, true, kVariant == kVariantFluxes>(s, p, rain_i, peva_i, o);
                            done(o.q_riv, o.q_gw, o.q_all);
                        }
                    }
                }
            };
            if (a.gap == 1) {
                if (!in_main) rows(std::true_type{}, std::false_type{}, std::false_type{});
                else if (a.obs != nullptr) rows(std::true_type{}, std::true_type{}, std::true_type{});
                else rows(std::true_type{}, std::true_type{}, std::false_type{});
                if (!kWide && kFast && in_main) {   // binary32 state: the chunk's sum of Q_out joins the binary64 sum
                    GD += static_cast<double>(acc);
                    acc = R(0);
                }
            } else {
                rows(std::false_type{}, std::false_type{}, std::false_type{});   // (tags unused: run-time tests in after_hour)
            }
        } else {
            // wet/dry driver of the fast step, formed one step ahead of the state
            double ex_next = kFast ? __dsub_rn(__dmul_rn(fr[0], *tdp), fp[0]) : 0.0;
#pragma unroll kStepUnroll
            for (int i = 0; i < n; ++i) {
                if (kFast) {
                    const double ex_d = ex_next;
                    fr += kc;
                    fp += kc;
                    ex_next = __dsub_rn(__dmul_rn(fr[0], *tdp), fp[0]);   // last step of a stage: padding row, unused
                    smart_step_fast<R, BLOCK>(s, fp_, kconst, carry, ex_d, o);
                } else {
                    smart_step<R, true, kVariant == kVariantFluxes>(s, p, fr[0], fp[0], o);
                    fr += kc;
                    fp += kc;
                }
                acc += o.q_riv;
                // groundwater share (structure.py:191, :194-195): 'summary' sums every step, 'raw'
                // only the sampled ones.  Fast form: the pathway total is recovered from the river's
                // mass balance (sum q_in = sum q_out + dV_river), so only sum q_out is accumulated
                // (binary64 state: in the otherwise unused register of the pathway total).
                if (summary) {
                    agw += o.q_gw;
                    if (!kFast) aall += o.q_all;
                }
                if (--countdown == 0) {
                    countdown = a.gap;
                    const R sval = (summary ? acc : o.q_riv) * static_cast<R>(SCALE);   // structure.py:190 | :193
                    if (summary) {
                        if (kFast) {
                            if (kWide) aall += acc;
                            else GD += static_cast<double>(acc);    // once per report step: shared memory
                        }
                    } else {
                        agw += o.q_gw;
                        aall += o.q_all;
                    }
                    acc = R(0);
                    report(sval);
                }
            }
        }
        __syncthreads();   // every thread is done with stage b before it is refilled
    }
    if (ci_end < nTot) {   // relay: hand the member over to the CTA of the next segment
        double *pk = park;
#pragma unroll
        for (int k = 0; k < 6; ++k) __stcg(pk + k * BLOCK, static_cast<double>(s.ly[k]));
        __stcg(pk + 6 * BLOCK, static_cast<double>(s.ove));
        __stcg(pk + 7 * BLOCK, static_cast<double>(s.dra));
        __stcg(pk + 8 * BLOCK, static_cast<double>(s.itf));
        __stcg(pk + 9 * BLOCK, static_cast<double>(s.sgw));
        __stcg(pk + 10 * BLOCK, static_cast<double>(s.dgw));
        __stcg(pk + 11 * BLOCK, static_cast<double>(s.riv));
        __stcg(pk + 12 * BLOCK, static_cast<double>(acc));
        __stcg(pk + 13 * BLOCK, static_cast<double>(agw));
        __stcg(pk + 14 * BLOCK, static_cast<double>(aall));
        __stcg(pk + 15 * BLOCK, static_cast<double>(carry.tot));
#pragma unroll
        for (int k = 0; k < kAccSlots; ++k) __stcg(pk + (16 + k) * BLOCK, sm.acc[k * BLOCK + tid]);
        __stcg(pk + 24 * BLOCK, __hiloint2double(countdown, r));
        __stcg(pk + 25 * BLOCK, carry.valid ? 1.0 : 0.0);
        return false;
    }
    double gn = kWide ? static_cast<double>(agw) : GN;
    double gd = kWide ? static_cast<double>(aall) : GD;
    if (kFast && summary)   // sum of Q_out (block mode and binary32 state keep it in GD) + what the river gained
        gd = ((kWide && kMode != kModeBlock) ? static_cast<double>(aall) : GD) + (static_cast<double>(s.riv) - RIV0);
    gw_out = gn / gd;
    return true;
}

// Final objective functions from the shifted sums (montecarlo.py:199-203; formulas of
// spotpy.objectivefunctions nashsutcliffe / kge / pbias / rmse restated in DESIGN.md).
__device__ __forceinline__ void finish_scores(const double *st, double A, double B, double Cc, double E,
                                              double *sc)
{
    const double n = st[0], sum_e = st[1], ebar = st[2], sde = st[3], sse = st[4];
    const double var_s = B - A * A / n;           // n * variance of the simulation
    const double var_e = sse - sde * sde / n;     // n * variance of the observations
    const double cov = Cc - A * sde / n;
    sc[0] = 1.0 - E / sse;                                    // NSE
    sc[2] = cov / sqrt(var_s * var_e);                        // KGEc (Pearson r)
    sc[3] = sqrt(var_s / var_e);                              // KGEa
    sc[4] = (A + n * ebar) / sum_e;                           // KGEb
    sc[1] = 1.0 - sqrt((sc[2] - 1.0) * (sc[2] - 1.0) + (sc[3] - 1.0) * (sc[3] - 1.0) +
                       (sc[4] - 1.0) * (sc[4] - 1.0));        // KGE
    sc[5] = 100.0 * ((A - sde) / sum_e);                      // PBias
    sc[6] = sqrt(E / n);                                      // RMSE
}

// Scores (montecarlo.py:193-209) and groundwater share of one member to global memory.  `acc` is
// the member's column of the shared-memory accumulators (slot k at acc[k * kStride]).  Returns the
// member's key for the best-member selection (-inf when it cannot win).
template <int kStride>
__device__ __forceinline__ double write_member_results(const KArgs &a, const double *acc, long long m, bool active, int c,
                                                       double gw)
{
    double target = -CUDART_INF;
    if (a.obs != nullptr) {
        const double *st = a.obs_stats + c * SMART_OBS_STATS;
        double sc[SMART_N_SCORES];
        finish_scores(st, acc[0 * kStride], acc[1 * kStride], acc[2 * kStride], acc[3 * kStride], sc);
        const bool gw_on = a.gw_constraint == a.gw_constraint && a.gw_constraint != 0.0;
        sc[7] = gw_on ? ((a.gw_constraint - 0.1 <= gw && gw <= a.gw_constraint + 0.1) ? 1.0 : 0.0)
                      : CUDART_NAN;                               // objfunctions.py:20-24
        if (a.scores != nullptr && active) {
#pragma unroll
            for (int k = 0; k < SMART_N_SCORES; ++k) a.scores[m * a.ld_s + k] = sc[k];
        }
        if (a.best_sign != 0 && active) {
            double t = sc[0];
#pragma unroll
            for (int k = 1; k < SMART_N_SCORES; ++k) t = (a.best_col == k) ? sc[k] : t;
            t = a.best_sign > 0 ? t : -t;
            target = (t == t) ? t : -CUDART_INF;
        }
    }
    if (a.gw != nullptr && active) a.gw[m * a.ld_g] = gw;
    return target;
}

// Arg-max of (target, idx) over the CTA with warp shuffles; ties go to the lower member index.
// `scratch` = the accumulator slots, free to reuse once every thread has finished its scores.
template <int kStride>
__device__ __forceinline__ void cta_best(const KArgs &a, double *scratch, double target, long long idx, int blk)
{
    const int tid = threadIdx.x;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ot = __shfl_xor_sync(0xffffffffu, target, off);
        const long long oi = __shfl_xor_sync(0xffffffffu, idx, off);
        if (ot > target || (ot == target && oi < idx)) {
            target = ot;
            idx = oi;
        }
    }
    __syncthreads();
    double *s_t = scratch;
    long long *s_idx = reinterpret_cast<long long *>(scratch + kStride);
    if ((tid & 31) == 0) {
        s_t[tid >> 5] = target;
        s_idx[tid >> 5] = idx;
    }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < static_cast<int>(blockDim.x) / 32; ++w) {
            const double ot = s_t[w];
            const long long oi = s_idx[w];
            if (ot > target || (ot == target && oi < idx)) {
                target = ot;
                idx = oi;
            }
        }
        a.blk_best_score[blk] = target;
        a.blk_best_index[blk] = idx;
    }
}

template <typename R, int kVariant, int BLOCK, bool kSingle, int kMode>
__device__ __forceinline__ void run_member(const KArgs &a, unsigned char *smem_raw, const double *par,
                                           long long m, bool active, int c, int col, int c_base, int kc_cta, int chunk,
                                           double area, int blk, int seg)
{
    constexpr bool kFast = kVariant == kVariantFast;
    const int tid = threadIdx.x;
    const Smem<R, BLOCK, kMode> sm(smem_raw, stage_doubles(chunk, kSingle ? 1 : kc_cta));
    const double T = par[0], C = par[1], H = par[2], D = par[3], S = par[4], Z = par[5];
    const double SK = par[6], FK = par[7], GK = par[8], RK = par[9];
    MemberPar<R> p;
    p.Td = T;
    p.C = static_cast<R>(C);
    p.D = static_cast<R>(D);
    p.omD = static_cast<R>(1.0 - D);
    p.Hz = static_cast<R>(H / Z);
    p.Sz = static_cast<R>(S / Z);
    p.z = static_cast<R>(Z / 6.0);
    const double r_sk = a.dt / (SK * 3600.0), r_fk = a.dt / (FK * 3600.0);
    const double r_gk = a.dt / (GK * 3600.0), r_rk = a.dt / (RK * 3600.0);
    p.r_sk = static_cast<R>(r_sk);
    p.r_fk = static_cast<R>(r_fk);
    p.r_gk = static_cast<R>(r_gk);
    p.r_rk = static_cast<R>(r_rk);
    FastPar<R> fp_;
    fp_.Hz = p.Hz;
    fp_.Sz = p.Sz;
    fp_.z = p.z;
    fp_.c_sk = static_cast<R>(1.0 - r_sk);
    fp_.c_fk = static_cast<R>(1.0 - r_fk);
    fp_.c_gk = static_cast<R>(1.0 - r_gk);
    fp_.c_rk = static_cast<R>(1.0 - r_rk);
    if (kFast) {
        R *kconst = sm.kconst + tid;
        kconst[0 * BLOCK] = p.C;
        kconst[1 * BLOCK] = p.D;
        kconst[2 * BLOCK] = p.omD;
        kconst[3 * BLOCK] = p.r_sk;
        kconst[4 * BLOCK] = p.r_fk;
        kconst[5 * BLOCK] = p.r_gk;
        kconst[6 * BLOCK] = p.r_rk;
        if (kMode == kModeBlock) {
            // closed form of a dry block of a.rep steps (smart_block_fast): c_x^rep and
            // K_x = r_x * sum_{h<rep} c_w^(rep-1-h) c_x^h by Horner, no cancellation; binary64
            const double cx[3] = {1.0 - r_sk, 1.0 - r_fk, 1.0 - r_gk}, rx[3] = {r_sk, r_fk, r_gk};
            const double cw = 1.0 - r_rk;
            double *kb = sm.kblock + tid;
            double pw_w = 1.0;
            for (int h = 0; h < a.rep; ++h) pw_w *= cw;
            kb[3 * BLOCK] = pw_w;
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                double G = 0.0, pxh = 1.0;
                for (int h = 0; h < a.rep; ++h) {
                    G = fma(cw, G, pxh);
                    pxh *= cx[x];
                }
                kb[x * BLOCK] = pxh;
                kb[(4 + x) * BLOCK] = rx[x] * G;
            }
        }
    }
    sm.td[tid] = T;

    // initial conditions in m3 exactly as the reference writes them, then to mm
    const double to_mm = 1e3 / area;
    double v[12];
    if (a.initial_state != nullptr) {
#pragma unroll
        for (int k = 0; k < 12; ++k) v[k] = a.initial_state[m * SMART_N_VARS + 7 + k];
    } else {
        const double kk[5] = {SK, SK, FK, GK, GK};
#pragma unroll
        for (int k = 0; k < 5; ++k)
            v[k] = a.has_extra ? a.aar_ro * a.split[k] / 1000 * area / 8766 * kk[k] : 0.0;   // structure.py:100-110
        v[11] = a.has_extra ? a.aar_ro / 1000 * area / 8766 * RK : 0.0;                         // :111-112
#pragma unroll
        for (int k = 0; k < 6; ++k) v[5 + k] = (Z / 12) / 1000 * area;                          // :115-116
    }
    MemberState<R> s;
    s.ove = static_cast<R>(v[0] * to_mm);
    s.dra = static_cast<R>(v[1] * to_mm);
    s.itf = static_cast<R>(v[2] * to_mm);
    s.sgw = static_cast<R>(v[3] * to_mm);
    s.dgw = static_cast<R>(v[4] * to_mm);
#pragma unroll
    for (int k = 0; k < 6; ++k) s.ly[k] = static_cast<R>(v[5 + k] * to_mm);
    s.riv = static_cast<R>(v[11] * to_mm);
    if (kFast) {   // same routing constant => one linear reservoir
        s.ove = s.ove + s.dra;
        s.sgw = s.sgw + s.dgw;
        s.dra = s.dgw = R(0);
        if (sizeof(R) == 4) {   // binary32 fast form: the soil as deficits z - level (fast_wet_soil_deficit)
#pragma unroll
            for (int k = 0; k < 6; ++k) s.ly[k] = static_cast<R>(static_cast<double>(p.z) - v[5 + k] * to_mm);
        }
    }

    double gw = 0.0;
    StepOut<R> o;
    double *park = a.n_seg > 1 ? a.relay_state + (static_cast<long long>(blk) * kRelaySlots) * BLOCK + tid : nullptr;
    const bool finished = run_timeline<R, kVariant, BLOCK, kSingle, kMode>(a, s, p, fp_, sm, m, active, c, col, c_base,
                                                                           kc_cta, chunk, area, gw, o, seg, park);
    if (!finished) {
        // relay: the group's state is parked; publish that segment `seg` is done (release: the CTA
        // that takes the next segment acquires this counter before it reads the state)
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.relay_progress + blk), "r"(seg + 1) : "memory");
        }
        return;
    }

    // ---- epilogue: scores (montecarlo.py:193-209), gw, last state, best member
    const double target = write_member_results<BLOCK>(a, sm.acc + tid, m, active, c, gw);
    if (kVariant == kVariantFluxes && a.last_state != nullptr && active) {
        // the 7 fluxes of the last step in m3/s, the 12 states back in m3 (structure.py:259-264)
        double *ls = a.last_state + m * SMART_N_VARS;
        const double to_m3 = area / 1e3;
        const double to_flux = area / (1e3 * a.dt);
        ls[0] = static_cast<double>(o.aeva) * to_flux;
        ls[1] = static_cast<double>(o.q_ove) * to_flux;
        ls[2] = static_cast<double>(o.q_dra) * to_flux;
        ls[3] = static_cast<double>(o.q_int) * to_flux;
        ls[4] = static_cast<double>(o.q_sgw) * to_flux;
        ls[5] = static_cast<double>(o.q_dgw) * to_flux;
        ls[6] = static_cast<double>(o.q_riv) * to_flux;
        ls[7] = static_cast<double>(s.ove) * to_m3;
        ls[8] = static_cast<double>(s.dra) * to_m3;
        ls[9] = static_cast<double>(s.itf) * to_m3;
        ls[10] = static_cast<double>(s.sgw) * to_m3;
        ls[11] = static_cast<double>(s.dgw) * to_m3;
#pragma unroll
        for (int k = 0; k < 6; ++k) ls[12 + k] = static_cast<double>(s.ly[k]) * to_m3;
        ls[18] = static_cast<double>(s.riv) * to_m3;
    }
    if (a.best_sign != 0) cta_best<BLOCK>(a, sm.acc, target, active ? m : 0x7fffffffffffffffLL, blk);
}

// Variant of the step a kernel instantiation carries.  The choice depends on the members'
// parameters, which live on the device, so the host launches the fast kernel and the general
// kernel side by side (two streams, see launch()): every CTA votes (one __syncthreads_or) on
// whether all of its members qualify for the merged form and runs in exactly one of the two
// launches; in the other it exits at once.  Separate kernels keep the fast variant's register
// count (and so its occupancy) independent of the branch-faithful code.  With a member_order from
// smart_member_order() the members that need the branch-faithful form sit in CTAs of their own
// (the order pads the fast group to a CTA boundary with idle threads), so no member's variant --
// and therefore no member's bits -- depends on its neighbours.
template <typename R, int kVariant, int BLOCK, int MAX_REGS, bool kSingle, int kMode>
__global__ void __maxnreg__(MAX_REGS) smart_batch_kernel(const KArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    // Which group of members, which segment of the timeline.  Plain launch: group = blockIdx.x, the
    // whole timeline.  Relay (a.n_seg > 1; the grid holds n_groups * n_seg CTAs): a batch that fills
    // about one wave leaves every SM sub-partition with the 5 or 6 warps it was dealt for the whole
    // run, and the launch ends with the slowest of them while the others idle (C2: 28 ms against
    // 21 ms of evenly shared work).  So the timeline is cut into segments and a CTA advances one
    // group through ONE segment, parking the state for the CTA of the next one.  The hardware's CTA
    // scheduler then is the work queue: whichever SM frees a slot takes the next (segment, group)
    // unit.  Units are numbered by a ticket drawn when the CTA comes to life, segment-major, so the
    // unit a CTA has to wait for (same group, previous segment) always holds a lower ticket: it is
    // running or done, never waiting for a slot -- no deadlock whatever order the CTAs are placed in.
    int blk = blockIdx.x, seg = 0;
    if (a.n_seg > 1) {
        unsigned *ticket = reinterpret_cast<unsigned *>(smem_raw + 32);
        if (tid == 0) *ticket = atomicAdd(a.relay_ticket, 1u);
        __syncthreads();
        const unsigned t = *ticket;
        seg = static_cast<int>(t / static_cast<unsigned>(a.n_groups));
        blk = static_cast<int>(t - static_cast<unsigned>(seg) * static_cast<unsigned>(a.n_groups));
    }
    const long long m_raw = static_cast<long long>(blk) * BLOCK + tid;
    bool active = m_raw < a.n_threads;
    long long m = active ? m_raw : a.n_threads - 1;   // tail threads shadow the last one, store nothing
    int c = 0, c_base = 0, kc_cta = a.kc, chunk = a.chunk;
    if (!kSingle) {
        // thread -> (catchment, member).  A warp that holds members of two catchments walks two
        // weathers: on most days one is wet and the other dry, the warp pays for both paths, and
        // in a wider CTA the other warps wait for it at the staging barrier (C4a: 22 % of the warp
        // time at barriers, 27 of 32 lanes active).  So the launch is laid out in two regions:
        //   [0, n_full_threads)  whole warps of ONE catchment: warp w serves catchment w / full,
        //                        members 32 (w % full) .. + 31 of it;
        //   behind them          the mpc % 32 leftover members of every catchment, packed.
        // Only the second region (mpc % 32 of every mpc members) still mixes catchments; with
        // one-warp CTAs (launch()) no warp waits for another.
        const long long cta_first = static_cast<long long>(blk) * BLOCK;
        if (m < a.n_full_threads) {
            const long long w = m >> 5;
            c = static_cast<int>(w / a.full);
            m = static_cast<long long>(c) * a.mpc + ((w - static_cast<long long>(c) * a.full) << 5) + (m & 31);
            c_base = static_cast<int>((cta_first >> 5) / a.full);
        } else {
            const long long r = m - a.n_full_threads;
            c = static_cast<int>(r / a.rem);
            m = static_cast<long long>(c) * a.mpc + 32LL * a.full + (r - static_cast<long long>(c) * a.rem);
            c_base = static_cast<int>((cta_first - a.n_full_threads) / a.rem);
            kc_cta = a.kc_rem;
            chunk = a.chunk_rem;
        }
    }
    if (a.order != nullptr) {
        // (single catchment, no [t][member] output: validate().)  Idle slots hold -1; an idle thread
        // shadows the member of its CTA's first thread, a CTA whose first slot is idle has no member.
        const long long first = a.order[static_cast<long long>(blk) * BLOCK];
        if (first < 0) {
            if (a.best_sign != 0 && tid == 0) {     // no candidate from this CTA
                a.blk_best_score[blk] = -CUDART_INF;
                a.blk_best_index[blk] = 0x7fffffffffffffffLL;
            }
            return;
        }
        m = a.order[m];
        if (m < 0) {
            active = false;
            m = first;
        }
    }

    double par[SMART_N_PARAMS];
#pragma unroll
    for (int k = 0; k < SMART_N_PARAMS; ++k) par[k] = a.params[m * SMART_N_PARAMS + k];

    if (kVariant != kVariantFluxes) {
        const bool fast_ok = fast_form_ok(par, a.dt) && !a.force_general && a.initial_state == nullptr;
        const int need_general = __syncthreads_or(fast_ok ? 0 : 1);
        if ((need_general != 0) != (kVariant == kVariantGeneral)) return;   // the other launch owns this CTA
    }

    const int col = c - c_base;
    const double area = a.area[c];

    if (seg > 0) {   // relay: the previous segment of this group must be done (its state parked)
        if (tid == 0) {
            // Polling costs issue slots the working warps of the SM want (with fewer groups than
            // slots every unit waits for its predecessor: at a fixed 256 ns the polls were 10 % of all
            // instructions the FP32 kernel executed on C5, 4 % on C2), so the pause doubles up to 8 us --
            // a unit lasts ~0.5 ms.
            int done;
            unsigned pause = 256;
            for (unsigned spins = 0;; ++spins) {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(done) : "l"(a.relay_progress + blk) : "memory");
                if (done >= seg) break;
                __nanosleep(pause);
                if (pause < 8192) pause <<= 1;
                if (spins > (1u << 23)) __trap();   // (a minute: a lost hand-over becomes an error, not a hung GPU)
            }
        }
        __syncthreads();
    }

    if (a.use_tma) {
        uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
        if (tid == 0) {
            mbar_init(&full[0], 1);
            mbar_init(&full[1], 1);
            mbar_fence_init();
        }
        __syncthreads();
    }

    run_member<R, kVariant, BLOCK, kSingle, kMode>(a, smem_raw, par, m, active, c, col, c_base, kc_cta, chunk, area, blk,
                                                   seg);
}

__global__ void best_finalize_kernel(const double *blk_score, const long long *blk_index, int n_blocks, int sign,
                                     double *best_score, long long *best_index)
{
    double t = -CUDART_INF;
    long long idx = 0x7fffffffffffffffLL;
    for (int i = threadIdx.x; i < n_blocks; i += blockDim.x) {
        const double ot = blk_score[i];
        const long long oi = blk_index[i];
        if (ot > t || (ot == t && oi < idx)) {
            t = ot;
            idx = oi;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ot = __shfl_xor_sync(0xffffffffu, t, off);
        const long long oi = __shfl_xor_sync(0xffffffffu, idx, off);
        if (ot > t || (ot == t && oi < idx)) {
            t = ot;
            idx = oi;
        }
    }
    if (threadIdx.x == 0) {
        if (best_score) *best_score = sign > 0 ? t : -t;
        if (best_index) *best_index = idx;
    }
}

// One CTA per catchment; fixed-order tree so the statistics are run-to-run deterministic.
__global__ void obs_stats_kernel(const double *obs, long long n_report, int C, double *stats)
{
    __shared__ double sh[3][256];
    const int c = blockIdx.x, tid = threadIdx.x;
    double n = 0.0, se = 0.0;
    for (long long r = tid; r < n_report; r += 256) {
        const double e = obs[r * C + c];
        if (e == e) {
            n += 1.0;
            se += e;
        }
    }
    sh[0][tid] = n;
    sh[1][tid] = se;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (tid < off) {
            sh[0][tid] += sh[0][tid + off];
            sh[1][tid] += sh[1][tid + off];
        }
        __syncthreads();
    }
    n = sh[0][0];
    se = sh[1][0];
    const double ebar = se / n;
    __syncthreads();
    double sd = 0.0, ss = 0.0;
    for (long long r = tid; r < n_report; r += 256) {
        const double e = obs[r * C + c];
        if (e == e) {
            const double de = e - ebar;
            sd += de;
            ss = fma(de, de, ss);
        }
    }
    sh[1][tid] = sd;
    sh[2][tid] = ss;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (tid < off) {
            sh[1][tid] += sh[1][tid + off];
            sh[2][tid] += sh[2][tid + off];
        }
        __syncthreads();
    }
    if (tid == 0) {
        double *st = stats + static_cast<long long>(c) * SMART_OBS_STATS;
        st[0] = n;
        st[1] = se;
        st[2] = ebar;
        st[3] = sh[1][0];
        st[4] = sh[2][0];
        st[5] = 0.0;
    }
}

// out[(i * repeat + k) * C + c] = in[i * C + c] / div   (timeframe.py:180-183; div = repeat or 1)
__global__ void disaggregate_kernel(const double *in, long long n_in, int C, int repeat, double div, double *out)
{
    const long long total = n_in * repeat * C;
    for (long long j = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; j < total;
         j += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long row = j / C;
        const int c = static_cast<int>(j - row * C);
        out[j] = in[(row / repeat) * C + c] / div;
    }
}

template <typename R>
__global__ void score_discharge_kernel(const R *q, long long ld, long long N, long long n_report, const double *obs,
                                       const double *obs_stats, int C, int mpc, double *scores)
{
    const long long m = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (m >= N) return;
    const int c = C > 1 ? static_cast<int>(m / mpc) : 0;
    const double *st = obs_stats + c * SMART_OBS_STATS;
    const double ebar = st[2];
    double A = 0.0, B = 0.0, Cc = 0.0, E = 0.0;
    for (long long r = 0; r < n_report; ++r) {
        const double e = __ldg(&obs[r * C + c]);
        if (e == e) {
            const double ds = static_cast<double>(q[r * ld + m]) - ebar;
            const double de = e - ebar;
            const double df = ds - de;
            A += ds;
            B = fma(ds, ds, B);
            Cc = fma(ds, de, Cc);
            E = fma(df, df, E);
        }
    }
    double sc[SMART_N_SCORES];
    finish_scores(st, A, B, Cc, E, sc);
    sc[7] = CUDART_NAN;
#pragma unroll
    for (int k = 0; k < SMART_N_SCORES; ++k) scores[m * SMART_N_SCORES + k] = sc[k];
}

template <typename R, int kChains>
__global__ void fma_peak_kernel(long long iters, double *out)
{
    R x[kChains];
    const R a = R(1.0) - R(1e-7) * R(threadIdx.x + 1), b = R(1e-7);
#pragma unroll
    for (int j = 0; j < kChains; ++j) x[j] = R(j + 1);
    for (long long i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < kChains; ++j) x[j] = fma(x[j], a, b);
    }
    R sum = R(0);
#pragma unroll
    for (int j = 0; j < kChains; ++j) sum += x[j];
    out[static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x] = static_cast<double>(sum);
}

// ------------------------------------------------------------------ host side
int64_t n_report_of(const smart_batch_desc *d)
{
    if (d->report_gap <= 0) return 0;
    return d->report_type == SMART_REPORT_SUMMARY ? d->n_steps / d->report_gap
                                                  : (d->n_steps + d->report_gap - 1) / d->report_gap;
}

constexpr int kBlockTiny = 32, kBlockSmall = 64, kBlockLarge = 128;

// Members per CTA.  Members cost the same, so a launch finishes when the SM with the most
// members does: 64-thread CTAs spread a batch of a few waves more evenly over the 148 SMs
// (config C2: 1e5 members = 675.7 per SM); 128-thread CTAs halve the per-CTA staging for
// batches of many waves.  SMART_B200_BLOCK overrides the choice (kernel tuning).
int block_of(const smart_batch_desc *d)
{
    static const int forced = [] {
        const char *e = getenv("SMART_B200_BLOCK");
        return e ? atoi(e) : 0;
    }();
    if (forced == kBlockSmall || forced == kBlockLarge) return forced;
    // multi-catchment batches with whole warps per catchment: one warp per CTA, so that catchments
    // with different weather never wait for each other at the staging barrier
    if (d->n_catchments > 1 && d->members_per_catchment >= 32 && !getenv("SMART_B200_NO_WARP_CTAS")) return kBlockTiny;
    return d->n_members <= 148LL * 736 * 8 ? kBlockSmall : kBlockLarge;
}

// threads of a launch that may carry a member: one per member, or one per slot of member_order
int64_t n_threads_of(const smart_batch_desc *d)
{
    return d->member_order && d->member_order_len > 0 ? d->member_order_len : d->n_members;
}

int n_blocks_of(const smart_batch_desc *d, int block) { return static_cast<int>((n_threads_of(d) + block - 1) / block); }

int mode_of(const smart_batch_desc *d)
{
    if (d->forcing_repeat <= 1) return kModeStep;
    return (d->report_gap == d->forcing_repeat && d->report_type == SMART_REPORT_SUMMARY) ? kModeBlock : kModeBlockSub;
}

// ---- relay (see smart_batch_kernel): which batches may run as one, and the scratch they need
// workspace layout: [best member: blocks * 16 bytes][relay header: 2 tickets, padded to 128 bytes]
//                   [progress: int per group, padded to 128][state: groups * kRelaySlots * block doubles]
constexpr int64_t kRelayMaxThreads = 148LL * 640 * 32;  // (scratch: 208 bytes per thread)
bool relay_possible(const smart_batch_desc *d)
{
    static const bool off = [] {
        const char *e = getenv("SMART_B200_RELAY");
        return e != nullptr && e[0] == '0';
    }();
    return !off && d->n_catchments == 1 && !d->last_state && !d->initial_state && n_threads_of(d) <= kRelayMaxThreads;
}
size_t best_workspace_bytes(const smart_batch_desc *d, int blocks)
{
    return d->best_sign != 0 ? static_cast<size_t>(blocks) * (sizeof(double) + sizeof(long long)) : 0;
}
size_t pad128(size_t n) { return (n + 127) & ~static_cast<size_t>(127); }
size_t relay_workspace_bytes(int blocks, int block)
{
    return 128 + pad128(sizeof(int) * static_cast<size_t>(blocks)) +
           sizeof(double) * static_cast<size_t>(blocks) * kRelaySlots * block;
}
// forcing stages the timeline of a launch is made of (run_timeline: nWc + main stages)
int n_stages_of(const smart_batch_desc *d, int chunk, int rep)
{
    const int64_t W = d->initial_state ? 0 : d->n_warmup;
    const int64_t rowsW = W / rep, rowsT = d->n_steps / rep;
    return static_cast<int>((rowsW + chunk - 1) / chunk + (rowsT + chunk - 1) / chunk);
}

int validate(const smart_batch_desc *d, bool host_mode = false)
{
    if (!d) return fail(SMART_ERR_BAD_ARG, "descriptor is NULL");
    if (d->n_members < 1 || d->n_steps < 1 || d->n_steps >= 0x7fffffffLL || d->n_warmup < 0)
        return fail(SMART_ERR_BAD_ARG, "n_members, n_steps must be >= 1 (n_steps < 2^31), n_warmup >= 0");
    if (d->n_members > 0x7fffffffLL * kBlockSmall) return fail(SMART_ERR_BAD_ARG, "n_members too large for one launch");
    if (d->n_catchments < 1) return fail(SMART_ERR_BAD_ARG, "n_catchments must be >= 1");
    if (d->n_catchments > 1 &&
        (d->members_per_catchment < 1 ||
         static_cast<int64_t>(d->members_per_catchment) * d->n_catchments != d->n_members))
        return fail(SMART_ERR_BAD_ARG, "n_members must equal n_catchments * members_per_catchment");
    if (d->report_type != SMART_REPORT_SUMMARY && d->report_type != SMART_REPORT_RAW)
        return fail(SMART_ERR_BAD_ARG, "Reporting type unknown (1 = summary, 2 = raw)");   // structure.py:70
    if (d->report_gap < 1) return fail(SMART_ERR_BAD_ARG, "report_gap must be >= 1");
    if (!(d->dt_sec > 0.0)) return fail(SMART_ERR_BAD_ARG, "dt_sec must be > 0");
    if (!d->params || !d->rain || !d->peva || !d->area_m2)
        return fail(SMART_ERR_BAD_ARG, "params, rain, peva and area_m2 are required");
    if (d->obs && !d->obs_stats && !host_mode) return fail(SMART_ERR_BAD_ARG, "obs given without obs_stats");
    if (d->n_warmup > d->n_steps)   // structure.py:90-95
        return fail(SMART_ERR_WARMUP_TOO_LONG,
                    "The warm-up duration cannot exceed the length of the simulation period");
    if (d->report_type == SMART_REPORT_SUMMARY) {   // np.reshape(-1, gap) at structure.py:190
        if (d->n_steps % d->report_gap != 0)
            return fail(SMART_ERR_GAP, "cannot reshape the simulation into (-1, report_gap)");
        if (!d->initial_state && d->n_warmup % d->report_gap != 0)
            return fail(SMART_ERR_GAP, "cannot reshape the warm-up run into (-1, report_gap)");
    }
    if (d->forcing_repeat > 1) {
        // block-constant forcing (what the reference's disaggregation produces): whole blocks only,
        // reports at the end of a block or at equal distances inside it
        if (d->n_steps % d->forcing_repeat != 0 || d->n_warmup % d->forcing_repeat != 0 ||
            d->forcing_repeat % d->report_gap != 0 || d->initial_state || d->last_state)
            return fail(SMART_ERR_BAD_ARG,
                        "forcing_repeat > 1 needs a report_gap that divides it, run and warm-up lengths that are "
                        "multiples of it, and no initial_state/last_state");
    }
    if (d->discharge && d->ld_discharge < d->n_members)
        return fail(SMART_ERR_BAD_ARG, "ld_discharge must be >= n_members");
    if (d->member_order && (d->n_catchments != 1 || d->discharge || d->last_state || d->initial_state || host_mode))
        return fail(SMART_ERR_BAD_ARG,
                    "member_order needs one catchment, device pointers and no discharge/last_state/initial_state");
    if (d->member_order && d->member_order_len != 0 && d->member_order_len < d->n_members)
        return fail(SMART_ERR_BAD_ARG, "member_order_len must be 0 (= n_members) or >= n_members");
    if ((d->scores && d->ld_scores != 0 && d->ld_scores < SMART_N_SCORES) || (d->gw && d->ld_gw < 0))
        return fail(SMART_ERR_BAD_ARG, "ld_scores must be 0 (= 8) or >= 8, ld_gw 0 (= 1) or >= 1");
    if (d->best_sign != 0) {
        if (!d->obs || (!d->workspace && !host_mode))
            return fail(SMART_ERR_BAD_ARG, "best member needs obs and workspace");
        if (d->best_column < 0 || d->best_column >= SMART_N_SCORES)
            return fail(SMART_ERR_BAD_ARG, "best_column out of range");
    }
    return SMART_OK;
}

template <typename R>
int launch(const smart_batch_desc *d, cudaStream_t stream)
{
    int rc = validate(d);
    if (rc) return rc;
    KArgs a;
    memset(&a, 0, sizeof a);
    a.params = d->params;
    a.rain = d->rain;
    a.peva = d->peva;
    a.area = d->area_m2;
    a.obs = d->obs;
    a.obs_stats = d->obs_stats;
    a.initial_state = d->initial_state;
    a.discharge = d->discharge;
    a.scores = d->scores;
    a.gw = d->gw;
    a.last_state = d->last_state;
    a.N = d->n_members;
    a.n_threads = n_threads_of(d);
    a.ld_s = d->ld_scores > 0 ? d->ld_scores : SMART_N_SCORES;
    a.ld_g = d->ld_gw > 0 ? d->ld_gw : 1;
    a.T = d->n_steps;
    a.W = d->initial_state ? 0 : d->n_warmup;
    a.ld_q = d->ld_discharge;
    a.q_stride_bytes = d->ld_discharge * static_cast<long long>(sizeof(R));
    a.C = d->n_catchments;
    a.mpc = d->n_catchments > 1 ? d->members_per_catchment : 1;

    a.gap = d->report_gap;
    a.report_type = d->report_type;
    a.has_extra = d->has_extra;
    a.dt = d->dt_sec;
    a.aar_ro = d->aar * d->ro_ratio;   // (extra['aar'] * extra['r-o_ratio']), structure.py:102
    for (int k = 0; k < 5; ++k) a.split[k] = d->ro_split[k];
    a.gw_constraint = d->gw_constraint;
    a.force_general = (d->flags & SMART_FLAG_FORCE_GENERAL) ? 1 : 0;
    a.best_col = d->best_column;
    a.best_sign = d->best_sign;
    a.order = reinterpret_cast<const long long *>(d->member_order);
    // first reporting step of the main run, 1-based: [::-gap][::-1] counts back from the end
    const int64_t n_rep = n_report_of(d);
    a.first_report = static_cast<int>(d->n_steps - (n_rep - 1) * d->report_gap);

    const int block = block_of(d);
    const int blocks = n_blocks_of(d, block);
    const int mode = mode_of(d);
    const bool daily = mode != kModeStep;
    a.mode = mode;
    a.rep = daily ? d->forcing_repeat : 1;
    if (a.C == 1) {
        a.kc = 1;
        a.chunk = daily ? 64 : (block == kBlockLarge ? kChunkSingle : kChunkSingle / 2);
        const bool aligned = (reinterpret_cast<uintptr_t>(d->rain) % 16 == 0) &&
                             (reinterpret_cast<uintptr_t>(d->peva) % 16 == 0);
        a.use_tma = (aligned && !(d->flags & SMART_FLAG_NO_TMA)) ? 1 : 0;
    } else {
        // forcing stages per CTA: 12 KB for the wide CTAs, 3 KB for the one-warp CTAs (20 and more of
        // them share an SM; the stages must not be what limits the resident warps)
        const int stage_bytes = block == kBlockTiny ? 3 * 1024 : 12 * 1024;
        auto tile_rows = [&](int kc) {
            int chunk = stage_bytes / (2 * 2 * 8 * kc);
            chunk = chunk > (daily ? 64 : 512) ? (daily ? 64 : 512) : chunk;
            return chunk & ~7;
        };
        // thread layout (kernel prologue): whole warps per catchment when the CTA is one warp,
        // else the members in their own order (every thread in the "leftover" region)
        a.full = block == kBlockTiny ? a.mpc / 32 : 0;
        a.rem = a.mpc - 32 * a.full;
        a.n_full_threads = 32LL * a.full * a.C;
        a.kc = 1;                                       // whole-warp region: one catchment per CTA
        a.chunk = tile_rows(1);
        a.kc_rem = a.rem > 0 ? (block - 1) / a.rem + 2 : 1;   // catchments one CTA of the leftover region can straddle
        if (a.kc_rem > a.C) a.kc_rem = a.C;
        a.chunk_rem = tile_rows(a.kc_rem);
        if (a.rem == 0) a.rem = 1;                      // (region is empty: keep the divisions defined)
        if (a.chunk_rem < 8) return fail(SMART_ERR_BAD_ARG, "members_per_catchment too small for one CTA tile");
        a.use_tma = 0;
    }
    if (d->best_sign != 0) {
        a.blk_best_score = static_cast<double *>(d->workspace);
        a.blk_best_index = reinterpret_cast<long long *>(a.blk_best_score + blocks);
    }
    a.n_seg = 1;
    a.n_groups = blocks;
    size_t tile = static_cast<size_t>(stage_doubles(a.chunk, a.kc));
    if (a.C > 1 && static_cast<size_t>(stage_doubles(a.chunk_rem, a.kc_rem)) > tile) tile = stage_doubles(a.chunk_rem, a.kc_rem);
    const size_t smem = kSmemHeader + sizeof(double) * (4 * tile + kAccSlots * block) +
                        sizeof(R) * (kConstSlots + 1) * block +
                        sizeof(double) * (1 + (mode == kModeBlock ? kBlockSlots : 0)) * block;
    using Kernel = void (*)(const KArgs);
    unsigned *tickets = nullptr;       // relay: one ticket counter per kernel of the call
    auto go = [&](Kernel kernel, cudaStream_t st, int which = 0) -> int {
        if (smem > 48 * 1024)   // above the default dynamic shared memory limit: opt in per kernel
            SMART_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        KArgs b = a;
        b.relay_ticket = tickets ? tickets + which : nullptr;
        const long long grid = static_cast<long long>(blocks) * (a.n_seg > 1 ? a.n_seg : 1);
        kernel<<<static_cast<unsigned>(grid), block, smem, st>>>(b);
        SMART_CUDA(cudaGetLastError());
        count_launches(1);
        return SMART_OK;
    };
    // Fast kernel: two register budgets are compiled.  The roomy one (no spills) is quicker per
    // wave; the lean one keeps more CTAs resident, which wins when it saves the batch a ragged
    // last wave (config C2: 1e5 members are 1.06 waves at 90 registers, 0.88 at 80).
    // Relay: decided from how many waves the launch would be with `kernel`; lays out the scratch
    // behind the best-member slots and clears the tickets and progress counters on the caller's stream.
    auto try_relay = [&](Kernel kernel) -> int {     // 1 = on, 0 = off, < 0 = error
        static const double min_waves = [] {
            const char *e = getenv("SMART_B200_RELAY_MIN_WAVES");
            return e ? atof(e) : 0.15;
        }();
        static const int forced_segs = [] {
            const char *e = getenv("SMART_B200_RELAY_SEGS");
            return e ? atoi(e) : 0;
        }();
        if (!relay_possible(d) || d->workspace == nullptr) return 0;
        const size_t head = pad128(best_workspace_bytes(d, blocks));
        if (d->workspace_bytes < 0 || static_cast<size_t>(d->workspace_bytes) < head + relay_workspace_bytes(blocks, block)) return 0;
        int dev = 0, sms = 0, per_sm = 0;
        SMART_CUDA(cudaGetDevice(&dev));
        SMART_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        SMART_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem));
        if (per_sm < 1) return 0;
        const double waves = static_cast<double>(blocks) / (static_cast<double>(per_sm) * sms);
        const int asked = static_cast<int>(SMART_FLAG_RELAY_SEGS_OF(d->flags));   // tests and tuning: any batch size
        static const double max_waves = [] {
            const char *e = getenv("SMART_B200_RELAY_MAX_WAVES");
            return e ? atof(e) : 64.0;
        }();
        if (asked == 0 && (waves < min_waves || waves >= max_waves)) return 0;
        const int n_stages = n_stages_of(d, a.chunk, a.rep);
        // about 32 units per slot of the GPU; batches of many waves are balanced by the CTA scheduler
        // already and gain a last 2 % from 8 segments (C3: 764 -> 750 ms)
        int want = asked > 0 ? asked : forced_segs > 0 ? forced_segs : static_cast<int>(ceil(32.0 / waves));
        if (asked == 0 && forced_segs == 0 && want < 8) want = 8;
        if (want > n_stages) want = n_stages;
        if (want < 2) return 0;
        a.seg_chunks = (n_stages + want - 1) / want;
        a.n_seg = (n_stages + a.seg_chunks - 1) / a.seg_chunks;
        if (a.n_seg < 2) {
            a.n_seg = 1;
            return 0;
        }
        char *base = static_cast<char *>(d->workspace) + head;
        tickets = reinterpret_cast<unsigned *>(base);
        a.relay_progress = reinterpret_cast<int *>(base + 128);
        const size_t clear = 128 + pad128(sizeof(int) * static_cast<size_t>(blocks));
        a.relay_state = reinterpret_cast<double *>(base + clear);
        SMART_CUDA(cudaMemsetAsync(base, 0, clear, stream));
        return 1;
    };
    constexpr int kLeanRegs = sizeof(R) == 8 ? SMART_FAST_REGS_F64_LEAN : SMART_FAST_REGS_F32;
    constexpr int kRoomyRegs = sizeof(R) == 8 ? SMART_FAST_REGS_F64 : SMART_FAST_REGS_F32;
    constexpr int kSlowRegs = SMART_SLOW_REGS;
    Kernel fast_lean = nullptr, fast_roomy = nullptr, general = nullptr, fluxes = nullptr;
    auto pick = [&](auto block_tag, auto single_tag, auto mode_tag) {
        constexpr int B = decltype(block_tag)::value;
        constexpr bool S = decltype(single_tag)::value;
        constexpr int M = decltype(mode_tag)::value;
        if constexpr (B == kBlockTiny) {     // one register budget for the one-warp CTAs
            constexpr int kMultiRegs = sizeof(R) == 8 ? SMART_MULTI_REGS_F64 : SMART_FAST_REGS_F32;
            fast_lean = fast_roomy = smart_batch_kernel<R, kVariantFast, B, kMultiRegs, S, M>;
        } else {
            fast_lean = smart_batch_kernel<R, kVariantFast, B, kLeanRegs, S, M>;
            fast_roomy = smart_batch_kernel<R, kVariantFast, B, kRoomyRegs, S, M>;
        }
        general = smart_batch_kernel<R, kVariantGeneral, B, kSlowRegs, S, M>;
        fluxes = smart_batch_kernel<R, kVariantFluxes, B, kSlowRegs, S, kModeStep>;
    };
    auto pick_mode = [&](auto block_tag, auto single_tag) {
        if (mode == kModeStep) pick(block_tag, single_tag, std::integral_constant<int, kModeStep>{});
        else if (mode == kModeBlock) pick(block_tag, single_tag, std::integral_constant<int, kModeBlock>{});
        else pick(block_tag, single_tag, std::integral_constant<int, kModeBlockSub>{});
    };
    using BL = std::integral_constant<int, kBlockLarge>;
    using BS = std::integral_constant<int, kBlockSmall>;
    using BT = std::integral_constant<int, kBlockTiny>;
    if (block == kBlockTiny) {                           // (multi-catchment only: block_of)
        pick_mode(BT{}, std::false_type{});
    } else {
        switch ((block == kBlockLarge ? 2 : 0) | (a.C == 1 ? 1 : 0)) {
            case 0: pick_mode(BS{}, std::false_type{}); break;
            case 1: pick_mode(BS{}, std::true_type{}); break;
            case 2: pick_mode(BL{}, std::false_type{}); break;
            default: pick_mode(BL{}, std::true_type{}); break;
        }
    }
    if (d->last_state) {
        if ((rc = go(fluxes, stream))) return rc;
    } else if (a.force_general || d->initial_state) {
        if ((rc = try_relay(general)) < 0) return rc;
        if ((rc = go(general, stream, 1))) return rc;
    } else {
        Kernel fast = fast_roomy;
        const int relayed = try_relay(fast_roomy);    // a relay is many waves of short CTAs: the roomy kernel
        if (relayed < 0) return relayed;
        if (!relayed && fast_lean != fast_roomy) {
            int dev = 0, sms = 0, per_sm_lean = 0, per_sm_roomy = 0;
            SMART_CUDA(cudaGetDevice(&dev));
            SMART_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            SMART_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_lean, fast_lean, block, smem));
            SMART_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_roomy, fast_roomy, block, smem));
            // cost model: full waves at the measured steady rates (lean is ~6 % slower per
            // member), a ragged last wave costs at least 40 % of a full one (latency bound)
            auto cost = [&](int per_sm, double rate) {
                const double slots = static_cast<double>(per_sm) * sms;
                const double waves = blocks / slots;
                const double whole = floor(waves), frac = waves - whole;
                const double tail = frac == 0.0 ? 0.0 : (frac > 0.4 ? frac : 0.4);
                return (whole + (whole == 0.0 ? frac : tail)) * slots / rate;
            };
            if (per_sm_lean > 0 && per_sm_roomy > 0 && cost(per_sm_lean, kLeanRate) < cost(per_sm_roomy, 1.0))
                fast = fast_lean;
        }
        if (fast_lean != fast_roomy) {
            static const int forced = [] {
                const char *e = getenv("SMART_B200_FAST_REGS");   // kernel tuning: "lean" | "roomy"
                return e ? (e[0] == 'l' ? 1 : 2) : 0;
            }();
            if (forced) fast = forced == 1 ? fast_lean : fast_roomy;
        }
        // The two forms side by side: the branch-faithful kernel goes to a stream of its own
        // (forked from and joined back into the caller's, so the call stays stream-ordered) and
        // its CTAs share the SMs with the fast kernel's instead of queueing behind the whole fast
        // launch -- members outside the fast form's domain are then paid for by their own work,
        // not by a second pass of the timeline at a few CTAs' worth of occupancy.
        SideStream *side = side_stream();
        if (side != nullptr) {
            std::lock_guard<std::mutex> hold(side->mu);
            SMART_CUDA(cudaEventRecord(side->fork, stream));
            SMART_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
            if ((rc = go(general, side->stream, 1))) return rc;
            SMART_CUDA(cudaEventRecord(side->join, side->stream));
            if ((rc = go(fast, stream, 0))) return rc;
            SMART_CUDA(cudaStreamWaitEvent(stream, side->join, 0));
        } else {
            if ((rc = go(fast, stream, 0))) return rc;
            if ((rc = go(general, stream, 1))) return rc;
        }
    }
    if (d->best_sign != 0) {
        best_finalize_kernel<<<1, 32, 0, stream>>>(a.blk_best_score, a.blk_best_index, blocks, d->best_sign,
                                                   d->best_score, reinterpret_cast<long long *>(d->best_index));
        SMART_CUDA(cudaGetLastError());
        count_launches(1);
    }
    return SMART_OK;
}

}  // namespace

#ifdef SMART_TU_F32
extern "C" int smart_batch_run_f32(const smart_batch_desc *d, void *stream)
{
    return launch<float>(d, static_cast<cudaStream_t>(stream));
}
#else
// error channel shared with smart_kernels_f32.cu, smart_select.cu and smart_sample.cu
int smart_internal_fail(int code, const char *msg) { return fail(code, msg); }
void smart_internal_count(int n) { count_launches(n); }
void *smart_internal_side_stream() { return side_stream_impl(); }

// =================================================================== C ABI
extern "C" {

int smart_version(void) { return SMART_B200_VERSION; }

int64_t smart_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char *smart_last_error(void) { return g_err.c_str(); }

int64_t smart_batch_n_report(const smart_batch_desc *d) { return d ? n_report_of(d) : 0; }

size_t smart_batch_workspace_bytes(const smart_batch_desc *d)
{
    if (!d || d->n_members < 1) return 0;
    const int block = block_of(d), blocks = n_blocks_of(d, block);
    const size_t best = best_workspace_bytes(d, blocks);
    return relay_possible(d) ? pad128(best) + relay_workspace_bytes(blocks, block) : best;
}

int smart_obs_stats(const double *obs, int64_t n_report, int32_t n_catchments, double *stats, void *stream)
{
    if (!obs || !stats || n_report < 1 || n_catchments < 1)
        return fail(SMART_ERR_BAD_ARG, "smart_obs_stats: bad argument");
    obs_stats_kernel<<<n_catchments, 256, 0, static_cast<cudaStream_t>(stream)>>>(obs, n_report, n_catchments, stats);
    SMART_CUDA(cudaGetLastError());
    count_launches(1);
    return SMART_OK;
}

int smart_batch_run_f64(const smart_batch_desc *d, void *stream)
{
    return launch<double>(d, static_cast<cudaStream_t>(stream));
}

static int stamp_rows(const double *in, int64_t n_in, int32_t n_catchments, int32_t repeat, double div, double *out,
                      void *stream)
{
    if (!in || !out || n_in < 1 || n_catchments < 1 || repeat < 1)
        return fail(SMART_ERR_BAD_ARG, "smart_disaggregate / smart_expand: bad argument");
    const long long total = static_cast<long long>(n_in) * repeat * n_catchments;
    const int threads = 256;
    const long long want = (total + threads - 1) / threads;
    const int blocks = static_cast<int>(want < 148 * 16 ? want : 148 * 16);
    disaggregate_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(in, n_in, n_catchments, repeat, div,
                                                                                   out);
    SMART_CUDA(cudaGetLastError());
    count_launches(1);
    return SMART_OK;
}

int smart_disaggregate(const double *in, int64_t n_in, int32_t n_catchments, int32_t repeat, double *out,
                       void *stream)
{
    return stamp_rows(in, n_in, n_catchments, repeat, static_cast<double>(repeat), out, stream);
}

int smart_expand(const double *in, int64_t n_in, int32_t n_catchments, int32_t repeat, double *out, void *stream)
{
    return stamp_rows(in, n_in, n_catchments, repeat, 1.0, out, stream);
}

int smart_stamp(const double *in, int64_t n_in, int32_t n_catchments, int32_t repeat, double divisor, double *out,
                void *stream)
{
    return stamp_rows(in, n_in, n_catchments, repeat, divisor, out, stream);
}

int smart_score_discharge(const void *discharge, int64_t ld_discharge, int64_t n_members, int64_t n_report,
                          const double *obs, const double *obs_stats, int32_t n_catchments,
                          int32_t members_per_catchment, int precision, double *scores, void *stream)
{
    if (!discharge || !obs || !obs_stats || !scores || n_members < 1 || n_report < 1 || n_catchments < 1 ||
        ld_discharge < n_members || (n_catchments > 1 && members_per_catchment < 1))
        return fail(SMART_ERR_BAD_ARG, "smart_score_discharge: bad argument");
    const int threads = 128;
    const int blocks = static_cast<int>((n_members + threads - 1) / threads);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int mpc = n_catchments > 1 ? members_per_catchment : 1;
    if (precision == 64)
        score_discharge_kernel<double><<<blocks, threads, 0, st>>>(static_cast<const double *>(discharge), ld_discharge,
                                                                   n_members, n_report, obs, obs_stats, n_catchments,
                                                                   mpc, scores);
    else if (precision == 32)
        score_discharge_kernel<float><<<blocks, threads, 0, st>>>(static_cast<const float *>(discharge), ld_discharge,
                                                                  n_members, n_report, obs, obs_stats, n_catchments,
                                                                  mpc, scores);
    else
        return fail(SMART_ERR_BAD_ARG, "precision must be 64 or 32");
    SMART_CUDA(cudaGetLastError());
    count_launches(1);
    return SMART_OK;
}

// Device arena of the *_host entry points: one grow-only allocation and one stream per host thread
// and device, reused from call to call (a per-call cudaMalloc/cudaFree pair costs more than a small
// batch does).  smart_host_arena_release() gives the memory back.
namespace {
struct HostArena {
    int device = -1;
    char *base = nullptr;
    size_t cap = 0, used = 0;
    cudaStream_t stream = nullptr;
    void release()
    {
        if (base) cudaFree(base);
        if (stream) cudaStreamDestroy(stream);
        base = nullptr;
        stream = nullptr;
        cap = used = 0;
        device = -1;
    }
    ~HostArena() {}   // (process exit: the CUDA context may already be gone; nothing to do)
    void *take(size_t bytes)
    {
        void *p = base + used;
        used += (bytes + 255) & ~static_cast<size_t>(255);
        return p;
    }
};
thread_local HostArena g_arena;
}  // namespace

int smart_host_arena_release(void)
{
    g_arena.release();
    return SMART_OK;
}

int smart_batch_run_host(const smart_batch_desc *h, int precision, int device)
{
    if (precision != 64 && precision != 32) return fail(SMART_ERR_BAD_ARG, "precision must be 64 or 32");
    int rc = validate(h, true);
    if (rc) return rc;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1)
        return fail(SMART_ERR_NO_DEVICE, "no CUDA device: the SMART hot path has no CPU fallback");
    SMART_CUDA(cudaSetDevice(device));

    const int64_t N = h->n_members, T = h->n_steps, C = h->n_catchments;
    const int64_t n_rep = n_report_of(h);
    const size_t q_elem = precision == 64 ? sizeof(double) : sizeof(float);
    const int64_t rows = h->forcing_repeat > 1 ? T / h->forcing_repeat : T;   // one row per block of steps
    const size_t b_params = sizeof(double) * N * SMART_N_PARAMS, b_forcing = sizeof(double) * rows * C;
    const size_t b_area = sizeof(double) * C, b_init = h->initial_state ? sizeof(double) * N * SMART_N_VARS : 0;
    const size_t b_obs = h->obs ? sizeof(double) * n_rep * C : 0, b_stats = h->obs ? sizeof(double) * C * SMART_OBS_STATS : 0;
    const size_t b_q = h->discharge ? q_elem * n_rep * N : 0, b_scores = h->scores ? sizeof(double) * N * SMART_N_SCORES : 0;
    const size_t b_gw = h->gw ? sizeof(double) * N : 0, b_last = h->last_state ? sizeof(double) * N * SMART_N_VARS : 0;
    const size_t b_ws_lib = smart_batch_workspace_bytes(h);
    const size_t b_ws = b_ws_lib > 0 ? pad128(b_ws_lib) + 128 : 0;
    const size_t sizes[] = {b_params, b_forcing, b_forcing, b_area, b_init, b_obs, b_stats, b_q, b_scores, b_gw, b_last, b_ws};
    size_t need = 0;
    for (size_t b : sizes) need += (b + 255) & ~static_cast<size_t>(255);

    HostArena &ar = g_arena;
    if (ar.device != device) {
        ar.release();
        ar.device = device;
    }
    if (ar.stream == nullptr) SMART_CUDA(cudaStreamCreateWithFlags(&ar.stream, cudaStreamNonBlocking));
    if (ar.cap < need) {
        if (ar.base) SMART_CUDA(cudaFree(ar.base));
        ar.base = nullptr;
        ar.cap = 0;
        const size_t grow = need + need / 4;          // head room: the next, slightly larger batch fits too
        SMART_CUDA(cudaMalloc(reinterpret_cast<void **>(&ar.base), grow));
        ar.cap = grow;
    }
    ar.used = 0;
    cudaStream_t st = ar.stream;

    smart_batch_desc d = *h;
    auto up = [&](const void *src, size_t bytes, const void **dst) -> int {
        void *p = ar.take(bytes);
        SMART_CUDA(cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, st));
        *dst = p;
        return SMART_OK;
    };
    if ((rc = up(h->params, b_params, (const void **)&d.params))) return rc;
    if ((rc = up(h->rain, b_forcing, (const void **)&d.rain))) return rc;
    if ((rc = up(h->peva, b_forcing, (const void **)&d.peva))) return rc;
    if ((rc = up(h->area_m2, b_area, (const void **)&d.area_m2))) return rc;
    if (h->initial_state && (rc = up(h->initial_state, b_init, (const void **)&d.initial_state))) return rc;
    if (h->obs) {
        if ((rc = up(h->obs, b_obs, (const void **)&d.obs))) return rc;
        double *stats = static_cast<double *>(ar.take(b_stats));
        d.obs_stats = stats;
        if ((rc = smart_obs_stats(d.obs, n_rep, static_cast<int32_t>(C), stats, st))) return rc;
    }
    void *q = nullptr;
    if (h->discharge) {
        d.ld_discharge = N;
        d.discharge = q = ar.take(b_q);
    }
    if (h->scores) {
        d.scores = static_cast<double *>(ar.take(b_scores));
        d.ld_scores = SMART_N_SCORES;
    }
    if (h->gw) {
        d.gw = static_cast<double *>(ar.take(b_gw));
        d.ld_gw = 1;
    }
    if (h->last_state) d.last_state = static_cast<double *>(ar.take(b_last));
    if (b_ws > 0) {
        char *ws = static_cast<char *>(ar.take(b_ws));
        d.best_score = reinterpret_cast<double *>(ws);
        d.best_index = reinterpret_cast<int64_t *>(ws + 8);
        d.workspace = ws + 128;
        d.workspace_bytes = static_cast<int64_t>(b_ws - 128);
    }
    rc = precision == 64 ? launch<double>(&d, st) : smart_batch_run_f32(&d, st);
    if (rc) return rc;
    if (h->discharge) {
        // host layout keeps the caller's leading dimension
        SMART_CUDA(cudaMemcpy2DAsync(h->discharge, q_elem * h->ld_discharge, q, q_elem * N, q_elem * N, n_rep,
                                     cudaMemcpyDeviceToHost, st));
    }
    if (h->scores) {
        const int64_t ld = h->ld_scores > 0 ? h->ld_scores : SMART_N_SCORES;
        if (ld == SMART_N_SCORES)
            SMART_CUDA(cudaMemcpyAsync(h->scores, d.scores, b_scores, cudaMemcpyDeviceToHost, st));
        else
            SMART_CUDA(cudaMemcpy2DAsync(h->scores, sizeof(double) * ld, d.scores, sizeof(double) * SMART_N_SCORES,
                                         sizeof(double) * SMART_N_SCORES, N, cudaMemcpyDeviceToHost, st));
    }
    if (h->gw) {
        const int64_t ld = h->ld_gw > 0 ? h->ld_gw : 1;
        if (ld == 1)
            SMART_CUDA(cudaMemcpyAsync(h->gw, d.gw, b_gw, cudaMemcpyDeviceToHost, st));
        else
            SMART_CUDA(cudaMemcpy2DAsync(h->gw, sizeof(double) * ld, d.gw, sizeof(double), sizeof(double), N,
                                         cudaMemcpyDeviceToHost, st));
    }
    if (h->last_state)
        SMART_CUDA(cudaMemcpyAsync(h->last_state, d.last_state, b_last, cudaMemcpyDeviceToHost, st));
    if (h->best_sign != 0) {
        if (h->best_score) SMART_CUDA(cudaMemcpyAsync(h->best_score, d.best_score, sizeof(double), cudaMemcpyDeviceToHost, st));
        if (h->best_index) SMART_CUDA(cudaMemcpyAsync(h->best_index, d.best_index, sizeof(long long), cudaMemcpyDeviceToHost, st));
    }
    SMART_CUDA(cudaStreamSynchronize(st));
    return SMART_OK;
}

int smart_allsteps_host(double area_m2, double delta_sec, int64_t length_simu, const double *nd_rain,
                        const double *nd_peva, const double *nd_parameters, const double *nd_initial,
                        int32_t report_type, int32_t report_gap, double *discharge_out, double *gw_out,
                        double *last_out, int device)
{
    smart_batch_desc d;
    memset(&d, 0, sizeof d);
    d.n_members = 1;
    d.n_steps = length_simu;
    d.n_warmup = 0;
    d.n_catchments = 1;
    d.members_per_catchment = 1;
    d.report_gap = report_gap;
    d.report_type = report_type;
    d.flags = SMART_FLAG_FORCE_GENERAL;
    d.dt_sec = delta_sec;
    d.params = nd_parameters;
    d.rain = nd_rain;
    d.peva = nd_peva;
    d.area_m2 = &area_m2;
    d.initial_state = nd_initial;
    d.discharge = discharge_out;
    d.ld_discharge = 1;
    d.gw = gw_out;
    d.last_state = last_out;
    return smart_batch_run_host(&d, 64, device);
}

int smart_fma_peak_probe(int precision, int blocks, int threads, int64_t iters, double *out, void *stream)
{
    if (blocks < 1 || threads < 1 || threads > 1024 || iters < 1 || !out)
        return fail(SMART_ERR_BAD_ARG, "smart_fma_peak_probe: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (precision == 64)
        fma_peak_kernel<double, 8><<<blocks, threads, 0, st>>>(iters, out);
    else if (precision == 32)
        fma_peak_kernel<float, 8><<<blocks, threads, 0, st>>>(iters, out);
    else
        return fail(SMART_ERR_BAD_ARG, "precision must be 64 or 32");
    SMART_CUDA(cudaGetLastError());
    count_launches(1);
    return SMART_OK;
}

}  // extern "C"
#endif  // SMART_TU_F32
