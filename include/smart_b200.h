/*
 * smart_b200.h -- C ABI of the B200-native SMART hot path (libsmart_b200.so).
 *
 * What it replaces.  The reference (ThibHlln/smartpy v0.2.2) has exactly one native plug
 * point for this path: an optional module `smartcpp` (smartpy/structure.py:22-27) whose
 * `allsteps` replaces run_all_steps (call sites structure.py:118-121 and :143-146) and
 * whose `onestep` replaces run_one_step (structure.py:171-174, :182-187).  That contract is
 * one member at a time.  The entry points below keep it (smart_allsteps_host) and add the
 * batch form the Monte-Carlo layer needs (smartpy/montecarlo/montecarlo.py:179-209 is a
 * per-sample Python loop in the reference): N members (parameter set x catchment) stepped
 * in one launch, with the objective functions of montecarlo.py:193-209 fused in.
 *
 * Conventions
 *   - plain C types only; every pointer in smart_batch_desc is a DEVICE pointer borrowed
 *     for the duration of the (stream-ordered) call, except in the *_host entry points
 *     where every pointer is a HOST pointer;
 *   - no hidden allocation and no implicit synchronisation in the device entry points;
 *   - return value 0 = success, negative = SMART_ERR_*; smart_last_error() gives the text;
 *   - parameter order is the reference's Parameters.names (smartpy/parameters.py:25):
 *     T, C, H, D, S, Z, SK, FK, GK, RK;
 *   - the 19-vector order is the reference's model_variables (structure.py:78-82):
 *     Q_aeva, Q_ove, Q_dra, Q_int, Q_sgw, Q_dgw, Q_out, V_ove, V_dra, V_int, V_sgw, V_dgw,
 *     V_ly1..V_ly6, V_river  (fluxes m3/s, volumes m3).
 */
#ifndef SMART_B200_H
#define SMART_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMART_B200_VERSION 200 /* 0.2.0 */

#define SMART_N_PARAMS 10
#define SMART_N_VARS 19
#define SMART_N_SCORES 8    /* NSE, KGE, KGEc, KGEa, KGEb, PBias, RMSE, GW (montecarlo.py:71-74) */
#define SMART_OBS_STATS 6   /* n_valid, sum_e, mean_e, sum(e-mean_e), sum((e-mean_e)^2), reserved */

#define SMART_REPORT_SUMMARY 1 /* structure.py:65-66, :190 */
#define SMART_REPORT_RAW 2     /* structure.py:67-68, :193 */

#define SMART_OK 0
#define SMART_ERR_BAD_ARG (-1)       /* null/misaligned pointer, non-positive size, unknown report type */
#define SMART_ERR_WARMUP_TOO_LONG (-2) /* structure.py:90-95 */
#define SMART_ERR_GAP (-3)           /* 'summary' with a run length not a multiple of report_gap:
                                        the reference's np.reshape raises (structure.py:190) */
#define SMART_ERR_CUDA (-4)          /* a CUDA runtime call failed; see smart_last_error() */
#define SMART_ERR_NO_DEVICE (-5)

/* Flags */
#define SMART_FLAG_FORCE_GENERAL 0x1u /* always take the branch-faithful step (clamps, 95% river cap,
                                         leak predicates); default: chosen per CTA from the parameters */
#define SMART_FLAG_NO_TMA 0x2u        /* stage forcing with plain loads instead of cp.async.bulk */
/* Tests and tuning: bits 8..15 ask for a relay (see workspace_bytes) of about that many segments
 * whatever the batch size; 0 = the library decides from the number of waves. */
#define SMART_FLAG_RELAY_SEGS(n) (((unsigned)(n) & 0xffu) << 8)
#define SMART_FLAG_RELAY_SEGS_OF(flags) (((flags) >> 8) & 0xffu)

/*
 * One batch of members.  Member m uses params[m][0..9] and catchment
 *   c(m) = (n_catchments == 1) ? 0 : m / members_per_catchment.
 */
typedef struct smart_batch_desc {
    /* ---- sizes */
    int64_t n_members;              /* N >= 1 */
    int64_t n_steps;                /* T: simulation steps of the main run (len(timeseries) - 1) */
    int64_t n_warmup;               /* W = int(warm_up_days * 86400 / dt) (structure.py:88); 0 = none */
    int32_t n_catchments;           /* C >= 1 */
    int32_t members_per_catchment;  /* used when C > 1; N must equal C * members_per_catchment */
    int32_t report_gap;             /* simulation steps per reporting step (structure.py:75) */
    int32_t report_type;            /* SMART_REPORT_SUMMARY | SMART_REPORT_RAW */
    uint32_t flags;                 /* SMART_FLAG_* */
    int32_t forcing_repeat;         /* 0 or 1: rain/peva hold one row per step.  k > 1: one row per block
                                       of k steps with constant forcing (what the reference's daily ->
                                       hourly split produces, timeframe.py:167-186), values already per
                                       STEP (daily total / k); rows = n_steps / k.  Needs n_steps % k ==
                                       n_warmup % k == 0 and a report_gap that divides k ('summary' or
                                       'raw'; gap == k with 'summary' advances dry blocks in closed form,
                                       any other gap walks the stores hour by hour and reports inside the
                                       block, as structure.py:181-195 does for every gap with one loop). */
    double dt_sec;                  /* simulation time step in seconds */

    /* ---- inputs */
    const double *params;           /* [N][10] row-major, always binary64 */
    const double *rain;             /* [T][C]  mm per step  (C == 1: plain [T]); [T / forcing_repeat][C] */
    const double *peva;             /* [T][C]  mm per step */
    const double *area_m2;          /* [C] */
    const double *obs;              /* [n_report][C], NaN = missing; NULL = no scoring */
    const double *obs_stats;        /* [C][SMART_OBS_STATS] from smart_obs_stats(); required iff obs */
    const double *initial_state;    /* optional [N][19]: start the MAIN run from these states
                                       (no guess, no warm-up) -- the run_all_steps contract */

    /* ---- initial-condition guess, structure.py:100-116 / :125-140 */
    int32_t has_extra;              /* truthiness of the reference's `extra` dict */
    int32_t reserved1;
    double aar;                     /* extra['aar'] */
    double ro_ratio;                /* extra['r-o_ratio'] */
    double ro_split[5];             /* extra['r-o_split'] */
    double gw_constraint;           /* settings 'gw_constraint'; 0 or NaN = off (montecarlo.py:71-74) */

    /* ---- outputs (any may be NULL) */
    void *discharge;                /* [n_report][ld_discharge] m3/s; double (f64 entry) or float (f32 entry) */
    int64_t ld_discharge;           /* >= N */
    double *scores;                 /* [N][ld_scores]; GW column is NaN when the constraint is off */
    double *gw;                     /* [N * ld_gw] groundwater share of runoff (structure.py:191, :194-195) */
    double *last_state;             /* [N][19] state after the last step (structure.py:197) */

    /* ---- best member (optional): arg-max (best_sign > 0) or arg-min (< 0) of scores column
     *      best_column over the batch, reduced with warp shuffles in the same pass */
    int32_t best_column;
    int32_t best_sign;              /* 0 = off */
    double *best_score;             /* [1] */
    int64_t *best_index;            /* [1] */
    void *workspace;                /* device scratch of smart_batch_workspace_bytes() bytes: required when
                                       best_sign != 0, optional otherwise (see workspace_bytes) */

    /* ---- grouping of members into warps (optional, device): a permutation of 0..n_members-1;
     *      thread i of the launch advances member member_order[i] and writes that member's
     *      outputs at the member's own index.  Outputs are unchanged bit for bit (a member's
     *      result does not depend on its neighbours); only the divergence inside a warp changes --
     *      sorted by T, the members of a warp agree on which steps are wet.  Needs
     *      n_catchments == 1 and no discharge / last_state / initial_state. */
    const int64_t *member_order;

    /* ---- leading dimensions of the score outputs (0 = the packed defaults 8 and 1).  With
     *      ld_scores = ld_gw = 9, scores = block and gw = block + 8 the kernel writes the
     *      [N][9] block that a sharded run all-gathers, with no packing pass in between. */
    int64_t ld_scores;
    int64_t ld_gw;
    /* ---- slots in member_order (0 = n_members).  smart_member_order() emits
     *      smart_member_order_len(N) slots: the members that qualify for the fast form first, padded
     *      with idle slots (-1) to a CTA boundary, then the members that need the branch-faithful
     *      form -- so that no CTA mixes the two. */
    int64_t member_order_len;
    /* ---- size of `workspace` in bytes.  0 = the caller sized it for the best-member search only
     *      (the contract before this field existed).  With smart_batch_workspace_bytes() bytes the
     *      library may run a batch that does not fill many waves as a RELAY: the timeline is cut
     *      into segments, one CTA advances one group of members through one segment and parks the
     *      group's state in the workspace for whichever CTA takes the next segment, so that the SMs
     *      share the work evenly instead of waiting for the slowest warp of a single wave.  Results
     *      are the same bits either way. */
    int64_t workspace_bytes;
} smart_batch_desc;

int smart_version(void);
const char *smart_last_error(void);
/* Kernels this library has launched since it was loaded (all entry points, all threads). */
int64_t smart_launch_count(void);

/* n_report implied by a descriptor: T/gap (summary) or ceil(T/gap) (raw). */
int64_t smart_batch_n_report(const smart_batch_desc *d);
size_t smart_batch_workspace_bytes(const smart_batch_desc *d);

/* Per-catchment observation statistics used by the fused scores (device pointers). */
int smart_obs_stats(const double *obs, int64_t n_report, int32_t n_catchments, double *stats, void *stream);

/* The hot path.  FP64 state / FP32 state (scores are always accumulated in binary64). */
int smart_batch_run_f64(const smart_batch_desc *d, void *stream);
int smart_batch_run_f32(const smart_batch_desc *d, void *stream);

/*
 * Daily -> sub-daily disaggregation of cumulative forcing on the device, exactly as
 * smartpy/timeframe.py:167-186 does on the host: every low-resolution value is divided once by
 * `repeat` (IEEE divide) and stamped on `repeat` consecutive steps.
 * in[n_in][C] -> out[n_in * repeat][C].  Device pointers.
 */
int smart_disaggregate(const double *in, int64_t n_in, int32_t n_catchments, int32_t repeat, double *out,
                       void *stream);
/* Same stamping without the division: out[(i * repeat + k)][c] = in[i][c]. */
int smart_expand(const double *in, int64_t n_in, int32_t n_catchments, int32_t repeat, double *out, void *stream);
/* The general form: out[(i * repeat + k)][c] = in[i][c] / divisor (one IEEE division per value). */
int smart_stamp(const double *in, int64_t n_in, int32_t n_catchments, int32_t repeat, double divisor, double *out,
                void *stream);

/*
 * Objective functions of already simulated series (montecarlo.py:193-209 on its own):
 * discharge[n_report][ld] (double when precision == 64, float when 32) against
 * obs[n_report][C] -> scores[N][8] with the GW column left NaN.  Device pointers.
 */
int smart_score_discharge(const void *discharge, int64_t ld_discharge, int64_t n_members, int64_t n_report,
                          const double *obs, const double *obs_stats, int32_t n_catchments,
                          int32_t members_per_catchment, int precision, double *scores, void *stream);

/*
 * Same, with HOST pointers everywhere in *d (workspace/obs_stats ignored: handled inside).
 * Copies in, runs, copies out, synchronises.  precision: 64 | 32.  Device buffers come from a
 * grow-only arena kept per host thread and device (no cudaMalloc on the steady path);
 * smart_host_arena_release() frees it.
 */
int smart_batch_run_host(const smart_batch_desc *d, int precision, int device);
int smart_host_arena_release(void);

/*
 * Grouping of the members of a launch (device pointers, stream-ordered, no allocation): writes
 * smart_member_order_len(n_members) slots into order_out -- see smart_batch_desc.member_order_len.
 * Inside each group the members are sorted so that the lanes of a warp take the same branches:
 * first by slice of T (the wet/dry predicate rain * T - peva >= 0 depends on the member through T
 * only), then by S * Z (how much room the leaks open in the top soil layer each hour).  Results
 * of a run do not change by a bit.  workspace: smart_member_order_workspace_bytes(n_members).
 */
int64_t smart_member_order_len(int64_t n_members);
size_t smart_member_order_workspace_bytes(int64_t n_members);
int smart_member_order(const double *params, int64_t n_members, double dt_sec, int64_t *order_out, void *workspace,
                       void *stream);

/*
 * Is rows[n_rows][C] constant inside aligned blocks of k rows?  If so writes the folded series
 * out[n_rows / k][C] (one row per block) and sets *flag_out (device int32) to 1, else to 0.
 * n_rows must be a multiple of k.  This is how a per-step series produced by the reference's
 * daily -> hourly split (timeframe.py:167-186) is recognised and handed to the kernel as one row
 * per block (forcing_repeat = k).
 */
int smart_fold_blocks(const double *rows, int64_t n_rows, int32_t n_catchments, int32_t k, double *out,
                      int32_t *flag_out, void *stream);

/*
 * Drop-in for smartcpp.allsteps (structure.py:118-121, :143-146), host pointers:
 * discharge_out[length/gap | ceil(length/gap)], gw_out[1], last_out[19].
 */
int smart_allsteps_host(double area_m2, double delta_sec, int64_t length_simu,
                        const double *nd_rain, const double *nd_peva,
                        const double *nd_parameters, const double *nd_initial,
                        int32_t report_type, int32_t report_gap,
                        double *discharge_out, double *gw_out, double *last_out, int device);

/*
 * Measurement aid: register-resident dependent-chain DFMA (fp64) or FFMA (fp32) loop,
 * `chains` independent chains per thread, `iters` FMAs per chain.  out[blocks*threads].
 * FMA count = blocks * threads * chains * iters.  Used by bench.py to measure the FP64/FP32
 * pipe peak that the roofline fraction is quoted against.
 */
int smart_fma_peak_probe(int precision, int blocks, int threads, int64_t iters, double *out, void *stream);

/*
 * Conditioning of a scored sample on the device.  `scores` is the
 * [n_rows][ld] table of doubles a batch run wrote (ld >= the columns used), device memory.
 * A condition keeps the rows whose `column` value is == lo (EQUAL), >= lo (MIN), <= lo (MAX),
 * in [lo, hi] (INSIDE) or "<= lo and >= hi" (OUTSIDE, the reference's literal rule); NaN fails
 * every comparison.  Replaces the numpy masks of smartpy/montecarlo/glue.py:246-289 and the
 * mask + full argsort of smartpy/montecarlo/best.py:243-287.
 */
#define SMART_MAX_CONDITIONS 8
enum { SMART_COND_EQUAL = 0, SMART_COND_MIN = 1, SMART_COND_MAX = 2, SMART_COND_INSIDE = 3, SMART_COND_OUTSIDE = 4 };

typedef struct smart_condition {
    int32_t column;
    int32_t kind;
    double lo;
    double hi;
} smart_condition;

/* bytes of device scratch both calls below need for n_rows rows and (smart_best_rows) k winners */
size_t smart_condition_workspace_bytes(int64_t n_rows, int64_t k);

/*
 * GLUE: rows_out[0..*count_out) = ascending indices of the rows passing all conditions
 * (rows_out: device, room for n_rows; count_out: device, one int64).  `conds` is HOST memory
 * (copied into the launch); stream-ordered, no allocation, no synchronisation.
 */
int smart_condition_rows(const double *scores, int64_t n_rows, int32_t ld, const smart_condition *conds,
                         int32_t n_conds, int64_t *rows_out, int64_t *count_out, void *workspace, void *stream);

/*
 * Best: rows_out[0..k) = the rows holding the k largest values of `target_column` among the
 * rows passing the conditions, in ascending order of (value, row) -- the best set last, ties
 * resolved like numpy.argsort(kind='stable')[-k:], NaN counted as the largest value like numpy
 * does.  *kept_out (device int64, may be NULL) = rows passing the conditions; when it is < k
 * the output is meaningless and the caller raises the reference's "restrained sample size"
 * exception (best.py:284-285).  Radix select + sort of the k winners only.
 */
int smart_best_rows(const double *scores, int64_t n_rows, int32_t ld, int32_t target_column,
                    const smart_condition *conds, int32_t n_conds, int64_t k, int64_t *rows_out,
                    int64_t *kept_out, void *workspace, void *stream);

/*
 * Latin Hypercube sample on the device (the construction of
 * smartpy/montecarlo/lhs.py:133-167 with the host permutation replaced by a keyed bijection, so
 * any row range can be generated on its own: a rank writes rows [row_first, row_first + n_rows)
 * of the [n_total][n_params] sample into out[n_rows][n_params] (device) and the union over the
 * ranks is ONE stratified sample).  `bounds` is HOST memory, [n_params][2] = (lower, upper).
 * Not the reference's random stream; the algorithm is spelled out in csrc/smart_sample.cu.
 */
#define SMART_LHS_MAX_PARAMS 16
int smart_lhs_rows(uint64_t seed, int64_t n_total, int64_t row_first, int64_t n_rows, int32_t n_params,
                   const double *bounds, double *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SMART_B200_H */
