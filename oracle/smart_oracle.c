/*
 * oracle/smart_oracle.c -- CPU restatement of the SMART hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker for the CUDA path; it is never shipped or measured as the
 * product.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.
 *
 * It restates, operation for operation and in the reference's evaluation order, the
 * algorithm of smartpy v0.2.2 (binary64, no FMA contraction: build with
 * -ffp-contract=off, no -ffast-math):
 *
 *   smart_oracle_onestep    <- smartpy/structure.py:200-264 (run_one_step), which calls
 *                              structure.py:267-458 (run_one_step_catchment) and
 *                              structure.py:461-503 (run_one_step_river)
 *   smart_oracle_allsteps   <- smartpy/structure.py:149-197 (run_all_steps)
 *   smart_oracle_run        <- smartpy/structure.py:30-146  (run)
 *
 * Parity pin: tests/test_oracle_golden.py checks these functions against outputs of the
 * reference itself (the .npz files in tests/golden/, produced by tests/golden/make_golden.py) --
 * discharge bit-for-bit -- and against the reference's printed known answers
 * (tests/test_run_daily_to_hourly.py:31-121, examples/out/ExampleDaily/ExampleDaily.{mod,obs}.flow).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define N_VARS 19 /* 7 outputs + 12 states, structure.py:78-82 */

/* structure.py:461-503 */
static void one_step_river(double dt, double q_in, double rk, double v_riv, double *q_out, double *v_out)
{
    rk *= 3600.0;                                   /* :482 */
    double q = v_riv / rk;                          /* :487 */
    double v_old = v_riv;                           /* :489 */
    double v_tmp = v_old + (q_in - q) * dt;         /* :490 */
    if (v_tmp < 0.0) {                              /* :492 */
        q = 0.95 * (q_in + v_old / dt);             /* :494 */
        v_riv += (q_in - q) * dt;                   /* :496 */
    } else {
        v_riv = v_tmp;                              /* :498 */
    }
    *q_out = q;
    *v_out = v_riv;
}

/* structure.py:267-458; out = 6 fluxes then 11 states in the order of :453-458 */
static void one_step_catchment(double area, double dt, double in_rain, double in_peva,
                               double p_t, double p_c, double p_h, double p_d, double p_s, double p_z,
                               double p_sk, double p_fk, double p_gk,
                               double v_ove, double v_dra, double v_int, double v_sgw, double v_dgw,
                               const double v_ly[6], double out[17])
{
    const double nb_soil_layers = 6.0;              /* :317 */
    p_sk *= 3600.0;                                 /* :320-322 */
    p_fk *= 3600.0;
    p_gk *= 3600.0;

    double z[7], lvl[7];
    z[0] = 0.0;
    lvl[0] = 0.0;
    for (int i = 1; i <= 6; ++i) {
        z[i] = p_z / nb_soil_layers;                /* :329-337 */
        lvl[i] = v_ly[i - 1] / area * 1e3;          /* :339-347 */
    }
    /* Python sum(): starts from int 0, left to right (:350) */
    double lvl_total_start = 0.0;
    for (int i = 0; i <= 6; ++i) lvl_total_start = lvl_total_start + lvl[i];

    double rain = in_rain * p_t;                    /* :353 */
    double excess_rain = rain - in_peva;            /* :355 */
    double aeva = 0.0;                              /* :357 */
    double overland_flow, drain_flow, inter_flow, shallow_flow, deep_flow;

    if (excess_rain >= 0.0) {                       /* :359 */
        aeva += in_peva;                            /* :361 */
        double h_prime = p_h * (lvl_total_start / p_z);   /* :363 */
        overland_flow = h_prime * excess_rain;      /* :364 */
        excess_rain -= overland_flow;               /* :365 */
        for (int i = 1; i <= 6; ++i) {              /* :367-374 */
            double space_in_lyr = z[i] - lvl[i];
            if (excess_rain <= space_in_lyr) {
                lvl[i] += excess_rain;
                excess_rain = 0.0;
            } else {
                lvl[i] = z[i];
                excess_rain -= space_in_lyr;
            }
        }
        drain_flow = p_d * excess_rain;             /* :376 */
        inter_flow = (1.0 - p_d) * excess_rain;     /* :377 */
        double s_prime = p_s * (lvl_total_start / p_z);   /* :379 */
        for (int i = 1; i <= 6; ++i) {              /* :381-385 */
            double leak = lvl[i] * pow(s_prime, (double)i);
            if (leak < lvl[i]) {
                inter_flow += leak;
                lvl[i] -= leak;
            }
        }
        shallow_flow = 0.0;                         /* :387-392 */
        for (int i = 1; i <= 6; ++i) {
            double leak = lvl[i] * (s_prime / (double)i);
            if (leak < lvl[i]) {
                shallow_flow += leak;
                lvl[i] -= leak;
            }
        }
        deep_flow = 0.0;                            /* :394-399 */
        for (int i = 6; i >= 1; --i) {
            double leak = lvl[i] * pow(s_prime, (double)(7 - i));
            if (leak < lvl[i]) {
                deep_flow += leak;
                lvl[i] -= leak;
            }
        }
    } else {                                        /* :400-419 */
        overland_flow = 0.0;
        drain_flow = 0.0;
        inter_flow = 0.0;
        shallow_flow = 0.0;
        deep_flow = 0.0;
        double deficit_rain = excess_rain * (-1.0);
        aeva += rain;
        for (int i = 1; i <= 6; ++i) {
            if (lvl[i] >= deficit_rain) {
                lvl[i] -= deficit_rain;
                aeva += deficit_rain;
                deficit_rain = 0.0;
            } else {
                aeva += lvl[i];
                deficit_rain = p_c * (deficit_rain - lvl[i]);
                lvl[i] = 0.0;
            }
        }
    }

    out[0] = aeva / 1e3 * area / dt;                /* :424 */

    double q;
    q = v_ove / p_sk;                               /* :427-430 */
    v_ove += (overland_flow / 1e3 * area) - (q * dt);
    if (v_ove < 0.0) v_ove = 0.0;
    out[1] = q;
    q = v_dra / p_sk;                               /* :432-435 */
    v_dra += (drain_flow / 1e3 * area) - (q * dt);
    if (v_dra < 0.0) v_dra = 0.0;
    out[2] = q;
    q = v_int / p_fk;                               /* :437-440 */
    v_int += (inter_flow / 1e3 * area) - (q * dt);
    if (v_int < 0.0) v_int = 0.0;
    out[3] = q;
    q = v_sgw / p_gk;                               /* :442-445 */
    v_sgw += (shallow_flow / 1e3 * area) - (q * dt);
    if (v_sgw < 0.0) v_sgw = 0.0;
    out[4] = q;
    q = v_dgw / p_gk;                               /* :447-450 */
    v_dgw += (deep_flow / 1e3 * area) - (q * dt);
    if (v_dgw < 0.0) v_dgw = 0.0;
    out[5] = q;

    out[6] = v_ove;                                 /* :453-458 */
    out[7] = v_dra;
    out[8] = v_int;
    out[9] = v_sgw;
    out[10] = v_dgw;
    for (int i = 1; i <= 6; ++i) out[10 + i] = lvl[i] / 1e3 * area;
}

/* structure.py:200-264.  in[26] = area, dt, rain, peva, T,C,H,D,S,Z,SK,FK,GK,RK, 12 states;
 * out[19] in the order of :259-264. */
void smart_oracle_onestep(const double *in, double *out)
{
    double c[17], q_riv, v_riv;
    one_step_catchment(in[0], in[1], in[2], in[3],
                       in[4], in[5], in[6], in[7], in[8], in[9], in[10], in[11], in[12],
                       in[14], in[15], in[16], in[17], in[18], &in[19], c);
    /* inflow summed left to right (:254) */
    one_step_river(in[1], c[1] + c[2] + c[3] + c[4] + c[5], in[13], in[25], &q_riv, &v_riv);
    for (int i = 0; i < 6; ++i) out[i] = c[i];
    out[6] = q_riv;
    for (int i = 0; i < 11; ++i) out[7 + i] = c[6 + i];
    out[18] = v_riv;
}

/* numpy's pairwise summation for one contiguous-or-strided run of n doubles
 * (the order np.mean(..., axis=-1) uses at structure.py:190 -- verified bit-for-bit against
 * the reference for gaps 1, 13 and 24 by tests/test_oracle_golden.py). */
static double np_pairwise_sum(const double *a, long n, long stride)
{
    if (n < 8) {
        double res = 0.;
        for (long i = 0; i < n; ++i) res += a[i * stride];
        return res;
    } else if (n <= 128) {
        double r[8], res;
        long i;
        for (int j = 0; j < 8; ++j) r[j] = a[j * stride];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[(i + j) * stride];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i * stride];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum(a, n2, stride) + np_pairwise_sum(a + n2 * stride, n - n2, stride);
    }
}

/* structure.py:149-197.  Returns 0, or -1 on allocation failure, -2 if 'summary' is asked
 * with length % report_gap != 0 (the reference's np.reshape raises there, :190).
 * storage_out (optional) receives the full [(length+1)][19] history. */
int smart_oracle_allsteps(double area, double dt, long length,
                          const double *rain, const double *peva,
                          const double *params, const double *initial,
                          int report_type, long report_gap,
                          double *discharge, double *gw, double *last, double *storage_out)
{
    if (report_type == 1 && report_gap > 0 && length % report_gap != 0) return -2;
    double *st = storage_out ? storage_out : (double *)malloc(sizeof(double) * (size_t)(length + 1) * N_VARS);
    if (!st) return -1;
    memcpy(st, initial, sizeof(double) * N_VARS);                      /* :179 */
    double in[26];
    in[0] = area;
    in[1] = dt;
    for (int k = 0; k < 10; ++k) in[4 + k] = params[k];
    for (long i = 1; i <= length; ++i) {                                /* :181-187 */
        in[2] = rain[i - 1];
        in[3] = peva[i - 1];
        for (int k = 0; k < 12; ++k) in[14 + k] = st[(i - 1) * N_VARS + 7 + k];
        smart_oracle_onestep(in, &st[i * N_VARS]);
    }
    long n_rep = report_gap > 0 ? length / report_gap : 0;
    double num = 0.0, den = 0.0;
    if (report_type == 1) {                                             /* :190-191 */
        for (long r = 0; r < n_rep; ++r)
            discharge[r] = np_pairwise_sum(&st[(1 + r * report_gap) * N_VARS + 6], report_gap, N_VARS)
                           / (double)report_gap;
        for (long i = 1; i <= length; ++i) {
            const double *row = &st[i * N_VARS];
            num += row[4] + row[5];
            den += row[1] + row[2] + row[3] + row[4] + row[5];
        }
    } else {                                                            /* :193-195 */
        /* [::-gap][::-1]: rows length, length-gap, ... while >= 1; ceil(length/gap) of them */
        long n_raw = report_gap > 0 ? (length + report_gap - 1) / report_gap : 0;
        for (long r = 0; r < n_raw; ++r) {
            long i = length - (n_raw - 1 - r) * report_gap;
            const double *row = &st[i * N_VARS];
            discharge[r] = row[6];
            num += row[4] + row[5];
            den += row[1] + row[2] + row[3] + row[4] + row[5];
        }
    }
    *gw = num / den;
    if (last) memcpy(last, &st[length * N_VARS], sizeof(double) * N_VARS);   /* :197 */
    if (!storage_out) free(st);
    return 0;
}

/* structure.py:30-146.  report_type 1 = 'summary', 2 = 'raw' (:65-70).
 * has_extra = truthiness of the reference's `extra` dict (:100, :125).
 * Returns 0; -3 when the warm-up exceeds the simulation period (:90-95); other codes as
 * smart_oracle_allsteps. */
int smart_oracle_run(double area, double dt, long simu_length,
                     const double *rain, const double *peva, const double *params,
                     int has_extra, double aar, double ro_ratio, const double *ro_split,
                     double warm_up_days, int report_type, long report_gap,
                     double *discharge, double *gw)
{
    double initial[N_VARS], guess[N_VARS], tmp_gw;
    memset(initial, 0, sizeof initial);
    memset(guess, 0, sizeof guess);
    if (has_extra) {                                                    /* :100-112 / :125-137 */
        const int kidx[5] = {6, 6, 7, 8, 8};
        for (int j = 0; j < 5; ++j)
            guess[7 + j] = (aar * ro_ratio) * ro_split[j] / 1000 * area / 8766 * params[kidx[j]];
        guess[18] = (aar * ro_ratio) / 1000 * area / 8766 * params[9];
    }
    for (int j = 0; j < 6; ++j) guess[12 + j] = (params[5] / 12) / 1000 * area;   /* :115-116 / :139-140 */

    if (warm_up_days != 0) {                                            /* :87 */
        long warm_up_length = (long)(warm_up_days * 86400 / dt);       /* :88 */
        if (warm_up_length > simu_length) return -3;                    /* :90-95 */
        long n_rep = report_type == 1 ? warm_up_length / report_gap
                                      : (warm_up_length + report_gap - 1) / report_gap;
        double *scratch = (double *)malloc(sizeof(double) * (size_t)(n_rep + 1));
        if (!scratch) return -1;
        int rc = smart_oracle_allsteps(area, dt, warm_up_length, rain, peva, params, guess,
                                       report_type, report_gap, scratch, &tmp_gw, initial, NULL);   /* :118-121 */
        free(scratch);
        if (rc) return rc;
    } else {
        memcpy(initial, guess, sizeof initial);
    }
    return smart_oracle_allsteps(area, dt, simu_length, rain, peva, params, initial,
                                 report_type, report_gap, discharge, gw, NULL, NULL);   /* :143-146 */
}
