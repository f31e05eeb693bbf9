/*
 * examples/abi_client.c -- the C ABI of include/smart_b200.h used from plain C, no Python, no torch.
 *
 * Reads a small binary case file (written by tests/test_gpu_abi_client.py), runs it through
 * smart_batch_run_host() and writes discharge / scores / gw back as raw doubles, so the test can
 * compare them with the oracle.  Without arguments it only checks that the library links and
 * validates its arguments (no GPU needed).
 *
 *   gcc -I include -o abi_client examples/abi_client.c -L smartpy_b200 -lsmart_b200 -Wl,-rpath,$PWD/smartpy_b200
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "smart_b200.h"

static void *read_doubles(FILE *f, size_t n)
{
    double *p = (double *)malloc(sizeof(double) * (n ? n : 1));
    if (!p || fread(p, sizeof(double), n, f) != n) {
        fprintf(stderr, "short read\n");
        exit(2);
    }
    return p;
}

int main(int argc, char **argv)
{
    printf("smart_version %d\n", smart_version());
    smart_batch_desc d;
    memset(&d, 0, sizeof d);
    if (argc < 3) {
        int rc = smart_batch_run_f64(&d, NULL);      /* must be refused before any CUDA call */
        printf("empty descriptor -> %d (%s)\n", rc, smart_last_error());
        return rc == SMART_ERR_BAD_ARG ? 0 : 1;
    }
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    /* header: N, T, W, gap, report_type, forcing_repeat, has_obs as int64; dt, area, gw_constraint as double */
    long long h[7];
    double g[3];
    if (fread(h, sizeof(long long), 7, f) != 7 || fread(g, sizeof(double), 3, f) != 3) return 2;
    const long long N = h[0], T = h[1], rows = h[5] > 1 ? T / h[5] : T;
    d.n_members = N;
    d.n_steps = T;
    d.n_warmup = h[2];
    d.n_catchments = 1;
    d.members_per_catchment = 1;
    d.report_gap = (int)h[3];
    d.report_type = (int)h[4];
    d.forcing_repeat = (int)h[5];
    d.dt_sec = g[0];
    d.area_m2 = &g[1];
    d.gw_constraint = g[2];
    d.has_extra = 1;
    d.aar = 1200.0;
    d.ro_ratio = 0.45;
    const double split[5] = {0.10, 0.15, 0.15, 0.30, 0.30};
    memcpy(d.ro_split, split, sizeof split);
    d.params = (const double *)read_doubles(f, (size_t)N * SMART_N_PARAMS);
    d.rain = (const double *)read_doubles(f, (size_t)rows);
    d.peva = (const double *)read_doubles(f, (size_t)rows);
    const long long n_rep = smart_batch_n_report(&d);
    if (h[6]) d.obs = (const double *)read_doubles(f, (size_t)n_rep);
    fclose(f);
    double *q = (double *)malloc(sizeof(double) * (size_t)(n_rep * N));
    double *sc = (double *)malloc(sizeof(double) * (size_t)N * SMART_N_SCORES);
    double *gw = (double *)malloc(sizeof(double) * (size_t)N);
    d.discharge = q;
    d.ld_discharge = N;
    d.scores = h[6] ? sc : NULL;
    d.gw = gw;
    int rc = smart_batch_run_host(&d, 64, 0);
    if (rc) {
        fprintf(stderr, "smart_batch_run_host -> %d: %s\n", rc, smart_last_error());
        return 1;
    }
    FILE *o = fopen(argv[2], "wb");
    if (!o) return 2;
    fwrite(q, sizeof(double), (size_t)(n_rep * N), o);
    fwrite(gw, sizeof(double), (size_t)N, o);
    if (h[6]) fwrite(sc, sizeof(double), (size_t)N * SMART_N_SCORES, o);
    fclose(o);
    printf("ran %lld members x %lld steps, %lld reporting steps\n", N, T + d.n_warmup, n_rep);
    return 0;
}
