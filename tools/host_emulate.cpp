// tools/host_emulate.cpp -- development aid: compile smart_step.cuh for the HOST (g++) to study
// the rounding behaviour of the step formulations against the golden fixtures without a GPU.
// Not part of the product and not an execution path of smartpy_b200 (which has no CPU fallback).
//
//   g++ -O2 -ffp-contract=off -DSMART_HOST_EMULATION -shared -fPIC -o build_exp/libemul.so tools/host_emulate.cpp
#include <cmath>
#include <cstring>
#define __device__
#define __forceinline__ inline
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline int __double2hiint(double x) { long long v; std::memcpy(&v, &x, 8); return (int)(v >> 32); }
static inline int __float_as_int(float x) { int v; std::memcpy(&v, &x, 4); return v; }
using std::fma;
using std::floor;
static inline unsigned __activemask() { return 1u; }
static inline int __any_sync(unsigned, int p) { return p; }
#include "../smartpy_b200/csrc/smart_step.cuh"

using namespace smart;

// mode 0: general, 1: fast (first form), 2: fast per-step, 3: block mode (one report per block of
// `gap` steps), 4: block-sub mode (blocks of `rep` steps, reports every `gap` steps inside them).
// Mirrors run_member/run_timeline of smart_kernels.cu for one member, summary or raw reporting.
template <typename R>
static int emulate(int mode, int rep, double area, double dt, long T, long W, const double *rain,
                   const double *peva, const double *par, int has_extra, double aar_ro,
                   const double *split, int report_type, int gap, double *discharge, double *gw_out)
{
    const double Tt = par[0], C = par[1], H = par[2], D = par[3], S = par[4], Z = par[5];
    const double SK = par[6], FK = par[7], GK = par[8], RK = par[9];
    MemberPar<R> p;
    p.Td = Tt; p.C = C; p.D = D; p.omD = 1.0 - D; p.Hz = H / Z; p.Sz = S / Z; p.z = Z / 6.0;
    p.r_sk = dt / (SK * 3600.0); p.r_fk = dt / (FK * 3600.0); p.r_gk = dt / (GK * 3600.0); p.r_rk = dt / (RK * 3600.0);
    R kc[7] = {R(C), R(D), R(1.0 - D), p.r_sk, p.r_fk, p.r_gk, p.r_rk};
    double kb[7];
    {
        const double cx[3] = {1.0 - p.r_sk, 1.0 - p.r_fk, 1.0 - p.r_gk}, rx[3] = {p.r_sk, p.r_fk, p.r_gk};
        const double cw = 1.0 - p.r_rk;
        double pw_w = 1.0;
        for (int h = 0; h < gap; ++h) pw_w *= cw;
        kb[3] = pw_w;
        for (int x = 0; x < 3; ++x) {
            double G = 0.0, pxh = 1.0;
            for (int h = 0; h < gap; ++h) { G = fma(cw, G, pxh); pxh *= cx[x]; }
            kb[x] = pxh; kb[4 + x] = rx[x] * G;
        }
    }
    FastPar<R> fp; fp.Hz = p.Hz; fp.Sz = p.Sz; fp.z = p.z;
    fp.c_sk = 1.0 - p.r_sk; fp.c_fk = 1.0 - p.r_fk; fp.c_gk = 1.0 - p.r_gk; fp.c_rk = 1.0 - p.r_rk;
    const double to_mm = 1e3 / area;
    double v[12];
    const double kk[5] = {SK, SK, FK, GK, GK};
    for (int k = 0; k < 5; ++k) v[k] = has_extra ? aar_ro * split[k] / 1000 * area / 8766 * kk[k] : 0.0;
    v[11] = has_extra ? aar_ro / 1000 * area / 8766 * RK : 0.0;
    for (int k = 0; k < 6; ++k) v[5 + k] = (Z / 12) / 1000 * area;
    MemberState<R> s;
    s.ove = v[0] * to_mm; s.dra = v[1] * to_mm; s.itf = v[2] * to_mm; s.sgw = v[3] * to_mm; s.dgw = v[4] * to_mm;
    for (int k = 0; k < 6; ++k) s.ly[k] = v[5 + k] * to_mm;
    s.riv = v[11] * to_mm;
    if (mode != 0) { s.ove += s.dra; s.sgw += s.dgw; s.dra = s.dgw = 0; }
    if (mode >= 2 && sizeof(R) == 4)      // binary32 fast form: the soil as deficits (fast_wet_soil_deficit)
        for (int k = 0; k < 6; ++k) s.ly[k] = (R)((double)p.z - v[5 + k] * to_mm);
    FastCarry<R> carry; carry.tot = carry.part = 0; carry.valid = false;
    StepOut<R> o;
    const R qscale = area / (1e3 * dt), mean_scale = area / (1e3 * dt) / (double)gap;
    long n_rep = report_type == 1 ? T / gap : (T + gap - 1) / gap;
    int countdown = 0x7fffffff; long r = 0;
    R acc = 0, agw = 0, aall = 0; double GN = 0, GD = 0, riv0 = s.riv;
    for (long seg = 0; seg < 2; ++seg) {
        long n = seg == 0 ? W : T;
        if (seg == 1) { countdown = (int)(T - (n_rep - 1) * gap); r = 0; acc = agw = aall = 0; GN = GD = 0; riv0 = s.riv; }
        if (mode == 4) {   // block-sub mode: kModeBlockSub of run_timeline
            const bool summary = report_type == 1 || gap == 1;
            if (seg == 1) countdown = gap;
            for (long day = 0; day < n / rep; ++day) {
                const double ex_d = __dsub_rn(__dmul_rn(rain[day * rep], Tt), peva[day * rep]);
                const BlockPar<R> bp = block_par<R, 1>(kc);
                const R r_rk = kc[6];
                const bool wet = ex_d >= 0.0;
                const R ex = wet ? ex_d : 0.0, hex = fp.Hz * ex;
                if (wet) {
                    if (!carry.valid) carry_form(s, carry);
                } else {
                    dry_block_soil<R>(s, kc[0], fp.z, ex_d, rep);
                    carry.valid = false;
                }
                for (int h = 0; h < rep; ++h) {
                    const R q_riv = s.riv * r_rk;
                    R q_gw, q_in;
                    if (wet) fast_wet_hour<R, 1>(s, fp, kc, bp, carry, ex, hex, 1u, q_gw, q_in);
                    else fast_dry_hour<R, 1>(s, fp, kc, bp, q_gw, q_in);
                    acc += q_riv;
                    if (summary) agw += q_gw;
                    if (--countdown == 0) {
                        countdown = gap;
                        const R sval = (summary ? acc : q_riv) * (summary ? mean_scale : qscale);
                        if (summary) aall += acc; else { agw += q_gw; aall += q_in; }
                        acc = 0;
                        if (seg == 1) discharge[r++] = sval;
                    }
                }
            }
            if (seg == 1) { *gw_out = summary ? agw / (aall + (s.riv - riv0)) : agw / aall; return 0; }
            riv0 = s.riv; acc = 0; agw = 0; aall = 0;
            continue;
        }
        if (mode == 3) {   // block mode: forcing constant inside aligned blocks of `gap` steps
            for (long day = 0; day < n / gap; ++day) {
                const double ex_d = __dsub_rn(__dmul_rn(rain[day * gap], Tt), peva[day * gap]);
                smart_block_fast<R, 1>(s, fp, kc, kb, carry, ex_d, gap, acc, agw);
                if (seg == 1) { discharge[r++] = acc * mean_scale; GD += acc; acc = 0; }
            }
            if (seg == 1) { *gw_out = agw / (GD + (s.riv - riv0)); return 0; }
            riv0 = s.riv; acc = 0; agw = 0;
            continue;
        }
        for (long i = 0; i < n; ++i) {
            if (mode == 0) smart_step<R, true, false>(s, p, rain[i], peva[i], o);
            else if (mode == 1) smart_step<R, false, false>(s, p, rain[i], peva[i], o);
            else smart_step_fast<R, 1>(s, fp, kc, carry, __dsub_rn(__dmul_rn(rain[i], Tt), peva[i]), o);
            acc += o.q_riv; agw += o.q_gw; aall += o.q_all;
            if (--countdown == 0) {
                countdown = gap;
                R sval;
                if (report_type == 1) sval = acc * mean_scale; else { sval = o.q_riv * qscale; agw = o.q_gw; aall = o.q_all; }
                GN += agw; GD += aall; acc = agw = aall = 0;
                discharge[r++] = sval;
            }
        }
    }
    *gw_out = GN / GD;
    return 0;
}

extern "C" int emulate_run(int mode, double area, double dt, long T, long W, const double *rain,
                           const double *peva, const double *par, int has_extra, double aar_ro,
                           const double *split, int report_type, int gap, double *discharge, double *gw_out)
{
    return emulate<double>(mode, gap, area, dt, T, W, rain, peva, par, has_extra, aar_ro, split, report_type, gap,
                           discharge, gw_out);
}

// binary32 state (the FP32 mode of the kernels; the GPU unit may contract a * b + c where this
// build does not, so this reproduces the size of the FP32 error, not its bits)
extern "C" int emulate_run_f32(int mode, double area, double dt, long T, long W, const double *rain,
                               const double *peva, const double *par, int has_extra, double aar_ro,
                               const double *split, int report_type, int gap, double *discharge, double *gw_out)
{
    return emulate<float>(mode, gap, area, dt, T, W, rain, peva, par, has_extra, aar_ro, split, report_type, gap,
                          discharge, gw_out);
}

// block-sub mode: forcing constant inside aligned blocks of `rep` steps, reports every `gap` steps
extern "C" int emulate_run_sub(int rep, double area, double dt, long T, long W, const double *rain,
                               const double *peva, const double *par, int has_extra, double aar_ro,
                               const double *split, int report_type, int gap, double *discharge, double *gw_out)
{
    return emulate<double>(4, rep, area, dt, T, W, rain, peva, par, has_extra, aar_ro, split, report_type, gap, discharge,
                           gw_out);
}
