// smart_step.cuh -- one SMART time step for one member, state held in registers.
//
// Restates smartpy/structure.py:267-503 of the reference (run_one_step_catchment +
// run_one_step_river, glued at :200-264) with these deliberate changes of FORM (results
// agree with the reference to 1e-12 .. 6e-12 relative on discharge, measured in
// profiles/r01_parity_report.txt; bar 1e-10, tests/test_gpu_parity.py):
//   * every store is kept in millimetres over the catchment (V / area * 1e3) instead of m3,
//     so the per-step m3<->mm conversions of :341-346, :424-457 disappear;
//   * per-member constants are hoisted: h' = (H/Z) * tot, s' = (S/Z) * tot, r_x = dt / k_x;
//   * s'^i by repeated multiplication instead of pow() (:382, :396), s'/i by reciprocal (:389);
//   * the dead output Q_aeva (:424) is not computed.
// The one true discontinuity of the model -- the wet/dry predicate `rain * T - peva >= 0`
// (:353-359) -- is evaluated exactly as the reference does: a rounded binary64 multiply, then
// a rounded subtract (no FMA contraction), in binary64 even in FP32 mode.
#pragma once
#ifndef SMART_HOST_EMULATION   // tools/host_emulate.cpp compiles this header with g++ to study rounding
#include <cuda_runtime.h>
#endif

namespace smart {

template <typename R>
struct MemberPar {
    double Td;            // T in binary64 for the wet/dry predicate
    R C, D, omD;          // C, D, (1 - D)
    R Hz, Sz;             // H / Z, S / Z
    R z;                  // Z / 6
    R r_sk, r_fk, r_gk, r_rk;   // dt / (k * 3600)
};

// mm-equivalent stores.  In the merged ("fast") form `ove` carries V_ove + V_dra and `sgw`
// carries V_sgw + V_dgw (same routing constant => same linear reservoir), `dra`/`dgw` unused.
template <typename R>
struct MemberState {
    R ly[6];
    R ove, dra, itf, sgw, dgw;
    R riv;
};

template <typename R>
struct StepOut {
    R q_riv;   // river outflow, mm per step
    R q_gw;    // Q_sgw + Q_dgw, mm per step
    R q_all;   // sum of the five pathway outflows, mm per step
    // only filled when kFluxes (the run_all_steps / run_one_step contract wants all 19 columns)
    R aeva, q_ove, q_dra, q_int, q_sgw, q_dgw;   // mm per step
};

// 1/(i+1) for the shallow-groundwater leak (structure.py:389).  The binary64 values that are not
// exact FP64 immediates live in constant memory so that they are operands of the multiply itself
// instead of being rebuilt with UMOV pairs every wet step.
#ifndef SMART_HOST_EMULATION
static __device__ __constant__ double smart_inv3 = 1.0 / 3.0, smart_inv5 = 0.2, smart_inv6 = 1.0 / 6.0;
#else
static const double smart_inv3 = 1.0 / 3.0, smart_inv5 = 0.2, smart_inv6 = 1.0 / 6.0;
#endif
template <typename R> __device__ __forceinline__ R inv_const(int i);
template <> __device__ __forceinline__ double inv_const<double>(int i)
{
    return i == 0 ? 1.0 : i == 1 ? 0.5 : i == 2 ? smart_inv3 : i == 3 ? 0.25 : i == 4 ? smart_inv5 : smart_inv6;
}
template <> __device__ __forceinline__ float inv_const<float>(int i)
{
    return i == 0 ? 1.0f : i == 1 ? 0.5f : i == 2 ? (1.0f / 3.0f) : i == 3 ? 0.25f : i == 4 ? 0.2f : (1.0f / 6.0f);
}

// x >= 0 for a freshly computed difference x = a - b of finite values, read from the sign bit on
// the integer pipe instead of an FP64 compare.  Exact: rounding never changes the sign of a
// difference, and a - b is +0 (sign clear) when a == b.
__device__ __forceinline__ bool sign_clear(double x) { return __double2hiint(x) >= 0; }
__device__ __forceinline__ bool sign_clear(float x) { return __float_as_int(x) >= 0; }

// The merged ("fast") form is exact only when no clamp, cap or leak predicate can fire:
// every routing constant >= dt, inflows >= 0 (0 <= D <= 1, 0 <= H < 1), s' < 1.
// par = T, C, H, D, S, Z, SK, FK, GK, RK (smartpy/parameters.py:25).
__device__ __forceinline__ bool fast_form_ok(const double *par, double dt)
{
    return par[6] * 3600.0 >= dt && par[7] * 3600.0 >= dt && par[8] * 3600.0 >= dt && par[9] * 3600.0 >= dt &&
           par[4] >= 0.0 && par[4] <= 0.5 && par[5] > 0.0 && par[3] >= 0.0 && par[3] <= 1.0 && par[2] >= 0.0 &&
           par[2] <= 0.99 && par[0] > 0.0;
}

// kGeneral = true : branch-faithful form -- five separate reservoirs with the >= 0 clamps
//                   (:427-450), the 95 % river cap (:492-498) and the `leak < level`
//                   predicates (:383, :390, :397).  Needed when dt > k for some store, when
//                   S/Z*tot can reach 1, or when the caller wants the 19-vector back.
// kGeneral = false: merged reservoirs, no clamps/cap/predicates -- valid exactly when none of
//                   them can fire (decided per CTA in the kernel from the parameters).
// kFluxes (needs kGeneral): also return actual evapotranspiration and the five pathway
//                   outflows of this step (structure.py:424-447), for the 19-vector.
template <typename R, bool kGeneral, bool kFluxes = false>
__device__ __forceinline__ void smart_step(MemberState<R> &s, const MemberPar<R> &p,
                                           double rain, double peva, StepOut<R> &o)
{
    const R zero = R(0);

    // structure.py:353-359 -- binary64, separately rounded
    const double rain_c = __dmul_rn(rain, p.Td);
    const double ex_d = __dsub_rn(rain_c, peva);

    R in_ove = zero, in_dra = zero, in_int = zero, in_sgw = zero, in_dgw = zero;

    if (ex_d >= 0.0) {
        // ---- wet branch, structure.py:360-399
        R ex = static_cast<R>(ex_d);
        if (kFluxes) o.aeva = static_cast<R>(peva);   // :361
        // structure.py:350 -- left-to-right sum of the six layer levels (only the wet branch reads it)
        const R tot = ((((s.ly[0] + s.ly[1]) + s.ly[2]) + s.ly[3]) + s.ly[4]) + s.ly[5];
        in_ove = (p.Hz * tot) * ex;                 // :363-364
        ex = ex - in_ove;                           // :365
        auto fill = [&](int i) {                    // :367-374, `ex <= space` as the sign of space - ex
            const R space = p.z - s.ly[i];
            const R t = space - ex;
            const bool fits = sign_clear(t);
            s.ly[i] = fits ? s.ly[i] + ex : p.z;
            ex = fits ? zero : -t;                  // -(space - ex) is exactly ex - space
            return fits;
        };
        // (no early-out here: with caller-supplied initial states a lower layer may be over-full,
        // and the reference then moves water down even when nothing arrives from above)
        fill(0); fill(1); fill(2); fill(3); fill(4); fill(5);
        in_dra = p.D * ex;                          // :376
        in_int = p.omD * ex;                        // :377
        const R sp = p.Sz * tot;                    // :379 (start-of-step total)
        R pw[6];
        pw[0] = sp;
        pw[1] = sp * sp;
        pw[2] = pw[1] * sp;
        pw[3] = pw[1] * pw[1];
        pw[4] = pw[3] * sp;
        pw[5] = pw[2] * pw[2];
#pragma unroll
        for (int i = 0; i < 6; ++i) {               // :381-385 interflow leak, top down
            const R leak = s.ly[i] * pw[i];
            if (!kGeneral || leak < s.ly[i]) {
                in_int += leak;
                s.ly[i] -= leak;
            }
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {               // :388-392 shallow groundwater leak
            const R leak = s.ly[i] * (i == 0 ? sp : sp * inv_const<R>(i));
            if (!kGeneral || leak < s.ly[i]) {
                in_sgw += leak;
                s.ly[i] -= leak;
            }
        }
#pragma unroll
        for (int i = 5; i >= 0; --i) {              // :395-399 deep groundwater leak, bottom up
            const R leak = s.ly[i] * pw[5 - i];
            if (!kGeneral || leak < s.ly[i]) {
                in_dgw += leak;
                s.ly[i] -= leak;
            }
        }
    } else {
        // ---- dry branch, structure.py:400-419 (aeva is dead downstream and not tracked)
        R d = static_cast<R>(-ex_d);
        R aeva = static_cast<R>(rain_c);            // :408
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const R t = s.ly[i] - d;
            const bool enough = sign_clear(t);      // level >= deficit
            if (kFluxes) aeva += enough ? d : s.ly[i];   // :412, :415
            s.ly[i] = enough ? t : zero;
            d = enough ? zero : p.C * (-t);
        }
        if (kFluxes) o.aeva = aeva;
    }

    if (kGeneral) {
        // structure.py:427-450 -- outflow from the OLD storage, then update, then clamp
        const R q_ove = s.ove * p.r_sk;
        s.ove = s.ove + (in_ove - q_ove);
        if (!sign_clear(s.ove)) s.ove = zero;
        const R q_dra = s.dra * p.r_sk;
        s.dra = s.dra + (in_dra - q_dra);
        if (!sign_clear(s.dra)) s.dra = zero;
        const R q_int = s.itf * p.r_fk;
        s.itf = s.itf + (in_int - q_int);
        if (!sign_clear(s.itf)) s.itf = zero;
        const R q_sgw = s.sgw * p.r_gk;
        s.sgw = s.sgw + (in_sgw - q_sgw);
        if (!sign_clear(s.sgw)) s.sgw = zero;
        const R q_dgw = s.dgw * p.r_gk;
        s.dgw = s.dgw + (in_dgw - q_dgw);
        if (!sign_clear(s.dgw)) s.dgw = zero;
        o.q_gw = q_sgw + q_dgw;
        if (kFluxes) {
            o.q_ove = q_ove;
            o.q_dra = q_dra;
            o.q_int = q_int;
            o.q_sgw = q_sgw;
            o.q_dgw = q_dgw;
        }
        const R q_in = (((q_ove + q_dra) + q_int) + q_sgw) + q_dgw;    // :254
        o.q_all = q_in;
        // structure.py:482-498
        R q = s.riv * p.r_rk;
        const R tmp = s.riv + (q_in - q);
        if (!sign_clear(tmp)) {
            q = R(0.95) * (q_in + s.riv);
            s.riv = s.riv + (q_in - q);
        } else {
            s.riv = tmp;
        }
        o.q_riv = q;
    } else {
        const R q_quick = s.ove * p.r_sk;
        s.ove = (s.ove - q_quick) + (in_ove + in_dra);
        const R q_int = s.itf * p.r_fk;
        s.itf = (s.itf - q_int) + in_int;
        const R q_gw = s.sgw * p.r_gk;
        s.sgw = (s.sgw - q_gw) + (in_sgw + in_dgw);
        o.q_gw = q_gw;
        const R q_in = (q_quick + q_int) + q_gw;
        o.q_all = q_in;
        const R q = s.riv * p.r_rk;
        s.riv = (s.riv - q) + q_in;
        o.q_riv = q;
    }
}

// ------------------------------------------------------------------------------------------
// The merged ("fast") step: same model, fewer FP64-pipe instructions.  Valid under the same
// conditions as kGeneral = false above (decided per CTA from the parameters).  Differences of
// form with respect to smart_step<R, false>:
//   * the soil total is only formed on wet steps (the dry branch never reads it) and is carried
//     over from the end of the previous wet step when no dry step intervened (it is the same
//     left-to-right sum of the same six values, so the carried number is bit-identical);
//   * fill ladder in two additions per layer: w = level + excess, t = z - w; the layer takes
//     everything iff t >= 0, else it is full and -t flows on (structure.py:367-374);
//   * `x >= 0` tests on freshly computed differences read the sign bit (integer pipe) instead
//     of issuing an FP64 compare; the wet/dry predicate itself stays an exact FP64 compare;
//   * binary64 only: the three leak passes (structure.py:381-399) update each layer with one
//     FMA per pass and obtain the interflow and groundwater inflows as differences of the soil
//     total before/after (mass conserving by construction; absolute error of a few ulp of the
//     soil total, i.e. ~1e-14 mm per wet step);
//   * reservoirs advance with one FMA: V' = V * (1 - dt/k) + inflow, Q = V * (dt/k).
// Registers hold what the end of the step needs (1 - dt/k, on the critical path of the store
// updates); shared memory (kc[], one column per thread) holds what can be fetched early:
//   kc[0] = C, kc[1] = D, kc[2] = 1 - D, kc[3..6] = dt/k for SK, FK, GK, RK.
template <typename R>
struct FastPar {
    R Hz, Sz, z;                  // H / Z, S / Z, Z / 6
    R c_sk, c_fk, c_gk, c_rk;     // 1 - dt / (k * 3600)   (binary64 form only)
};

template <typename R>
struct FastCarry {
    R tot;          // soil total at the end of the previous step, valid iff that step was wet
    R part;         // binary64 only: the part of it that layers 2..6 hold (soil_lower), valid with tot
    bool valid;
};

// Soil total of the fast form.  Binary64: from the bottom layer up, the top layer LAST -- on most
// wet hours the rain only reaches the top layer, and the total after the fill is then the carried
// sum of the other five plus the new top level: one addition instead of five, and the very number
// a full re-summation gives (so that a zero leak still comes out as exactly zero).  The
// branch-faithful step keeps the reference's order (structure.py:350); the two orders differ by an
// ulp of the total.  Binary32 (sum of deficits): plain left to right, as before.
template <typename R>
__device__ __forceinline__ R soil_lower(const MemberState<R> &s)
{
    return (((s.ly[5] + s.ly[4]) + s.ly[3]) + s.ly[2]) + s.ly[1];
}
template <typename R>
__device__ __forceinline__ R soil_total(const MemberState<R> &s)
{
    if constexpr (sizeof(R) == 8) return soil_lower(s) + s.ly[0];
    else return ((((s.ly[0] + s.ly[1]) + s.ly[2]) + s.ly[3]) + s.ly[4]) + s.ly[5];
}
// (Re)form the carried sums from the state: the same numbers the carry holds after a wet step.
template <typename R>
__device__ __forceinline__ void carry_form(const MemberState<R> &s, FastCarry<R> &c)
{
    if constexpr (sizeof(R) == 8) {
        c.part = soil_lower(s);
        c.tot = c.part + s.ly[0];
    } else {
        c.tot = soil_total(s);
    }
    c.valid = true;
}


// ---- pieces of the fast step ---------------------------------------------------------------

// Wet hour, soil part (structure.py:360-399): overland split, fill ladder, saturation excess,
// three leak passes.  ex = rain * T - peva >= 0.  Returns the three merged inflows.
// hex = (H / Z) * ex, formed by the caller (once per block when the forcing is block-constant).
// kTotKnown: the caller guarantees carry.valid (block mode forms the total once per wet block), so
// the hour starts from carry.tot with no test.  Otherwise the total is re-formed behind a
// warp-uniform branch: left to predication, the five additions would be issued -- and waited for,
// the scoreboard does not look at predicates -- in every wet hour of every member.
// mask: the lanes of the warp that are in this call together (__activemask() taken by the caller
// where it is known to be stable, e.g. once per wet block).
// ---- binary32 state: the soil as DEFICITS ----------------------------------------------------
// In binary32 the six soil values hold d = z - level, how far each layer is from full, not the
// level.  Why: a wet column sits at or near capacity (18 mm per layer for Z = 109), the hourly
// leaks are 1e-3 mm, and level - leak rounds to the level's ulp (2e-6 mm): the water the soil
// actually gives up each hour differs from the leak by up to 5e-4 of it, and under constant forcing
// in a saturated column it is the SAME rounding every hour -- a bias, not noise (measured on an
// LHS sample of 1e5 members: |dNSE| up to 2.3e-5 against the binary64 kernel, 157 members above the
// 1e-5 bar; profiles/r02_fp32_*.json).  The deficit of a wet column is itself of the size of the
// leaks, so binary32 resolves it to 1e-10 mm; in a dry column the deficit is large and coarse
// (2e-6 mm) but so is every flux that touches it.  Water balance then holds to 1e-7 relative in
// both regimes, with the same number of operations: the fill ladder needs ONE addition per layer
// (t = d - water: the layer takes everything iff t >= 0), the three leak passes nine per layer
// as before.  carry.tot holds the SUM OF DEFICITS (soil total = 6 z - that).
template <int kStride, bool kTotKnown>
__device__ __forceinline__ void fast_wet_soil_deficit(MemberState<float> &s, const FastPar<float> &p, float D, float omD,
                                                      FastCarry<float> &carry, float ex, float hex, unsigned mask,
                                                      float &in_quick, float &in_int, float &in_gw)
{
    float sd = carry.tot;
    if (!kTotKnown) {
        if (__any_sync(mask, !carry.valid)) {
            if (!carry.valid) sd = soil_total(s);
        }
    }
    const float tot = fma(6.0f, p.z, -sd);          // structure.py:350
    in_quick = hex * tot;                           // :363-364
    const float u0 = fma(hex, tot, -ex);            // u = -(excess rain still to place) <= 0
    float u = u0;
    auto fill = [&](int i) {                        // :367-374: room = the deficit itself
        const float t = s.ly[i] + u;
        const bool fits = sign_clear(t);
        s.ly[i] = fits ? t : 0.0f;
        u = fits ? 0.0f : t;
        return fits;
    };
    // most often the first layer takes everything; a lane's own branch (round 1 voted over the warp:
    // four more instructions per hour, and the lanes that are done would run the ladder as no-ops),
    // the top layer written out so that the usual path issues no selects
    const float t0 = s.ly[0] + u;
    if (sign_clear(t0)) {
        s.ly[0] = t0;
        u = 0.0f;
    } else {
        s.ly[0] = 0.0f;
        u = t0;
        fill(1); fill(2); fill(3); fill(4); fill(5);
        in_quick = fma(D, -u, in_quick);            // + D * saturation excess (:376)
    }
    const float sp = p.Sz * tot;                    // :379
    float pw[6];
    pw[0] = sp;
    pw[1] = sp * sp;
    pw[2] = pw[1] * sp;
    pw[3] = pw[1] * pw[1];
    pw[4] = pw[3] * sp;
    pw[5] = pw[2] * pw[2];
    // the three passes act on the LEVEL z - d of each layer (:381-399); what they take out together
    // is added to the deficit, and the same numbers -- no second rounding -- go to the stores
    float leak_int = 0.0f, leak_all = 0.0f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float f2 = i == 0 ? sp : sp * inv_const<float>(i);
        const float l0 = p.z - s.ly[i];
        const float l1 = fma(-l0, pw[i], l0);                       // after the interflow pass
        const float l2 = fma(-l1, f2, l1);                          // after the shallow groundwater pass
        leak_int = fma(l0, pw[i], leak_int);
        const float out = fma(l2, pw[5 - i], fma(l1, f2, l0 * pw[i]));   // all three passes
        s.ly[i] += out;
        leak_all += out;
    }
    in_int = fma(omD, -u, leak_int);                // (1 - D) * saturation excess + interflow (:377, :381-385)
    in_gw = leak_all - leak_int;
    // sum of deficits after this hour: minus what the fill placed (u - u0 >= 0), plus the leaks
    carry.tot = ((sd + u0) - u) + leak_all;
    carry.valid = true;
}

// kNestedLadder: one early exit per layer of the fill ladder instead of one for the lower five.
// Measured on B200 (profiles/r02_variant_sweep.txt): +1.7 % for single-catchment block mode (C3), -4 % for
// the one-warp CTAs of multi-catchment batches (C4b), -1 % per step -- so only the first uses it.
template <typename R, int kStride, bool kTotKnown = false, bool kNestedLadder = false>
__device__ __forceinline__ void fast_wet_soil(MemberState<R> &s, const FastPar<R> &p, R D, R omD,
                                              FastCarry<R> &carry, R ex, R hex, unsigned mask, R &in_quick, R &in_int,
                                              R &in_gw)
{
    if constexpr (sizeof(R) == 4) {
        fast_wet_soil_deficit<kStride, kTotKnown>(s, p, D, omD, carry, ex, hex, mask, in_quick, in_int, in_gw);
        return;
    }
    const R zero = R(0);
    R tot = carry.tot, lower = carry.part;
    if (!kTotKnown) {
        if (__any_sync(mask, !carry.valid)) {
            if (!carry.valid) {
                lower = soil_lower(s);
                tot = lower + s.ly[0];
            }
        }
    }
    in_quick = hex * tot;                           // :363-364, h' * excess = (H/Z * excess) * total
    const R u0 = fma(hex, tot, -ex);                // u = -(excess rain still to place) <= 0
    R u = u0;
    auto fill = [&](int i) {                        // :367-374
        const R w = s.ly[i] - u;                    // level if the layer took everything
        const R t = p.z - w;
        const bool fits = sign_clear(t);
        s.ly[i] = fits ? w : p.z;
        u = fits ? zero : t;
        return fits;
    };
    // most often the first layer takes everything; a lane's own branches (no vote, no mask to carry).
    // Saturation excess (-u after the ladder, :376-377) only exists on that path, so its two
    // products are formed there and the usual hour does not issue them.
    R sat_int = zero;
    // (the top layer written out as a branch: on the usual path the level is w and nothing flows on,
    // with no selects to issue)
    const R w0 = s.ly[0] - u;
    const R t0 = p.z - w0;
    if (sign_clear(t0)) {
        s.ly[0] = w0;
        u = zero;
    } else {
        s.ly[0] = p.z;
        u = t0;
        if (kNestedLadder) {
            // one branch per layer: the water stops at the first layer with room
            if (!fill(1)) {
                if (!fill(2)) {
                    if (!fill(3)) {
                        if (!fill(4)) fill(5);
                    }
                }
            }
        } else {
            fill(1); fill(2); fill(3); fill(4); fill(5);
        }
        lower = soil_lower(s);                      // the rain went below the top layer: re-sum the lower five
        in_quick = fma(D, -u, in_quick);            // + D * saturation excess (:376)
        sat_int = omD * (-u);                       // (1 - D) * saturation excess (:377)
    }
    const R sp = p.Sz * tot;                        // :379
    R pw[6];
    pw[0] = sp;
    pw[1] = sp * sp;
    pw[2] = pw[1] * sp;
    pw[3] = pw[1] * pw[1];
    pw[4] = pw[3] * sp;
    pw[5] = pw[2] * pw[2];
    {
        // soil total after the fill: the lower five layers as carried (or re-summed above when the
        // rain reached them) plus the new top level -- the number soil_total(s) would give
        const R tot_f = lower + s.ly[0];
#pragma unroll
        for (int i = 0; i < 6; ++i) s.ly[i] = fma(-s.ly[i], pw[i], s.ly[i]);            // :381-385
        const R tot1 = soil_total(s);
        in_int = sat_int + (tot_f - tot1);                                              // :377 + interflow
#pragma unroll
        for (int i = 0; i < 6; ++i) {                                                   // :388-399
            const R f2 = i == 0 ? sp : sp * inv_const<R>(i);
            s.ly[i] = fma(-s.ly[i], f2, s.ly[i]);
        }
#pragma unroll
        for (int i = 5; i >= 0; --i) s.ly[i] = fma(-s.ly[i], pw[5 - i], s.ly[i]);
        const R lower3 = soil_lower(s);
        const R tot3 = lower3 + s.ly[0];
        in_gw = tot1 - tot3;
        carry.tot = tot3;
        carry.part = lower3;
        carry.valid = true;
    }
}

// Dry hour, soil part (structure.py:407-419): the deficit d = peva - rain * T > 0 is taken from
// the layers top down, decayed by C each time a layer runs empty.
template <typename R>
__device__ __forceinline__ void fast_dry_soil(MemberState<R> &s, R C, R d, R z)
{
    const R zero = R(0);
    auto take = [&](int i) {
        if constexpr (sizeof(R) == 8) {
            const R t = s.ly[i] - d;
            const bool enough = sign_clear(t);      // level >= deficit
            s.ly[i] = enough ? t : zero;
            d = enough ? zero : C * (-t);
            return enough;
        } else {                                    // binary32: s.ly holds z - level (see fast_wet_soil_deficit)
            const R t = (z - s.ly[i]) - d;
            const bool enough = sign_clear(t);
            s.ly[i] = enough ? s.ly[i] + d : z;
            d = enough ? zero : C * (-t);
            return enough;
        }
    };
    if (!take(0)) {                                 // (a lane's own branch, see fast_wet_soil)
        take(1); take(2); take(3); take(4); take(5);
    }
}

// One hour of the fast step.  ex_d = rain * T - peva (structure.py:353-355, two separately
// rounded binary64 operations) is formed by the caller one step ahead: it depends on the
// forcing only, which takes the shared-memory load and two FP64 latencies off the head of
// every step.
template <typename R, int kStride>
__device__ __forceinline__ void smart_step_fast(MemberState<R> &s, const FastPar<R> &p, const R *kc,
                                                FastCarry<R> &carry, double ex_d, StepOut<R> &o)
{
    constexpr bool kOneFma = sizeof(R) == 8;
    const R zero = R(0);
    R in_quick = zero, in_int = zero, in_gw = zero;

    // Outflows leave from the OLD storage (structure.py:427-447, :487), so they and the whole
    // river update do not wait for the soil: issue them first.  SK/SK and GK/GK stores merged.
    const R r_sk = kc[3 * kStride], r_fk = kc[4 * kStride];
    const R q_gw = s.sgw * kc[5 * kStride];
    const R q = s.riv * kc[6 * kStride];
    const R q_in = fma(r_sk, s.ove, fma(r_fk, s.itf, q_gw));            // :254 as one fused dot product
    s.riv = kOneFma ? fma(s.riv, p.c_rk, q_in) : (s.riv - q) + q_in;   // :487-498, cap cannot fire
    o.q_gw = q_gw;
    o.q_all = q_in;
    o.q_riv = q;

    if (ex_d >= 0.0) {
        const R ex = static_cast<R>(ex_d);
        fast_wet_soil<R, kStride>(s, p, kc[1 * kStride], kc[2 * kStride], carry, ex, p.Hz * ex, __activemask(), in_quick,
                                  in_int, in_gw);
    } else {
        fast_dry_soil<R>(s, kc[0], static_cast<R>(-ex_d), p.z);
        carry.valid = false;
    }

    // binary64 advances a store with one FMA, V' = V * (1 - dt/k) + inflow; in binary32 the
    // rounding of (1 - dt/k) would bias the recession constant by up to 1e-4, so the two-step
    // form V' = (V - Q) + inflow is kept there.
    s.ove = kOneFma ? fma(s.ove, p.c_sk, in_quick) : (s.ove - s.ove * r_sk) + in_quick;
    s.itf = kOneFma ? fma(s.itf, p.c_fk, in_int) : (s.itf - s.itf * r_fk) + in_int;
    s.sgw = kOneFma ? fma(s.sgw, p.c_gk, in_gw) : (s.sgw - q_gw) + in_gw;
}

// ---- a whole block of `rep` steps under constant forcing -----------------------------------
// The reference's own daily -> hourly disaggregation (timeframe.py:167-186) gives every hour of
// a day the same rain and PET, so the wet/dry predicate holds for the whole block.
//  * wet block: `rep` hourly steps as above (the soil is nonlinear in its own state);
//  * dry block: the soil ladder over the whole block in closed-form phases (dry_block_soil); the
//    stores receive nothing for `rep` steps, so they, the river and the two running sums evolve
//    LINEARLY.  When one block is one reporting step they are advanced in closed form:
//        V_x(rep) = c_x^rep V_x,    W(rep) = c_w^rep W + sum_x K_x V_x,
//        K_x = r_x * sum_{h<rep} c_w^(rep-1-h) c_x^h   (Horner sum at set-up, no cancellation),
//    and what left each store gives the sums by mass balance: sum Q_gw = G - G(rep),
//    sum Q_out = (W - W(rep)) + sum_x (V_x - V_x(rep))  (smart_block_fast).  When the run reports
//    inside the block (report_gap < rep, or 'raw') the stores are walked hour by hour with the
//    hour's own arithmetic (fast_dry_hour: 7 instructions, no soil, no forcing) and every hour's
//    outflow is available to the reporting code (run_timeline, block-sub mode).
// kb[] (binary64 in both precisions, one column per thread): c_sk^rep, c_fk^rep, c_gk^rep,
// c_rk^rep, K_sk, K_fk, K_gk.  The closed forms are evaluated in binary64 also for binary32
// state: ~40 instructions per BLOCK, and binary32 powers of (1 - dt/k) would bias the recession.
#ifndef SMART_WET_UNROLL
#define SMART_WET_UNROLL 4
#endif
constexpr int kWetUnroll = SMART_WET_UNROLL;   // unroll factor of the wet-block hour loop

// Soil over a whole dry block, per member (a member's arithmetic never depends on the other
// lanes of its warp).  Nothing refills the layers, so each one runs empty at most once: while
// layers 0..k-1 are empty the demand reaching layer k is C^k d0 per step (:418).  A layer that is
// already empty only passes the demand on, decayed by C; otherwise it serves floor(level /
// demand) whole steps in one multiplication, then one ordinary ladder step (from k down) empties
// it and the demand decays by C again.  Always binary64 arithmetic, whatever R is.
template <typename R>
__device__ __forceinline__ void dry_block_soil(MemberState<R> &s, R Cpar, R zpar, double ex_d, int rep)
{
    constexpr bool kDeficits = sizeof(R) == 4;   // binary32 state keeps z - level (fast_wet_soil_deficit)
    const double z = static_cast<double>(zpar);
    double left = static_cast<double>(rep);      // steps of the block still to account for
    double dem = -ex_d;                          // demand arriving at layer k in each of them
    const double C = static_cast<double>(Cpar);
    double ly[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) ly[k] = kDeficits ? z - static_cast<double>(s.ly[k]) : static_cast<double>(s.ly[k]);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        {
            if (left > 0.0) {                    // (a lane's own branch: lanes that are done skip the layer)
                bool moves_on = true;            // the demand reaches layer k + 1
                if (ly[k] > 0.0) {
                    // Level if the layer served every remaining step.  Not negative <=> the exact
                    // quotient level / demand is >= left <=> floor(level / dem) >= left (left is a
                    // whole number, rounding is monotonic): the usual case, decided without dividing.
                    const double served = fma(-left, dem, ly[k]);
                    if (sign_clear(served)) {
                        ly[k] = served;
                        left = 0.0;
                        moves_on = false;
                    } else {
                        // whole steps this layer can serve before it runs empty.  The rounded
                        // quotient can land on the next whole number when the exact one lies just
                        // below it; the level it would leave is then (slightly) negative: one step
                        // fewer in that case, so that a layer never goes below zero.
                        double can = floor(ly[k] / dem);
                        if (!sign_clear(fma(-can, dem, ly[k]))) can -= 1.0;
                        const double n_full = can < left ? can : left;
                        ly[k] = fma(-n_full, dem, ly[k]);
                        left -= n_full;
                        if (left > 0.0) {        // transition step: layer k cannot meet the demand
                            double d = dem;
#pragma unroll
                            for (int j = k; j < 6; ++j) {
                                const double t = ly[j] - d;
                                const bool enough = sign_clear(t);
                                ly[j] = enough ? t : 0.0;
                                d = enough ? 0.0 : C * (-t);
                            }
                            left -= 1.0;
                        } else {
                            moves_on = false;
                        }
                    }
                }
                if (moves_on) dem *= C;
            }
        }
    }
    // (after layer 5 nothing is left to take: remaining steps of the block change nothing)
#pragma unroll
    for (int k = 0; k < 6; ++k) s.ly[k] = static_cast<R>(kDeficits ? z - ly[k] : ly[k]);
}

// Dry block with one report at its end: soil as above, stores + river + the two running sums by
// the linear recurrences.  acc += sum of river outflow over the block, agw += sum of groundwater
// outflow (mm).
template <typename R, int kStride>
__device__ __forceinline__ void dry_block_fast(MemberState<R> &s, const R *kc, const double *kb, R zpar, double ex_d,
                                               int rep, R &acc, R &agw)
{
    dry_block_soil<R>(s, kc[0], zpar, ex_d, rep);
    const double a = static_cast<double>(s.ove), b = static_cast<double>(s.itf);
    const double g = static_cast<double>(s.sgw), w = static_cast<double>(s.riv);
    const double a_n = a * kb[0 * kStride], b_n = b * kb[1 * kStride], g_n = g * kb[2 * kStride];
    const double w_n = fma(kb[4 * kStride], a, fma(kb[5 * kStride], b, fma(kb[6 * kStride], g, w * kb[3 * kStride])));
    const double out_g = g - g_n;
    agw += static_cast<R>(out_g);
    acc += static_cast<R>(((w - w_n) + (a - a_n)) + ((b - b_n) + out_g));
    s.ove = static_cast<R>(a_n);
    s.itf = static_cast<R>(b_n);
    s.sgw = static_cast<R>(g_n);
    s.riv = static_cast<R>(w_n);
}

// Routing constants of a block, fetched once from the per-thread shared-memory column.
template <typename R>
struct BlockPar {
    R D, omD, r_sk, r_fk, r_gk;     // (dt / k of the river stays in shared memory: kc[6], read where it is used)
};
template <typename R, int kStride>
__device__ __forceinline__ BlockPar<R> block_par(const R *kc)
{
    BlockPar<R> b;
    b.D = kc[1 * kStride];
    b.omD = kc[2 * kStride];
    b.r_sk = kc[3 * kStride];
    b.r_fk = kc[4 * kStride];
    b.r_gk = kc[5 * kStride];
    return b;
}

// One wet hour of a block: outflows from the OLD storage (:427-447) -- Q_gw alone (groundwater
// share) and the river inflow as one fused dot product r_sk V_quick + r_fk V_int + Q_gw (:254) --
// river, soil, stores.  The caller reads the river store BEFORE the call (its outflow is
// r_rk * that value).  carry.tot must be valid (kTotKnown).
template <typename R, int kStride, bool kNestedLadder = false>
__device__ __forceinline__ void fast_wet_hour(MemberState<R> &s, const FastPar<R> &p, const R *kc, const BlockPar<R> &b,
                                              FastCarry<R> &carry, R ex, R hex, unsigned mask, R &q_gw, R &q_in)
{
    constexpr bool kOneFma = sizeof(R) == 8;
    q_gw = s.sgw * b.r_gk;
    q_in = fma(b.r_sk, s.ove, fma(b.r_fk, s.itf, q_gw));
    s.riv = kOneFma ? fma(s.riv, p.c_rk, q_in) : (s.riv - s.riv * kc[6 * kStride]) + q_in;
    R in_quick, in_int, in_gw;
    fast_wet_soil<R, kStride, true, kNestedLadder>(s, p, b.D, b.omD, carry, ex, hex, mask, in_quick, in_int, in_gw);
    s.ove = kOneFma ? fma(s.ove, p.c_sk, in_quick) : (s.ove - s.ove * b.r_sk) + in_quick;
    s.itf = kOneFma ? fma(s.itf, p.c_fk, in_int) : (s.itf - s.itf * b.r_fk) + in_int;
    s.sgw = kOneFma ? fma(s.sgw, p.c_gk, in_gw) : (s.sgw - q_gw) + in_gw;
}

// One dry hour of a block, routing only (the soil of the whole block is dry_block_soil's): the
// same arithmetic as the hour above with zero inflows (fma(V, c, 0) == V * c).
template <typename R, int kStride>
__device__ __forceinline__ void fast_dry_hour(MemberState<R> &s, const FastPar<R> &p, const R *kc, const BlockPar<R> &b,
                                              R &q_gw, R &q_in)
{
    constexpr bool kOneFma = sizeof(R) == 8;
    q_gw = s.sgw * b.r_gk;
    q_in = fma(b.r_sk, s.ove, fma(b.r_fk, s.itf, q_gw));
    s.riv = kOneFma ? fma(s.riv, p.c_rk, q_in) : (s.riv - s.riv * kc[6 * kStride]) + q_in;
    s.ove = kOneFma ? s.ove * p.c_sk : s.ove - s.ove * b.r_sk;
    s.itf = kOneFma ? s.itf * p.c_fk : s.itf - s.itf * b.r_fk;
    s.sgw = kOneFma ? s.sgw * p.c_gk : s.sgw - q_gw;
}

template <typename R, int kStride, bool kNestedLadder = false>
__device__ __forceinline__ void smart_block_fast(MemberState<R> &s, const FastPar<R> &p, const R *kc, const double *kb,
                                                 FastCarry<R> &carry, double ex_d, int rep, R &acc, R &agw)
{
    constexpr bool kOneFma = sizeof(R) == 8;
    if (ex_d >= 0.0) {
        const R ex = static_cast<R>(ex_d);
        const BlockPar<R> b = block_par<R, kStride>(kc);
        const R hex = p.Hz * ex;                    // constant over the block
        const unsigned mask = __activemask();       // the lanes walking this wet block together
        if (!carry.valid) carry_form(s, carry);   // soil total (binary32: sum of deficits) once per block, then carried hour to hour
        R sum_riv = R(0);                           // river outflow of the block = r_rk * sum of the store
#pragma unroll kWetUnroll
        for (int h = 0; h < rep; ++h) {
            R q_gw, q_in;
            sum_riv += s.riv;
            fast_wet_hour<R, kStride, kNestedLadder>(s, p, kc, b, carry, ex, hex, mask, q_gw, q_in);
            agw += q_gw;
        }
        acc = kOneFma ? fma(sum_riv, kc[6 * kStride], acc) : acc + sum_riv * kc[6 * kStride];
    } else {
        dry_block_fast<R, kStride>(s, kc, kb, p.z, ex_d, rep, acc, agw);
        carry.valid = false;
    }
}

}  // namespace smart
