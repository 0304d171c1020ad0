/*
 * emc_physics.cuh — the 6-DOF flight physics of the engine, written for one trajectory per
 * thread with all state in registers.  Everything here is `__host__ __device__` so the SAME code
 * can also be compiled by g++ into the test seam (tests/hostseam) and checked against the oracle
 * in a container without a GPU.  The shipped library (libemc.so) only ever runs it on the device.
 *
 * What it mirrors (file:line under rocket_simulation/ of the reference):
 *   derivative            simulator.py:295-460
 *   rail phase            simulator.py:42-125
 *   RK4 step + events     simulator.py:209-264
 *   summary               simulator.py:474-494,579-582 (+ per-state diagnostics :511-552)
 *   atmosphere/gravity    environment.py:26-108        wind interp environment.py:267-276
 *   mass / aero           rocket.py:105-218            thrust / mdot motor.py:54-93,152-169
 *   quaternion helpers    utils.py:76-121,139-205
 *
 * It is NOT a transliteration.  Per-run constants are folded on the host (DevModel), table slopes
 * are precomputed, reciprocals are shared, sin/cos of the aerodynamic angles come from velocity
 * ratios instead of sincos(atan2()), the layered pressure law is evaluated as per-layer polynomials in
 * altitude (exp(e*log()) only outside them), the Cd and CP tables share one Mach search, and the four
 * RK4 stages share one copy of the derivative code.  These changes
 * move results by a few ulp per evaluation; the parity bar is 1e-6 relative per flight (tests/).
 * Python/NumPy NaN semantics of max()/min()/np.interp/np.argmax are kept where they decide control
 * flow.
 */
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/emc.h"

#if defined(__CUDACC__)
#define EMC_HD __device__ __forceinline__          /* libemc.so never runs the physics on the host */
#define EMC_KDECL static __constant__ double       /* constant bank: used as instruction operands, no register */
#else
#define EMC_HD inline
#define EMC_KDECL static const double
#endif
#define EMC_LIKELY(x) __builtin_expect(!!(x), 1)
/* rarely executed helper kept OUT of line: ptxas lays a rarely taken block where it stands in the program, so the common
 * path would otherwise take a branch across it every time (an instruction-fetch bubble); as a call, the block shrinks to a
 * predicated call instruction */
#if defined(__CUDACC__)
#define EMC_COLD __device__ __noinline__
#else
#define EMC_COLD EMC_HD
#endif

namespace emc {

/* Minimax coefficients (tools/fit_minimax.py).  On the device they sit in the constant bank so that
 * every DFMA reads its coefficient as an operand instead of materialising a 64-bit immediate with two
 * moves (which cost 13 % of all issued instructions in the first version, profiles/). */
EMC_KDECL K_ATAN[19] = {   /* atan(t) = t + t*u*Q(u), u = t^2 in [0,1]; degree 18, 2.8e-17 */
    -0.3333333333333186, 0.19999999999755017, -0.14285714271334715, 0.11111110678746688, -0.0909090123535805,
    0.07692212930161942, -0.0666586036040998, 0.058773077574447684, -0.05239232950492257, 0.04673949464358622,
    -0.04092637904494918, 0.03406780544224133, -0.02582678957209261, 0.01697802875462528, -0.009184554046120928,
    0.0038559722045492582, -0.0011640707910558462, 0.00022302218576995646, -2.025853092076122e-05 };
EMC_KDECL K_EXP[11] = {    /* exp(r) = 1 + r + r^2 Q(r), |r| <= ln2/2; degree 10, 6e-20 */
    0.5, 0.16666666666666685, 0.04166666666666603, 0.008333333333315703, 0.001388888888913419, 0.0001984126989972697,
    2.4801586903844624e-05, 2.755723148198774e-06, 2.755759237788011e-07, 2.5113226396698607e-08, 2.0825635355315893e-09 };
EMC_KDECL K_LOG[7] = {     /* atanh(s) = s + s*w*Q(w), w = s^2 <= 0.0295; degree 6, 1.4e-18 */
    0.3333333333333372, 0.19999999999670642, 0.1428571438008729, 0.11111098305913143, 0.09091836369173024,
    0.0765556191960003, 0.07413327838146679 };
EMC_KDECL K_MISC[10] = {
    1.5707963267948966, 6.123233995736766e-17,     /* 0,1: pi/2 hi, lo */
    3.141592653589793, 1.2246467991473532e-16,     /* 2,3: pi hi, lo */
    1.4426950408889634,                            /* 4: log2(e) */
    0.6931471805599453, 2.3190468138462996e-17,    /* 5,6: ln2 hi, lo */
    0.7142857142857143,                            /* 7: 1/1.4 (utils.py:154) */
    6.371e6,                                       /* 8: earth radius (environment.py:107) */
    1.7976931348623157e308 };                      /* 9: DBL_MAX */

/* ------------------------------------------------------------------------------------------------
 * Run constants (device: __constant__ memory, so they are instruction operands, not registers)
 * ---------------------------------------------------------------------------------------------- */
#define EMC_ATM_SEG 16
struct DevModel {
    /* atmosphere, environment.py:13-24 + literals of :52-90 */
    double T0, lapse, inv_T0, p0, h_tropo, h_strat, T_strat, inv_T_strat;
    double expo_tropo;   /* g/(R*L)                        :33 */
    double p11;          /* p0*(Ts/T0)^expo                :38-40 */
    double k_iso;        /* -g/(R*Ts)                      :43-44 */
    double p20, p25;     /* :56-62, :72-75 */
    double expo_25;      /* g/(R*0.0028)                   :81 */
    double R_gas, g0;
    double mach_k;       /* R_gas/(1.4*287.053): utils.mach_number hardcodes gamma and R (utils.py:152-157) whatever the atmosphere holds */
    double gamma;        /* atmosphere.gamma: speed_of_sound of get_properties only (environment.py:96) */
    double a2_k, rho_k;  /* 1.4*287.053 (a^2 = a2_k*T, utils.py:152-157) and a2_k/R_gas (1/(R T) = rho_k/a^2) */
    /* troposphere pressure as ONE polynomial: p0*(1 - L z/T0)^(g/(R L)) expanded about the middle of [tp_lo, tp_hi]
     * (binomial series in zeta = (z - tp_zc)*tp_inv_zh, |zeta| <= 1, truncation < 1e-18 relative; built in long double
     * by build_dev_model).  Replaces exp(e*log(T/T0)) where nearly every sounding-rocket step is flown; an atmosphere
     * whose constants do not allow it has tp_lo = +inf and takes the exp/log path. */
    double tp_lo, tp_hi, tp_zc, tp_inv_zh, tp_c[17];
    /* the layers above it, each cut into segments (at_lo, at_hi] that carry a degree-16 polynomial of the pressure in
     * zeta = (z - at_zc)*at_izh (Chebyshev interpolant of the layer's own formula, fitted and verified in long double by
     * build_dev_model) and the layer's temperature line T = clamp(at_tb + at_ts*(z - at_tz0), at_tmin, at_tmax).
     * Altitudes no segment covers (above 100 km, NaN, a layer the fit could not represent) take the exp/log path. */
    double at_lo[EMC_ATM_SEG], at_hi[EMC_ATM_SEG], at_zc[EMC_ATM_SEG], at_izh[EMC_ATM_SEG];
    double at_tb[EMC_ATM_SEG], at_ts[EMC_ATM_SEG], at_tz0[EMC_ATM_SEG], at_tmin[EMC_ATM_SEG], at_tmax[EMC_ATM_SEG];
    double at_c[EMC_ATM_SEG][17];
    /* mass properties, rocket.py:110-136 */
    double cg_dry, prop_cg, d4sq, len2_12, Ixx_dry, Iyy_dry;
    /* aerodynamics, rocket.py:138-218 */
    double ref_area, ref_diam, inv_ref_diam, area_diam, cp_location;
    double AR_over_cos2, two_pi_AR_cos, power_off_factor;   /* (AR/max(cos,1e-6))^2, 2*pi*AR*cos(sweep) */
    double stall_angle, inv_stall_span;
    /* the same quantities unfolded, for the strict continuation (emc_strict.cuh): rocket.py:167-180 as written */
    double fin_AR, fin_cos, fin_cos_floor, two_pi_AR, stall_span;
    double chute_cd, chute_area, chute_alt;
    /* simulator knobs, simulator.py:19-37,42,209 */
    double max_time, dt_rail, dt, half_dt, dt_over_6, pitch_damping, yaw_damping, rail_length;
    double stage_t[4], stage_c[4];   /* RK4 stage time offsets {0, dt/2, dt/2, dt} and next-state coefficients {dt/2, dt/2, dt, -} (simulator.py:218-222) */
    /* wind grid */
    double wind_alt0, wind_inv_dz;
    int32_t motor_kind, n_cd, n_cp, n_thrust, has_wind, n_wind, wind_uniform, n_mb;   /* n_mb: brackets of the Mach union grid */
    int32_t n_atm, pad_;                                                              /* segments in at_* */
};

/* Tables staged into shared memory by the kernels (host seam: plain struct), stored as BRACKETS so
 * that np.interp's clamping needs no special case: a table with n knots has n+1 brackets
 *   b = 0      (-inf, x[0])       value f[0],    slope 0
 *   b = 1..n-1 [x[b-1], x[b])     value f[b-1] + s*(x - x[b-1]),  s = (f[b]-f[b-1])/(x[b]-x[b-1])
 *   b = n      [x[n-1], +inf)     value f[n-1],  slope 0
 * Slopes are computed on the host with the expression np.interp uses (bit-identical).  Each lane
 * remembers its bracket (Mach and burn time move slowly), so a lookup is two compares and an FMA. */
/* The Cd and CP tables are both functions of Mach: their brackets are stored on the UNION of the two knot vectors
 * (m_lo/m_hi), each union bracket carrying the anchor, value and slope of the ORIGINAL Cd bracket and of the original
 * CP bracket it lies in — the same numbers the separate tables would use, found with one search instead of two. */
#define EMC_BRK_M (EMC_MAX_CD_KNOTS + EMC_MAX_CP_KNOTS + 2)
#define EMC_BRK_TH (EMC_MAX_THRUST_KNOTS + 2)
struct DevTables {
    double m_lo[EMC_BRK_M], m_hi[EMC_BRK_M];
    double cd_x0[EMC_BRK_M], cd0_f[EMC_BRK_M], cd0_s[EMC_BRK_M], cda_f[EMC_BRK_M], cda_s[EMC_BRK_M];
    double cp_x0[EMC_BRK_M], cp_f[EMC_BRK_M], cp_s[EMC_BRK_M];
    double th_lo[EMC_BRK_TH], th_hi[EMC_BRK_TH], th_x0[EMC_BRK_TH], th_f[EMC_BRK_TH], th_s[EMC_BRK_TH];
};

/* Per-sample parameters (registers) */
struct Sample {
    double dry_mass, prop_mass, dry_cg;      /* dry_cg = dry_mass*cg_dry, rocket.py:117 */
    double thrust_a, nozzle_area, burn_time;
    double pf_rate;                          /* -mdot/propellant_mass, simulator.py:444 */
    double cd_scale;
    const double *wind;                      /* this sample's [n_wind][3] table (global memory) */
};

/* Cached wind bracket: valid while lo <= z < hi (altitude moves <= ~10 m per step against
 * hundreds of metres of knot spacing, so reloads are rare). */
struct WindBracket {
    double lo, hi, x0;
    double f0[3], s[3];
    int32_t j_m, j_th, j_atm;      /* remembered brackets: Mach (Cd + CP) table, thrust-vs-time table, atmosphere segment */
};

struct State {
    double x, y, z, vx, vy, vz, q0, q1, q2, q3, wx, wy, wz, pf;
};

/* What stage 0 of a step exports about the stored state it is evaluated at (simulator.py:511-552) */
struct Diag {
    double mach2, qdyn, abs_aoa, stab;
};

EMC_HD bool finite_d(double v) { return fabs(v) <= 1.7976931348623157e308; }

/* a product the compiler may not contract into a following add: where one value is computed by two code paths (the
 * regular / general forms of aero_angles, chosen by a warp vote), both must round it the same way or a trajectory would
 * depend on its warp neighbours */
EMC_HD double mul_nc(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}

/* ---------------- NaN-faithful selects ---------------- */
EMC_HD double py_max(double a, double b) { return (b > a) ? b : a; }   /* Python max(a, b) */
EMC_HD double py_min(double a, double b) { return (b < a) ? b : a; }   /* Python min(a, b) */
/* running np.max / np.min over a series: NaN propagates and sticks */
EMC_HD void np_max_acc(double &m, double v) { if (v > m || v != v) m = v; }   /* a NaN m stays: both tests are False */
EMC_HD void np_min_acc(double &m, double v) { if (v < m || v != v) m = v; }

/* ---------------- cheap reciprocal / reciprocal square root / atan2 ----------------
 * Device: MUFU seed + Newton steps in DFMA, no slow-path branches (operands here are positive, normal
 * numbers: masses, inertias, R*T, squared speeds).  ~1 ulp; tests/test_gpu_parity.py measures it.
 * Host (test seam): plain IEEE division / sqrt. */
EMC_HD double fast_rcp(double x)
{
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    /* one cubic step from the ~20-bit seed: r (1 + e + e^2), e = 1 - x r; error e^3 ~ 2^-60 */
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
#else
    return 1.0 / x;
#endif
}

EMC_HD double fast_rsqrt(double x)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    /* one cubic step from the ~20-bit seed: y (1 + e/2 + 3 e^2/8), e = 1 - x y^2; error ~ e^3 */
    const double e = fma(-x * y, y, 1.0);
    const double t = fma(0.375, e, 0.5) * e;
    return fma(y, t, y);
#else
    return 1.0 / sqrt(x);
#endif
}

/* sqrt(x) for x >= 0 via x*rsqrt(x); 0 -> 0, NaN -> NaN */
EMC_HD double fast_sqrt(double x)
{
    const double r = x * fast_rsqrt(x);        /* x < 0 and NaN give NaN by themselves; +-0 and +inf (0*inf, inf*0) are passed through */
    return (x == 0.0 || x > 1.7976931348623157e308) ? x : r;
}

EMC_HD long long d2ll(double v)
{
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(v);
#else
    long long b; memcpy(&b, &v, sizeof b); return b;
#endif
}
EMC_HD double ll2d(long long b)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(b);
#else
    double v; memcpy(&v, &b, sizeof v); return v;
#endif
}

/* Degree-16 polynomial c[0..16] in x as two interleaved Horner chains (even / odd coefficients).  Estrin's scheme (5
 * dependent levels instead of 9, +3 multiplications) was measured SLOWER on the B200 (round 2, profiles/): with three
 * warps per scheduler the kernel follows the FP64-pipe instruction count, not the depth of the chain. */
EMC_HD double poly16(const double *c, double x)
{
    const double x2 = x * x;
    double pe = c[16], po = c[15];
    pe = fma(pe, x2, c[14]);  po = fma(po, x2, c[13]);
    pe = fma(pe, x2, c[12]);  po = fma(po, x2, c[11]);
    pe = fma(pe, x2, c[10]);  po = fma(po, x2, c[9]);
    pe = fma(pe, x2, c[8]);   po = fma(po, x2, c[7]);
    pe = fma(pe, x2, c[6]);   po = fma(po, x2, c[5]);
    pe = fma(pe, x2, c[4]);   po = fma(po, x2, c[3]);
    pe = fma(pe, x2, c[2]);   po = fma(po, x2, c[1]);
    pe = fma(pe, x2, c[0]);
    return fma(po, x, pe);
}

/* exp(x) = 2^k * (1 + r + r^2 Q(r)), k = rint(x/ln2), |r| <= ln2/2; Q: degree-10 minimax
 * (tools/fit_minimax.py exp 10, 6e-20).  Straight-line code: the atmosphere calls it once per
 * derivative for every layer (environment.py:42-45,59-62,66-69,90). */
EMC_HD double fast_exp(double x)
{
    const double magic = 6755399441055744.0;                    /* 1.5 * 2^52 */
    const double kd = fma(x, K_MISC[4], magic) - magic;
    double r = fma(kd, -K_MISC[5], x);
    r = fma(kd, -K_MISC[6], r);
    const double r2 = r * r;
    double qe = K_EXP[10], qo = K_EXP[9];
    qe = fma(qe, r2, K_EXP[8]);  qo = fma(qo, r2, K_EXP[7]);
    qe = fma(qe, r2, K_EXP[6]);  qo = fma(qo, r2, K_EXP[5]);
    qe = fma(qe, r2, K_EXP[4]);  qo = fma(qo, r2, K_EXP[3]);
    qe = fma(qe, r2, K_EXP[2]);  qo = fma(qo, r2, K_EXP[1]);
    qe = fma(qe, r2, K_EXP[0]);
    const double q = fma(qo, r, qe);
    const double p = fma(r2, q, r) + 1.0;
    const long long k = (long long)kd;
    const double scale = ll2d((k + 1023) << 52);
    const double y = p * scale;
    return (x > 709.0) ? INFINITY : ((x < -708.0) ? 0.0 : y);   /* NaN falls through as NaN */
}

/* log(x) = k ln2 + 2 atanh(s), s = (m-1)/(m+1), m in [sqrt(1/2), sqrt(2)); atanh(s) = s + s w Q(w), w = s^2,
 * Q: degree-6 minimax (tools/fit_minimax.py log 6, 1.4e-18).  Used for the two pow() layers of the
 * atmosphere (environment.py:31-33,79-81) where x = T/T_base stays within [0.7, 1.4], i.e. k = 0. */
EMC_HD double fast_log(double x)
{
    if (!(x >= 2.2250738585072014e-308 && x <= 1.7976931348623157e308)) return log(x);   /* 0, <0, denormal, inf, NaN */
    const long long b = d2ll(x);
    unsigned int hx = (unsigned int)(b >> 32);
    hx += 0x3ff00000u - 0x3fe6a09eu;
    const int k = (int)(hx >> 20) - 0x3ff;
    hx = (hx & 0x000fffffu) + 0x3fe6a09eu;
    const double m = ll2d(((long long)hx << 32) | (b & 0xffffffffLL));
    const double f = m - 1.0;
    const double s = f * fast_rcp(2.0 + f);
    const double w = s * s;
    double q = K_LOG[6];
    q = fma(q, w, K_LOG[5]);
    q = fma(q, w, K_LOG[4]);
    q = fma(q, w, K_LOG[3]);
    q = fma(q, w, K_LOG[2]);
    q = fma(q, w, K_LOG[1]);
    q = fma(q, w, K_LOG[0]);
    const double at = fma(s * w, q, s);
    const double kd = (double)k;
    return fma(kd, K_MISC[5], fma(2.0, at, kd * K_MISC[6]));
}

/* atan2 with one division and a degree-18 minimax polynomial in t^2 (tools/fit_atan.py, relative error
 * 2.8e-17 before rounding): atan(t) = t + t*u*Q(u), u = t^2, t = min(|x|,|y|)/max(|x|,|y|) in [0,1]. */
/* |x| off the FP64 pipe (where the value feeds a select, fabs() would be materialised by a DADD) */
EMC_HD double fabs_bits(double x)
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(x) & 0x7fffffff, __double2loint(x));
#else
    return fabs(x);
#endif
}

/* x > 0 ? x : 0.0 (NaN -> 0.0): one compare and one select (the ternary is pattern-matched into a seven-instruction
 * fmax emulation with NaN quieting, and rematerialised wherever the value is needed again) */
EMC_HD double pos_part(double x)
{
#if defined(__CUDA_ARCH__)
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, 0d0000000000000000;\n\tselp.f64 %0, %1, 0d0000000000000000, p;\n\t}" : "=d"(r) : "d"(x));
    return r;
#else
    return (x > 0.0) ? x : 0.0;
#endif
}

/* XPOS: x >= 0 is known (sideslip: x = |v_xz|), the half-plane fix-up is dropped.
 * NONZERO: max(|x|, |y|) > 0 is known (the caller has excluded the dead zone), the atan2(0, 0) guard is dropped. */
template <bool XPOS = false, bool NONZERO = false>
EMC_HD double fast_atan2(double y, double x)
{
    const double ax = fabs_bits(x), ay = fabs_bits(y);
    const bool swap = ay > ax;
    const double num = swap ? ax : ay, den = swap ? ay : ax;
    double t = num * fast_rcp(den);
    if (!NONZERO) t = (den == 0.0) ? 0.0 : t;            /* atan2(0, 0) = 0 (0 * rcp(0) is NaN); NaN stays NaN */
    const double u = t * t, u2 = u * u;
    /* two interleaved Horner chains (even / odd coefficients).  A shorter polynomial for |t| <= 1/4 behind a per-lane
     * branch was measured slower (round 2): the branch diverges inside warps that hold a tumbling flight. */
    double pe = K_ATAN[18], po = K_ATAN[17];
    pe = fma(pe, u2, K_ATAN[16]);  po = fma(po, u2, K_ATAN[15]);
    pe = fma(pe, u2, K_ATAN[14]);  po = fma(po, u2, K_ATAN[13]);
    pe = fma(pe, u2, K_ATAN[12]);  po = fma(po, u2, K_ATAN[11]);
    pe = fma(pe, u2, K_ATAN[10]);  po = fma(po, u2, K_ATAN[9]);
    pe = fma(pe, u2, K_ATAN[8]);   po = fma(po, u2, K_ATAN[7]);
    pe = fma(pe, u2, K_ATAN[6]);   po = fma(po, u2, K_ATAN[5]);
    pe = fma(pe, u2, K_ATAN[4]);   po = fma(po, u2, K_ATAN[3]);
    pe = fma(pe, u2, K_ATAN[2]);   po = fma(po, u2, K_ATAN[1]);
    pe = fma(pe, u2, K_ATAN[0]);
    const double q = fma(po, u, pe);
    double r = fma(t * u, q, t);
    /* octant fix-ups as straight-line selects: r <- c_hi - r + c_lo */
    const double r1 = (K_MISC[0] - r) + K_MISC[1];
    r = swap ? r1 : r;
    if (!XPOS) {
        const double r2 = (K_MISC[2] - r) + K_MISC[3];
        r = (x < 0.0) ? r2 : r;
    }
    return copysign(r, y);
}

/* The angle of attack atan2(y1, x1) and the sideslip atan2(y2, x2), x2 >= 0, of one derivative evaluation, computed
 * TOGETHER: the two evaluations are independent, and written side by side their four Horner chains interleave, so the
 * ~8-cycle latency of a dependent DFMA is covered by the other chains instead of being waited out twice (the flight
 * kernel holds three warps per scheduler; fixed-latency dependency waits are its largest stall, profiles/).  Same
 * arithmetic per value as fast_atan2<., true>: both dead zones have been excluded by the caller. */
EMC_HD void fast_atan2_pair(double y1, double x1, double y2, double x2, double &r1, double &r2)
{
    const double ax1 = fabs_bits(x1), ay1 = fabs_bits(y1), ay2 = fabs_bits(y2);
    const bool sw1 = ay1 > ax1, sw2 = ay2 > x2;
    const double n1 = sw1 ? ax1 : ay1, d1 = sw1 ? ay1 : ax1;
    const double n2 = sw2 ? x2 : ay2, d2 = sw2 ? ay2 : x2;
    const double t1 = n1 * fast_rcp(d1), t2 = n2 * fast_rcp(d2);
    const double u1 = t1 * t1, u2 = t2 * t2, v1 = u1 * u1, v2 = u2 * u2;
    double pe1 = K_ATAN[18], po1 = K_ATAN[17], pe2 = K_ATAN[18], po2 = K_ATAN[17];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 16; i >= 2; i -= 2) {
        pe1 = fma(pe1, v1, K_ATAN[i]); pe2 = fma(pe2, v2, K_ATAN[i]);
        po1 = fma(po1, v1, K_ATAN[i - 1]); po2 = fma(po2, v2, K_ATAN[i - 1]);
    }
    pe1 = fma(pe1, v1, K_ATAN[0]); pe2 = fma(pe2, v2, K_ATAN[0]);
    const double q1 = fma(po1, u1, pe1), q2 = fma(po2, u2, pe2);
    double a = fma(t1 * u1, q1, t1), b = fma(t2 * u2, q2, t2);
    const double a1 = (K_MISC[0] - a) + K_MISC[1], b1 = (K_MISC[0] - b) + K_MISC[1];
    a = sw1 ? a1 : a; b = sw2 ? b1 : b;
    const double a2 = (K_MISC[2] - a) + K_MISC[3];
    a = (x1 < 0.0) ? a2 : a;
    r1 = copysign(a, y1); r2 = copysign(b, y2);
}

/* true if the predicate holds for any lane of the warp that is executing this code */
EMC_HD bool any_lane(bool p)
{
#if defined(__CUDA_ARCH__)
    return __any_sync(__activemask(), p) != 0;
#else
    return p;
#endif
}

/* remembered-bracket lookup: j stays valid while lo[j] <= x < hi[j]; NaN leaves j alone (the FMA that
 * follows then yields NaN, which is np.interp's answer) */
EMC_COLD int brk_search(const double *lo, const double *hi, int nb, int j, double x)
{
    while (j < nb - 1 && x >= hi[j]) ++j;
    while (j > 0 && x < lo[j]) --j;
    return j;
}
EMC_HD int brk_find(const double *lo, const double *hi, int nb, int j, double x)
{
    if (!(x >= lo[j] && x < hi[j])) j = brk_search(lo, hi, nb, j, x);      /* rare; a NaN x fails both searches and keeps j */
    return j;
}

/* ---------------- atmosphere: T and 1/(R*T), p  (environment.py:26-103) ---------------- */
/* T and p; 1/(R T) is left to the caller (the flight derivative takes 1/a = rsqrt(1.4*287.053*T) instead and derives
 * 1/(R T) from it, the other users take fast_rcp(R T), see the wrapper below) */
EMC_HD void atmosphere_upper(const DevModel &M, double z, int &j_atm, double &T, double &p);

/* SPEC: the troposphere value is computed before the range test, in the caller's straight-line code (where its
 * eleven-deep dependent chain overlaps the quaternion and mass-property work), and the other layers overwrite it
 * behind one rarely taken branch. */
template <bool SPEC = false>
EMC_HD void atmosphere_Tp(const DevModel &M, double z, int &j_atm, double &T, double &p)
{
    if (SPEC) {
        T = M.T0 - M.lapse * z;
        p = poly16(M.tp_c, (z - M.tp_zc) * M.tp_inv_zh);
        if (!EMC_LIKELY(z >= M.tp_lo && z <= M.tp_hi)) {
            double T2, p2;
            atmosphere_upper(M, z, j_atm, T2, p2);
            T = T2; p = p2;
        }
        return;
    }
    if (EMC_LIKELY(z >= M.tp_lo && z <= M.tp_hi)) {          /* troposphere (environment.py:28-33): T linear, p by the series above */
        T = M.T0 - M.lapse * z;
        p = poly16(M.tp_c, (z - M.tp_zc) * M.tp_inv_zh);
        return;
    }
    atmosphere_upper(M, z, j_atm, T, p);
}

/* everything outside the troposphere polynomial's range */
EMC_HD void atmosphere_upper(const DevModel &M, double z, int &j_atm, double &T, double &p)
{
    if (M.n_atm > 0) {                          /* upper layers (environment.py:35-103) by segment polynomial */
        int j = j_atm;
        bool ok = (z > M.at_lo[j]) && (z <= M.at_hi[j]);
        if (!ok) {
            for (int k = 0; k < M.n_atm; ++k)
                if (z > M.at_lo[k] && z <= M.at_hi[k]) { j = k; ok = true; break; }
        }
        if (ok) {
            j_atm = j;
            T = M.at_tb[j] + M.at_ts[j] * (z - M.at_tz0[j]);
            T = py_min(T, M.at_tmax[j]);
            T = py_max(T, M.at_tmin[j]);
            p = poly16(M.at_c[j], (z - M.at_zc[j]) * M.at_izh[j]);
            return;
        }
    }
    /* every layer is p = base * exp(arg); the two pow() layers use arg = e*log(T/Tb) */
    double base, arg, lx = 1.0, le = 0.0;
    bool use_log = false;
    if (z <= M.h_tropo) {
        T = M.T0 - M.lapse * z;
        base = M.p0; lx = T * M.inv_T0; le = M.expo_tropo; use_log = true; arg = 0.0;
    } else if (z <= M.h_strat) {
        T = M.T_strat;
        base = M.p11; arg = M.k_iso * (z - M.h_tropo);
    } else if (z <= 32000.0) {
        T = M.T_strat + 0.001 * (z - M.h_strat);
        T = py_min(T, 228.65);
        if (z <= 25000.0) { base = M.p20; arg = M.k_iso * (z - M.h_strat); }
        else { base = M.p25; lx = T * M.inv_T_strat; le = M.expo_25; use_log = true; arg = 0.0; }
    } else {
        T = 228.65 - 0.0028 * (z - 32000.0);
        T = py_max(T, 180.0);
        base = 868.02; arg = 0.0;     /* filled below once 1/(R*T) is known */
    }
    if (use_log) arg = le * fast_log(lx);
    if (z > 32000.0 || z != z) arg = -(z - 32000.0) * (M.g0 * fast_rcp(M.R_gas * T));   /* -(z-32000)/(R*T/g) */
    p = base * fast_exp(arg);
}

EMC_HD void atmosphere(const DevModel &M, double z, int &j_atm, double &T, double &inv_RT, double &p)
{
    atmosphere_Tp<false>(M, z, j_atm, T, p);
    inv_RT = fast_rcp(M.R_gas * T);
}

EMC_HD void atmosphere(const DevModel &M, double z, double &T, double &inv_RT, double &p)
{
    int j = 0;
    atmosphere(M, z, j, T, inv_RT, p);
}

/* environment.py:105-108 */
EMC_HD double gravity(const DevModel &M, double z)
{
    const double Re = K_MISC[8];
    const double r = Re * fast_rcp(Re + z);
    return M.g0 * (r * r);
}

/* Compile-time knowledge of the run (flight kernel instances): MK = 0 liquid, 1 solid; WK = 0 no wind table, 1 wind
 * table; -1 = read the flag from the model at run time (test seam, rail / series / debug kernels). */
template <int MK> EMC_HD bool cfg_solid(const DevModel &M) { return MK < 0 ? (M.motor_kind == EMC_MOTOR_SOLID) : (MK == 1); }
template <int WK> EMC_HD bool cfg_wind(const DevModel &M) { return WK < 0 ? (M.has_wind != 0) : (WK == 1); }

/* ---------------- wind table (environment.py:267-276 -> three np.interp on one grid) ------------- */
/* returns true when the altitude is +-inf: np.interp gives the end value there, and the bracket form s*(z - x0) + f0 would
 * give 0*inf = NaN, so the caller takes f0 as it is (the bracket is left invalid: an infinite altitude reloads every time) */
EMC_COLD bool wind_bracket_load(const DevModel &M, const double *alt, const double *w, double z, WindBracket &B)
{
    const int n = M.n_wind;
    if (z != z) {                               /* NaN altitude -> NaN wind, never valid */
        B.lo = z; B.hi = z; B.x0 = 0.0;
        B.f0[0] = B.f0[1] = B.f0[2] = z; B.s[0] = B.s[1] = B.s[2] = 0.0;
        return false;
    }
    if (!finite_d(z)) {
        const int e = (z > 0.0) ? n - 1 : 0;
        B.lo = 1.0; B.hi = 0.0; B.x0 = 0.0;     /* empty interval */
        for (int k = 0; k < 3; ++k) { B.f0[k] = w[3 * e + k]; B.s[k] = 0.0; }
        return true;
    }
    int j; bool clamp = false;
    if (z >= alt[n - 1]) { j = n - 1; clamp = true; B.lo = alt[n - 1]; B.hi = INFINITY; }
    else if (z < alt[0]) { j = 0; clamp = true; B.lo = -INFINITY; B.hi = alt[0]; }
    else {
        int lo = 0, hi = n - 1;
        if (M.wind_uniform) {                   /* guess, then fix up against the real knots */
            int g = (int)((z - M.wind_alt0) * M.wind_inv_dz);
            g = g < 0 ? 0 : (g > n - 2 ? n - 2 : g);
            while (g > 0 && alt[g] > z) --g;
            while (g < n - 2 && alt[g + 1] <= z) ++g;
            lo = g;
        } else {
            while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (alt[mid] <= z) lo = mid; else hi = mid; }
        }
        j = lo; B.lo = alt[j]; B.hi = alt[j + 1];
    }
    B.x0 = alt[j];
    for (int k = 0; k < 3; ++k) {
        double f0 = w[3 * j + k];
        B.f0[k] = f0;
        B.s[k] = clamp ? 0.0 : (w[3 * (j + 1) + k] - f0) / (alt[j + 1] - alt[j]);
    }
    return false;
}

template <int WK = -1>
EMC_HD void wind_at(const DevModel &M, const double *alt, const Sample &S, double z, WindBracket &B, double w[3])
{
    if (!cfg_wind<WK>(M)) { w[0] = w[1] = w[2] = 0.0; return; }
    /* every value of the remembered bracket is read before the range test (no short circuit: both bounds and the seven
     * coefficients are loads the scheduler may issue at once); the rare miss reloads them */
    const double lo = B.lo, hi = B.hi;
    double x0 = B.x0, s0 = B.s[0], s1 = B.s[1], s2 = B.s[2], f0 = B.f0[0], f1 = B.f0[1], f2 = B.f0[2];
    if (!((z >= lo) & (z < hi))) {
        if (wind_bracket_load(M, alt, S.wind, z, B)) { w[0] = B.f0[0]; w[1] = B.f0[1]; w[2] = B.f0[2]; return; }
        x0 = B.x0; s0 = B.s[0]; s1 = B.s[1]; s2 = B.s[2]; f0 = B.f0[0]; f1 = B.f0[1]; f2 = B.f0[2];
    }
    const double dz = z - x0;
    w[0] = s0 * dz + f0;
    w[1] = s1 * dz + f1;
    w[2] = s2 * dz + f2;
}

/* ---------------- thrust (motor.py:54-76, 152-156), caller has checked pf>0 && t<=burn ---------- */
/* thrust inside the burn window (the caller has established 0 <= t <= burn_time) */
/* th_safe: the caller has established th_lo[j_th] <= t < th_hi[j_th] (rk4_step tests [t, t + dt] once per step) */
/* the part of the thrust that depends on the time alone (thrust curve x sample factor; a liquid motor has none): the
 * derivative evaluates it first, off the chain that leads through the atmosphere to the pressure term */
template <int MK = -1>
EMC_HD double thrust_time_part(const DevModel &M, const DevTables &Tb, const Sample &S, WindBracket &C, double t, bool th_safe)
{
    if (!cfg_solid<MK>(M)) return S.thrust_a;
    int j = C.j_th;
    double ts = Tb.th_s[j], tx = Tb.th_x0[j], tf = Tb.th_f[j];
    if (!th_safe) {
        const int j2 = brk_find(Tb.th_lo, Tb.th_hi, M.n_thrust + 1, j, t);
        if (j2 != j) { C.j_th = j2; ts = Tb.th_s[j2]; tx = Tb.th_x0[j2]; tf = Tb.th_f[j2]; }
    }
    return fma(ts, t - tx, tf) * S.thrust_a;
}
template <int MK = -1>
EMC_HD double thrust_finish(const DevModel &M, const Sample &S, double time_part, double p)
{
    if (cfg_solid<MK>(M)) return time_part + S.nozzle_area * (101325.0 - p);
    return time_part - S.nozzle_area * p;
}

template <int MK = -1>
EMC_HD double thrust_core(const DevModel &M, const DevTables &Tb, const Sample &S, WindBracket &C, double t, double p, bool th_safe = false)
{
    if (cfg_solid<MK>(M)) {
        int j = C.j_th;
        if (!th_safe) { j = brk_find(Tb.th_lo, Tb.th_hi, M.n_thrust + 1, j, t); C.j_th = j; }
        const double f = fma(Tb.th_s[j], t - Tb.th_x0[j], Tb.th_f[j]) * S.thrust_a;
        return f + S.nozzle_area * (101325.0 - p);
    }
    return S.thrust_a - S.nozzle_area * p;
}

template <int MK = -1>
EMC_HD double thrust_at(const DevModel &M, const DevTables &Tb, const Sample &S, WindBracket &C, double t, double p)
{
    if (t < 0.0 || t > S.burn_time) return 0.0;
    return thrust_core<MK>(M, Tb, S, C, t, p);
}

/* t < 0 for a time value (never NaN, never -0): the sign bit, tested off the FP64 pipe */
EMC_HD bool time_negative(double t)
{
#if defined(__CUDA_ARCH__)
    return __double2hiint(t) < 0;
#else
    return t < 0.0;
#endif
}

/* Mach number, aerodynamic angles (utils.py:160-172) and their sin/cos (utils.py:194-197, from the velocity ratios) of
 * one derivative evaluation.  GENERAL = true is the reference's semantics in full: dead zones (|vbx|, |vbz| < 1e-6 ->
 * alpha = 0; |v_xz| < 1e-6 -> beta = 0), sqrt(0) = 0 and sqrt(inf) = inf, Mach clamped like np.interp's right end.
 * GENERAL = false assumes !a_dead and finite vb2, under which those guards are dead code: |v_xz| >= 1e-6 follows from
 * !a_dead, vb2 > 0, and Mach <= 1e152.  Where any lane of the warp needs the sideslip polynomial (3-D flights: always;
 * planar flights: never, atan2(+-0, vxz) = +-0) the two angles are evaluated together. */
struct AeroAngles { double mach, alpha, beta, ca, sa, cb, sb; };

template <bool GENERAL>
EMC_HD void aero_angles(double vbx, double vby, double vbz, double vxz2, double vb2, double rvb, double ya, bool aero, bool a_dead,
                        AeroAngles &A, double rvxz)
{
    double speed = mul_nc(vb2, rvb);
    if (GENERAL) speed = (vb2 == 0.0 || vb2 > 1.7976931348623157e308) ? vb2 : speed;        /* sqrt(0) = 0, sqrt(inf) = inf */
    double mach = mul_nc(speed, ya);
    if (GENERAL) mach = (mach > 1e300) ? 1e300 : mach;     /* +inf clamps like np.interp (right value), NaN stays NaN */
    A.mach = mach;
    const double vxz = (GENERAL && !(vxz2 > 0.0)) ? vxz2 : mul_nc(vxz2, rvxz);
    const bool b_dead = GENERAL ? (vxz < 1e-6) : false;
    const bool need_beta = aero && !b_dead && vby != 0.0;
    double alpha, beta;
    if (any_lane(need_beta)) {
        fast_atan2_pair(vbz, vbx, vby, vxz, alpha, beta);
        if (GENERAL) alpha = a_dead ? 0.0 : alpha;
        beta = need_beta ? beta : ((GENERAL && b_dead) ? 0.0 : vby);
    } else {
        alpha = fast_atan2<false, true>(vbz, vbx);      /* not dead: max(|vbx|, |vbz|) >= 1e-6 */
        if (GENERAL) alpha = a_dead ? 0.0 : alpha;
        beta = (GENERAL && b_dead) ? 0.0 : vby;
    }
    A.alpha = alpha; A.beta = beta;
    A.ca = (GENERAL && a_dead) ? 1.0 : mul_nc(vbx, rvxz); A.sa = (GENERAL && a_dead) ? 0.0 : mul_nc(vbz, rvxz);
    A.cb = (GENERAL && b_dead) ? 1.0 : mul_nc(vxz, rvb); A.sb = (GENERAL && b_dead) ? 0.0 : mul_nc(vby, rvb);
}

/* ------------------------------------------------------------------------------------------------
 * The derivative, simulator.py:295-460.
 *   chute      : sticky parachute flag (self.parachute_deployed), may be latched by any stage (F12)
 *   want_diag  : stage 0 only — export Mach^2, q_inf, |alpha|, stability margin of this state
 * ---------------------------------------------------------------------------------------------- */
/* REG: the sample is REGULAR — dry mass positive and finite, propellant mass non-negative and finite, inertia constants
 * positive (sample_regular below).  Then `mass < dry_mass` (simulator.py:315-318) and `Ixx > 0`, `Iyy > 0` (:431-436) are
 * invariants of the flight, and the flight kernel drops the tests (irregular samples never enter its fast path: they are
 * handed to the strict continuation before their first step). */
template <int MK = -1, int WK = -1, bool REG = false>
EMC_HD void derivative(const DevModel &M, const DevTables &Tb, const double *wind_alt, const Sample &S,
                       WindBracket &WB, double t, const State &s, bool &chute, double &chute_time,
                       State &k, bool want_diag, Diag &dg, bool th_safe = false)
{
    /* :305  pf = max(0.0, pf)  (NaN -> 0.0) */
    const double pf = pos_part(s.pf);
    const bool burning = (pf > 0.0) & (t <= S.burn_time);
    const double thr_t = thrust_time_part<MK>(M, Tb, S, WB, t, th_safe);
    double w[3];
    wind_at<WK>(M, wind_alt, S, s.z, WB, w);
    /* the remembered Mach bracket is read here, long before the Mach number exists */
    const int jm0 = WB.j_m;
    const double m_lo0 = Tb.m_lo[jm0], m_hi0 = Tb.m_hi[jm0];

    /* :308  the quaternion is normalised by the reference (identity if |q| <= 1e-12 or NaN, utils.py:76-82).  Here the
     * components stay as they are and the norm goes into two scalars: a = 2 v / |q|^2 for the rotations and
     * hq = 1 / (2 |q|) for q_dot. */
    double qw = s.q0, qx = s.q1, qy = s.q2, qz = s.q3;
    double n2 = qw * qw + qx * qx + qy * qy + qz * qz;
    if (!(n2 > 1e-24)) { qw = 1.0; qx = 0.0; qy = 0.0; qz = 0.0; n2 = 1.0; }
    const double rn = fast_rsqrt(n2);
    const double hq = 0.5 * rn;
    const double s2 = (rn + rn) * rn;
    const double ax = s2 * qx, ay = s2 * qy, az = s2 * qz;

    /* :311-321  mass properties, rocket.py:110-136 */
    double mp = mul_nc(S.prop_mass, pf);        /* its own rounding in every instance (with REG nothing stands between it and the sum) */
    double mass = S.dry_mass + mp;
    if (!REG && mass < S.dry_mass) { mass = S.dry_mass; mp = S.prop_mass * 0.0; }      /* :315-318 */
    const double inv_m = fast_rcp(mass);
    const double cg = (S.dry_cg + mp * M.prop_cg) * inv_m;
    const double dcg = M.prop_cg - cg;
    const double Ixx = M.Ixx_dry + mp * M.d4sq;
    const double Iyy = M.Iyy_dry + mp * (M.len2_12 + dcg * dcg);

    /* :328-338  atmosphere + wind.  ya = 1/a with a^2 = 1.4*287.053*T (utils.py:152-157): Mach = |v| ya, and
     * 1/(R T) = ya^2 * (1.4*287.053/R) gives the density, so one reciprocal square root serves both. */
    double T, p;
#ifndef EMC_NO_SPEC_ATM
    atmosphere_Tp<true>(M, s.z, WB.j_atm, T, p);
#else
    atmosphere_Tp<false>(M, s.z, WB.j_atm, T, p);
#endif
    const double ya = fast_rsqrt(M.a2_k * T);
    const double ya2 = ya * ya;
    const double rho = p * (ya2 * M.rho_k);

    /* :341-352  v_body = R(q)^T u as a quaternion rotation: u - w tb + v x tb, tb = a x u (utils.py:100-111,129-136) */
    const double ux = s.vx - w[0], uy = s.vy - w[1], uz = s.vz - w[2];
    const double tbx = ay * uz - az * uy, tby = az * ux - ax * uz, tbz = ax * uy - ay * ux;
    const double vbx = fma(-qz, tby, fma(qy, tbz, fma(-qw, tbx, ux)));
    const double vby = fma(-qx, tbz, fma(qz, tbx, fma(-qw, tby, uy)));
    const double vbz = fma(-qy, tbx, fma(qx, tby, fma(-qw, tbz, uz)));
    /* |u| = |v_body| (a rotation): the squared speed is taken from the body components, whose partial sum the
     * aerodynamic angles need anyway */
    const double vxz2 = vbx * vbx + vbz * vbz;
    const double vb2 = vxz2 + vby * vby;
    const double rvb = fast_rsqrt(vb2);                      /* 1/|v|: Mach, sideslip ratios, parachute drag direction */
    const double mach2 = vb2 * ya2;                          /* (|v|/sqrt(1.4*287.053*T))^2 */
    const double qdyn = 0.5 * rho * vb2;

    /* :359-363 thrust along body x */
    /* :359-363 the time part was evaluated first (above); np.interp clamps outside the curve, so any time is a valid
     * argument and the value is selected afterwards: no branch region around the table reads */
    const double thrust = thrust_finish<MK>(M, S, thr_t, p);
    double fbx = (burning && !time_negative(t)) ? thrust : 0.0;                                          /* motor.py:55,153: 0 outside [0, burn_time] */
    double fby = 0.0, fbz = 0.0;
    double mx = 0.0, my = 0.0, mz = 0.0;

    /* :366-369 sticky parachute latch */
    {
        const bool latch = (!chute) & (s.z <= M.chute_alt) & (s.vz < 0.0);
        chute_time = latch ? t : chute_time;
        chute = chute | latch;
    }

    const bool aero = (!chute) && (qdyn > 0.0);
    if (aero || want_diag) {
        /* Mach, angle of attack, sideslip and their sin/cos.  REGULAR state: the body velocity is outside the dead zone of
         * utils.py:160-172 and |v|^2 is finite — every step of a real flight.  Then none of the guards of the general
         * form can fire (no dead-zone selects, sqrt(0) / sqrt(inf) fix-ups or Mach clamp), and a warp whose lanes are all
         * regular takes the short form; one irregular lane sends the whole warp through the general one (a vote, so the
         * branch never diverges). */
        const bool a_dead = (fabs(vbx) < 1e-6) && (fabs(vbz) < 1e-6);
        const bool regular = (!a_dead) && (vb2 <= 1.7976931348623157e308);
        AeroAngles A;
        const double rvxz = fast_rsqrt(vxz2);
        if (!any_lane(!regular)) aero_angles<false>(vbx, vby, vbz, vxz2, vb2, rvb, ya, aero, a_dead, A, rvxz);
        else aero_angles<true>(vbx, vby, vbz, vxz2, vb2, rvb, ya, aero, a_dead, A, rvxz);
        const double mach = A.mach, alpha = A.alpha, beta = A.beta;
        /* Mach-table brackets (rocket.py:105-108,156-157): shared by Cd0/Cda, separate knots for CP */
        int jm = jm0;
        double cps = Tb.cp_s[jm0], cpx = Tb.cp_x0[jm0], cpf = Tb.cp_f[jm0];
        double cdx = Tb.cd_x0[jm0], cd0s = Tb.cd0_s[jm0], cd0f = Tb.cd0_f[jm0], cdas = Tb.cda_s[jm0], cdaf = Tb.cda_f[jm0];
        if (!((mach >= m_lo0) & (mach < m_hi0))) {          /* rare; a NaN Mach number fails both searches and keeps the bracket */
            jm = brk_search(Tb.m_lo, Tb.m_hi, M.n_mb, jm0, mach);
            if (jm != jm0) {
                WB.j_m = jm;
                cps = Tb.cp_s[jm]; cpx = Tb.cp_x0[jm]; cpf = Tb.cp_f[jm];
                cdx = Tb.cd_x0[jm]; cd0s = Tb.cd0_s[jm]; cd0f = Tb.cd0_f[jm]; cdas = Tb.cda_s[jm]; cdaf = Tb.cda_f[jm];
            }
        }
        const double cp = M.cp_location + fma(cps, mach - cpx, cpf);
        const double sm = cp - cg;
        if (want_diag) { dg.mach2 = mach2; dg.qdyn = qdyn; dg.abs_aoa = fabs(alpha); dg.stab = sm * M.inv_ref_diam; }
        if (aero) {
            const double ca = A.ca, sa = A.sa, cb = A.cb, sb = A.sb;

            /* rocket.py:138-218 */
            const double dm = mach - cdx;
            const double cd0 = fma(cd0s, dm, cd0f) * S.cd_scale;
            const double cda = fma(cdas, dm, cdaf);
            double cd = cd0 + cda * (alpha * alpha);
            if (!(pf > 0.0)) cd *= M.power_off_factor;
            const double abs_alpha = fabs(alpha);
            /* cl_alpha = 2*pi*AR / (2 + sqrt(4 + (AR*beta_M/cos)^2)) * cos,  beta_M^2 = |1 - M^2|  (:178-180) */
            const double rad = 4.0 + M.AR_over_cos2 * fabs(1.0 - mach2);
            const double cl_alpha = M.two_pi_AR_cos * fast_rcp(2.0 + rad * fast_rsqrt(rad));
            double cl = cl_alpha * alpha;
            double cy = cl_alpha * beta;
            {
                /* rocket.py:186-196 as selects, not a branch: the kernel is bound by dependent-instruction latency, and a
                 * (divergent) branch here ends the basic block in which the force and moment products overlap
                 * (measured: +1.8 % on C3, +1.3 % on launch->landing, lone trajectory 5.02 -> 4.92 us/step) */
                const bool stalled = abs_alpha > M.stall_angle;
                const double over = (abs_alpha - M.stall_angle) * M.inv_stall_span;
                const double sf = pos_part(1.0 - over);          /* max(0.0, 1 - over), NaN -> 0 */
                const double sgn = (alpha > 0.0) ? 1.0 : ((alpha < 0.0) ? -1.0 : alpha);
                const double cl_s = cl_alpha * M.stall_angle * sf * sgn;
                const double cd_s = cd * (1.0 + 0.5 * over);
                const double cy_s = cy * sf;
                cl = stalled ? cl_s : cl; cd = stalled ? cd_s : cd; cy = stalled ? cy_s : cy;
            }
            const double cm = -cl_alpha * sm * alpha;
            const double cyaw = -cl_alpha * sm * beta;

            /* :385-391  F_b += W2B(alpha,beta) @ [-D,-S,-L], utils.py:199-205 */
            const double qa = qdyn * M.ref_area;
            const double D = qa * cd, L = qa * cl, Sd = qa * cy;
            fbx += (ca * cb) * (-D) + (-sb) * (-Sd) + (sa * cb) * (-L);
            fby += (ca * sb) * (-D) + cb * (-Sd) + (sa * sb) * (-L);
            fbz += (-sa) * (-D) + ca * (-L);
            /* :394-411; croll = 0.0 but q*0.0 keeps the reference's NaN/Inf propagation into roll */
            const double qad = qdyn * M.area_diam;
            mx = qdyn * 0.0;
            my = qad * cm;
            mz = qad * cyaw;
        }
    }
    if (chute) {
        /* :372-377 */
        if (vb2 > 0.0) {                                   /* rel_speed > 0 (False for NaN) */
            const double drag = (0.5 * rho * vb2 * M.chute_cd) * M.chute_area;
            const double f = -drag * rvb;
            fbx += f * vbx; fby += f * vby; fbz += f * vbz;
        }
    }

    /* :414-415 damping */
    my += -M.pitch_damping * s.wy;
    mz += -M.yaw_damping * s.wz;

    /* :418-425  F_inertial = R(q) F_body = F + w tf + v x tf, tf = a x F */
    const double g = gravity(M, s.z);
    const double tfx = ay * fbz - az * fby, tfy = az * fbx - ax * fbz, tfz = ax * fby - ay * fbx;
    const double fix = fma(-qz, tfy, fma(qy, tfz, fma(qw, tfx, fbx)));
    const double fiy = fma(-qx, tfz, fma(qz, tfx, fma(qw, tfy, fby)));
    const double fiz = fma(-qy, tfx, fma(qx, tfy, fma(qw, tfz, fbz)));
    k.x = s.vx; k.y = s.vy; k.z = s.vz;
    k.vx = fix * inv_m; k.vy = fiy * inv_m; k.vz = fma(fiz, inv_m, -g);      /* (F_z - m g)/m */

    /* :431-436  Euler equations with Izz == Iyy (rocket.py:127); Ixx, Iyy > 0 */
    const double inv_Iyy = fast_rcp(Iyy);
    /* roll: (mx - (Izz-Iyy)*wy*wz)/Ixx with Izz-Iyy == 0 and mx in {0, NaN}: no division needed */
    k.wx = (REG || Ixx > 0.0) ? (mx - 0.0 * (s.wy * s.wz)) : 0.0;
    k.wy = (REG || Iyy > 0.0) ? (my - (Ixx - Iyy) * s.wz * s.wx) * inv_Iyy : 0.0;
    k.wz = (REG || Iyy > 0.0) ? (mz - (Iyy - Ixx) * s.wx * s.wy) * inv_Iyy : 0.0;

    /* :439  q_dot = 0.5 * qn (x) (0,w) - 0.5*(qn.qn - 1)*qn with qn = q/|q|, utils.py:114-121.  qn.qn - 1 is a
     * rounding residue (<= 3e-16): its term is 1e-16 of q_dot and, times dt, 1e-3 ulp of q — dropped (a non-finite q
     * makes every product below non-finite as well, so the NaN behaviour is the same). */
    k.q0 = hq * (-qx * s.wx - qy * s.wy - qz * s.wz);
    k.q1 = hq * (qw * s.wx + qy * s.wz - qz * s.wy);
    k.q2 = hq * (qw * s.wy - qx * s.wz + qz * s.wx);
    k.q3 = hq * (qw * s.wz + qx * s.wy - qy * s.wx);

    /* :442-450 propellant; the 10 ms taper test pf/|rate| < 0.01 is written as pf < 0.01*|rate| (the two branches are
     * continuous at the boundary; a zero rate gives 0 and the test fails like the reference's `rate != 0`) */
    const double pf_full = S.pf_rate, pf_taper = -pf * 100.0;
    double pfr = (pf < 0.01 * fabs(pf_full)) ? pf_taper : pf_full;
    pfr = burning ? pfr : 0.0;
    k.pf = pfr;
}

/* ------------------------------------------------------------------------------------------------
 * Flight bookkeeping: everything the loop of simulator.py:216-264 and the summary of :474-494 need
 * ---------------------------------------------------------------------------------------------- */
/* HOT part: read or written in every stage / every step — registers or shared memory. */
struct TrackHot {
    double t, t_rail;
    double apogee_alt;                /* running np.argmax(altitudes), :488 */
    int32_t n_steps, first_nan;
    bool chute, apogee_detected, burnout_found;
    bool finishing;   /* loop ended: one more stage-0 pass exports the last stored state's diagnostics */
    int8_t replay;    /* NaN fast-forward mode (0 none, 1 all-NaN, 2 altitude-NaN ballistic): t is replayed to max_time */
    int8_t term;      /* emc_termination */
    uint8_t om_half;  /* flight kernel: the attitude-rate amplitude half way to the lane hand-back point (emc_counters.yielded), on
                       * a logarithmic scale (omega_code) */
};
/* COLD part: touched once per step at most (running maxima, event times, tape cursor).  The flight kernel may keep it
 * in global memory, field-major over the resident lanes (coalesced, L2-resident), to fit 16 warps per SM into the
 * shared memory; everything else keeps it next to the hot part.  Access goes through a small accessor (ColdStruct
 * here, GlobalCold in emc_engine.cu) with compile-time field indices. */
enum { TC_APOGEE_T = 0, TC_LATCH, TC_MAX_COAST,         /* :488-490, :247-257 */
       TC_BURNOUT_TIME, TC_CHUTE_TIME,
       TC_MAX_MACH2, TC_MAX_Q, TC_MAX_V2, TC_MAX_OM, TC_MIN_STAB, TC_MAX_STAB, TC_MAX_AOA, TC_DCOUNT };
enum { TI_APOGEE_INDEX = 0, TI_BT_SLOT, TI_BT_NEXT, TI_SAMPLE, TI_ICOUNT };   /* BT_*: batch tape row block (-1: not recorded), next stored-state index to record; SAMPLE: index of the sample this record flies */
struct TrackCold { double d[TC_DCOUNT]; int32_t i[TI_ICOUNT]; };
struct ColdStruct {
    TrackCold &c;
    EMC_HD explicit ColdStruct(TrackCold &r) : c(r) {}
    EMC_HD double getd(int f) const { return c.d[f]; }
    EMC_HD void setd(int f, double v) const { c.d[f] = v; }
    EMC_HD int32_t geti(int f) const { return c.i[f]; }
    EMC_HD void seti(int f, int32_t v) const { c.i[f] = v; }
};
struct Track { TrackHot h; TrackCold c; };      /* both parts side by side (test seam, register / shared-memory variants) */

template <class CA>
EMC_HD void track_init(TrackHot &K, const CA &C, const State &s, double t_rail)
{
    K.t = t_rail; K.t_rail = t_rail;
    K.apogee_alt = s.z; C.setd(TC_APOGEE_T, t_rail); C.seti(TI_APOGEE_INDEX, 0);
    K.first_nan = (s.z != s.z) ? 0 : -1;
    C.setd(TC_LATCH, 0.0); C.setd(TC_MAX_COAST, 0.0);
    C.setd(TC_BURNOUT_TIME, 0.0); K.burnout_found = false;
    C.setd(TC_CHUTE_TIME, NAN); K.chute = false; K.apogee_detected = false;
    K.n_steps = 0; K.term = EMC_TERM_NONE; K.finishing = false; K.replay = 0; K.om_half = 0;
    C.seti(TI_BT_SLOT, -1); C.seti(TI_BT_NEXT, 0);
    C.setd(TC_MAX_MACH2, -INFINITY); C.setd(TC_MAX_Q, -INFINITY); C.setd(TC_MAX_V2, -INFINITY); C.setd(TC_MAX_OM, -INFINITY);
    C.setd(TC_MIN_STAB, INFINITY); C.setd(TC_MAX_STAB, -INFINITY); C.setd(TC_MAX_AOA, -INFINITY);
}

/* running np.max / np.min of one cold field: NaN propagates and sticks (a NaN maximum fails both tests) */
template <class CA> EMC_HD void cold_max(const CA &C, int f, double v) { const double m = C.getd(f); if (v > m || v != v) C.setd(f, v); }
template <class CA> EMC_HD void cold_min(const CA &C, int f, double v) { const double m = C.getd(f); if (v < m || v != v) C.setd(f, v); }

/* diagnostics of a stored state (stage 0 of the next step, or the final extra evaluation) */
template <class CA>
EMC_HD void track_diag(const CA &C, const State &s, const Diag &d)
{
    cold_max(C, TC_MAX_MACH2, d.mach2);
    cold_max(C, TC_MAX_Q, d.qdyn);
    cold_max(C, TC_MAX_V2, s.vx * s.vx + s.vy * s.vy + s.vz * s.vz);
    double om = fabs(s.wx);
    np_max_acc(om, fabs(s.wy));
    np_max_acc(om, fabs(s.wz));
    cold_max(C, TC_MAX_OM, om);
    cold_min(C, TC_MIN_STAB, d.stab);
    cold_max(C, TC_MAX_STAB, d.stab);
    cold_max(C, TC_MAX_AOA, d.abs_aoa);
}

/* After an accepted step: t already advanced, s is the new stored state (:229-264).
 * Returns true when the loop of :216 ends (break or guard). */
template <class CA>
EMC_HD bool track_post_step(const DevModel &M, const Sample &S, TrackHot &K, const CA &C, const State &s)
{
    K.n_steps += 1;
    const double z = s.z, vz = s.vz;
    if (!(K.apogee_alt != K.apogee_alt) && ((z != z) || z > K.apogee_alt)) {
        K.apogee_alt = z; C.seti(TI_APOGEE_INDEX, K.n_steps); C.setd(TC_APOGEE_T, K.t);
    }
    if (K.first_nan < 0 && (z != z)) K.first_nan = K.n_steps;
    if (!K.burnout_found && (K.t - K.t_rail) > S.burn_time) { K.burnout_found = true; C.setd(TC_BURNOUT_TIME, K.t - K.t_rail); }
    if (z <= 0.5 && vz <= 0.0) { K.term = EMC_TERM_GROUND; return true; }
    if (z > 100000.0) { K.term = EMC_TERM_ALTITUDE; return true; }
    if (z > 1000.0 && vz < 0.0 && !K.apogee_detected) {
        K.apogee_detected = true;
        C.setd(TC_LATCH, K.t);
        C.setd(TC_MAX_COAST, (z > 50000.0) ? 60.0 : ((z > 25000.0) ? 120.0 : 300.0));
    }
    if (K.apogee_detected && z > 25000.0) {
        if (K.t - C.getd(TC_LATCH) > C.getd(TC_MAX_COAST)) { K.term = EMC_TERM_COAST; return true; }
    }
    if (!(K.t < M.max_time)) { K.term = EMC_TERM_MAX_TIME; return true; }
    return false;
}

/* Where a lane keeps its base state s and the RK4 accumulator between stages.  RegStore: registers
 * (host seam, and the register-resident kernel variants).  The flight kernel can instead keep them in
 * shared memory (SharedStore in emc_engine.cu): they are touched only at stage boundaries, and the ~56
 * registers they would pin for the whole derivative are better spent on instruction-level parallelism.
 * The 14 state words are moved as 7 PAIRS (Pair = 16 bytes: one LDS.128 / STS.128 per pair on the device). */
struct alignas(16) Pair { double a, b; };
struct RegStore {
    State s_, a_;
    EMC_HD double s(int i) const { return reinterpret_cast<const double *>(&s_)[i]; }
    EMC_HD void set_s(int i, double v) { reinterpret_cast<double *>(&s_)[i] = v; }
    EMC_HD Pair s2(int p) const { Pair r; r.a = s(2 * p); r.b = s(2 * p + 1); return r; }
    EMC_HD void set_s2(int p, Pair v) { set_s(2 * p, v.a); set_s(2 * p + 1, v.b); }
    EMC_HD Pair acc2(int p) const { const double *q = reinterpret_cast<const double *>(&a_); Pair r; r.a = q[2 * p]; r.b = q[2 * p + 1]; return r; }
    EMC_HD void set_acc2(int p, Pair v) { double *q = reinterpret_cast<double *>(&a_); q[2 * p] = v.a; q[2 * p + 1] = v.b; }
};

template <class Store>
EMC_HD void store_put(Store &st, const State &s)
{
    const double *p = reinterpret_cast<const double *>(&s);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 7; ++i) { Pair v; v.a = p[2 * i]; v.b = p[2 * i + 1]; st.set_s2(i, v); }
}
template <class Store>
EMC_HD void store_get(const Store &st, State &s)
{
    double *p = reinterpret_cast<double *>(&s);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 7; ++i) { const Pair v = st.s2(i); p[2 * i] = v.a; p[2 * i + 1] = v.b; }
}

/* One classical RK4 step (simulator.py:217-229): the four stages share ONE copy of the derivative.
 * Stage 0 doubles as the diagnostics pass of the stored state s (simulator.py:511-552 evaluates the
 * same quantities at the same state).  A lane whose loop has ended (K.finishing) runs stage 0 only,
 * for the diagnostics of its last stored state, with the sticky flag protected: the reference never
 * evaluates the derivative there.  Returns true if a full step was taken. */
template <class Store, class CA, int MK = -1, int WK = -1, bool REG = false>
EMC_HD bool rk4_step(const DevModel &M, const DevTables &Tb, const double *wind_alt, const Sample &S,
                     WindBracket &WB, TrackHot &K, const CA &C, Store &st)
{
    State ys, k;
    double *y = reinterpret_cast<double *>(&ys);
    const double *kk = reinterpret_cast<const double *>(&k);
    store_get(st, ys);
    Diag dg;
    dg.mach2 = dg.qdyn = dg.abs_aoa = dg.stab = 0.0;
    const bool fin = K.finishing;
    /* the sticky parachute flag (F12) is carried in registers through the four stages and written back once */
    const bool chute_keep = K.chute;
    bool chute = chute_keep;
    double chute_time = 0.0;                        /* only read back when a stage of THIS step latches the flag */
    /* the stage times of this step lie in [t, t + dt]: if the remembered thrust-curve bracket holds both ends, the
     * four stages skip the bracket test (the time axis is the one table argument that is known in advance) */
    bool th_safe = false;
    if (cfg_solid<MK>(M)) { const int j = WB.j_th; th_safe = (K.t >= Tb.th_lo[j]) && (K.t + M.dt < Tb.th_hi[j]); }
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int stage = 0; stage < 4; ++stage) {
        const double ts = K.t + M.stage_t[stage];      /* warp-uniform offset */
        derivative<MK, WK, REG>(M, Tb, wind_alt, S, WB, ts, ys, chute, chute_time, k, stage == 0, dg, th_safe);
        if (stage == 0) {
            track_diag(C, ys, dg);                      /* ys == s at stage 0 */
            if (fin) return false;                      /* diagnostic pass only: a latch by this evaluation is dropped */
        }
        if (stage < 3) {
            /* acc = k1 + 2 k2 + 2 k3 (+ k4 below), in the reference's order ((k1 + 2k2) + 2k3) + k4 (:224);
             * next stage state = s + c k  (:218-222) */
            const double c = M.stage_c[stage];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int i = 0; i < 7; ++i) {
                const Pair sv = st.s2(i);
                y[2 * i] = fma(c, kk[2 * i], sv.a); y[2 * i + 1] = fma(c, kk[2 * i + 1], sv.b);
            }
            if (stage == 0) {                         /* warp-uniform: stage 0 only stores, stages 1-2 accumulate */
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int i = 0; i < 7; ++i) { Pair v; v.a = kk[2 * i]; v.b = kk[2 * i + 1]; st.set_acc2(i, v); }
            } else {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int i = 0; i < 7; ++i) {
                    Pair v = st.acc2(i);
                    v.a = fma(2.0, kk[2 * i], v.a); v.b = fma(2.0, kk[2 * i + 1], v.b);
                    st.set_acc2(i, v);
                }
            }
        } else {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int i = 0; i < 7; ++i) {
                const Pair av = st.acc2(i);
                Pair sv = st.s2(i);
                sv.a = fma(M.dt_over_6, av.a + kk[2 * i], sv.a); sv.b = fma(M.dt_over_6, av.b + kk[2 * i + 1], sv.b);
                if (i == 3 || i == 4) { y[2 * i] = sv.a; y[2 * i + 1] = sv.b; }      /* q0..q3: renormalised below */
                else st.set_s2(i, sv);
            }
        }
    }
    /* :227 renormalise (q0..q3 = words 6..9 = pairs 3 and 4, still in registers) */
    {
        const double q0 = y[6], q1 = y[7], q2 = y[8], q3 = y[9];
        const double n2 = q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3;
        Pair u, v;
        if (n2 > 1e-24) { const double rn = fast_rsqrt(n2); u.a = q0 * rn; u.b = q1 * rn; v.a = q2 * rn; v.b = q3 * rn; }
        else { u.a = 1.0; u.b = 0.0; v.a = 0.0; v.b = 0.0; }
        st.set_s2(3, u); st.set_s2(4, v);
    }
    if (chute != chute_keep) { K.chute = true; C.setd(TC_CHUTE_TIME, chute_time); }
    K.t += M.dt;
    return true;
}

/* NaN fast-forward.  Every break test of simulator.py:238-264 reads the altitude, so
 * once z (and vz) are NaN the loop can only end at the guard `t < max_time` (:216); the reference
 * grinds through ~57 k more steps.  Two cases let the engine replay that tail without the derivative:
 *   mode 1  x, y, vx, vy are NaN as well: nothing but t changes any more.
 *   mode 2  the motor is off for good and the attitude is finite: the body force is exactly zero (no
 *           thrust; q_inf = NaN fails `q_dynamic > 0`, :378; a deployed parachute sees |v_body| = NaN
 *           and fails `rel_speed > 0`, :374; a NaN altitude cannot latch the chute, :366-369), so
 *           vx, vy keep their values (finite, +-inf or NaN) and x, y advance by the RK4 update of :224
 *           with k = (vx, vy) in all four stages; the damped body rates stay finite.  Only x, y and t
 *           are replayed, with the same IEEE arithmetic the full step would apply to them.
 * Anything else (still burning, non-finite attitude) turns fully NaN within a few steps and is
 * integrated until it reaches mode 1.  Apogee (first NaN, :488), the NaN maxima and the step count
 * are what the reference produces; max|omega| keeps its value at the fast-forward point (NaN runs:
 * category parity). */

EMC_HD int nan_mode(const DevModel &M, const Sample &S, const TrackHot &K, const State &s)
{
    if (!((s.z != s.z) && (s.vz != s.vz))) return 0;
    if (!(K.t + 2.0 * M.dt < M.max_time)) return 0;          /* too close to the guard: just integrate */
    if ((s.x != s.x) && (s.y != s.y) && (s.vx != s.vx) && (s.vy != s.vy)) return 1;
    const double pf = (s.pf > 0.0) ? s.pf : 0.0;
    const bool burning = (pf > 0.0) && (K.t <= S.burn_time);
    if (burning) return 0;
    const bool att_ok = finite_d(s.q0) && finite_d(s.q1) && finite_d(s.q2) && finite_d(s.q3) &&
                        finite_d(s.wx) && finite_d(s.wy) && finite_d(s.wz) && finite_d(s.pf) &&
                        (s.q0 * s.q0 + s.q1 * s.q1 + s.q2 * s.q2 + s.q3 * s.q3 > 1e-24);
    return att_ok ? 2 : 0;
}

/* Plain replay of `t += dt` (:229) with the burnout-index test of :479-480 (reads only the time). */
template <class CA>
EMC_HD int64_t replay_time_loop(const DevModel &M, const Sample &S, TrackHot &K, const CA &C, int64_t max_iter)
{
    int64_t n = 0;
    while (K.t < M.max_time && n < max_iter) {
        K.t += M.dt; K.n_steps += 1; ++n;
        if (!K.burnout_found && (K.t - K.t_rail) > S.burn_time) { K.burnout_found = true; C.setd(TC_BURNOUT_TIME, K.t - K.t_rail); }
    }
    return n;
}

/* 2^e as a double, e in the normal range */
EMC_HD double pow2_d(int e)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)(e + 1023) << 52);
#else
    return ldexp(1.0, e);
#endif
}
EMC_HD int exponent_d(double v)    /* floor(log2(v)) for normal positive v */
{
#if defined(__CUDA_ARCH__)
    return (int)((__double_as_longlong(v) >> 52) & 0x7ff) - 1023;
#else
    int e; frexp(v, &e); return e - 1;
#endif
}

/*
 * The same replay in closed form, bit-exact.  While t stays inside one binade [2^e, 2^(e+1)) every
 * t_k is a multiple of ulp_e, so fl(t + dt) = t + d with ONE constant d = fl(t + dt) - t (unless
 * dt sits exactly on a rounding tie, which is detected and sent to the plain loop).  Whole binades are
 * therefore advanced with integer arithmetic (~10 segments for t in [0.5, 300]) instead of ~57 k
 * dependent additions that would stall the other 31 lanes of the warp; the step that crosses a
 * binade edge is taken as a real floating-point addition.  Mode 2 also advances x, y by
 * n * (dt/6 * acc): 1e-16-level difference to the step-by-step sum, far inside the parity bar.
 */
template <class CA>
EMC_HD int64_t replay_time(const DevModel &M, const Sample &S, TrackHot &K, const CA &C, State &s)
{
    int64_t n = 0;
    const double t_begin = K.t;
    while (K.t < M.max_time) {
        const double t = K.t;
        const double t1 = t + M.dt;
        int64_t seg = 0;          /* steps that can be taken in closed form inside this binade */
        double d = 0.0;
        if (t >= 4.450147717014403e-308 && t1 > t && finite_d(t1)) {
            const int e = exponent_d(t);
            const double hi = pow2_d(e + 1);
            if (t1 < hi) {
                d = t1 - t;                                   /* exact: both are multiples of ulp_e */
                const double ulp = pow2_d(e - 52);
                const double resid = M.dt - d;                /* exact low part of dt */
                if (fabs(resid) != 0.5 * ulp) {               /* not a rounding tie */
                    const double sc = pow2_d(52 - e);         /* 1/ulp: scaling by 2^k is exact */
                    const long long D = (long long)(d * sc);
                    const long long A = (long long)((hi - t) * sc);
                    long long jb = (A + D - 1) / D - 1;       /* largest j with t + j*d < hi */
                    if (M.max_time < hi) {
                        const long long G = (long long)((M.max_time - t) * sc);
                        const long long jg = (G + D - 1) / D; /* smallest j with t + j*d >= max_time */
                        if (jg < jb) jb = jg;
                    }
                    seg = jb;
                }
            }
        }
        if (seg < 1) {            /* binade crossing, tie, or degenerate operands: one real step */
            if (!(t1 > t)) { break; }                         /* dt no longer advances t (reference would spin) */
            n += replay_time_loop(M, S, K, C, 1);
            continue;
        }
        if (!K.burnout_found && (S.burn_time == S.burn_time)) {
            /* first j in [1, seg] with fl((t + j*d) - t_rail) > burn_time; the left side is monotone in j */
            double est = ceil((S.burn_time + K.t_rail - t) / d);
            long long j = (est < 1.0) ? 1 : ((est > (double)seg) ? seg : (long long)est);
            while (j > 1 && ((t + (double)(j - 1) * d) - K.t_rail) > S.burn_time) --j;
            while (j <= seg && !(((t + (double)j * d) - K.t_rail) > S.burn_time)) ++j;
            if (j <= seg) { K.burnout_found = true; C.setd(TC_BURNOUT_TIME, (t + (double)j * d) - K.t_rail); }
        }
        K.t = t + (double)seg * d;                            /* exact */
        K.n_steps += (int32_t)seg;
        n += seg;
    }
    if (K.replay == 2) {
        const double ax = ((s.vx + 2.0 * s.vx) + 2.0 * s.vx) + s.vx;     /* :224 with k = vx in all stages */
        const double ay = ((s.vy + 2.0 * s.vy) + 2.0 * s.vy) + s.vy;
        s.x += (double)n * (M.dt_over_6 * ax);
        s.y += (double)n * (M.dt_over_6 * ay);
    }
    (void)t_begin;
    K.term = EMC_TERM_MAX_TIME;
    return n;
}

/* One scheduling quantum of a lane, shared by the flight kernel and the test seam: a full RK4 step
 * plus the event logic, or the closing diagnostics pass.  Returns true when the lane retires (its
 * outputs are final).  `stepped` reports whether a stored state was produced (tape). */
/* A stored state that shows the reference's blow-up in its last steps (|v| > 1e7 m/s or |omega| > 1000 rad/s;
 * the speed then goes 1e7 -> 1e9 -> 1e80 -> overflow): the trajectory leaves the fast path here — every value still far
 * from overflow — and is finished by the strict continuation (emc_strict.cuh), whose operation order reproduces the
 * reference's inf / NaN patterns.  Measured (tests/, round 2): at most 4 strict steps per flight with these thresholds,
 * integer outputs identical to the reference on 50 000 of 50 000 samples; earlier thresholds (1e6 m/s, 100 rad/s: up to
 * 134 steps) change nothing but the cost. */
#ifndef EMC_STRICT_V2
#define EMC_STRICT_V2 1e14
#define EMC_STRICT_OMEGA 1000.0
#endif
#define EMC_REPLAY_PARK 3            /* TrackHot.replay: parked for the strict continuation */
EMC_HD bool strict_trigger(const State &s)
{
    const double v2 = s.vx * s.vx + s.vy * s.vy + s.vz * s.vz;
    return (v2 > EMC_STRICT_V2) || (fabs(s.wx) > EMC_STRICT_OMEGA) || (fabs(s.wy) > EMC_STRICT_OMEGA) || (fabs(s.wz) > EMC_STRICT_OMEGA);
}

template <class Store, class CA, int MK = -1, int WK = -1, bool REG = false>
EMC_HD bool lane_advance(const DevModel &M, const DevTables &Tb, const double *wind_alt, const Sample &S,
                         WindBracket &WB, TrackHot &K, const CA &C, Store &st, bool nan_ff, bool &stepped, int64_t &replayed,
                         bool park = false)
{
    stepped = rk4_step<Store, CA, MK, WK, REG>(M, Tb, wind_alt, S, WB, K, C, st);
    if (!stepped) {                       /* closing pass done */
        if (K.replay) {
            State s; store_get(st, s);
            replayed += replay_time(M, S, K, C, s);
            st.set_s(0, s.x); st.set_s(1, s.y);
        }
        return true;
    }
    State s; store_get(st, s);
    bool done = track_post_step(M, S, K, C, s);
    if (!done && park && strict_trigger(s)) { K.replay = EMC_REPLAY_PARK; return true; }      /* retire into the strict queue */
    if (!done && nan_ff) { K.replay = (int8_t)nan_mode(M, S, K, s); done = (K.replay != 0); }
    K.finishing = done;
    return false;
}

/* write the flight part of the summary (strided SoA column) */
template <class CA>
EMC_HD void write_flight_outputs(const TrackHot &K, const CA &C, const State &s, double *out, int32_t *iout, int64_t ld)
{
    out[EMC_OUT_APOGEE_ALTITUDE * ld] = K.apogee_alt;
    out[EMC_OUT_APOGEE_TIME * ld] = C.getd(TC_APOGEE_T) - K.t_rail;
    out[EMC_OUT_RANGE * ld] = sqrt(s.x * s.x + s.y * s.y);
    out[EMC_OUT_FLIGHT_TIME * ld] = K.t - K.t_rail;
    out[EMC_OUT_FINAL_X * ld] = s.x; out[EMC_OUT_FINAL_Y * ld] = s.y; out[EMC_OUT_FINAL_Z * ld] = s.z;
    out[EMC_OUT_FINAL_VX * ld] = s.vx; out[EMC_OUT_FINAL_VY * ld] = s.vy; out[EMC_OUT_FINAL_VZ * ld] = s.vz;
    out[EMC_OUT_MAX_MACH * ld] = sqrt(C.getd(TC_MAX_MACH2));
    out[EMC_OUT_MAX_Q * ld] = C.getd(TC_MAX_Q);
    out[EMC_OUT_MAX_SPEED * ld] = sqrt(C.getd(TC_MAX_V2));
    out[EMC_OUT_MAX_ABS_OMEGA * ld] = C.getd(TC_MAX_OM);
    out[EMC_OUT_MIN_STABILITY * ld] = C.getd(TC_MIN_STAB);
    out[EMC_OUT_MAX_STABILITY * ld] = C.getd(TC_MAX_STAB);
    out[EMC_OUT_MAX_ABS_AOA * ld] = C.getd(TC_MAX_AOA);
    out[EMC_OUT_BURNOUT_TIME * ld] = C.getd(TC_BURNOUT_TIME);
    out[EMC_OUT_CHUTE_TIME * ld] = C.getd(TC_CHUTE_TIME);
    iout[EMC_IOUT_N_STEPS * ld] = K.n_steps;
    iout[EMC_IOUT_TERMINATION * ld] = K.term;
    iout[EMC_IOUT_APOGEE_INDEX * ld] = C.geti(TI_APOGEE_INDEX);
    iout[EMC_IOUT_FIRST_NAN_STEP * ld] = K.first_nan;
}

/* ------------------------------------------------------------------------------------------------
 * Sample loading and the rail phase (simulator.py:42-125)
 * ---------------------------------------------------------------------------------------------- */
EMC_HD void load_sample(const DevModel &M, const double *col, int64_t ld, const double *wind, Sample &S)
{
    S.dry_mass = col[EMC_IN_DRY_MASS * ld];
    S.prop_mass = col[EMC_IN_PROP_MASS * ld];
    S.dry_cg = S.dry_mass * M.cg_dry;
    S.thrust_a = col[EMC_IN_THRUST_A * ld];
    S.nozzle_area = col[EMC_IN_NOZZLE_AREA * ld];
    S.burn_time = col[EMC_IN_BURN_TIME * ld];
    S.pf_rate = -col[EMC_IN_MDOT * ld] / S.prop_mass;
    S.cd_scale = col[EMC_IN_CD_SCALE * ld];
    S.wind = wind;
}

/* see derivative<., ., REG>: what the flight kernel's fast path assumes about a sample (and about the inertia constants of
 * the model: Ixx = Ixx_dry + mp d^2/4, Iyy = Iyy_dry + mp (L^2/12 + dcg^2) with mp >= 0) */
EMC_HD bool sample_regular(const DevModel &M, const Sample &S)
{
    return (S.dry_mass > 0.0) && finite_d(S.dry_mass) && (S.prop_mass >= 0.0) && finite_d(S.prop_mass) &&
           (M.Ixx_dry > 0.0) && (M.Iyy_dry > 0.0) && (M.d4sq >= 0.0) && (M.len2_12 >= 0.0) && finite_d(M.prop_cg) && finite_d(M.cg_dry) &&
           finite_d(M.Ixx_dry) && finite_d(M.Iyy_dry) && finite_d(M.d4sq) && finite_d(M.len2_12);
}

EMC_HD void wind_bracket_reset(WindBracket &B)
{
    B.lo = 1.0; B.hi = 0.0; B.x0 = 0.0;     /* empty interval: first use loads */
    B.f0[0] = B.f0[1] = B.f0[2] = 0.0; B.s[0] = B.s[1] = B.s[2] = 0.0;
    B.j_m = 1; B.j_th = 1; B.j_atm = 0;
}

/* motor.py:86-93 */
EMC_HD double propellant_remaining(const Sample &S, double t)
{
    if (t <= 0.0) return 1.0;
    if (t >= S.burn_time) return 0.0;
    double r = 1.0 - t / S.burn_time;
    return (r > 0.0) ? r : 0.0;
}

/* Rail phase for one sample; writes the rail_* outputs and returns the number of Euler steps.
 * The state at rail exit is (out[RAIL_EXIT_X..VZ], q and omega unchanged, pf = remaining(t)). */
EMC_HD int rail_phase(const DevModel &M, const DevTables &Tb, const double *wind_alt, const Sample &S,
                      const double *col, int64_t ld, double *out, int64_t old)
{
    double px = col[EMC_IN_X * ld], py = col[EMC_IN_Y * ld], pz = col[EMC_IN_Z * ld];
    double vx = col[EMC_IN_VX * ld], vy = col[EMC_IN_VY * ld], vz = col[EMC_IN_VZ * ld];
    const double q0 = col[EMC_IN_Q0 * ld], q1 = col[EMC_IN_Q1 * ld], q2 = col[EMC_IN_Q2 * ld], q3 = col[EMC_IN_Q3 * ld];
    double n2 = q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3;
    double qw, qx, qy, qz;
    if (n2 > 1e-24) { double rn = 1.0 / sqrt(n2); qw = q0 * rn; qx = q1 * rn; qy = q2 * rn; qz = q3 * rn; }
    else { qw = 1.0; qx = 0.0; qy = 0.0; qz = 0.0; }
    /* the factors of two are folded into one operand (exact), so 2*(a*b - c*d) costs one product and one FMA */
    const double x2 = qx + qx, y2 = qy + qy, z2 = qz + qz;
    const double r00 = 1.0 - (qy * y2 + qz * z2), r01 = qx * y2 - qw * z2, r02 = qx * z2 + qw * y2;
    const double r10 = qx * y2 + qw * z2, r11 = 1.0 - (qx * x2 + qz * z2), r12 = qy * z2 - qw * x2;
    const double r20 = qx * z2 - qw * y2, r21 = qy * z2 + qw * x2, r22 = 1.0 - (qx * x2 + qy * y2);
    const double dx = r00, dy = r10, dz = r20;                /* :57 body x in inertial axes */
    WindBracket WB;
    wind_bracket_reset(WB);
    double dist = 0.0, t = 0.0, pf = 1.0;
    const double dt = M.dt_rail;
    int steps = 0;
    while (dist < M.rail_length && t < S.burn_time) {          /* :63 */
        const double pfc = pf;
        const double mp = S.prop_mass * pfc;
        const double mass = S.dry_mass + mp;
        double T, inv_RT, p;
        atmosphere(M, pz, T, inv_RT, p);
        const double rho = p * inv_RT;
        double w[3];
        wind_at(M, wind_alt, S, pz, WB, w);
        double speed = vx * dx + vy * dy + vz * dz;            /* :75 */
        const double rx = dx * speed - w[0], ry = dy * speed - w[1], rz = dz * speed - w[2];
        const double rel_speed = rx * dx + ry * dy + rz * dz;  /* :80 */
        const double mach = fast_sqrt((rx * rx + ry * ry + rz * rz) * (inv_RT * M.mach_k));
        const int jd = brk_find(Tb.m_lo, Tb.m_hi, M.n_mb, WB.j_m, mach);
        WB.j_m = jd;
        const double dm = mach - Tb.cd_x0[jd];
        const double cd = fma(Tb.cd0_s[jd], dm, Tb.cd0_f[jd]) * S.cd_scale
                          + fma(Tb.cda_s[jd], dm, Tb.cda_f[jd]) * 0.0;                 /* alpha = 0, :82-83 */
        const double drag = 0.5 * rho * (rel_speed * rel_speed) * cd * M.ref_area;     /* :84 */
        const double thrust = thrust_at(M, Tb, S, WB, t, p);   /* :86 */
        const double g = gravity(M, pz);
        const double accel = (thrust - mass * g - drag) / mass; /* :88 */
        speed += accel * dt;
        px += dx * speed * dt; py += dy * speed * dt; pz += dz * speed * dt;
        dist += speed * dt;
        vx = dx * speed; vy = dy * speed; vz = dz * speed;
        t += dt;
        pf = propellant_remaining(S, t);                       /* :96 */
        ++steps;
    }
    /* :103-123 */
    out[EMC_OUT_RAIL_EXIT_TIME * old] = t;
    out[EMC_OUT_RAIL_EXIT_X * old] = px; out[EMC_OUT_RAIL_EXIT_Y * old] = py; out[EMC_OUT_RAIL_EXIT_Z * old] = pz;
    out[EMC_OUT_RAIL_EXIT_VX * old] = vx; out[EMC_OUT_RAIL_EXIT_VY * old] = vy; out[EMC_OUT_RAIL_EXIT_VZ * old] = vz;
    out[EMC_OUT_RAIL_EXIT_SPEED * old] = sqrt(vx * vx + vy * vy + vz * vz);
    {   /* utils.py:139-144,46-69 on the RAW quaternion */
        const double x = q1, y = q2, z = q3, ww = q0;
        out[EMC_OUT_RAIL_EXIT_ROLL * old] = atan2(2.0 * (ww * x + y * z), 1.0 - 2.0 * (x * x + y * y));
        const double sinp = 2.0 * (ww * y - z * x);
        out[EMC_OUT_RAIL_EXIT_PITCH * old] = (fabs(sinp) >= 1.0) ? copysign(1.5707963267948966, sinp) : asin(sinp);
        out[EMC_OUT_RAIL_EXIT_YAW * old] = atan2(2.0 * (ww * z + x * y), 1.0 - 2.0 * (y * y + z * z));
    }
    double w[3];
    wind_at(M, wind_alt, S, pz, WB, w);
    const double ux = vx - w[0], uy = vy - w[1], uz = vz - w[2];
    const double vbx = r00 * ux + r10 * uy + r20 * uz;
    const double vby = r01 * ux + r11 * uy + r21 * uz;
    const double vbz = r02 * ux + r12 * uy + r22 * uz;
    out[EMC_OUT_RAIL_EXIT_AOA * old] = ((fabs(vbx) < 1e-6) && (fabs(vbz) < 1e-6)) ? 0.0 : atan2(vbz, vbx);
    const double vxz = sqrt(vbx * vbx + vbz * vbz);
    out[EMC_OUT_RAIL_EXIT_SIDESLIP * old] = (vxz < 1e-6) ? 0.0 : atan2(vby, vxz);
    out[EMC_OUT_WIND_AT_EXIT_U * old] = w[0]; out[EMC_OUT_WIND_AT_EXIT_V * old] = w[1]; out[EMC_OUT_WIND_AT_EXIT_W * old] = w[2];
    return steps;
}

/* State at rail exit, rebuilt from the inputs and the rail outputs (simulator.py:98-100,161) */
EMC_HD void load_flight_state(const Sample &S, const double *col, int64_t ld, const double *out, int64_t old,
                              State &s, double &t_rail)
{
    t_rail = out[EMC_OUT_RAIL_EXIT_TIME * old];
    s.x = out[EMC_OUT_RAIL_EXIT_X * old]; s.y = out[EMC_OUT_RAIL_EXIT_Y * old]; s.z = out[EMC_OUT_RAIL_EXIT_Z * old];
    s.vx = out[EMC_OUT_RAIL_EXIT_VX * old]; s.vy = out[EMC_OUT_RAIL_EXIT_VY * old]; s.vz = out[EMC_OUT_RAIL_EXIT_VZ * old];
    s.q0 = col[EMC_IN_Q0 * ld]; s.q1 = col[EMC_IN_Q1 * ld]; s.q2 = col[EMC_IN_Q2 * ld]; s.q3 = col[EMC_IN_Q3 * ld];
    s.wx = col[EMC_IN_WX * ld]; s.wy = col[EMC_IN_WY * ld]; s.wz = col[EMC_IN_WZ * ld];
    s.pf = propellant_remaining(S, t_rail);
}

/* rocket.py:138-218 as a standalone function (series extraction and the component test seam; the hot derivative has
 * its own fused form).  c = {cd, cl, cm(=cpitch), cy, cyaw, cp, cn}.  `mach` is already clamped to 1e300. */
EMC_HD void aero_coefficients(const DevModel &M, const DevTables &Tb, double mach, double mach2, double alpha, double beta,
                              double cg, bool power_on, double cd_scale, double c[7])
{
    const int jc = brk_find(Tb.m_lo, Tb.m_hi, M.n_mb, 1, mach);
    const double cp = M.cp_location + fma(Tb.cp_s[jc], mach - Tb.cp_x0[jc], Tb.cp_f[jc]);
    const int jd = jc;
    const double dm = mach - Tb.cd_x0[jd];
    const double cd0 = fma(Tb.cd0_s[jd], dm, Tb.cd0_f[jd]) * cd_scale;
    const double cda = fma(Tb.cda_s[jd], dm, Tb.cda_f[jd]);
    double cd = cd0 + cda * (alpha * alpha);
    if (!power_on) cd *= M.power_off_factor;
    const double rad = 4.0 + M.AR_over_cos2 * fabs(1.0 - mach2);
    const double cl_alpha = M.two_pi_AR_cos / (2.0 + fast_sqrt(rad));
    double cl = cl_alpha * alpha, cy = cl_alpha * beta, cn = cl_alpha * alpha;
    const double abs_alpha = fabs(alpha);
    if (abs_alpha > M.stall_angle) {
        const double over = (abs_alpha - M.stall_angle) * M.inv_stall_span;
        double sf = 1.0 - over;
        sf = (sf > 0.0) ? sf : 0.0;
        const double sgn = (alpha > 0.0) ? 1.0 : ((alpha < 0.0) ? -1.0 : alpha);
        cl = cl_alpha * M.stall_angle * sf * sgn;
        cn = cl;
        cd *= 1.0 + 0.5 * over;
        cy *= sf;
    }
    const double sm = cp - cg;
    c[0] = cd; c[1] = cl; c[2] = -cl_alpha * sm * alpha; c[3] = cy; c[4] = -cl_alpha * sm * beta; c[5] = cp; c[6] = cn;
}

/* component evaluation for one element (emc_component_debug / host seam): in/out are field-major with stride ld */
EMC_HD void component_eval(const DevModel &M, const DevTables &Tb, int comp, const double *in, double *out, int64_t ld)
{
    if (comp == 0) {
        double T, inv_RT, p;
        atmosphere(M, in[0], T, inv_RT, p);
        out[0] = T; out[ld] = p; out[2 * ld] = p * inv_RT; out[3 * ld] = sqrt(M.gamma * M.R_gas * T); out[4 * ld] = gravity(M, in[0]);
    } else if (comp == 1) {
        const double pf = in[0], dry = in[ld], prop = in[2 * ld];
        const double mp = prop * pf, mass = dry + mp;
        const double cg = (dry * M.cg_dry + mp * M.prop_cg) / mass, dcg = M.prop_cg - cg;
        out[0] = mass; out[ld] = cg; out[2 * ld] = M.Ixx_dry + mp * M.d4sq;
        out[3 * ld] = M.Iyy_dry + mp * (M.len2_12 + dcg * dcg); out[4 * ld] = out[3 * ld];
    } else if (comp == 2) {
        double c[7];
        const double mach = in[0], mc = (mach > 1e300) ? 1e300 : mach;
        aero_coefficients(M, Tb, mc, mach * mach, in[ld], in[2 * ld], in[3 * ld], in[4 * ld] != 0.0, in[5 * ld], c);
        for (int k = 0; k < 7; ++k) out[k * ld] = c[k];
    } else {
        Sample S;
        S.dry_mass = S.prop_mass = S.dry_cg = S.pf_rate = 0.0; S.cd_scale = 1.0; S.wind = nullptr;
        S.thrust_a = in[2 * ld]; S.nozzle_area = in[3 * ld]; S.burn_time = in[4 * ld];
        WindBracket B; wind_bracket_reset(B);
        out[0] = thrust_at(M, Tb, S, B, in[0], in[ld]);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Per-stored-state result series, simulator.py:511-552 (_extract_results): one call per stored state.
 * `t_shift` is time[i] = t - t_rail: the reference evaluates the thrust history at the SHIFTED time (:543).
 * ---------------------------------------------------------------------------------------------- */
EMC_HD void series_state(const DevModel &M, const DevTables &Tb, const double *wind_alt, const Sample &S,
                         const double *row /*t, state[14]*/, double t_shift, double *out, int64_t ld)
{
    State s;
    s.x = row[1]; s.y = row[2]; s.z = row[3]; s.vx = row[4]; s.vy = row[5]; s.vz = row[6];
    s.q0 = row[7]; s.q1 = row[8]; s.q2 = row[9]; s.q3 = row[10]; s.wx = row[11]; s.wy = row[12]; s.wz = row[13]; s.pf = row[14];
    /* :512 euler angles of the stored quaternion, utils.py:139-144,46-69 */
    {
        const double x = s.q1, y = s.q2, z = s.q3, w = s.q0;
        out[EMC_SER_EULER_ROLL * ld] = fast_atan2(2.0 * (w * x + y * z), 1.0 - 2.0 * (x * x + y * y));
        const double sinp = 2.0 * (w * y - z * x);
        out[EMC_SER_EULER_PITCH * ld] = (fabs(sinp) >= 1.0) ? copysign(K_MISC[0], sinp) : asin(sinp);
        out[EMC_SER_EULER_YAW * ld] = fast_atan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z));
    }
    /* :515-520 mass properties at the UNCLAMPED propellant fraction, rocket.py:110-136 */
    const double mp = S.prop_mass * s.pf;
    const double mass = S.dry_mass + mp;
    const double cg = (S.dry_cg + mp * M.prop_cg) / mass;
    const double dcg = M.prop_cg - cg;
    const double Ixx = M.Ixx_dry + mp * M.d4sq;
    const double Iyy = M.Iyy_dry + mp * (M.len2_12 + dcg * dcg);
    out[EMC_SER_MASS * ld] = mass; out[EMC_SER_CENTER_OF_MASS * ld] = cg;
    out[EMC_SER_IXX * ld] = Ixx; out[EMC_SER_IYY * ld] = Iyy; out[EMC_SER_IZZ * ld] = Iyy;
    /* :522-534 */
    double T, inv_RT, p;
    atmosphere(M, s.z, T, inv_RT, p);
    const double rho = p * inv_RT;
    WindBracket WB; wind_bracket_reset(WB);
    double w[3];
    wind_at(M, wind_alt, S, s.z, WB, w);
    const double ux = s.vx - w[0], uy = s.vy - w[1], uz = s.vz - w[2];
    const double n2 = s.q0 * s.q0 + s.q1 * s.q1 + s.q2 * s.q2 + s.q3 * s.q3;
    const bool q_ok = n2 > 1e-24;
    const double rn = fast_rsqrt(q_ok ? n2 : 1.0);
    const double qw = q_ok ? s.q0 * rn : 1.0, qx = q_ok ? s.q1 * rn : 0.0, qy = q_ok ? s.q2 * rn : 0.0, qz = q_ok ? s.q3 * rn : 0.0;
    /* the factors of two are folded into one operand (exact), so 2*(a*b - c*d) costs one product and one FMA */
    const double x2 = qx + qx, y2 = qy + qy, z2 = qz + qz;
    const double r00 = 1.0 - (qy * y2 + qz * z2), r01 = qx * y2 - qw * z2, r02 = qx * z2 + qw * y2;
    const double r10 = qx * y2 + qw * z2, r11 = 1.0 - (qx * x2 + qz * z2), r12 = qy * z2 - qw * x2;
    const double r20 = qx * z2 - qw * y2, r21 = qy * z2 + qw * x2, r22 = 1.0 - (qx * x2 + qy * y2);
    const double vbx = r00 * ux + r10 * uy + r20 * uz, vby = r01 * ux + r11 * uy + r21 * uz, vbz = r02 * ux + r12 * uy + r22 * uz;
    const double v2 = ux * ux + uy * uy + uz * uz;
    const double mach2 = v2 * (inv_RT * M.mach_k);
    double mach = fast_sqrt(mach2);
    const double mach_c = (mach > 1e300) ? 1e300 : mach;
    const bool a_dead = (fabs(vbx) < 1e-6) && (fabs(vbz) < 1e-6);
    const double alpha = a_dead ? 0.0 : fast_atan2(vbz, vbx);
    const double vxz = fast_sqrt(vbx * vbx + vbz * vbz);
    const double beta = (vxz < 1e-6) ? 0.0 : fast_atan2(vby, vxz);
    /* :535-539 rocket.py:105-108,138-218 */
    double co[7];
    aero_coefficients(M, Tb, mach_c, mach2, alpha, beta, cg, s.pf > 0.0, S.cd_scale, co);
    const double cd = co[0], cl = co[1], cm = co[2], cp = co[5];
    const double sm = cp - cg;
    const double qdyn = 0.5 * rho * v2;                                   /* :541 */
    out[EMC_SER_DRAG * ld] = qdyn * cd * M.ref_area;                      /* :542 */
    WindBracket TB; wind_bracket_reset(TB);
    out[EMC_SER_THRUST * ld] = thrust_at(M, Tb, S, TB, t_shift, p);       /* :543 */
    out[EMC_SER_CD * ld] = cd; out[EMC_SER_CL * ld] = cl; out[EMC_SER_CM * ld] = cm;
    out[EMC_SER_CP_DYNAMIC * ld] = cp;
    out[EMC_SER_STABILITY_MARGIN * ld] = sm / M.ref_diam;                 /* :549 */
    out[EMC_SER_AOA * ld] = alpha; out[EMC_SER_SIDESLIP * ld] = beta;
    out[EMC_SER_SPEED * ld] = sqrt(s.vx * s.vx + s.vy * s.vy + s.vz * s.vz);
    out[EMC_SER_MACH * ld] = mach; out[EMC_SER_QDYN * ld] = qdyn;
}

}  // namespace emc
