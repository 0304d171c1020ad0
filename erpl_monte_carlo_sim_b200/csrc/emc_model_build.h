/*
 * emc_model_build.h — host-side folding of the raw reference attributes (emc_model, include/emc.h)
 * into the constants the kernels read (DevModel / DevTables).  Each derived value is computed with
 * the expression the reference itself evaluates (file:line cited), in double, on the host.
 */
#pragma once
#include <math.h>
#include <string.h>
#include <cmath>

#include "emc_physics.cuh"

namespace emc {

inline const char *validate_model(const emc_model &m)
{
    if (m.n_cd < 2 || m.n_cd > EMC_MAX_CD_KNOTS) return "Cd_data needs 2..16 knots";
    if (m.n_cp < 2 || m.n_cp > EMC_MAX_CP_KNOTS) return "CP_shift_data needs 2..16 knots";
    for (int i = 1; i < m.n_cd; ++i) if (!(m.cd_mach[i] > m.cd_mach[i - 1])) return "Cd_data['mach'] must be strictly increasing";
    for (int i = 1; i < m.n_cp; ++i) if (!(m.cp_mach[i] > m.cp_mach[i - 1])) return "CP_shift_data['mach'] must be strictly increasing";
    if (m.motor_kind != EMC_MOTOR_LIQUID && m.motor_kind != EMC_MOTOR_SOLID) return "motor_kind must be 0 (liquid) or 1 (solid)";
    if (m.motor_kind == EMC_MOTOR_SOLID) {
        if (m.n_thrust < 2 || m.n_thrust > EMC_MAX_THRUST_KNOTS) return "thrust curve needs 2..32 knots";
        for (int i = 1; i < m.n_thrust; ++i) if (!(m.thrust_time[i] > m.thrust_time[i - 1])) return "thrust_curve_time must be strictly increasing";
    }
    if (m.has_wind) {
        if (m.n_wind < 2 || m.n_wind > EMC_MAX_WIND_KNOTS) return "altitude_profile needs 2..1024 knots";
        if (!m.wind_altitudes) return "wind_altitudes is NULL";
        for (int i = 1; i < m.n_wind; ++i) if (!(m.wind_altitudes[i] > m.wind_altitudes[i - 1])) return "altitude_profile must be strictly increasing";
    }
    if (!(m.dt_initial > 0.0)) return "dt_initial must be > 0";
    if (!(m.reference_diameter != 0.0)) return "reference_diameter must be non-zero";
    return nullptr;
}

/* Pressure above the troposphere exactly as atmosphere()'s general path defines it (environment.py:35-103), in long
 * double, from the folded double constants the kernels use.  layer: 1 (h_tropo, h_strat], 2 (h_strat, 25 km],
 * 3 (25 km, 32 km], 4 above 32 km. */
inline long double atm_pressure_ld(const DevModel &D, int layer, long double z)
{
    if (layer == 1) return (long double)D.p11 * expl((long double)D.k_iso * (z - (long double)D.h_tropo));
    if (layer == 2) return (long double)D.p20 * expl((long double)D.k_iso * (z - (long double)D.h_strat));
    if (layer == 3) {
        long double T = (long double)D.T_strat + 0.001L * (z - (long double)D.h_strat);
        if (T > 228.65L) T = 228.65L;
        return (long double)D.p25 * expl((long double)D.expo_25 * logl(T * (long double)D.inv_T_strat));
    }
    long double T = 228.65L - 0.0028L * (z - 32000.0L);
    if (T < 180.0L) T = 180.0L;
    return 868.02L * expl(-(z - 32000.0L) * ((long double)D.g0 / ((long double)D.R_gas * T)));
}

/* degree-16 Chebyshev interpolant of f on [a, b] as monomial coefficients in zeta = (z - zc)/zh; returns the largest
 * relative error of the DOUBLE even/odd Horner evaluation the kernel performs, sampled on 97 points */
template <class F>
inline double atm_fit_segment(F f, long double a, long double b, double c[17])
{
    const int N = 17;
    const long double zc = 0.5L * (a + b), zh = 0.5L * (b - a), PI = 3.14159265358979323846264338327950288L;
    long double fx[N], ck[N];
    for (int i = 0; i < N; ++i) fx[i] = f(zc + zh * cosl(PI * (i + 0.5L) / N));
    for (int k = 0; k < N; ++k) {
        long double acc = 0.0L;
        for (int i = 0; i < N; ++i) acc += fx[i] * cosl(PI * k * (i + 0.5L) / N);
        ck[k] = acc * (k == 0 ? 1.0L : 2.0L) / N;
    }
    /* Chebyshev -> monomial: T_0 = 1, T_1 = x, T_{k+1} = 2 x T_k - T_{k-1} */
    long double mono[N] = { 0 }, t0[N] = { 0 }, t1[N] = { 0 }, t2[N];
    t0[0] = 1.0L; t1[1] = 1.0L;
    mono[0] += ck[0];
    for (int j = 0; j < N; ++j) mono[j] += ck[1] * t1[j];
    for (int k = 2; k < N; ++k) {
        for (int j = 0; j < N; ++j) t2[j] = (j > 0 ? 2.0L * t1[j - 1] : 0.0L) - t0[j];
        for (int j = 0; j < N; ++j) { mono[j] += ck[k] * t2[j]; t0[j] = t1[j]; t1[j] = t2[j]; }
    }
    for (int j = 0; j < N; ++j) c[j] = (double)mono[j];
    const double dzc = (double)zc, dizh = (double)(1.0L / zh);
    double worst = 0.0;
    for (int i = 0; i <= 96; ++i) {
        const long double zl = a + (b - a) * i / 96.0L;
        const double z = (double)zl;
        const double zeta = (z - dzc) * dizh, z2 = zeta * zeta;
        double pe = c[16], po = c[15];
        for (int k = 14; k >= 2; k -= 2) { pe = fma(pe, z2, c[k]); po = fma(po, z2, c[k - 1]); }
        pe = fma(pe, z2, c[0]);
        const double p = fma(po, zeta, pe);
        const long double ref = f((long double)z);
        const double err = (double)fabsl(((long double)p - ref) / ref);
        if (!(err <= worst)) worst = err;          /* NaN counts as failure */
    }
    return worst;
}

inline void build_atmosphere_segments(DevModel &D)
{
    D.n_atm = 0;
    const double z180 = 32000.0 + (228.65 - 180.0) / 0.0028;
    struct Layer { int id; double lo, hi, tb, ts, tz0, tmin, tmax; };
    const Layer layers[5] = {
        { 1, D.h_tropo, D.h_strat, D.T_strat, 0.0, D.h_tropo, -INFINITY, INFINITY },
        { 2, D.h_strat, 25000.0, D.T_strat, 0.001, D.h_strat, -INFINITY, 228.65 },
        { 3, 25000.0, 32000.0, D.T_strat, 0.001, D.h_strat, -INFINITY, 228.65 },
        { 4, 32000.0, z180, 228.65, -0.0028, 32000.0, 180.0, INFINITY },
        { 4, z180, 100000.0, 228.65, -0.0028, 32000.0, 180.0, INFINITY },
    };
    if (!(D.h_tropo < D.h_strat && D.h_strat < 25000.0)) return;      /* non-standard layer order: keep the general path */
    for (const Layer &Ly : layers) {
        auto f = [&](long double z) { return atm_pressure_ld(D, Ly.id, z); };
        bool done = false;
        for (int pieces = 1; pieces <= 8 && !done; ++pieces) {
            if (D.n_atm + pieces > EMC_ATM_SEG) break;
            double cc[8][17];
            bool good = true;
            for (int q = 0; q < pieces && good; ++q) {
                const long double a = Ly.lo + (long double)(Ly.hi - Ly.lo) * q / pieces, b = Ly.lo + (long double)(Ly.hi - Ly.lo) * (q + 1) / pieces;
                good = atm_fit_segment(f, a, b, cc[q]) < 4e-16;
            }
            if (!good) continue;
            for (int q = 0; q < pieces; ++q) {
                const long double a = Ly.lo + (long double)(Ly.hi - Ly.lo) * q / pieces, b = Ly.lo + (long double)(Ly.hi - Ly.lo) * (q + 1) / pieces;
                const int j = D.n_atm++;
                D.at_lo[j] = (q == 0) ? Ly.lo : (double)a; D.at_hi[j] = (q == pieces - 1) ? Ly.hi : (double)b;
                D.at_zc[j] = (double)(0.5L * (a + b)); D.at_izh[j] = (double)(1.0L / (0.5L * (b - a)));
                D.at_tb[j] = Ly.tb; D.at_ts[j] = Ly.ts; D.at_tz0[j] = Ly.tz0; D.at_tmin[j] = Ly.tmin; D.at_tmax[j] = Ly.tmax;
                for (int k = 0; k < 17; ++k) D.at_c[j][k] = cc[q][k];
            }
            done = true;
        }
        if (!done) break;       /* this layer and everything above it keep the exp/log path */
    }
}

inline void build_dev_model(const emc_model &m, DevModel &D, DevTables &T)
{
    memset(&D, 0, sizeof D);
    memset(&T, 0, sizeof T);
    const double g = m.gravity, R = m.gas_constant, L = m.temperature_lapse_rate;
    D.T0 = m.sea_level_temperature; D.lapse = L; D.inv_T0 = 1.0 / m.sea_level_temperature;
    D.p0 = m.sea_level_pressure; D.h_tropo = m.troposphere_height; D.h_strat = m.stratosphere_height;
    D.T_strat = m.stratosphere_temp; D.inv_T_strat = 1.0 / m.stratosphere_temp;
    D.expo_tropo = g / (R * L);                                                       /* environment.py:33 */
    D.p11 = m.sea_level_pressure * pow(m.stratosphere_temp / m.sea_level_temperature, g / (R * L)); /* :38-40 */
    D.k_iso = -g / (R * m.stratosphere_temp);                                         /* :43-44 */
    D.p20 = D.p11 * exp(-g * (m.stratosphere_height - m.troposphere_height) / (R * m.stratosphere_temp)); /* :56-62 */
    D.p25 = D.p20 * exp(-g * 5000.0 / (R * m.stratosphere_temp));                     /* :72-75 */
    D.expo_25 = g / (R * 0.0028);                                                     /* :76,81 */
    D.R_gas = R; D.g0 = g;
    D.mach_k = (R == 287.053) ? 1.0 / 1.4 : R / (1.4 * 287.053);                      /* utils.py:152-157 */
    D.gamma = (m.gamma > 0.0) ? m.gamma : 1.4;                                        /* environment.py:19,96 */
    D.a2_k = 1.4 * 287.053;                                                           /* a^2 = 1.4*287.053*T, utils.py:152-157 */
    D.rho_k = (1.4 * 287.053) / R;                                                    /* 1/(R T) = rho_k / a^2 */
    {
        /* p0*(1 - a z)^e, a = L/T0, about zc: p0*(1 - a zc)^e * (1 - ap zeta)^e with ap = a zh/(1 - a zc); binomial series
         * c_k = c_{k-1} * (e - k + 1)/k * (-ap).  Used on [-2 km, troposphere_height] when the series has converged to
         * 1e-18 by degree 16 (standard atmosphere: ap = 0.16, last kept term 1e-20). */
        D.tp_lo = INFINITY; D.tp_hi = -INFINITY; D.tp_zc = 0.0; D.tp_inv_zh = 0.0;
        const long double zlo = -2000.0L, zhi = m.troposphere_height;
        const long double zc = 0.5L * (zlo + zhi), zh = 0.5L * (zhi - zlo);
        const long double a = (long double)L / (long double)m.sea_level_temperature, base = 1.0L - a * zc;
        const long double e = (long double)g / ((long double)R * (long double)L);
        if (zh > 0.0L && base > 0.0L && std::isfinite((double)e) && std::isfinite((double)a)) {
            const long double ap = a * zh / base;
            long double c = (long double)m.sea_level_pressure * powl(base, e);
            long double ck[18];
            for (int k = 0; k <= 17; ++k) {
                if (k > 0) c *= (e - (long double)(k - 1)) / (long double)k * (-ap);
                ck[k] = c;
            }
            /* the dropped tail (k >= 17, at |zeta| = 1) against the value itself */
            if (fabsl(ap) < 0.25L && std::isfinite((double)ck[0]) && ck[0] != 0.0L && fabsl(ck[17]) < 1e-18L * fabsl(ck[0])) {
                for (int k = 0; k <= 16; ++k) D.tp_c[k] = (double)ck[k];
                D.tp_lo = (double)zlo; D.tp_hi = (double)zhi; D.tp_zc = (double)zc; D.tp_inv_zh = (double)(1.0L / zh);
            }
        }
    }

    build_atmosphere_segments(D);

    D.cg_dry = m.center_of_mass_dry; D.prop_cg = m.center_of_mass_dry - 0.5;          /* rocket.py:116 */
    const double d4 = m.diameter / 4;
    D.d4sq = d4 * d4;                                                                 /* rocket.py:122 */
    D.len2_12 = 2.0 * 2.0 / 12;                                                       /* rocket.py:121,123 */
    D.Ixx_dry = m.Ixx_dry; D.Iyy_dry = m.Iyy_dry;

    D.ref_area = m.reference_area; D.ref_diam = m.reference_diameter; D.inv_ref_diam = 1.0 / m.reference_diameter;
    D.area_diam = m.reference_area * m.reference_diameter;
    D.cp_location = m.cp_location;
    const double cr = m.fin_root_chord, ct = m.fin_tip_chord, s = m.fin_span;
    const double fin_area = 0.5 * (cr + ct) * s;                                      /* rocket.py:176 */
    const double AR = (fin_area > 0) ? 2 * (s * s) / fin_area : 0.0;                  /* rocket.py:177 */
    const double cs = cos(m.fin_sweep_angle);
    const double aoc = AR / ((1e-6 > cs) ? 1e-6 : cs);                                /* rocket.py:179 */
    D.AR_over_cos2 = aoc * aoc;
    D.two_pi_AR_cos = (2 * M_PI * AR) * cs;                                           /* rocket.py:180 */
    D.fin_AR = AR; D.fin_cos = cs; D.fin_cos_floor = (1e-6 > cs) ? 1e-6 : cs; D.two_pi_AR = 2 * M_PI * AR;
    D.stall_span = 45.0 * (M_PI / 180.0) - 15.0 * (M_PI / 180.0);                     /* rocket.py:168,185 */
    D.power_off_factor = m.power_off_drag_factor;
    D.stall_angle = 15.0 * (M_PI / 180.0);                                            /* rocket.py:167 */
    D.inv_stall_span = 1.0 / (45.0 * (M_PI / 180.0) - 15.0 * (M_PI / 180.0));         /* rocket.py:168,185 */
    D.chute_cd = m.parachute_cd; D.chute_area = m.parachute_area; D.chute_alt = m.parachute_deployment_altitude;

    D.max_time = m.max_time; D.dt_rail = m.dt_initial;
    D.dt = (0.005 < m.dt_initial) ? 0.005 : m.dt_initial;                             /* simulator.py:209 */
    D.half_dt = 0.5 * D.dt; D.dt_over_6 = D.dt / 6.0;
    D.stage_t[0] = 0.0; D.stage_t[1] = D.half_dt; D.stage_t[2] = D.half_dt; D.stage_t[3] = D.dt;      /* simulator.py:218-222 */
    D.stage_c[0] = D.half_dt; D.stage_c[1] = D.half_dt; D.stage_c[2] = D.dt; D.stage_c[3] = 0.0;
    D.pitch_damping = m.pitch_damping; D.yaw_damping = m.yaw_damping; D.rail_length = m.rail_length;

    D.motor_kind = m.motor_kind; D.n_cd = m.n_cd; D.n_cp = m.n_cp;
    D.n_thrust = (m.motor_kind == EMC_MOTOR_SOLID) ? m.n_thrust : 0;
    D.has_wind = m.has_wind ? 1 : 0; D.n_wind = m.has_wind ? m.n_wind : 0;
    if (D.has_wind) {
        const double *a = m.wind_altitudes;
        const double dz = (a[m.n_wind - 1] - a[0]) / (m.n_wind - 1);
        bool uni = dz > 0;
        for (int i = 0; i < m.n_wind && uni; ++i) if (fabs(a[i] - (a[0] + i * dz)) > 1e-6 * dz) uni = false;
        D.wind_uniform = uni ? 1 : 0; D.wind_alt0 = a[0]; D.wind_inv_dz = uni ? 1.0 / dz : 0.0;
    }
    /* bracket tables (see DevTables): np.interp slope expression, compiled_base.c */
    auto fill = [](int n, const double *x, const double *f, double *lo, double *hi, double *x0, double *f0, double *sl) {
        for (int b = 0; b <= n; ++b) {
            if (b == 0) { lo[b] = -INFINITY; hi[b] = x[0]; x0[b] = x[0]; f0[b] = f[0]; sl[b] = 0.0; }
            else if (b == n) { lo[b] = x[n - 1]; hi[b] = INFINITY; x0[b] = x[n - 1]; f0[b] = f[n - 1]; sl[b] = 0.0; }
            else { lo[b] = x[b - 1]; hi[b] = x[b]; x0[b] = x[b - 1]; f0[b] = f[b - 1]; sl[b] = (f[b] - f[b - 1]) / (x[b] - x[b - 1]); }
        }
    };
    {
        /* Mach union grid: own brackets of each table first, then every union bracket copies the entries of the original
         * bracket that contains it (a union bracket never straddles an original knot) */
        const int BC = EMC_MAX_CD_KNOTS + 2, BP = EMC_MAX_CP_KNOTS + 2;
        double cd_lo[BC], cd_hi[BC], cd_x0[BC], cd0_f[BC], cd0_s[BC], cda_f[BC], cda_s[BC];
        double cp_lo[BP], cp_hi[BP], cp_x0[BP], cp_f[BP], cp_s[BP];
        fill(m.n_cd, m.cd_mach, m.cd0, cd_lo, cd_hi, cd_x0, cd0_f, cd0_s);
        fill(m.n_cd, m.cd_mach, m.cda, cd_lo, cd_hi, cd_x0, cda_f, cda_s);
        fill(m.n_cp, m.cp_mach, m.cp_shift, cp_lo, cp_hi, cp_x0, cp_f, cp_s);
        double u[EMC_MAX_CD_KNOTS + EMC_MAX_CP_KNOTS];
        int nu = 0, i = 0, j = 0;
        while (i < m.n_cd || j < m.n_cp) {
            double v;
            if (j >= m.n_cp || (i < m.n_cd && m.cd_mach[i] <= m.cp_mach[j])) v = m.cd_mach[i++];
            else v = m.cp_mach[j++];
            if (nu == 0 || v != u[nu - 1]) u[nu++] = v;
        }
        D.n_mb = nu + 1;
        for (int b = 0; b <= nu; ++b) {
            const double lo = (b == 0) ? -INFINITY : u[b - 1], hi = (b == nu) ? INFINITY : u[b];
            T.m_lo[b] = lo; T.m_hi[b] = hi;
            int bc = 0, bp = 0;                         /* original bracket with lo_orig <= lo (brackets are half-open [lo, hi)) */
            while (bc < m.n_cd && cd_hi[bc] <= lo) ++bc;
            while (bp < m.n_cp && cp_hi[bp] <= lo) ++bp;
            if (b == 0) { bc = 0; bp = 0; }
            T.cd_x0[b] = cd_x0[bc]; T.cd0_f[b] = cd0_f[bc]; T.cd0_s[b] = cd0_s[bc]; T.cda_f[b] = cda_f[bc]; T.cda_s[b] = cda_s[bc];
            T.cp_x0[b] = cp_x0[bp]; T.cp_f[b] = cp_f[bp]; T.cp_s[b] = cp_s[bp];
        }
    }
    if (D.n_thrust > 0) fill(D.n_thrust, m.thrust_time, m.thrust_curve, T.th_lo, T.th_hi, T.th_x0, T.th_f, T.th_s);
}

}  // namespace emc
