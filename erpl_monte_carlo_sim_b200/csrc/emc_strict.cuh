/*
 * emc_strict.cuh — the STRICT continuation of a flight: the reference's arithmetic operation by operation.
 *
 * Why it exists.  The reference's 3-D flights end in a super-exponential blow-up: within the last
 * two or three RK4 steps the speed goes 1e6 -> 1e8 -> 1e80 -> overflow, and whether a stage then produces +-inf or
 * NaN (inf - inf) — hence whether the loop of simulator.py:238-264 sees "z <= 0.5 and vz <= 0" and stops, or carries
 * a NaN to max_time, and at which stored state the first NaN appears — is decided by the ORDER of the floating-point
 * operations of that step.  The fast derivative (emc_physics.cuh: FMA contraction, folded constants, reciprocals,
 * polynomials) is accurate to an ulp or two while everything is finite, but it cannot reproduce those patterns:
 * round 1 measured 0.75 % of the headline workload's samples with a different step count / termination / first-NaN
 * index, all of them blown-up outliers.
 *
 * What it does.  The flight kernel parks a trajectory as soon as a stored state shows the blow-up under way
 * (|v| > 1e7 m/s or |omega| > 1000 rad/s: one to four steps before the end, nothing has overflowed yet) and
 * emc_strict_kernel finishes it with the code below: no FMA contraction (__dmul_rn / __dadd_rn), IEEE division and
 * square root, libm pow / exp / atan2 / sin / cos, the operation order of the reference as written
 * (rocket_simulation/ file:line cited per block).  The one difference to the reference that remains on the device:
 * CUDA's libm is within 1-2 ulp of glibc's.
 *
 * Compiled by g++ as well (tests/hostseam): there the same code uses glibc, and flown from the rail exit it must
 * reproduce the CPU checker of tests/ BIT FOR BIT on every golden flight — that pins the operation order.
 */
#pragma once
#include "emc_physics.cuh"

namespace emc {

#if defined(__CUDA_ARCH__)
#define S_MUL(a, b) __dmul_rn((a), (b))
#define S_ADD(a, b) __dadd_rn((a), (b))
#define S_SUB(a, b) __dsub_rn((a), (b))
#define S_DIV(a, b) __ddiv_rn((a), (b))
#define S_SQRT(a) __dsqrt_rn(a)
#else
#define S_MUL(a, b) ((a) * (b))
#define S_ADD(a, b) ((a) + (b))
#define S_SUB(a, b) ((a) - (b))
#define S_DIV(a, b) ((a) / (b))
#define S_SQRT(a) sqrt(a)
#endif
#if defined(__CUDACC__)
#define S_HD __device__ __noinline__       /* the strict kernel is not register-critical: keep its code compact */
#else
#define S_HD inline
#endif

/* np.interp on a bracket table (DevTables): slope*(x - x0) + f0, NaN stays NaN */
EMC_HD double strict_interp(const double *lo, const double *hi, const double *x0, const double *f, const double *sl, int nb, double x)
{
    if (x != x) return x;
    const int j = brk_find(lo, hi, nb, 1, x);
    return S_ADD(S_MUL(sl[j], S_SUB(x, x0[j])), f[j]);
}

/* environment.py:26-103 */
EMC_HD void strict_atmosphere(const DevModel &M, double z, double &T, double &p, double &rho)
{
    const double g = M.g0, R = M.R_gas;
    if (z <= M.h_tropo) {
        T = S_SUB(M.T0, S_MUL(M.lapse, z));                                   /* :30 */
        p = S_MUL(M.p0, pow(S_DIV(T, M.T0), M.expo_tropo));                   /* :31-33 */
    } else if (z <= M.h_strat) {
        T = M.T_strat;                                                        /* :37 */
        p = S_MUL(M.p11, exp(S_DIV(S_MUL(-g, S_SUB(z, M.h_tropo)), S_MUL(R, T))));      /* :42-45 */
    } else if (z <= 32000.0) {
        T = S_ADD(M.T_strat, S_MUL(0.001, S_SUB(z, M.h_strat)));              /* :52 */
        T = py_min(T, 228.65);                                                /* :53 */
        if (z <= 25000.0) p = S_MUL(M.p20, exp(S_DIV(S_MUL(-g, S_SUB(z, M.h_strat)), S_MUL(R, M.T_strat))));   /* :66-69 */
        else p = S_MUL(M.p25, pow(S_DIV(T, M.T_strat), M.expo_25));           /* :79-81 */
    } else {
        T = S_SUB(228.65, S_MUL(0.0028, S_SUB(z, 32000.0)));                  /* :84 */
        T = py_max(T, 180.0);                                                 /* :85 */
        const double scale_height = S_DIV(S_MUL(R, T), g);                    /* :88 */
        p = S_MUL(868.02, exp(S_DIV(-S_SUB(z, 32000.0), scale_height)));      /* :90 */
    }
    rho = S_DIV(p, S_MUL(R, T));                                              /* :93 */
}

/* environment.py:105-108 */
EMC_HD double strict_gravity(const DevModel &M, double z)
{
    const double r = S_DIV(6.371e6, S_ADD(6.371e6, z));
    return S_MUL(M.g0, S_MUL(r, r));
}

/* rocket.py:110-136: mp = {mass, cg, Ixx, Iyy} */
EMC_HD void strict_mass(const DevModel &M, const Sample &S, double pf, double mp[4])
{
    const double cur = S_MUL(S.prop_mass, pf);
    const double total = S_ADD(S.dry_mass, cur);
    const double cg = S_DIV(S_ADD(S_MUL(S.dry_mass, M.cg_dry), S_MUL(cur, M.prop_cg)), total);
    const double dcg = S_SUB(M.prop_cg, cg);
    mp[0] = total; mp[1] = cg;
    mp[2] = S_ADD(M.Ixx_dry, S_MUL(cur, M.d4sq));
    mp[3] = S_ADD(M.Iyy_dry, S_MUL(cur, S_ADD(M.len2_12, S_MUL(dcg, dcg))));
}

/* utils.py:76-82 */
EMC_HD void strict_normalize(const double q[4], double o[4])
{
    const double n = S_SQRT(S_ADD(S_ADD(S_ADD(S_MUL(q[0], q[0]), S_MUL(q[1], q[1])), S_MUL(q[2], q[2])), S_MUL(q[3], q[3])));
    if (n > 1e-12) { o[0] = S_DIV(q[0], n); o[1] = S_DIV(q[1], n); o[2] = S_DIV(q[2], n); o[3] = S_DIV(q[3], n); }
    else { o[0] = 1.0; o[1] = 0.0; o[2] = 0.0; o[3] = 0.0; }
}

/* utils.py:100-111 (normalises internally) */
EMC_HD void strict_rotation(const double qin[4], double R[3][3])
{
    double q[4];
    strict_normalize(qin, q);
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    R[0][0] = S_SUB(1.0, S_MUL(2.0, S_ADD(S_MUL(y, y), S_MUL(z, z)))); R[0][1] = S_MUL(2.0, S_SUB(S_MUL(x, y), S_MUL(w, z))); R[0][2] = S_MUL(2.0, S_ADD(S_MUL(x, z), S_MUL(w, y)));
    R[1][0] = S_MUL(2.0, S_ADD(S_MUL(x, y), S_MUL(w, z))); R[1][1] = S_SUB(1.0, S_MUL(2.0, S_ADD(S_MUL(x, x), S_MUL(z, z)))); R[1][2] = S_MUL(2.0, S_SUB(S_MUL(y, z), S_MUL(w, x)));
    R[2][0] = S_MUL(2.0, S_SUB(S_MUL(x, z), S_MUL(w, y))); R[2][1] = S_MUL(2.0, S_ADD(S_MUL(y, z), S_MUL(w, x))); R[2][2] = S_SUB(1.0, S_MUL(2.0, S_ADD(S_MUL(x, x), S_MUL(y, y))));
}

EMC_HD double strict_norm3(double a, double b, double c) { return S_SQRT(S_ADD(S_ADD(S_MUL(a, a), S_MUL(b, b)), S_MUL(c, c))); }

/* environment.py:267-276: three np.interp on the sample's table */
EMC_HD void strict_wind(const DevModel &M, const double *alt, const Sample &S, double z, WindBracket &B, double w[3])
{
    if (!M.has_wind) { w[0] = w[1] = w[2] = 0.0; return; }
    bool end_value = false;
    if (!(z >= B.lo && z < B.hi)) end_value = wind_bracket_load(M, alt, S.wind, z, B);
    if (end_value || B.lo == -INFINITY || B.hi == INFINITY) {              /* outside the grid np.interp returns the end value itself, also for z = -+inf */
        for (int k = 0; k < 3; ++k) w[k] = B.f0[k];
        return;
    }
    const double dz = S_SUB(z, B.x0);
    for (int k = 0; k < 3; ++k) w[k] = S_ADD(S_MUL(B.s[k], dz), B.f0[k]);
}

/* motor.py:54-76 / :152-156.  Solid: np.interp over the sample's SCALED knots (thrust_curve_thrust * k, motor.py:105):
 * the knots of bracket j are th_f[j] and th_f[j+1] of the base curve. */
EMC_HD double strict_thrust(const DevModel &M, const DevTables &Tb, const Sample &S, double t, double p)
{
    if (t < 0.0 || t > S.burn_time) return 0.0;
    if (M.motor_kind == EMC_MOTOR_SOLID) {
        const int nb = M.n_thrust + 1;
        double sl;
        if (t != t) sl = t;
        else {
            const int j = brk_find(Tb.th_lo, Tb.th_hi, nb, 1, t);
            const double f0 = S_MUL(Tb.th_f[j], S.thrust_a);
            if (j == 0 || j == nb - 1 || Tb.th_lo[j] == t) sl = f0;             /* outside the knots, or exactly on one */
            else {
                const double f1 = S_MUL(Tb.th_f[j + 1], S.thrust_a);
                const double slope = S_DIV(S_SUB(f1, f0), S_SUB(Tb.th_hi[j], Tb.th_lo[j]));
                sl = S_ADD(S_MUL(slope, S_SUB(t, Tb.th_lo[j])), f0);
            }
        }
        return S_ADD(sl, S_MUL(S.nozzle_area, S_SUB(101325.0, p)));
    }
    return S_SUB(S.thrust_a, S_MUL(S.nozzle_area, p));
}

/* rocket.py:138-218: c = {cd, cl, cm, cy, cyaw, cp} */
EMC_HD void strict_aero(const DevModel &M, const DevTables &Tb, double mach, double alpha, double beta, double cg, bool power_on,
                        double cd_scale, double c[6])
{
    const double mc = (mach > 1e300) ? 1e300 : mach;                 /* np.interp clamps +inf to the last knot */
    double cd0 = strict_interp(Tb.m_lo, Tb.m_hi, Tb.cd_x0, Tb.cd0_f, Tb.cd0_s, M.n_mb, mc);      /* :156 */
    if (cd_scale != 1.0) cd0 = S_MUL(cd0, cd_scale);
    const double cda = strict_interp(Tb.m_lo, Tb.m_hi, Tb.cd_x0, Tb.cda_f, Tb.cda_s, M.n_mb, mc); /* :157 */
    double cd = S_ADD(cd0, S_MUL(cda, S_MUL(alpha, alpha)));                                      /* :158 */
    if (!power_on) cd = S_MUL(cd, M.power_off_factor);                                            /* :159-160 */
    const double abs_alpha = fabs(alpha);
    const double beta_m = (mach < 1.0) ? S_SQRT(fabs(S_SUB(1.0, S_MUL(mach, mach)))) : S_SQRT(fabs(S_SUB(S_MUL(mach, mach), 1.0)));   /* :178 */
    const double t = S_DIV(S_MUL(M.fin_AR, beta_m), M.fin_cos_floor);
    const double denom = S_ADD(2.0, S_SQRT(S_ADD(4.0, S_MUL(t, t))));                             /* :179 */
    const double cl_alpha = S_MUL(S_DIV(M.two_pi_AR, denom), M.fin_cos);                          /* :180 */
    double cl = S_MUL(cl_alpha, alpha);                                                           /* :181 */
    double stall_factor = 0.0;
    const bool stalled = abs_alpha > M.stall_angle;
    if (stalled) {                                                                                /* :183-187 */
        const double over = S_DIV(S_SUB(abs_alpha, M.stall_angle), M.stall_span);
        stall_factor = py_max(0.0, S_SUB(1.0, over));
        const double sgn = (alpha > 0.0) ? 1.0 : ((alpha < 0.0) ? -1.0 : ((alpha == 0.0) ? 0.0 : alpha));
        cl = S_MUL(S_MUL(S_MUL(cl_alpha, M.stall_angle), stall_factor), sgn);
        cd = S_MUL(cd, S_ADD(1.0, S_DIV(S_MUL(0.5, S_SUB(abs_alpha, M.stall_angle)), M.stall_span)));
    }
    const double cp = S_ADD(M.cp_location, strict_interp(Tb.m_lo, Tb.m_hi, Tb.cp_x0, Tb.cp_f, Tb.cp_s, M.n_mb, mc));   /* :190 */
    const double sm = S_SUB(cp, cg);                                                              /* :195 */
    const double cm_alpha = S_MUL(-cl_alpha, sm);                                                 /* :196 */
    double cy = S_MUL(cl_alpha, beta);                                                            /* :200 */
    if (stalled) cy = S_MUL(cy, stall_factor);                                                    /* :203-204 */
    c[0] = cd; c[1] = cl; c[2] = S_MUL(cm_alpha, alpha); c[3] = cy; c[4] = S_MUL(S_MUL(-cl_alpha, sm), beta); c[5] = cp;
}

/* simulator.py:295-460 */
S_HD void strict_derivative(const DevModel &M, const DevTables &Tb, const double *wind_alt, const Sample &S, WindBracket &WB,
                            double t, const double st[14], bool &chute, double &chute_time, double sd[14])
{
    const double pf = py_max(0.0, st[13]);                                                        /* :305 */
    double q[4];
    strict_normalize(st + 6, q);                                                                  /* :308 */
    double mp[4];
    strict_mass(M, S, pf, mp);                                                                    /* :311 */
    double mass = mp[0];
    if (mass < S.dry_mass) { mass = S.dry_mass; strict_mass(M, S, 0.0, mp); }                    /* :315-318 */
    const double Ixx = mp[2], Iyy = mp[3], Izz = mp[3];
    double R[3][3];
    strict_rotation(q, R);                                                                        /* :324 */
    const double z = st[2];
    double T, p, rho;
    strict_atmosphere(M, z, T, p, rho);                                                           /* :328 */
    double w[3];
    strict_wind(M, wind_alt, S, z, WB, w);                                                        /* :333-338 */
    const double u[3] = { S_SUB(st[3], w[0]), S_SUB(st[4], w[1]), S_SUB(st[5], w[2]) };           /* :341 */
    double vb[3];
    for (int i = 0; i < 3; ++i) vb[i] = S_ADD(S_ADD(S_MUL(R[0][i], u[0]), S_MUL(R[1][i], u[1])), S_MUL(R[2][i], u[2]));   /* :344 */
    const double vn = strict_norm3(u[0], u[1], u[2]);
    const double mach = S_DIV(vn, S_SQRT(S_MUL(S_MUL(1.4, 287.053), T)));                         /* :347, utils.py:152-157 */
    const double alpha = (fabs(vb[0]) < 1e-6 && fabs(vb[2]) < 1e-6) ? 0.0 : atan2(vb[2], vb[0]);  /* :348 */
    const double vxz = S_SQRT(S_ADD(S_MUL(vb[0], vb[0]), S_MUL(vb[2], vb[2])));
    const double beta = (vxz < 1e-6) ? 0.0 : atan2(vb[1], vxz);                                   /* :349 */
    const double qdyn = S_MUL(S_MUL(0.5, rho), S_MUL(vn, vn));                                    /* :352 */
    double fb[3] = { 0.0, 0.0, 0.0 }, mb[3] = { 0.0, 0.0, 0.0 };
    const double thrust = (pf > 0.0 && t <= S.burn_time) ? strict_thrust(M, Tb, S, t, p) : 0.0;   /* :359-360 */
    fb[0] = S_ADD(fb[0], thrust);                                                                 /* :363 */
    if (!chute && z <= M.chute_alt && st[5] < 0.0) { chute = true; chute_time = t; }              /* :366-369 */
    if (chute) {                                                                                  /* :372-377 */
        const double rs = strict_norm3(vb[0], vb[1], vb[2]);
        if (rs > 0.0) {
            double drag = S_MUL(S_MUL(S_MUL(0.5, rho), S_MUL(rs, rs)), M.chute_cd);
            drag = S_MUL(drag, M.chute_area);
            for (int i = 0; i < 3; ++i) fb[i] = S_ADD(fb[i], S_DIV(S_MUL(-drag, vb[i]), rs));
        }
    } else if (qdyn > 0.0) {                                                                      /* :378-411 */
        double c[6];
        strict_aero(M, Tb, mach, alpha, beta, mp[1], pf > 0.0, S.cd_scale, c);
        const double drag = S_MUL(S_MUL(qdyn, c[0]), M.ref_area);
        const double lift = S_MUL(S_MUL(qdyn, c[1]), M.ref_area);
        const double side = S_MUL(S_MUL(qdyn, c[3]), M.ref_area);
        const double ca = cos(alpha), sa = sin(alpha), cb = cos(beta), sb = sin(beta);            /* utils.py:194-205 */
        const double W[3][3] = { { S_MUL(ca, cb), -sb, S_MUL(sa, cb) }, { S_MUL(ca, sb), cb, S_MUL(sa, sb) }, { -sa, 0.0, ca } };
        const double fw[3] = { -drag, -side, -lift };
        for (int i = 0; i < 3; ++i) fb[i] = S_ADD(fb[i], S_ADD(S_ADD(S_MUL(W[i][0], fw[0]), S_MUL(W[i][1], fw[1])), S_MUL(W[i][2], fw[2])));
        mb[0] = S_ADD(mb[0], S_MUL(S_MUL(S_MUL(qdyn, 0.0), M.ref_area), M.ref_diam));
        mb[1] = S_ADD(mb[1], S_MUL(S_MUL(S_MUL(qdyn, c[2]), M.ref_area), M.ref_diam));
        mb[2] = S_ADD(mb[2], S_MUL(S_MUL(S_MUL(qdyn, c[4]), M.ref_area), M.ref_diam));
    }
    mb[1] = S_ADD(mb[1], S_MUL(-M.pitch_damping, st[11]));                                        /* :414 */
    mb[2] = S_ADD(mb[2], S_MUL(-M.yaw_damping, st[12]));                                          /* :415 */
    double fi[3];
    for (int i = 0; i < 3; ++i) fi[i] = S_ADD(S_ADD(S_MUL(R[i][0], fb[0]), S_MUL(R[i][1], fb[1])), S_MUL(R[i][2], fb[2]));   /* :418 */
    fi[2] = S_SUB(fi[2], S_MUL(mass, strict_gravity(M, z)));                                      /* :421-422 */
    sd[0] = st[3]; sd[1] = st[4]; sd[2] = st[5];
    sd[3] = S_DIV(fi[0], mass); sd[4] = S_DIV(fi[1], mass); sd[5] = S_DIV(fi[2], mass);           /* :425 */
    const double wx = st[10], wy = st[11], wz = st[12];
    sd[10] = (Ixx > 0.0) ? S_DIV(S_SUB(mb[0], S_MUL(S_MUL(S_SUB(Izz, Iyy), wy), wz)), Ixx) : 0.0; /* :431-436 */
    sd[11] = (Iyy > 0.0) ? S_DIV(S_SUB(mb[1], S_MUL(S_MUL(S_SUB(Ixx, Izz), wz), wx)), Iyy) : 0.0;
    sd[12] = (Izz > 0.0) ? S_DIV(S_SUB(mb[2], S_MUL(S_MUL(S_SUB(Iyy, Ixx), wx), wy)), Izz) : 0.0;
    /* :439, utils.py:114-121 with :85-97: q (x) (0, w), then the norm correction */
    const double w1 = q[0], x1 = q[1], y1 = q[2], z1 = q[3], w2 = 0.0, x2 = wx, y2 = wy, z2 = wz;
    const double qm[4] = {
        S_SUB(S_SUB(S_SUB(S_MUL(w1, w2), S_MUL(x1, x2)), S_MUL(y1, y2)), S_MUL(z1, z2)),
        S_SUB(S_ADD(S_ADD(S_MUL(w1, x2), S_MUL(x1, w2)), S_MUL(y1, z2)), S_MUL(z1, y2)),
        S_ADD(S_ADD(S_SUB(S_MUL(w1, y2), S_MUL(x1, z2)), S_MUL(y1, w2)), S_MUL(z1, x2)),
        S_ADD(S_SUB(S_ADD(S_MUL(w1, z2), S_MUL(x1, y2)), S_MUL(y1, x2)), S_MUL(z1, w2)) };
    const double ne = S_SUB(S_ADD(S_ADD(S_ADD(S_MUL(q[0], q[0]), S_MUL(q[1], q[1])), S_MUL(q[2], q[2])), S_MUL(q[3], q[3])), 1.0);
    for (int i = 0; i < 4; ++i) sd[6 + i] = S_SUB(S_MUL(0.5, qm[i]), S_MUL(S_MUL(0.5, ne), q[i]));
    double pfr = 0.0;                                                                             /* :442-450 */
    if (pf > 0.0 && t <= S.burn_time) {
        pfr = (t < 0.0) ? S_DIV(-0.0, S.prop_mass) : S.pf_rate;     /* -mass_flow/propellant_mass, mass_flow = 0 outside [0, burn_time] (motor.py:78-84) */
        const double remaining = (pfr != 0.0) ? S_DIV(pf, fabs(pfr)) : INFINITY;
        if (remaining < 0.01) pfr = S_DIV(-pf, 0.01);
    }
    sd[13] = pfr;
}

/* per-stored-state diagnostics as _extract_results forms them (simulator.py:511-552), into the running maxima */
template <class CA>
S_HD void strict_diag(const DevModel &M, const DevTables &Tb, const double *wind_alt, const Sample &S, WindBracket &WB,
                      const double st[14], const CA &C)
{
    double mp[4];
    strict_mass(M, S, st[13], mp);                                                                /* :515 (pf unclamped) */
    double T, p, rho;
    strict_atmosphere(M, st[2], T, p, rho);
    double w[3];
    strict_wind(M, wind_alt, S, st[2], WB, w);
    const double u[3] = { S_SUB(st[3], w[0]), S_SUB(st[4], w[1]), S_SUB(st[5], w[2]) };
    double R[3][3];
    strict_rotation(st + 6, R);
    double vb[3];
    for (int i = 0; i < 3; ++i) vb[i] = S_ADD(S_ADD(S_MUL(R[0][i], u[0]), S_MUL(R[1][i], u[1])), S_MUL(R[2][i], u[2]));
    const double vn = strict_norm3(u[0], u[1], u[2]);
    const double mach = S_DIV(vn, S_SQRT(S_MUL(S_MUL(1.4, 287.053), T)));
    const double aoa = (fabs(vb[0]) < 1e-6 && fabs(vb[2]) < 1e-6) ? 0.0 : atan2(vb[2], vb[0]);
    const double mc = (mach > 1e300) ? 1e300 : mach;
    const double cp = S_ADD(M.cp_location, strict_interp(Tb.m_lo, Tb.m_hi, Tb.cp_x0, Tb.cp_f, Tb.cp_s, M.n_mb, mc));
    const double qd = S_MUL(S_MUL(0.5, rho), S_MUL(vn, vn));
    const double stab = S_DIV(S_SUB(cp, mp[1]), M.ref_diam);
    /* the fast bookkeeping keeps Mach^2 and |v|^2 (square roots are taken once, at the end) */
    cold_max(C, TC_MAX_MACH2, S_MUL(mach, mach));
    cold_max(C, TC_MAX_Q, qd);
    cold_max(C, TC_MAX_V2, S_ADD(S_ADD(S_MUL(st[3], st[3]), S_MUL(st[4], st[4])), S_MUL(st[5], st[5])));
    double om = fabs(st[10]);
    np_max_acc(om, fabs(st[11]));
    np_max_acc(om, fabs(st[12]));
    cold_max(C, TC_MAX_OM, om);
    cold_min(C, TC_MIN_STAB, stab);
    cold_max(C, TC_MAX_STAB, stab);
    cold_max(C, TC_MAX_AOA, fabs(aoa));
}

/* The rest of a flight from the stored state `st` (diagnostics of st not yet taken): simulator.py:216-264 with the
 * summary bookkeeping of the engine (track_post_step) and the closed-form NaN replay.  Returns the replayed steps. */
struct NoTape { EMC_HD void operator()(const TrackHot &, const double *) const {} };

template <class CA, class TAPE = NoTape>
EMC_HD int64_t strict_fly(const DevModel &M, const DevTables &Tb, const double *wind_alt, const Sample &S, TrackHot &K, const CA &C,
                          double st[14], bool nan_ff, int64_t *steps_out, const TAPE &tape = TAPE())
{
    WindBracket WB; wind_bracket_reset(WB);
    int64_t replayed = 0, steps = 0;
    strict_diag(M, Tb, wind_alt, S, WB, st, C);
    bool done = K.finishing;                      /* parked at its end state: only the diagnostics were missing */
    while (!done) {
        double k1[14], k2[14], k3[14], k4[14], y[14];
        bool chute = K.chute; double chute_time = 0.0;
        const double t = K.t;
        strict_derivative(M, Tb, wind_alt, S, WB, t, st, chute, chute_time, k1);
        for (int i = 0; i < 14; ++i) y[i] = S_ADD(st[i], S_MUL(M.half_dt, k1[i]));                /* :218 */
        strict_derivative(M, Tb, wind_alt, S, WB, S_ADD(t, M.half_dt), y, chute, chute_time, k2);
        for (int i = 0; i < 14; ++i) y[i] = S_ADD(st[i], S_MUL(M.half_dt, k2[i]));
        strict_derivative(M, Tb, wind_alt, S, WB, S_ADD(t, M.half_dt), y, chute, chute_time, k3);
        for (int i = 0; i < 14; ++i) y[i] = S_ADD(st[i], S_MUL(M.dt, k3[i]));
        strict_derivative(M, Tb, wind_alt, S, WB, S_ADD(t, M.dt), y, chute, chute_time, k4);
        for (int i = 0; i < 14; ++i)                                                              /* :224 */
            st[i] = S_ADD(st[i], S_MUL(M.dt_over_6, S_ADD(S_ADD(S_ADD(k1[i], S_MUL(2.0, k2[i])), S_MUL(2.0, k3[i])), k4[i])));
        double qn[4];
        strict_normalize(st + 6, qn);                                                             /* :227 */
        st[6] = qn[0]; st[7] = qn[1]; st[8] = qn[2]; st[9] = qn[3];
        if (chute != K.chute) { K.chute = true; C.setd(TC_CHUTE_TIME, chute_time); }
        K.t = S_ADD(K.t, M.dt);                                                                   /* :229 */
        ++steps;
        State s;
        memcpy(&s, st, sizeof s);
        done = track_post_step(M, S, K, C, s);
        tape(K, st);                                  /* a stored state (simulator.py:230-231) */
        strict_diag(M, Tb, wind_alt, S, WB, st, C);
        if (!done && nan_ff) {
            K.replay = (int8_t)nan_mode(M, S, K, s);
            if (K.replay) {
                replayed += replay_time(M, S, K, C, s);
                st[0] = s.x; st[1] = s.y;
                done = true;
            }
        }
    }
    if (steps_out) *steps_out = steps;
    return replayed;
}

}  // namespace emc
