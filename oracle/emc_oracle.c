/*
 * emc_oracle.c — TEST INFRASTRUCTURE.  CPU restatement (plain C, scalar, FP64) of the reference's
 * hot path.  It is the checker for the CUDA engine and the `cpu_baseline`/`--impl reference` arm of
 * bench.py; it is NOT shipped, NOT linked into libemc.so and NOT a fallback: only tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline legs may load it.
 *
 * Parity is PINNED: tests/test_oracle_golden.py checks this file against golden vectors produced by
 * running the unmodified Python reference in the build container (oracle/make_golden.py →
 * tests/golden/), since the reference's own tests hold no usable fixtures.
 *
 * Every function follows the reference's operation order as written; compile with
 * -ffp-contract=off so no FMA contraction changes the low bits (see oracle/Makefile).
 * Citations are file:line under /root/reference/rocket_simulation/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#include "../include/emc.h"

#define ORC_API __attribute__((visibility("default")))

/* ---------------- Python / NumPy scalar semantics ---------------- */
/* Python max(a,b) returns a unless b > a; min(a,b) returns a unless b < a (NaN comparisons False). */
static inline double py_max(double a, double b) { return (b > a) ? b : a; }
static inline double py_min(double a, double b) { return (b < a) ? b : a; }
/* np.sign */
static inline double np_sign(double x) { return (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : ((x == 0.0) ? 0.0 : x)); }

/* utils.py:147-149 -> np.interp (numpy/_core/src/multiarray/compiled_base.c arr_interp, scalar x).
 * fp is read with a stride so the (N,3) wind table columns need no copy. */
static double interp_strided(double x, const double *xp, const double *fp, int64_t fs, int n)
{
    if (isnan(x)) return x;
    if (n <= 0) return NAN;
    if (n == 1) return fp[0];
    if (x > xp[n - 1]) return fp[(int64_t)(n - 1) * fs];
    if (x < xp[0]) return fp[0];
    int lo = 0, hi = n - 1;                          /* largest j with xp[j] <= x */
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (xp[mid] <= x) lo = mid; else hi = mid; }
    int j = (xp[hi] <= x) ? hi : lo;
    if (j == n - 1) return fp[(int64_t)j * fs];
    if (xp[j] == x) return fp[(int64_t)j * fs];
    double f0 = fp[(int64_t)j * fs], f1 = fp[(int64_t)(j + 1) * fs];
    double slope = (f1 - f0) / (xp[j + 1] - xp[j]);
    double r = slope * (x - xp[j]) + f0;
    if (isnan(r)) {
        r = slope * (x - xp[j + 1]) + f1;
        if (isnan(r) && f0 == f1) r = f0;
    }
    return r;
}

ORC_API double orc_interp(double x, const double *xp, const double *fp, int n)
{
    return interp_strided(x, xp, fp, 1, n);
}

/* np.linalg.norm of a short vector: sqrt(dot(x,x)), sequential accumulation */
static inline double norm3(const double v[3]) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
static inline double norm4(const double v[4]) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]); }
static inline double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* utils.py:76-82 */
static void normalize_quaternion(const double q[4], double o[4])
{
    double n = norm4(q);
    if (n > 1e-12) { o[0] = q[0] / n; o[1] = q[1] / n; o[2] = q[2] / n; o[3] = q[3] / n; }
    else { o[0] = 1.0; o[1] = 0.0; o[2] = 0.0; o[3] = 0.0; }
}

/* utils.py:100-111 (normalises internally) */
static void quaternion_to_rotation_matrix(const double qin[4], double R[3][3])
{
    double q[4];
    normalize_quaternion(qin, q);
    double w = q[0], x = q[1], y = q[2], z = q[3];
    R[0][0] = 1 - 2 * (y * y + z * z); R[0][1] = 2 * (x * y - w * z);     R[0][2] = 2 * (x * z + w * y);
    R[1][0] = 2 * (x * y + w * z);     R[1][1] = 1 - 2 * (x * x + z * z); R[1][2] = 2 * (y * z - w * x);
    R[2][0] = 2 * (x * z - w * y);     R[2][1] = 2 * (y * z + w * x);     R[2][2] = 1 - 2 * (x * x + y * y);
}

/* utils.py:139-144 + :46-69 (SimpleRotation.as_euler) ; q = [w,x,y,z] */
static void quaternion_to_euler(const double q[4], double e[3])
{
    double x = q[1], y = q[2], z = q[3], w = q[0];
    double sinr_cosp = 2 * (w * x + y * z);
    double cosr_cosp = 1 - 2 * (x * x + y * y);
    e[0] = atan2(sinr_cosp, cosr_cosp);
    double sinp = 2 * (w * y - z * x);
    if (fabs(sinp) >= 1) e[1] = copysign(M_PI / 2, sinp);
    else e[1] = asin(sinp);
    double siny_cosp = 2 * (w * z + x * y);
    double cosy_cosp = 1 - 2 * (y * y + z * z);
    e[2] = atan2(siny_cosp, cosy_cosp);
}

/* utils.py:129-136 + :14-35 : euler (xyz) -> [w,x,y,z] */
ORC_API void orc_euler_to_quaternion(double roll, double pitch, double yaw, double q[4])
{
    double cr = cos(roll / 2), sr = sin(roll / 2);
    double cp = cos(pitch / 2), sp = sin(pitch / 2);
    double cy = cos(yaw / 2), sy = sin(yaw / 2);
    double x = sr * cp * cy - cr * sp * sy;
    double y = cr * sp * cy + sr * cp * sy;
    double z = cr * cp * sy - sr * sp * cy;
    double w = cr * cp * cy + sr * sp * sy;
    q[0] = w; q[1] = x; q[2] = y; q[3] = z;
}

/* utils.py:152-157 */
static double mach_number(const double v[3], double temperature)
{
    double gamma = 1.4, R = 287.053;
    double speed_of_sound = sqrt(gamma * R * temperature);
    return norm3(v) / speed_of_sound;
}

/* utils.py:160-164 */
static double angle_of_attack(const double vb[3])
{
    if (fabs(vb[0]) < 1e-6 && fabs(vb[2]) < 1e-6) return 0.0;
    return atan2(vb[2], vb[0]);
}

/* utils.py:167-172 */
static double sideslip_angle(const double vb[3])
{
    double V_xz = sqrt(vb[0] * vb[0] + vb[2] * vb[2]);
    if (V_xz < 1e-6) return 0.0;
    return atan2(vb[1], V_xz);
}

/* environment.py:26-103 */
ORC_API void orc_atmosphere(const emc_model *m, double altitude, double *T_out, double *p_out, double *rho_out)
{
    double temperature, pressure;
    double p0 = m->sea_level_pressure, T0 = m->sea_level_temperature, L = m->temperature_lapse_rate;
    double Rg = m->gas_constant, g = m->gravity, Ts = m->stratosphere_temp;
    if (altitude <= m->troposphere_height) {
        temperature = T0 - L * altitude;                                              /* :30 */
        pressure = p0 * pow(temperature / T0, g / (Rg * L));                          /* :31-33 */
    } else if (altitude <= m->stratosphere_height) {
        temperature = Ts;                                                             /* :37 */
        double pressure_11km = p0 * pow(Ts / T0, g / (Rg * L));                       /* :38-40 */
        pressure = pressure_11km * exp(-g * (altitude - m->troposphere_height) / (Rg * temperature)); /* :42-45 */
    } else {
        if (altitude <= 32000.0) {
            temperature = Ts + 0.001 * (altitude - m->stratosphere_height);           /* :52 */
            temperature = py_min(temperature, 228.65);                                /* :53 */
            double pressure_20km = p0 * pow(Ts / T0, g / (Rg * L));                   /* :56-58 */
            pressure_20km *= exp(-g * (m->stratosphere_height - m->troposphere_height) / (Rg * Ts)); /* :59-62 */
            if (altitude <= 25000.0) {
                pressure = pressure_20km * exp(-g * (altitude - m->stratosphere_height) / (Rg * Ts)); /* :66-69 */
            } else {
                double pressure_25km = pressure_20km * exp(-g * 5000.0 / (Rg * Ts));  /* :72-75 */
                double temp_gradient = 0.0028, temp_25km = Ts;                        /* :76-77 */
                pressure = pressure_25km * pow(temperature / temp_25km, g / (Rg * temp_gradient)); /* :79-81 */
            }
        } else {
            temperature = 228.65 - 0.0028 * (altitude - 32000.0);                     /* :84 */
            temperature = py_max(temperature, 180.0);                                 /* :85 */
            double scale_height = Rg * temperature / g;                               /* :88 */
            double pressure_32km = 868.02;                                            /* :89 */
            pressure = pressure_32km * exp(-(altitude - 32000.0) / scale_height);     /* :90 */
        }
    }
    *T_out = temperature;
    *p_out = pressure;
    *rho_out = pressure / (Rg * temperature);                                         /* :93 */
}

/* environment.py:105-108 */
ORC_API double orc_gravity(const emc_model *m, double altitude)
{
    double earth_radius = 6.371e6;
    double r = earth_radius / (earth_radius + altitude);
    return m->gravity * (r * r);
}

/* rocket.py:110-136 ; mp[4] = mass, center_of_mass, Ixx, Iyy(=Izz) */
ORC_API void orc_mass_properties(const emc_model *m, double dry_mass, double propellant_mass, double pf, double mp[4])
{
    double current_propellant = propellant_mass * pf;
    double total_mass = dry_mass + current_propellant;
    double propellant_cg = m->center_of_mass_dry - 0.5;
    double current_cg = (dry_mass * m->center_of_mass_dry + current_propellant * propellant_cg) / total_mass;
    double propellant_length = 2.0;
    double d4 = m->diameter / 4;
    double propellant_Ixx = current_propellant * (d4 * d4);
    double dcg = propellant_cg - current_cg;
    double propellant_Iyy = current_propellant * (propellant_length * propellant_length / 12 + dcg * dcg);
    mp[0] = total_mass;
    mp[1] = current_cg;
    mp[2] = m->Ixx_dry + propellant_Ixx;
    mp[3] = m->Iyy_dry + propellant_Iyy;
}

/* rocket.py:105-108 */
static double dynamic_cp(const emc_model *m, double mach)
{
    return m->cp_location + orc_interp(mach, m->cp_mach, m->cp_shift, m->n_cp);
}

/* rocket.py:138-218 ; c[6] = cd, cl, cm(=cpitch), cy, cyaw, cp */
ORC_API void orc_aero_coefficients(const emc_model *m, double mach, double alpha, double beta, double cg,
                                   int power_on, double cd_scale, double c[6])
{
    double cd0 = orc_interp(mach, m->cd_mach, m->cd0, m->n_cd);                        /* :156 */
    if (cd_scale != 1.0) cd0 = cd0 * cd_scale;      /* engine extension: scaled Cd_data['cd0'] */
    double cda = orc_interp(mach, m->cd_mach, m->cda, m->n_cd);                        /* :157 */
    double cd = cd0 + cda * (alpha * alpha);                                           /* :158 */
    if (!power_on) cd *= m->power_off_drag_factor;                                     /* :159-160 */
    double stall_angle = 15.0 * (M_PI / 180.0);                                        /* :167 np.radians */
    double max_angle = 45.0 * (M_PI / 180.0);                                          /* :168 */
    double abs_alpha = fabs(alpha);
    double cr = m->fin_root_chord, ct = m->fin_tip_chord, s = m->fin_span;
    double fin_area = 0.5 * (cr + ct) * s;                                             /* :176 */
    double AR = (fin_area > 0) ? 2 * (s * s) / fin_area : 0.0;                         /* :177 */
    double beta_m = (mach < 1) ? sqrt(fabs(1.0 - mach * mach)) : sqrt(fabs(mach * mach - 1)); /* :178 */
    double cs = cos(m->fin_sweep_angle);
    double t = AR * beta_m / py_max(cs, 1e-6);
    double denom = 2 + sqrt(4 + t * t);                                                /* :179 */
    double cl_alpha = (2 * M_PI * AR / denom) * cs;                                    /* :180 */
    double cl = cl_alpha * alpha;                                                      /* :181 */
    double stall_factor = 0.0;
    int stalled = abs_alpha > stall_angle;
    if (stalled) {                                                                     /* :183-187 */
        stall_factor = py_max(0.0, 1.0 - (abs_alpha - stall_angle) / (max_angle - stall_angle));
        cl = cl_alpha * stall_angle * stall_factor * np_sign(alpha);
        cd *= 1.0 + 0.5 * (abs_alpha - stall_angle) / (max_angle - stall_angle);
    }
    double cp_current = dynamic_cp(m, mach);                                           /* :190 */
    double static_margin = cp_current - cg;                                            /* :195 */
    double cm_alpha = -cl_alpha * static_margin;                                       /* :196 */
    double cm = cm_alpha * alpha;                                                      /* :197 */
    double cy = cl_alpha * beta;                                                       /* :200 */
    if (stalled) cy *= stall_factor;                                                   /* :203-204 */
    double cyaw = -cl_alpha * static_margin * beta;                                    /* :206 */
    c[0] = cd; c[1] = cl; c[2] = cm; c[3] = cy; c[4] = cyaw; c[5] = cp_current;
}

/* per-sample view of the scalar block */
typedef struct orc_sample {
    double dry_mass, propellant_mass, thrust_a, nozzle_area, mdot, burn_time, cd_scale;
    const double *wind;                 /* [n_wind][3] or NULL */
    double thrust_knots[EMC_MAX_THRUST_KNOTS]; /* Solid: base curve * multiplier, motor.py:105 */
} orc_sample;

static void sample_init(const emc_model *m, const double *col, int64_t stride, const double *wind, orc_sample *s)
{
    s->dry_mass = col[EMC_IN_DRY_MASS * stride];
    s->propellant_mass = col[EMC_IN_PROP_MASS * stride];
    s->thrust_a = col[EMC_IN_THRUST_A * stride];
    s->nozzle_area = col[EMC_IN_NOZZLE_AREA * stride];
    s->mdot = col[EMC_IN_MDOT * stride];
    s->burn_time = col[EMC_IN_BURN_TIME * stride];
    s->cd_scale = col[EMC_IN_CD_SCALE * stride];
    s->wind = m->has_wind ? wind : NULL;
    if (m->motor_kind == EMC_MOTOR_SOLID)
        for (int k = 0; k < m->n_thrust; ++k) s->thrust_knots[k] = m->thrust_curve[k] * s->thrust_a;
}

/* motor.py:54-76 (Solid) / :152-156 (Liquid); ambient pressure always supplied by the path */
ORC_API double orc_thrust_raw(const emc_model *m, const orc_sample *s, double time, double ambient_pressure)
{
    if (time < 0 || time > s->burn_time) return 0.0;
    if (m->motor_kind == EMC_MOTOR_SOLID) {
        double thrust_sl = orc_interp(time, m->thrust_time, s->thrust_knots, m->n_thrust);
        double pressure_correction = s->nozzle_area * (101325.0 - ambient_pressure);
        return thrust_sl + pressure_correction;
    }
    return s->thrust_a - s->nozzle_area * ambient_pressure;
}

/* motor.py:78-84 / :158-161 */
static double mass_flow_rate(const orc_sample *s, double time)
{
    if (time < 0 || time > s->burn_time) return 0.0;
    return s->mdot;
}

/* motor.py:86-93 / :163-169 */
static double propellant_remaining(const orc_sample *s, double time)
{
    if (time <= 0) return 1.0;
    else if (time >= s->burn_time) return 0.0;
    else return py_max(0.0, 1.0 - time / s->burn_time);
}

/* environment.py:267-276 */
static void wind_at_altitude(const emc_model *m, const orc_sample *s, double altitude, double w[3])
{
    if (!s->wind || m->n_wind == 0) { w[0] = w[1] = w[2] = 0.0; return; }
    /* three independent np.interp calls on the columns of the (N,3) table */
    for (int k = 0; k < 3; ++k)
        w[k] = interp_strided(altitude, m->wind_altitudes, s->wind + k, 3, m->n_wind);
}

/* utils.py:175-205 */
static void wind_to_body_matrix(double alpha, double beta, double M[3][3])
{
    double ca = cos(alpha), sa = sin(alpha), cb = cos(beta), sb = sin(beta);
    M[0][0] = ca * cb; M[0][1] = -sb; M[0][2] = sa * cb;
    M[1][0] = ca * sb; M[1][1] = cb;  M[1][2] = sa * sb;
    M[2][0] = -sa;     M[2][1] = 0.0; M[2][2] = ca;
}

/* simulator.py:295-460.  *chute is self.parachute_deployed (sticky); *chute_time gets the stage time
 * of the latching call when non-NULL. */
static void rocket_dynamics(const emc_model *m, const orc_sample *s, double t, const double state[14],
                            int *chute, double *chute_time, double state_dot[14])
{
    const double *position = state, *velocity = state + 3, *angular_velocity = state + 10;
    double propellant_fraction = state[13];
    propellant_fraction = py_max(0.0, propellant_fraction);                            /* :305 */
    double quaternion[4];
    normalize_quaternion(state + 6, quaternion);                                       /* :308 */
    double mp[4];
    orc_mass_properties(m, s->dry_mass, s->propellant_mass, propellant_fraction, mp);  /* :311 */
    double mass = mp[0];
    if (mass < s->dry_mass) {                                                          /* :315-318 */
        mass = s->dry_mass;
        orc_mass_properties(m, s->dry_mass, s->propellant_mass, 0.0, mp);
    }
    double Ixx = mp[2], Iyy = mp[3], Izz = mp[3];
    double R[3][3];
    quaternion_to_rotation_matrix(quaternion, R);                                      /* :324 */
    double altitude = position[2];
    double temperature, pressure, density;
    orc_atmosphere(m, altitude, &temperature, &pressure, &density);                    /* :328 */
    double wind_velocity[3];
    wind_at_altitude(m, s, altitude, wind_velocity);                                   /* :333-338 */
    double velocity_relative[3] = { velocity[0] - wind_velocity[0], velocity[1] - wind_velocity[1],
                                    velocity[2] - wind_velocity[2] };                 /* :341 */
    double velocity_body[3];                                                           /* :344  R.T @ v */
    for (int i = 0; i < 3; ++i)
        velocity_body[i] = R[0][i] * velocity_relative[0] + R[1][i] * velocity_relative[1] + R[2][i] * velocity_relative[2];
    double mach = mach_number(velocity_relative, temperature);                         /* :347 */
    double alpha = angle_of_attack(velocity_body);                                     /* :348 */
    double beta = sideslip_angle(velocity_body);                                       /* :349 */
    double vn = norm3(velocity_relative);
    double q_dynamic = 0.5 * density * (vn * vn);                                      /* :352 */
    double forces_body[3] = { 0, 0, 0 }, moments_body[3] = { 0, 0, 0 };
    double thrust;
    if (propellant_fraction > 0 && t <= s->burn_time) thrust = orc_thrust_raw(m, s, t, pressure); /* :359-360 */
    else thrust = 0.0;
    forces_body[0] += thrust;                                                          /* :363 */
    if (!*chute && altitude <= m->parachute_deployment_altitude && velocity[2] < 0) { /* :366-369 */
        *chute = 1;
        if (chute_time) *chute_time = t;
    }
    if (*chute) {                                                                      /* :372-377 */
        double rel_speed = norm3(velocity_body);
        if (rel_speed > 0) {
            double drag = 0.5 * density * (rel_speed * rel_speed) * m->parachute_cd;
            drag *= m->parachute_area;
            for (int i = 0; i < 3; ++i) forces_body[i] += -drag * velocity_body[i] / rel_speed;
        }
    } else if (q_dynamic > 0) {                                                        /* :378-411 */
        double c[6];
        orc_aero_coefficients(m, mach, alpha, beta, mp[1], propellant_fraction > 0, s->cd_scale, c);
        double drag = q_dynamic * c[0] * m->reference_area;
        double lift = q_dynamic * c[1] * m->reference_area;
        double side = q_dynamic * c[3] * m->reference_area;
        double W[3][3];
        wind_to_body_matrix(alpha, beta, W);
        double fw[3] = { -drag, -side, -lift };
        for (int i = 0; i < 3; ++i) forces_body[i] += W[i][0] * fw[0] + W[i][1] * fw[1] + W[i][2] * fw[2];
        moments_body[0] += q_dynamic * 0.0 * m->reference_area * m->reference_diameter;
        moments_body[1] += q_dynamic * c[2] * m->reference_area * m->reference_diameter;
        moments_body[2] += q_dynamic * c[4] * m->reference_area * m->reference_diameter;
    }
    moments_body[1] += -m->pitch_damping * angular_velocity[1];                        /* :414 */
    moments_body[2] += -m->yaw_damping * angular_velocity[2];                          /* :415 */
    double forces_inertial[3];                                                         /* :418 */
    for (int i = 0; i < 3; ++i)
        forces_inertial[i] = R[i][0] * forces_body[0] + R[i][1] * forces_body[1] + R[i][2] * forces_body[2];
    double gravity = orc_gravity(m, altitude);                                         /* :421 */
    forces_inertial[2] -= mass * gravity;                                              /* :422 */
    double acceleration[3] = { forces_inertial[0] / mass, forces_inertial[1] / mass, forces_inertial[2] / mass };
    double angular_acceleration[3] = { 0, 0, 0 };
    if (Ixx > 0) angular_acceleration[0] = (moments_body[0] - (Izz - Iyy) * angular_velocity[1] * angular_velocity[2]) / Ixx;
    if (Iyy > 0) angular_acceleration[1] = (moments_body[1] - (Ixx - Izz) * angular_velocity[2] * angular_velocity[0]) / Iyy;
    if (Izz > 0) angular_acceleration[2] = (moments_body[2] - (Iyy - Ixx) * angular_velocity[0] * angular_velocity[1]) / Izz;
    /* utils.py:114-121 with utils.py:85-97; q = normalised quaternion, omega_q = [0, wx, wy, wz] */
    double w1 = quaternion[0], x1 = quaternion[1], y1 = quaternion[2], z1 = quaternion[3];
    double w2 = 0.0, x2 = angular_velocity[0], y2 = angular_velocity[1], z2 = angular_velocity[2];
    double qm[4] = { w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
                     w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                     w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2 };
    double norm_error = (quaternion[0] * quaternion[0] + quaternion[1] * quaternion[1] +
                         quaternion[2] * quaternion[2] + quaternion[3] * quaternion[3]) - 1.0;
    double quaternion_rate[4];
    for (int i = 0; i < 4; ++i) quaternion_rate[i] = 0.5 * qm[i] - 0.5 * norm_error * quaternion[i];
    double propellant_fraction_rate;
    if (propellant_fraction > 0 && t <= s->burn_time) {                                /* :442-448 */
        double mass_flow = mass_flow_rate(s, t);
        propellant_fraction_rate = -mass_flow / s->propellant_mass;
        double remaining_time = (propellant_fraction_rate != 0) ? propellant_fraction / fabs(propellant_fraction_rate) : INFINITY;
        if (remaining_time < 0.01) propellant_fraction_rate = -propellant_fraction / 0.01;
    } else propellant_fraction_rate = 0.0;
    state_dot[0] = velocity[0]; state_dot[1] = velocity[1]; state_dot[2] = velocity[2];
    state_dot[3] = acceleration[0]; state_dot[4] = acceleration[1]; state_dot[5] = acceleration[2];
    for (int i = 0; i < 4; ++i) state_dot[6 + i] = quaternion_rate[i];
    for (int i = 0; i < 3; ++i) state_dot[10 + i] = angular_acceleration[i];
    state_dot[13] = propellant_fraction_rate;
}

/* NumPy max/min over a series: NaN propagates */
static inline void np_max_acc(double *m, double v) { if (v > *m || isnan(v)) { if (!isnan(*m)) *m = v; } }
static inline void np_min_acc(double *m, double v) { if (v < *m || isnan(v)) { if (!isnan(*m)) *m = v; } }

/* per-stored-state diagnostics exactly as _extract_results forms them (simulator.py:511-552) */
typedef struct orc_diag {
    double max_mach, max_q, max_speed, max_abs_omega, min_stab, max_stab, max_abs_aoa;
} orc_diag;

static void diag_state(const emc_model *m, const orc_sample *s, const double st[14], orc_diag *d, int first)
{
    double mp[4];
    orc_mass_properties(m, s->dry_mass, s->propellant_mass, st[13], mp);               /* :515 (pf unclamped) */
    double alt = st[2];
    double T, p, rho;
    orc_atmosphere(m, alt, &T, &p, &rho);                                              /* :523 */
    double w[3];
    wind_at_altitude(m, s, alt, w);                                                    /* :525-528 */
    double vel_rel[3] = { st[3] - w[0], st[4] - w[1], st[5] - w[2] };                 /* :530 */
    double R[3][3];
    quaternion_to_rotation_matrix(st + 6, R);
    double vb[3];
    for (int i = 0; i < 3; ++i) vb[i] = R[0][i] * vel_rel[0] + R[1][i] * vel_rel[1] + R[2][i] * vel_rel[2]; /* :531 */
    double mach = mach_number(vel_rel, T);                                             /* :532 */
    double aoa = angle_of_attack(vb);                                                  /* :533 */
    double cp_val = dynamic_cp(m, mach);                                               /* :535 */
    double vn = norm3(vel_rel);
    double q_dyn = 0.5 * rho * (vn * vn);                                              /* :541 */
    double stab = (cp_val - mp[1]) / m->reference_diameter;                            /* :549 */
    double speed = sqrt(st[3] * st[3] + st[4] * st[4] + st[5] * st[5]);                /* :476 */
    double om = fabs(st[10]);
    if (fabs(st[11]) > om || isnan(st[11])) { if (!isnan(om)) om = fabs(st[11]); }
    if (fabs(st[12]) > om || isnan(st[12])) { if (!isnan(om)) om = fabs(st[12]); }
    if (first) {
        d->max_mach = mach; d->max_q = q_dyn; d->max_speed = speed; d->max_abs_omega = om;
        d->min_stab = stab; d->max_stab = stab; d->max_abs_aoa = fabs(aoa);
    } else {
        np_max_acc(&d->max_mach, mach); np_max_acc(&d->max_q, q_dyn); np_max_acc(&d->max_speed, speed);
        np_max_acc(&d->max_abs_omega, om); np_min_acc(&d->min_stab, stab); np_max_acc(&d->max_stab, stab);
        np_max_acc(&d->max_abs_aoa, fabs(aoa));
    }
}

/*
 * One flight: simulator.py:127-293 (state0 :131-161, rail :42-125, RK4 loop :209-264, summary
 * :474-494,579-582).  out/iout are strided columns (stride = ld) of the emc_outputs blocks.
 * tape (optional): rows of [t, state[14]] for every stored state, up to tape_cap rows.
 */
static int flight(const emc_model *m, const double *col, int64_t stride, const double *wind,
                  double *out, int32_t *iout, int64_t ostride,
                  double *tape, int64_t tape_cap, int64_t *n_states_out, int want_diag)
{
    orc_sample s;
    sample_init(m, col, stride, wind, &s);
    double state[14];
    for (int i = 0; i < 13; ++i) state[i] = col[(EMC_IN_X + i) * stride];
    state[13] = 1.0;                                                                   /* :161 */
    int chute = 0;                                                                     /* :166 */
    double chute_time = NAN;

    /* ---- launch rail, simulator.py:42-125 ---- */
    double position[3] = { state[0], state[1], state[2] };
    double velocity[3] = { state[3], state[4], state[5] };
    const double *quaternion = state + 6;
    double prop_frac = state[13];
    double R[3][3];
    quaternion_to_rotation_matrix(quaternion, R);
    double direction[3] = { R[0][0], R[1][0], R[2][0] };                               /* :57 */
    double distance = 0.0, t = 0.0, dt = m->dt_initial;
    int rail_steps = 0;
    while (distance < m->rail_length && t < s.burn_time) {                             /* :63 */
        double mp[4];
        orc_mass_properties(m, s.dry_mass, s.propellant_mass, prop_frac, mp);
        double mass = mp[0];
        double temp, pres, density;
        orc_atmosphere(m, position[2], &temp, &pres, &density);
        double wind_vel[3];
        wind_at_altitude(m, &s, position[2], wind_vel);
        double speed = dot3(velocity, direction);                                      /* :75 */
        double rel_vel[3] = { direction[0] * speed - wind_vel[0], direction[1] * speed - wind_vel[1],
                              direction[2] * speed - wind_vel[2] };                   /* :76 */
        double rel_speed = dot3(rel_vel, direction);                                   /* :80 */
        double mach = mach_number(rel_vel, temp);                                      /* :81 */
        double c[6];
        orc_aero_coefficients(m, mach, 0.0, 0.0, mp[1], 1, s.cd_scale, c);             /* :82-83 */
        double drag = 0.5 * density * (rel_speed * rel_speed) * c[0] * m->reference_area; /* :84 */
        double thrust = orc_thrust_raw(m, &s, t, pres);                                /* :86 */
        double gravity = orc_gravity(m, position[2]);                                  /* :87 */
        double accel = (thrust - mass * gravity - drag) / mass;                        /* :88 */
        speed += accel * dt;                                                           /* :90 */
        for (int i = 0; i < 3; ++i) position[i] += direction[i] * speed * dt;          /* :91 */
        distance += speed * dt;                                                        /* :92 */
        for (int i = 0; i < 3; ++i) velocity[i] = direction[i] * speed;                /* :93 */
        t += dt;                                                                       /* :95 */
        prop_frac = propellant_remaining(&s, t);                                       /* :96 */
        ++rail_steps;
    }
    for (int i = 0; i < 3; ++i) { state[i] = position[i]; state[3 + i] = velocity[i]; }
    state[13] = prop_frac;
    double rail_time = t;
    {
        double e[3], wv[3], vb[3];
        quaternion_to_euler(quaternion, e);                                            /* :108 */
        wind_at_altitude(m, &s, position[2], wv);                                      /* :112-117 */
        double vel_rel[3] = { velocity[0] - wv[0], velocity[1] - wv[1], velocity[2] - wv[2] };
        for (int i = 0; i < 3; ++i) vb[i] = R[0][i] * vel_rel[0] + R[1][i] * vel_rel[1] + R[2][i] * vel_rel[2];
        out[EMC_OUT_RAIL_EXIT_TIME * ostride] = t;
        for (int i = 0; i < 3; ++i) {
            out[(EMC_OUT_RAIL_EXIT_X + i) * ostride] = position[i];
            out[(EMC_OUT_RAIL_EXIT_VX + i) * ostride] = velocity[i];
            out[(EMC_OUT_RAIL_EXIT_ROLL + i) * ostride] = e[i];
            out[(EMC_OUT_WIND_AT_EXIT_U + i) * ostride] = wv[i];
        }
        out[EMC_OUT_RAIL_EXIT_SPEED * ostride] = norm3(velocity);                      /* :107 */
        out[EMC_OUT_RAIL_EXIT_AOA * ostride] = angle_of_attack(vb);                    /* :121 */
        out[EMC_OUT_RAIL_EXIT_SIDESLIP * ostride] = sideslip_angle(vb);                /* :122 */
    }

    /* ---- RK4 loop, simulator.py:209-264 ---- */
    dt = py_min(m->dt_initial, 0.005);                                                 /* :209 */
    int64_t n_states = 1;
    if (tape && tape_cap > 0) { tape[0] = t; memcpy(tape + 1, state, 14 * sizeof(double)); }
    int apogee_detected = 0;
    double apogee_time = 0.0, max_coast_time = 0.0;
    /* running np.argmax(altitudes) (:488): first maximum, NaN counts as maximum */
    int64_t apogee_index = 0, first_nan = isnan(state[2]) ? 0 : -1;
    double apogee_alt = state[2], apogee_t = t;
    /* burnout_index = argmax(time > burn_time) on the SHIFTED time (:479-480) */
    double burnout_time = 0.0; int burnout_found = 0;
    orc_diag dg;
    if (want_diag) diag_state(m, &s, state, &dg, 1);
    int term = EMC_TERM_MAX_TIME;
    while (t < m->max_time) {                                                          /* :216 */
        double k1[14], k2[14], k3[14], k4[14], s2[14], s3[14], s4[14];
        rocket_dynamics(m, &s, t, state, &chute, &chute_time, k1);
        for (int i = 0; i < 14; ++i) s2[i] = state[i] + 0.5 * dt * k1[i];              /* :218 */
        rocket_dynamics(m, &s, t + 0.5 * dt, s2, &chute, &chute_time, k2);
        for (int i = 0; i < 14; ++i) s3[i] = state[i] + 0.5 * dt * k2[i];
        rocket_dynamics(m, &s, t + 0.5 * dt, s3, &chute, &chute_time, k3);
        for (int i = 0; i < 14; ++i) s4[i] = state[i] + dt * k3[i];
        rocket_dynamics(m, &s, t + dt, s4, &chute, &chute_time, k4);
        for (int i = 0; i < 14; ++i) state[i] += (dt / 6.0) * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]); /* :224 */
        double qn[4];
        normalize_quaternion(state + 6, qn);                                           /* :227 */
        memcpy(state + 6, qn, sizeof qn);
        t += dt;                                                                       /* :229 */
        if (tape && n_states < tape_cap) { tape[n_states * EMC_TAPE_WIDTH] = t; memcpy(tape + n_states * EMC_TAPE_WIDTH + 1, state, 14 * sizeof(double)); }
        double altitude = state[2], vertical_velocity = state[5];
        if (!isnan(apogee_alt) && (isnan(altitude) || altitude > apogee_alt)) { apogee_alt = altitude; apogee_index = n_states; apogee_t = t; }
        if (first_nan < 0 && isnan(altitude)) first_nan = n_states;
        if (!burnout_found && (t - rail_time) > s.burn_time) { burnout_found = 1; burnout_time = t - rail_time; }
        if (want_diag) diag_state(m, &s, state, &dg, 0);
        ++n_states;
        if (altitude <= 0.5 && vertical_velocity <= 0) { term = EMC_TERM_GROUND; break; }   /* :238 */
        if (altitude > 100000.0) { term = EMC_TERM_ALTITUDE; break; }                       /* :242 */
        if (altitude > 1000.0 && vertical_velocity < 0 && !apogee_detected) {               /* :247-257 */
            apogee_detected = 1;
            apogee_time = t;
            if (altitude > 50000.0) max_coast_time = 60.0;
            else if (altitude > 25000.0) max_coast_time = 120.0;
            else max_coast_time = 300.0;
        }
        if (apogee_detected && altitude > 25000.0) {                                        /* :260-264 */
            double coast_time = t - apogee_time;
            if (coast_time > max_coast_time) { term = EMC_TERM_COAST; break; }
        }
    }
    out[EMC_OUT_APOGEE_ALTITUDE * ostride] = apogee_alt;                               /* :490 */
    out[EMC_OUT_APOGEE_TIME * ostride] = apogee_t - rail_time;                         /* :464,489 */
    out[EMC_OUT_RANGE * ostride] = sqrt(state[0] * state[0] + state[1] * state[1]);    /* :494 */
    out[EMC_OUT_FLIGHT_TIME * ostride] = t - rail_time;                                /* :582 */
    for (int i = 0; i < 3; ++i) {
        out[(EMC_OUT_FINAL_X + i) * ostride] = state[i];
        out[(EMC_OUT_FINAL_VX + i) * ostride] = state[3 + i];
    }
    if (want_diag) {
        out[EMC_OUT_MAX_MACH * ostride] = dg.max_mach;
        out[EMC_OUT_MAX_Q * ostride] = dg.max_q;
        out[EMC_OUT_MAX_SPEED * ostride] = dg.max_speed;
        out[EMC_OUT_MAX_ABS_OMEGA * ostride] = dg.max_abs_omega;
        out[EMC_OUT_MIN_STABILITY * ostride] = dg.min_stab;
        out[EMC_OUT_MAX_STABILITY * ostride] = dg.max_stab;
        out[EMC_OUT_MAX_ABS_AOA * ostride] = dg.max_abs_aoa;
    } else {
        for (int f = EMC_OUT_MAX_MACH; f <= EMC_OUT_MAX_ABS_AOA; ++f) out[f * ostride] = NAN;
    }
    out[EMC_OUT_BURNOUT_TIME * ostride] = burnout_time;
    out[EMC_OUT_CHUTE_TIME * ostride] = chute_time;
    iout[EMC_IOUT_N_STEPS * ostride] = (int32_t)(n_states - 1);
    iout[EMC_IOUT_TERMINATION * ostride] = term;
    iout[EMC_IOUT_APOGEE_INDEX * ostride] = (int32_t)apogee_index;
    iout[EMC_IOUT_FIRST_NAN_STEP * ostride] = (int32_t)first_nan;
    iout[EMC_IOUT_RAIL_STEPS * ostride] = rail_steps;
    if (n_states_out) *n_states_out = n_states;
    return 0;
}

/* ---------------- exported entry points (ctypes from tests/ and bench.py) ---------------- */

/* out[i][14] = _rocket_dynamics(t[i], state[i]); chute[i] in/out */
ORC_API int emc_oracle_derivative(const emc_model *m, const emc_inputs *in, int64_t n, const double *t,
                                  const double *state, int32_t *chute, double *state_dot)
{
    for (int64_t i = 0; i < n; ++i) {
        orc_sample s;
        const double *wind = in->wind ? in->wind + i * in->wind_sample_stride : NULL;
        sample_init(m, in->scalars + i, in->ld, wind, &s);
        int c = chute ? chute[i] : 0;
        rocket_dynamics(m, &s, t[i], state + 14 * i, &c, NULL, state_dot + 14 * i);
        if (chute) chute[i] = c;
    }
    return 0;
}

/* n flights over n_threads pthreads (0 = all online cores) pulling sample indices from an atomic
 * counter; flags bit0: compute the per-state diagnostics */
typedef struct orc_job {
    const emc_model *m; const emc_inputs *in; const emc_outputs *o; int64_t n; int flags;
    atomic_llong next;
} orc_job;

static void *batch_worker(void *arg)
{
    orc_job *J = (orc_job *)arg;
    for (;;) {
        int64_t i = atomic_fetch_add(&J->next, 1);
        if (i >= J->n) break;
        const double *wind = J->in->wind ? J->in->wind + i * J->in->wind_sample_stride : NULL;
        flight(J->m, J->in->scalars + i, J->in->ld, wind, J->o->out + i, J->o->iout + i, J->o->ld,
               NULL, 0, NULL, J->flags & 1);
    }
    return NULL;
}

/* simulator.py:496-552 (_extract_results): derived series of ONE flight from its stored states.
 * series[EMC_SERIES_COUNT][n]; tape[i] = t (since ignition), state[14].  time[i] = t_i - t_0 (:464). */
ORC_API int emc_oracle_series(const emc_model *m, const emc_inputs *in, const double *tape, int64_t n, double *series)
{
    orc_sample s;
    sample_init(m, in->scalars, in->ld, in->wind, &s);
    const double rail_time = tape[0];
    for (int64_t i = 0; i < n; ++i) {
        const double *st = tape + i * EMC_TAPE_WIDTH + 1;
        double *o = series + i;
        double time_i = tape[i * EMC_TAPE_WIDTH] - rail_time;                           /* :464 */
        double e[3];
        quaternion_to_euler(st + 6, e);                                                 /* :512 */
        o[EMC_SER_EULER_ROLL * n] = e[0]; o[EMC_SER_EULER_PITCH * n] = e[1]; o[EMC_SER_EULER_YAW * n] = e[2];
        double mp[4];
        orc_mass_properties(m, s.dry_mass, s.propellant_mass, st[13], mp);              /* :515 */
        o[EMC_SER_MASS * n] = mp[0]; o[EMC_SER_CENTER_OF_MASS * n] = mp[1];
        o[EMC_SER_IXX * n] = mp[2]; o[EMC_SER_IYY * n] = mp[3]; o[EMC_SER_IZZ * n] = mp[3];
        double alt = st[2], T, p, rho;
        orc_atmosphere(m, alt, &T, &p, &rho);                                           /* :523 */
        double w[3];
        wind_at_altitude(m, &s, alt, w);                                                /* :525-528 */
        double vel_rel[3] = { st[3] - w[0], st[4] - w[1], st[5] - w[2] };              /* :530 */
        double R[3][3];
        quaternion_to_rotation_matrix(st + 6, R);
        double vb[3];
        for (int k = 0; k < 3; ++k) vb[k] = R[0][k] * vel_rel[0] + R[1][k] * vel_rel[1] + R[2][k] * vel_rel[2]; /* :531 */
        double mach = mach_number(vel_rel, T);                                          /* :532 */
        double aoa = angle_of_attack(vb), beta = sideslip_angle(vb);                    /* :533-534 */
        double cp_val = dynamic_cp(m, mach);                                            /* :535 */
        double c[6];
        orc_aero_coefficients(m, mach, aoa, beta, mp[1], st[13] > 0, s.cd_scale, c);    /* :536-539 */
        double vn = norm3(vel_rel);
        double q_dyn = 0.5 * rho * (vn * vn);                                           /* :541 */
        o[EMC_SER_DRAG * n] = q_dyn * c[0] * m->reference_area;                         /* :542 */
        o[EMC_SER_THRUST * n] = orc_thrust_raw(m, &s, time_i, p);                       /* :543 (shifted time) */
        o[EMC_SER_CD * n] = c[0]; o[EMC_SER_CL * n] = c[1]; o[EMC_SER_CM * n] = c[2];   /* :544-546 */
        o[EMC_SER_CP_DYNAMIC * n] = cp_val;                                             /* :548 */
        o[EMC_SER_STABILITY_MARGIN * n] = (cp_val - mp[1]) / m->reference_diameter;     /* :549 */
        o[EMC_SER_AOA * n] = aoa; o[EMC_SER_SIDESLIP * n] = beta;                       /* :551-552 */
        o[EMC_SER_SPEED * n] = sqrt(st[3] * st[3] + st[4] * st[4] + st[5] * st[5]);     /* :476 */
        o[EMC_SER_MACH * n] = mach; o[EMC_SER_QDYN * n] = q_dyn;
    }
    return 0;
}

ORC_API int emc_oracle_max_threads(void)
{
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    return c > 0 ? (int)c : 1;
}

ORC_API int emc_oracle_batch(const emc_model *m, const emc_inputs *in, int64_t n, const emc_outputs *o,
                             int n_threads, int flags)
{
    if (n_threads <= 0) n_threads = emc_oracle_max_threads();
    if (n_threads > n) n_threads = (int)(n > 0 ? n : 1);
    orc_job J = { m, in, o, n, flags, 0 };
    if (n_threads == 1) { batch_worker(&J); return 0; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    int started = 0;
    for (int k = 0; k < n_threads; ++k) { if (pthread_create(&th[k], NULL, batch_worker, &J) == 0) ++started; else break; }
    if (started == 0) batch_worker(&J);
    for (int k = 0; k < started; ++k) pthread_join(th[k], NULL);
    free(th);
    return 0;
}

ORC_API int emc_oracle_tape(const emc_model *m, const emc_inputs *in, const emc_outputs *o,
                            double *tape, int64_t cap, int64_t *n_states)
{
    return flight(m, in->scalars, in->ld, in->wind, o->out, o->iout, o->ld, tape, cap, n_states, 1);
}

ORC_API double orc_thrust(const emc_model *m, double thrust_a, double nozzle_area, double burn_time,
                          double time, double ambient_pressure)
{
    orc_sample s;
    memset(&s, 0, sizeof s);
    s.thrust_a = thrust_a; s.nozzle_area = nozzle_area; s.burn_time = burn_time;
    if (m->motor_kind == EMC_MOTOR_SOLID)
        for (int k = 0; k < m->n_thrust; ++k) s.thrust_knots[k] = m->thrust_curve[k] * thrust_a;
    return orc_thrust_raw(m, &s, time, ambient_pressure);
}

ORC_API void orc_quaternion_to_euler(const double q[4], double e[3]) { quaternion_to_euler(q, e); }
ORC_API void orc_rotation_matrix(const double q[4], double R[9]) { double M[3][3]; quaternion_to_rotation_matrix(q, M); memcpy(R, M, sizeof M); }
