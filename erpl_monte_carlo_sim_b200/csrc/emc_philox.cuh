/*
 * emc_philox.cuh — device-side dispersion draws and per-sample perturbation (reference
 * monte_carlo.py:156-201,225-288; motor.py:95-125,171-186; environment.py:125-200,218-265).
 *
 * Counter-based Philox4x32-10 (Salmon et al., Random123): sample i draws from counter (i, j, stream) with the run
 * seed as key, so any sample can be regenerated independently on any GPU.  The STRUCTURE of the reference's
 * streams is kept: one normal stream per sample is consumed from its start three times — by the
 * parameter dict (14 normals, then 2 uniforms, then 1 normal), by the motor perturbation and by the wind generator —
 * so thrust multiplier, position offsets and the first turbulence draws share their Gaussians exactly as in the
 * reference.  The draws themselves are Philox/Box-Muller, not MT19937/polar: this mode matches the reference in
 * distribution, the host-seeded mode matches it bit for bit.  With `gauss`/`unif` arrays supplied the same kernel
 * applies the perturbation to given draws (used to test it against the host-seeded inputs).
 */
#pragma once
#include <stdint.h>

#include "../../include/emc.h"

namespace emc {

struct u4 { uint32_t x, y, z, w; };

__host__ __device__ inline u4 philox4x32_10(u4 c, uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c.x, p1 = (uint64_t)M1 * c.z;
        u4 n;
        n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0; n.y = (uint32_t)p1;
        n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1; n.w = (uint32_t)p0;
        c = n; k0 += W0; k1 += W1;
    }
    return c;
}

/* two 32-bit words -> double in (0, 1): 53 random bits, never 0 */
__host__ __device__ inline double u01(uint32_t hi, uint32_t lo)
{
    const uint64_t b = (((uint64_t)hi << 32) | lo) >> 11;
    return ((double)b + 0.5) * (1.0 / 9007199254740992.0);
}

/* normals 2j and 2j+1 of sample `idx` (Box-Muller on one Philox block), stream 0 */
__device__ inline void normal_pair(uint64_t seed, uint64_t idx, uint32_t j, double &g0, double &g1)
{
    const u4 c = { (uint32_t)idx, (uint32_t)(idx >> 32), j, 0u };
    const u4 r = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u1 = u01(r.x, r.y), u2 = u01(r.z, r.w);
    const double rad = sqrt(-2.0 * log(u1));
    double s, co;
    sincospi(2.0 * u2, &s, &co);
    g0 = rad * co; g1 = rad * s;
}

__device__ inline void uniform_pair(uint64_t seed, uint64_t idx, double &a, double &b)
{
    const u4 c = { (uint32_t)idx, (uint32_t)(idx >> 32), 0u, 1u };
    const u4 r = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    a = u01(r.x, r.y); b = u01(r.z, r.w);
}

/* device copy of emc_dispersion with device table pointers */
struct DevDispersion {
    double base_pos[3], base_vel[3], base_att[3], base_omega[3];
    double sigma_pos[3], sigma_vel[3], sigma_att[3], sigma_omega[3];
    double mass_sigma, wind_speed_lo, wind_speed_hi, wind_dir_lo, wind_dir_hi;
    double dry_mass, propellant_mass;
    double thrust_vacuum, thrust_sea_level, mass_flow_rate, nozzle_exit_area, motor_propellant_mass, motor_burn_time;
    double thrust_sigma, flow_sigma, burn_sigma;
    int32_t motor_kind, wind_mode, n_knots, pad_;
    const double *shear, *base_wind, *rho, *innov;      /* device */
    double scale0;
};

struct NormalStream {
    uint64_t seed, idx; const double *given; int64_t stride;
    uint32_t cached_j; double c0, c1;
    __device__ NormalStream(uint64_t s, uint64_t i, const double *g, int64_t st) : seed(s), idx(i), given(g), stride(st), cached_j(0xffffffffu), c0(0), c1(0) {}
    __device__ double operator()(uint32_t k)
    {
        if (given) return given[k];
        const uint32_t j = k >> 1;
        if (j != cached_j) { normal_pair(seed, idx, j, c0, c1); cached_j = j; }
        return (k & 1u) ? c1 : c0;
    }
};

__global__ void __launch_bounds__(128) emc_generate_kernel(DevDispersion D, uint64_t seed, int64_t first, int64_t n,
                                                           const double *gauss /*[n][G] or null*/, int64_t G,
                                                           const double *unif /*[n][2] or null*/,
                                                           double *scal, int64_t ld, double *wind)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    NormalStream g(seed, (uint64_t)(first + i), gauss ? gauss + i * G : nullptr, 1);
    double u0, u1;
    if (unif) { u0 = unif[2 * i]; u1 = unif[2 * i + 1]; } else uniform_pair(seed, (uint64_t)(first + i), u0, u1);
    double *col = scal + i;
    /* monte_carlo.py:165-173,228-249: initial-condition offsets (normal(0, sigma) = 0 + sigma*g) */
    double att[3];
    for (int k = 0; k < 3; ++k) {
        col[(EMC_IN_X + k) * ld] = D.base_pos[k] + (0.0 + D.sigma_pos[k] * g(k));
        col[(EMC_IN_VX + k) * ld] = D.base_vel[k] + (0.0 + D.sigma_vel[k] * g(3 + k));
        att[k] = D.base_att[k] + (0.0 + D.sigma_att[k] * g(6 + k));
        col[(EMC_IN_WX + k) * ld] = D.base_omega[k] + (0.0 + D.sigma_omega[k] * g(9 + k));
    }
    {   /* utils.py:129-136,14-35 euler (xyz) -> [w,x,y,z] */
        double sr, cr, sp, cp, sy, cy;
        sincos(att[0] / 2, &sr, &cr); sincos(att[1] / 2, &sp, &cp); sincos(att[2] / 2, &sy, &cy);
        col[EMC_IN_Q1 * ld] = sr * cp * cy - cr * sp * sy;
        col[EMC_IN_Q2 * ld] = cr * sp * cy + sr * cp * sy;
        col[EMC_IN_Q3 * ld] = cr * cp * sy - sr * sp * cy;
        col[EMC_IN_Q0 * ld] = cr * cp * cy + sr * sp * sy;
    }
    const double mass_mult = 1.0 + D.mass_sigma * g(12);                               /* :169 */
    const double speed = D.wind_speed_lo + (D.wind_speed_hi - D.wind_speed_lo) * u0;   /* :171 */
    const double dir = D.wind_dir_lo + (D.wind_dir_hi - D.wind_dir_lo) * u1;           /* :172 */
    const double dry = D.dry_mass * mass_mult, prop = D.propellant_mass * mass_mult;   /* :315-316 */
    double mdot, own_burn;
    if (D.motor_kind == EMC_MOTOR_SOLID) {                                             /* motor.py:104-123 */
        const double k = 1.0 + D.thrust_sigma * g(0);
        mdot = 4.26 * k;
        col[EMC_IN_THRUST_A * ld] = k;
        col[EMC_IN_NOZZLE_AREA * ld] = D.nozzle_exit_area * k;
        own_burn = D.motor_burn_time * (1.0 + D.burn_sigma * g(1));
    } else {                                                                           /* motor.py:175-184 */
        const double kt = 1.0 + D.thrust_sigma * g(0), kf = 1.0 + D.flow_sigma * g(1);
        const double tv = D.thrust_vacuum * kt, tsl = D.thrust_sea_level * kt;
        mdot = D.mass_flow_rate * kf;
        col[EMC_IN_THRUST_A * ld] = tv;
        col[EMC_IN_NOZZLE_AREA * ld] = (tv - tsl) / 101325.0;
        own_burn = D.motor_propellant_mass / mdot;
    }
    col[EMC_IN_DRY_MASS * ld] = dry; col[EMC_IN_PROP_MASS * ld] = prop; col[EMC_IN_MDOT * ld] = mdot;
    col[EMC_IN_BURN_TIME * ld] = (mdot > 0.0) ? prop / mdot : own_burn;                /* monte_carlo.py:258-260 */
    col[EMC_IN_CD_SCALE * ld] = 1.0;
    if (!wind || D.n_knots <= 0) return;
    /* wind table: AR(1) turbulence about the mean profile, environment.py:156-198 (mode 0) / 242-263 (mode 1) */
    double sd, cd;
    sincos(dir, &sd, &cd);
    double *w = wind + i * (int64_t)D.n_knots * 3;
    double pu = 0.0, pv = 0.0, pw = 0.0, mu_prev = 0.0, mv_prev = 0.0, mw_prev = 0.0;
    const double off_u = speed * cd, off_v = speed * sd;                               /* monte_carlo.py:277-278 */
    for (int k = 0; k < D.n_knots; ++k) {
        double mu, mv, mw;
        if (D.wind_mode == 0) { const double level = speed * D.shear[k]; mu = level * cd; mv = level * sd; mw = 0.0; }
        else { mu = D.base_wind[3 * k]; mv = D.base_wind[3 * k + 1]; mw = D.base_wind[3 * k + 2]; }
        const double g0 = g(3 * k), g1 = g(3 * k + 1), g2 = g(3 * k + 2);
        double tu, tv, tw;
        if (k == 0) { tu = 0.0 + D.scale0 * g0; tv = 0.0 + D.scale0 * g1; tw = 0.0 + (D.scale0 * 0.3) * g2; }
        else {
            tu = D.rho[k] * (pu - mu_prev) + (0.0 + D.innov[k] * g0);
            tv = D.rho[k] * (pv - mv_prev) + (0.0 + D.innov[k] * g1);
            tw = D.rho[k] * (pw - mw_prev) + (0.0 + (D.innov[k] * 0.3) * g2);
        }
        pu = mu + tu; pv = mv + tv; pw = mw + tw;
        mu_prev = mu; mv_prev = mv; mw_prev = mw;
        w[3 * k] = (D.wind_mode == 0) ? pu : pu + off_u;
        w[3 * k + 1] = (D.wind_mode == 0) ? pv : pv + off_v;
        w[3 * k + 2] = pw;
    }
}

/* the draws the generator uses, for inspection: gauss[n][G], unif[n][2] */
__global__ void emc_philox_draws_kernel(uint64_t seed, int64_t first, int64_t n, int64_t G, double *gauss, double *unif)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    NormalStream g(seed, (uint64_t)(first + i), nullptr, 1);
    for (int64_t k = 0; k < G; ++k) gauss[i * G + k] = g((uint32_t)k);
    double a, b;
    uniform_pair(seed, (uint64_t)(first + i), a, b);
    unif[2 * i] = a; unif[2 * i + 1] = b;
}

/* ------------------------------------------------------------------------------------------------
 * NumPy-compatible draws on the device: MT19937 seeded like np.random.seed(i) / RandomState(i) (init_genrand) and
 * NumPy's legacy Gaussian (polar method with the cached second value) and 53-bit doubles.  One thread regenerates
 * the stream of ONE sample seed: gauss[i][0..G) are the first G standard normals of RandomState(seed) — the stream the
 * motor perturbation and the wind generator consume (monte_carlo.py:274,287,323) and whose first 14 values are also the
 * parameter draws (np.random.seed(i), monte_carlo.py:162-170) — and unif[i] the two uniforms that follow
 * those 14 normals in the parameter stream (:171-172), dens[i] the normal after them (:173).  Same bits as NumPy except where CUDA's log/sqrt differ from
 * the host libm in the last place.
 * ---------------------------------------------------------------------------------------------- */
struct MT19937 {
    uint32_t mt[624];
    int pos;
    __device__ void seed(uint32_t s)
    {
        for (int i = 0; i < 624; ++i) { mt[i] = s; s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)(i + 1); }
        pos = 624;
    }
    __device__ void twist()
    {
        const uint32_t UP = 0x80000000u, LO = 0x7fffffffu, A = 0x9908b0dfu;
        int k = 0;
        for (; k < 624 - 397; ++k) { const uint32_t y = (mt[k] & UP) | (mt[k + 1] & LO); mt[k] = mt[k + 397] ^ (y >> 1) ^ ((y & 1u) ? A : 0u); }
        for (; k < 623; ++k) { const uint32_t y = (mt[k] & UP) | (mt[k + 1] & LO); mt[k] = mt[k + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? A : 0u); }
        const uint32_t y = (mt[623] & UP) | (mt[0] & LO);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
        pos = 0;
    }
    __device__ uint32_t next32()
    {
        if (pos == 624) twist();
        uint32_t y = mt[pos++];
        y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
        return y;
    }
    __device__ double next_double()
    {
        const uint32_t a = next32() >> 5, b = next32() >> 6;
        return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
    }
    /* one polar-method pair: first = f*x2 (returned first by legacy_gauss), second = f*x1 (the cached value) */
    __device__ void gauss_pair(double &first, double &second)
    {
        double x1, x2, r2;
        do {
            x1 = 2.0 * next_double() - 1.0;
            x2 = 2.0 * next_double() - 1.0;
            r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));     /* no FMA contraction: near r2 = 1 one ulp of r2 is 1e-13 of log(r2) */
        } while (r2 >= 1.0 || r2 == 0.0);
        const double f = sqrt(-2.0 * log(r2) / r2);
        first = f * x2; second = f * x1;
    }
};

__global__ void __launch_bounds__(64) emc_numpy_draws_kernel(int64_t first_seed, int64_t n, int64_t G, double *gauss, double *unif, double *dens)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    MT19937 rng;
    rng.seed((uint32_t)((uint64_t)(first_seed + i) & 0xffffffffu));
    double *g = gauss + i * G;
    for (int64_t k = 0; k < G; k += 2) {
        double a, b;
        rng.gauss_pair(a, b);
        g[k] = a;
        if (k + 1 < G) g[k + 1] = b;
        if (k == 12) {
            /* after the 7th pair (14 normals) the PARAMETER stream draws two uniforms; the wind/motor stream goes on with
             * more normals from the same position.  Both read the same not-yet-overwritten first block of 624 words
             * (7 pairs consume ~36 words), so they are taken with a copy of the cursor. */
            const int keep = rng.pos;
            unif[2 * i] = rng.next_double();
            unif[2 * i + 1] = rng.next_double();
            double d0, d1;
            rng.gauss_pair(d0, d1);                 /* the parameter stream's 15th normal (density multiplier, :173) */
            dens[i] = d0;
            rng.pos = keep;
        }
    }
}

}  // namespace emc
