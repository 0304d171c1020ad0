/*
 * emc_stats.cuh — device-side Monte Carlo statistics (reference monte_carlo.py:337-473).
 *
 *   classify      _filter_physics_outliers (:348-390): non-finite, apogee > 80 km or < 100 m, range > 200 km,
 *                 flight time > 600 s, apogee > 1.2 * 1200^2/(2*9.81)
 *   moments1      counts by reason, sum/min/max of apogee, range, flight time and of the landing point (x, y)
 *                 over the VALID samples                                        -> one SUM block + MIN + MAX
 *   moments2      centred second moments about given means (np.std is two-pass, ddof = 0) and the landing
 *                 ellipse covariance                                             -> one SUM block
 *   select_hist   one digit pass of an exact radix select on the order-preserving 64-bit image of a
 *                 metric: for every requested prefix, the histogram of the next digit.  np.percentile's
 *                 order statistics are located in 6 passes; with several GPUs each pass's histogram is
 *                 all-reduced (NCCL) so the percentiles are exact over the whole job without gathering samples.
 *
 * Reductions are deterministic: per-block partials in a fixed order, then one block sums the partials.
 */
#pragma once
#include <stdint.h>

#include "../../include/emc.h"

namespace emc {

enum { ST_N = 0, ST_VALID, ST_OUTLIER, ST_NONFINITE, ST_AP_HIGH, ST_AP_LOW, ST_RANGE, ST_TIME, ST_ENERGY,
       ST_SUM_AP, ST_SUM_RG, ST_SUM_FT, ST_SUM_X, ST_SUM_Y, ST_SUM_COUNT };          /* SUM block */
enum { ST_MM_AP = 0, ST_MM_RG, ST_MM_FT, ST_MM_COUNT };                               /* MIN / MAX blocks */
enum { ST2_AP = 0, ST2_RG, ST2_FT, ST2_XX, ST2_XY, ST2_YY, ST2_COUNT };               /* centred second moments */
#define EMC_SELECT_BINS 2048
#define EMC_SELECT_MAX_PREFIX 16

__device__ __forceinline__ int classify_outlier(double ap, double rg, double ft, int *reasons)
{
    int r = 0;
    const double big = 1.7976931348623157e308;
    if (!(fabs(ap) <= big) || !(fabs(rg) <= big) || !(fabs(ft) <= big)) r |= 1;     /* :357 non-finite */
    if (ap > 80000.0) r |= 2;                                                       /* :362 */
    else if (ap < 100.0) r |= 4;                                                    /* :365 */
    if (rg > 200000.0) r |= 8;                                                      /* :370 */
    if (ft > 600.0) r |= 16;                                                        /* :375 */
    if (ap > (1200.0 * 1200.0 / (2 * 9.81)) * 1.2) r |= 32;                         /* :383-386 */
    *reasons = r;
    return r != 0;
}

/* order-preserving map double -> uint64 (NaN never reaches it: NaN samples are outliers) */
__device__ __forceinline__ unsigned long long ordered_key(double v)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double *partial /*[gridDim.x][NV]*/, int op /*0 sum,1 min,2 max*/)
{
    __shared__ double sh[32][NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = v[k];
        for (int o = 16; o > 0; o >>= 1) {
            const double y = __shfl_down_sync(0xffffffffu, x, o);
            x = (op == 0) ? x + y : ((op == 1) ? fmin(x, y) : fmax(x, y));
        }
        if (lane == 0) sh[warp][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double x = sh[0][threadIdx.x];
        for (int w = 1; w < nw; ++w) {
            const double y = sh[w][threadIdx.x];
            x = (op == 0) ? x + y : ((op == 1) ? fmin(x, y) : fmax(x, y));
        }
        partial[(size_t)blockIdx.x * NV + threadIdx.x] = x;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) emc_stats_moments1_kernel(const double *out, int64_t ld, int64_t n,
                                                                 double *p_sum, double *p_min, double *p_max)
{
    double s[ST_SUM_COUNT], mn[ST_MM_COUNT], mx[ST_MM_COUNT];
#pragma unroll
    for (int k = 0; k < ST_SUM_COUNT; ++k) s[k] = 0.0;
#pragma unroll
    for (int k = 0; k < ST_MM_COUNT; ++k) { mn[k] = INFINITY; mx[k] = -INFINITY; }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double ap = out[EMC_OUT_APOGEE_ALTITUDE * ld + i], rg = out[EMC_OUT_RANGE * ld + i], ft = out[EMC_OUT_FLIGHT_TIME * ld + i];
        int why;
        const int bad = classify_outlier(ap, rg, ft, &why);
        s[ST_N] += 1.0;
        if (bad) {
            s[ST_OUTLIER] += 1.0;
            s[ST_NONFINITE] += (why & 1) ? 1.0 : 0.0; s[ST_AP_HIGH] += (why & 2) ? 1.0 : 0.0; s[ST_AP_LOW] += (why & 4) ? 1.0 : 0.0;
            s[ST_RANGE] += (why & 8) ? 1.0 : 0.0; s[ST_TIME] += (why & 16) ? 1.0 : 0.0; s[ST_ENERGY] += (why & 32) ? 1.0 : 0.0;
        } else {
            s[ST_VALID] += 1.0;
            s[ST_SUM_AP] += ap; s[ST_SUM_RG] += rg; s[ST_SUM_FT] += ft;
            s[ST_SUM_X] += out[EMC_OUT_FINAL_X * ld + i]; s[ST_SUM_Y] += out[EMC_OUT_FINAL_Y * ld + i];
            mn[ST_MM_AP] = fmin(mn[ST_MM_AP], ap); mx[ST_MM_AP] = fmax(mx[ST_MM_AP], ap);
            mn[ST_MM_RG] = fmin(mn[ST_MM_RG], rg); mx[ST_MM_RG] = fmax(mx[ST_MM_RG], rg);
            mn[ST_MM_FT] = fmin(mn[ST_MM_FT], ft); mx[ST_MM_FT] = fmax(mx[ST_MM_FT], ft);
        }
    }
    block_reduce_store<ST_SUM_COUNT>(s, p_sum, 0);
    block_reduce_store<ST_MM_COUNT>(mn, p_min, 1);
    block_reduce_store<ST_MM_COUNT>(mx, p_max, 2);
}

__global__ void __launch_bounds__(256) emc_stats_moments2_kernel(const double *out, int64_t ld, int64_t n,
                                                                 const double *center /*[5] ap rg ft x y*/, double *p_sum)
{
    double s[ST2_COUNT];
#pragma unroll
    for (int k = 0; k < ST2_COUNT; ++k) s[k] = 0.0;
    const double c_ap = center[0], c_rg = center[1], c_ft = center[2], c_x = center[3], c_y = center[4];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double ap = out[EMC_OUT_APOGEE_ALTITUDE * ld + i], rg = out[EMC_OUT_RANGE * ld + i], ft = out[EMC_OUT_FLIGHT_TIME * ld + i];
        int why;
        if (classify_outlier(ap, rg, ft, &why)) continue;
        const double dx = out[EMC_OUT_FINAL_X * ld + i] - c_x, dy = out[EMC_OUT_FINAL_Y * ld + i] - c_y;
        s[ST2_AP] += (ap - c_ap) * (ap - c_ap); s[ST2_RG] += (rg - c_rg) * (rg - c_rg); s[ST2_FT] += (ft - c_ft) * (ft - c_ft);
        s[ST2_XX] += dx * dx; s[ST2_XY] += dx * dy; s[ST2_YY] += dy * dy;
    }
    block_reduce_store<ST2_COUNT>(s, p_sum, 0);
}

/* sums the per-block partials in block order: result[k] = op over b of partial[b][k] */
__global__ void emc_stats_finish_kernel(const double *partial, int nblocks, int nv, int op, double *result)
{
    const int k = threadIdx.x;
    if (k >= nv) return;
    double x = partial[k];
    for (int b = 1; b < nblocks; ++b) {
        const double y = partial[(size_t)b * nv + k];
        x = (op == 0) ? x + y : ((op == 1) ? fmin(x, y) : fmax(x, y));
    }
    result[k] = x;
}

struct SelectArgs {
    unsigned long long prefix[EMC_SELECT_MAX_PREFIX];
    int n_prefix, field, shift, prefix_shift;     /* digit = (key >> shift) & (BINS-1); prefix = key >> prefix_shift (64 -> all match) */
};

/* all three metrics in one pass over the samples: hist[f][u][digit] */
struct Select3Args {
    unsigned long long prefix[3][EMC_SELECT_MAX_PREFIX];
    int n_prefix[3], row0[3];          /* rows of metric f start at row0[f] in the compact histogram block */
    int shift, prefix_shift;
};

__global__ void __launch_bounds__(256) emc_stats_select3_kernel(const double *out, int64_t ld, int64_t n, Select3Args a,
                                                                unsigned long long *hist /*[n_prefix[0]+n_prefix[1]+n_prefix[2]][BINS], metric-major*/)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v[3] = { out[EMC_OUT_APOGEE_ALTITUDE * ld + i], out[EMC_OUT_RANGE * ld + i], out[EMC_OUT_FLIGHT_TIME * ld + i] };
        int why;
        if (classify_outlier(v[0], v[1], v[2], &why)) continue;
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            const unsigned long long key = ordered_key(v[f]);
            const unsigned long long pre = (a.prefix_shift >= 64) ? 0ull : (key >> a.prefix_shift);
            const unsigned digit = (unsigned)((key >> a.shift) & (EMC_SELECT_BINS - 1));
            for (int u = 0; u < a.n_prefix[f]; ++u)
                if (pre == a.prefix[f][u]) atomicAdd(&hist[(size_t)(a.row0[f] + u) * EMC_SELECT_BINS + digit], 1ull);
        }
    }
}

/* min, max and sum partials finished by ONE launch: result = [sum block | min block | max block] */
__global__ void emc_stats_finish3_kernel(const double *p_sum, const double *p_min, const double *p_max, int nblocks, double *result)
{
    const int k = threadIdx.x;
    if (k < ST_SUM_COUNT) {
        double x = p_sum[k];
        for (int b = 1; b < nblocks; ++b) x += p_sum[(size_t)b * ST_SUM_COUNT + k];
        result[k] = x;
    } else if (k < ST_SUM_COUNT + ST_MM_COUNT) {
        const int j = k - ST_SUM_COUNT;
        double x = p_min[j];
        for (int b = 1; b < nblocks; ++b) x = fmin(x, p_min[(size_t)b * ST_MM_COUNT + j]);
        result[k] = x;
    } else if (k < ST_SUM_COUNT + 2 * ST_MM_COUNT) {
        const int j = k - ST_SUM_COUNT - ST_MM_COUNT;
        double x = p_max[j];
        for (int b = 1; b < nblocks; ++b) x = fmax(x, p_max[(size_t)b * ST_MM_COUNT + j]);
        result[k] = x;
    }
}

__global__ void __launch_bounds__(256) emc_stats_select_kernel(const double *out, int64_t ld, int64_t n, SelectArgs a,
                                                               unsigned long long *hist /*[n_prefix][BINS]*/)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double ap = out[EMC_OUT_APOGEE_ALTITUDE * ld + i], rg = out[EMC_OUT_RANGE * ld + i], ft = out[EMC_OUT_FLIGHT_TIME * ld + i];
        int why;
        if (classify_outlier(ap, rg, ft, &why)) continue;
        const double v = (a.field == 0) ? ap : ((a.field == 1) ? rg : ft);
        const unsigned long long key = ordered_key(v);
        const unsigned long long pre = (a.prefix_shift >= 64) ? 0ull : (key >> a.prefix_shift);
        const unsigned digit = (unsigned)((key >> a.shift) & (EMC_SELECT_BINS - 1));
        for (int u = 0; u < a.n_prefix; ++u)
            if (pre == a.prefix[u]) atomicAdd(&hist[(size_t)u * EMC_SELECT_BINS + digit], 1ull);
    }
}

/* fixed-bin histogram of a metric over the valid samples: bin = floor((v - lo) / (hi - lo) * nbins), v == hi -> last bin
 * (numpy.histogram's convention) */
__global__ void __launch_bounds__(256) emc_stats_linear_hist_kernel(const double *out, int64_t ld, int64_t n, int field,
                                                                    double lo, double hi, int nbins, unsigned long long *hist)
{
    const double scale = (hi > lo) ? (double)nbins / (hi - lo) : 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double ap = out[EMC_OUT_APOGEE_ALTITUDE * ld + i], rg = out[EMC_OUT_RANGE * ld + i], ft = out[EMC_OUT_FLIGHT_TIME * ld + i];
        int why;
        if (classify_outlier(ap, rg, ft, &why)) continue;
        const double v = (field == 0) ? ap : ((field == 1) ? rg : ((field == 2) ? ft : out[(field == 3 ? EMC_OUT_FINAL_X : EMC_OUT_FINAL_Y) * ld + i]));
        if (!(v >= lo && v <= hi)) continue;
        int b = (int)((v - lo) * scale);
        b = b < 0 ? 0 : (b >= nbins ? nbins - 1 : b);
        atomicAdd(&hist[b], 1ull);
    }
}

/* ------------------------------------------------------------------------------------------------
 * The whole summary as ONE stream-ordered chain (single GPU): moments1 -> plan -> moments2 -> 6 x (digit histogram,
 * digit decision) -> values, with every intermediate (means, order-statistic ranks, radix-select prefixes) kept in
 * device memory, so the host synchronises once.  Same arithmetic as stats.compute_statistics / radix_select3.
 * Layout of the state block (8-byte words):
 *   [0, 20)            sum | min | max            (moments1)
 *   [20, 26)           centred second moments      (moments2)
 *   [26, 31)           means ap rg ft x y
 *   [32, 32 + 3 NT)    order statistics val[f][t], NT = 2 * n_pct targets per metric (lo and hi rank of each percentile)
 *   then prefix[3][NT] (u64), rem[3][NT] (i64), hist[3][NT][BINS] (u64)
 * ---------------------------------------------------------------------------------------------- */
#define EMC_SUMMARY_MAX_PCT 8
struct SummaryLayout {
    double *res;                  /* words [0, 32 + 3 NT) */
    unsigned long long *prefix;   /* [3][NT] */
    long long *rem;               /* [3][NT] */
    unsigned long long *hist;     /* [3][NT][BINS] */
    int nt;                       /* 2 * n_pct */
};
struct SummaryPct { double pct[EMC_SUMMARY_MAX_PCT]; int n_pct; };

/* np.percentile(method="linear") ranks: q = p/100, pos = (m-1) q, lo = floor(pos), hi = min(lo+1, m-1) */
__global__ void emc_stats_plan_kernel(SummaryLayout L, SummaryPct P)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double valid = L.res[ST_VALID];
    const long long m = llrint(valid);
    for (int k = 0; k < 5; ++k) L.res[26 + k] = (m > 0) ? L.res[ST_SUM_AP + k] / (double)m : NAN;
    for (int j = 0; j < P.n_pct; ++j) {
        const double q = P.pct[j] / 100.0;
        const double pos = (double)(m - 1) * q;
        long long lo = (long long)floor(pos), hi = lo + 1;
        if (hi > m - 1) hi = m - 1;
        for (int f = 0; f < 3; ++f) {
            L.prefix[f * L.nt + 2 * j] = 0ull;  L.rem[f * L.nt + 2 * j] = (m > 0) ? lo : -1;
            L.prefix[f * L.nt + 2 * j + 1] = 0ull;  L.rem[f * L.nt + 2 * j + 1] = (m > 0) ? hi : -1;
        }
    }
}

__global__ void __launch_bounds__(256) emc_stats_select_dev_kernel(const double *out, int64_t ld, int64_t n, SummaryLayout L,
                                                                   int shift, int prefix_shift)
{
    __shared__ unsigned long long pre_sh[3 * 2 * EMC_SUMMARY_MAX_PCT];
    __shared__ int live;
    if (threadIdx.x == 0) live = (L.rem[0] >= 0);
    if (threadIdx.x < 3 * L.nt) pre_sh[threadIdx.x] = L.prefix[threadIdx.x];
    __syncthreads();
    if (!live) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v[3] = { out[EMC_OUT_APOGEE_ALTITUDE * ld + i], out[EMC_OUT_RANGE * ld + i], out[EMC_OUT_FLIGHT_TIME * ld + i] };
        int why;
        if (classify_outlier(v[0], v[1], v[2], &why)) continue;
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            const unsigned long long key = ordered_key(v[f]);
            const unsigned long long pre = (prefix_shift >= 64) ? 0ull : (key >> prefix_shift);
            const unsigned digit = (unsigned)((key >> shift) & (EMC_SELECT_BINS - 1));
            for (int t = 0; t < L.nt; ++t)
                if (pre == pre_sh[f * L.nt + t]) atomicAdd(&L.hist[((size_t)(f * L.nt + t)) * EMC_SELECT_BINS + digit], 1ull);
        }
    }
}

/* one warp per (metric, target): b = first digit whose cumulative count exceeds the remaining rank
 * (np.searchsorted(cum, rem, side="right")), prefix <- prefix << width | b, rem -= cum[b-1]; the row is zeroed for the
 * next pass.  After the last pass the prefix is the full key: its value is written to val. */
__global__ void __launch_bounds__(32) emc_stats_select_finish_kernel(SummaryLayout L, int width, int last)
{
    const int row = blockIdx.x, lane = threadIdx.x;
    long long rem = L.rem[row];
    if (rem < 0) { if (last && lane == 0) L.res[32 + row] = NAN; return; }
    unsigned long long *h = L.hist + (size_t)row * EMC_SELECT_BINS;
    const int per = EMC_SELECT_BINS / 32;
    unsigned long long mine = 0;
    for (int k = 0; k < per; ++k) mine += h[lane * per + k];
    unsigned long long incl = mine;                                  /* inclusive scan over lanes */
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    const unsigned long long before = incl - mine;
    const bool here = (unsigned long long)rem >= before && (unsigned long long)rem < incl;     /* the target digit is in my chunk */
    const unsigned who = __ballot_sync(0xffffffffu, here);
    if (who == 0u) { if (last && lane == 0) L.res[32 + row] = NAN; return; }                   /* rank beyond the population */
    const int owner = __ffs(who) - 1;
    if (lane == owner) {
        unsigned long long cum = before;
        int b = lane * per;
        for (;; ++b) { const unsigned long long c = h[b]; if ((unsigned long long)rem < cum + c) break; cum += c; }
        const unsigned long long pre = (L.prefix[row] << width) | (unsigned long long)b;
        L.prefix[row] = pre;
        L.rem[row] = rem - (long long)cum;
        if (last) {
            const unsigned long long bits = (pre >> 63) ? (pre & 0x7fffffffffffffffull) : ~pre;
            L.res[32 + row] = __longlong_as_double((long long)bits);
        }
    }
    __syncwarp();
    for (int k = 0; k < per; ++k) h[lane * per + k] = 0ull;
}

}  // namespace emc
