/*
 * emc_engine.cu — libemc.so: the sm_100a kernels and the C ABI of include/emc.h.
 *
 * Kernels
 *   emc_rail_kernel        one thread per sample, convergent: the launch-rail Euler loop
 *                          (reference simulator.py:42-125) -> rail_* outputs = state at rail exit.
 *   emc_flight_kernel      PERSISTENT: one trajectory per lane — the stage state and the derivative in registers, the
 *                          base state, the RK4 accumulator and the per-lane bookkeeping in shared memory; lanes whose
 *                          trajectory ended (simulator.py:238-264) are found with a warp ballot and
 *                          refilled from a global atomic work queue, so a warp never idles behind its
 *                          longest flight.  Run-constant tables (Cd/CP vs Mach, thrust curve, wind
 *                          altitude grid) are staged once into shared memory; scalars and the atmosphere
 *                          polynomials sit in __constant__ memory.  Summaries are written as field-major SoA.
 *   emc_stats_*_kernel     classification, moments, exact percentiles (emc_stats.cuh); emc_generate_kernel /
 *                          emc_numpy_draws_kernel: dispersions drawn on the device (emc_philox.cuh).
 *   emc_derivative_kernel  test seam: one derivative evaluation per thread (simulator.py:295-460).
 *   emc_dfma_kernel        register-resident DFMA chains: measures the FP64 roofline denominator.
 *
 * There is no CPU implementation in this library: every entry point needs a CUDA device.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <new>
#include <string>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include "emc_model_build.h"
#include "emc_strict.cuh"
#include "emc_stats.cuh"
#include "emc_philox.cuh"

using namespace emc;

#ifdef EMC_YIELD_DEBUG
#include <vector>
#include <algorithm>
#endif
#define EMC_EXPORT extern "C" __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------ */
__constant__ DevModel c_model;
__constant__ DevTables c_tables;

struct KernelArgs {
    const double *scalars; int64_t ld;
    const double *wind; int64_t wind_stride;
    const double *wind_alt;            /* device copy of the altitude grid */
    double *out; int32_t *iout; int64_t old;
    int64_t n;
    unsigned long long *queue;         /* [0] next sample index */
    unsigned long long *counters;      /* [0] rk4 steps [1] replays [2] rail steps [3] refills */
    double *tape; int64_t tape_cap; int64_t *tape_n;
    int32_t refill_threshold; int32_t nan_ff;
    /* downsampled batch tape (emc_tape_request): slot of every sample (-1: not recorded), rows[n_sel][bt_max][4], counts */
    const int32_t *bt_slot; double *bt_rows; int32_t *bt_count; int32_t bt_stride, bt_max;
    /* once-per-step bookkeeping of the 16-warp variant: [TC_DCOUNT][gcold_ld] doubles, [TI_ICOUNT][gcold_ld] ints */
    double *gcold_d; int32_t *gcold_i; int64_t gcold_ld;
    int32_t sm_count, compact;         /* tail compaction (shared-memory variant): collector choice, on/off */
    /* strict continuation: trajectories parked by the flight kernel (emc_strict.cuh) */
    struct ParkRec *park; unsigned long long *park_count; int32_t park_on;
    /* streaming hand-over to emc_strict_kernel, which runs CONCURRENTLY on a second stream: a record is published by
     * writing the run's epoch into it after its contents; consumers draw tickets from park_next and wait for their record
     * or until no flight warp can publish any more: the sample queue is exhausted and every warp that has
     * started has finished (flight_started == flight_done; blocks that become resident later find the queue empty) */
    int32_t epoch, flight_warps;
    unsigned long long *park_next, *flight_started, *flight_done, *strict_err;
    /* lane hand-back (emc_counters.yielded): records of the flights that gave their lane to an unstarted sample, claimed in
     * order once the sample queue is empty; histogram of the attitude-rate amplitudes seen at the hand-back point */
    struct ParkRec *resume; int32_t resume_cap;      /* resume_cap records per warp, warp w owns [w cap, (w + 1) cap) */
    int32_t yield_step, yield_half;    /* hand-back point and half of it in stored states; -1: off */
#ifdef EMC_YIELD_DEBUG
    unsigned int *dbg;                 /* [n][4] event times (globaltimer >> 10): start, hand-back, resume, end */
#endif
};
#ifdef EMC_YIELD_DEBUG
__device__ __forceinline__ void dbg_mark(const KernelArgs &a, int64_t idx, int ev)
{
    if (!a.dbg) return;
    unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    a.dbg[idx * 4 + ev] = (unsigned int)(t >> 10);
}
#define DBG_MARK(a, idx, ev) dbg_mark(a, idx, ev)
#else
#define DBG_MARK(a, idx, ev) ((void)0)
#endif

/* everything the strict kernel needs to finish a parked trajectory (its sample index is C.i[TI_SAMPLE]); `epoch` is
 * written last (after a fence) and marks the record as belonging to this run */
struct ParkRec { State s; TrackHot K; TrackCold C; int32_t epoch, pad_; };

/* Stage the run-constant tables into shared memory.  The tables are a STATIC __shared__ object and the wind
 * altitude grid the only dynamic part: objects addressed as shared arrays are read with plain LDS offsets, whereas
 * pointers carved out of `extern __shared__` storage by pointer arithmetic are generic and made the compiler
 * rebuild the shared-window address (S2UR CgaCtaId + ULEA) before every table access (profiles/, round 1). */
__device__ __forceinline__ void stage_tables(DevTables &tb, double *alt, const KernelArgs &a)
{
    const double *src = reinterpret_cast<const double *>(&c_tables);
    double *dst = reinterpret_cast<double *>(&tb);
    constexpr int NT = sizeof(DevTables) / sizeof(double);
    for (int i = threadIdx.x; i < NT; i += blockDim.x) dst[i] = src[i];
    const int nw = c_model.n_wind;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) alt[i] = a.wind_alt[i];
    __syncthreads();
}

static size_t smem_bytes(int n_wind) { return sizeof(double) * (size_t)(n_wind > 0 ? n_wind : 0); }

/* ------------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(128) emc_rail_kernel(KernelArgs a)
{
    extern __shared__ double alt[];
    __shared__ DevTables Tb;
    stage_tables(Tb, alt, a);
    unsigned long long steps = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        Sample S;
        load_sample(c_model, a.scalars + i, a.ld, a.wind ? a.wind + i * a.wind_stride : nullptr, S);
        int rs = rail_phase(c_model, Tb, alt, S, a.scalars + i, a.ld, a.out + i, a.old);
        a.iout[EMC_IOUT_RAIL_STEPS * a.old + i] = rs;
        steps += (unsigned long long)rs;
    }
    for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
    if ((threadIdx.x & 31) == 0 && steps) atomicAdd(a.counters + 2, steps);
}

/* ------------------------------------------------------------------------------------------------ */
/* Per-lane state that is touched once per step or per derivative (event bookkeeping, sample constants,
 * the remembered table brackets) can live in shared memory instead of registers (COLD >= 1): the
 * derivative then fits in fewer registers and more warps are resident to hide the FP64 latency.
 * The structs span an odd number of 8-byte words, so lane-strided access is conflict-free.
 * COLD == 3 keeps only the HOT half of the bookkeeping here; the once-per-step half (running maxima, event times,
 * tape cursor) goes to global memory, field-major over the resident lanes (GlobalCold): 200 B + 224 B of state store
 * per lane let 512 lanes (16 warps) share an SM. */
struct alignas(8) ColdLaneFull { TrackHot K; TrackCold C; Sample S; WindBracket WB; };
struct alignas(8) ColdLaneHot { TrackHot K; Sample S; WindBracket WB; };
template <bool PAD> struct ColdPad { double pad_; };
template <> struct ColdPad<false> {};
template <class Raw> struct Padded : Raw, ColdPad<(sizeof(Raw) / 8) % 2 == 0> {};
static_assert(sizeof(Padded<ColdLaneFull>) % 8 == 0 && (sizeof(Padded<ColdLaneFull>) / 8) % 2 == 1, "lane state must span an odd number of 8-byte words");
static_assert(sizeof(Padded<ColdLaneHot>) % 8 == 0 && (sizeof(Padded<ColdLaneHot>) / 8) % 2 == 1, "lane state must span an odd number of 8-byte words");

/* cold bookkeeping in global memory: d[field][lane], i[field][lane], lane = resident lane of the grid */
struct GlobalCold {
    double *d; int32_t *i; int64_t ld;
    __device__ __forceinline__ double getd(int f) const { return d[f * ld]; }
    __device__ __forceinline__ void setd(int f, double v) const { d[f * ld] = v; }
    __device__ __forceinline__ int32_t geti(int f) const { return i[f * ld]; }
    __device__ __forceinline__ void seti(int f, int32_t v) const { i[f * ld] = v; }
};

/* Dynamic shared memory of the flight kernel: [14][BLOCK] PAIRS of state words (7 pairs of the base state, 7 of the RK4
 * accumulator; COLD >= 2) followed by the wind altitude grid.  It is indexed directly (never through a pointer carved out
 * of it), so the accesses are plain LDS/STS — 128-bit ones for the pairs: a warp reads 32 consecutive 16-byte words. */
extern __shared__ __align__(16) double emc_dyn[];

/* one row {t - t_rail, x, y, z} of the downsampled batch tape: a 32-byte store per lane (one sector) */
__device__ __forceinline__ void bt_write(const KernelArgs &a, int32_t slot, int32_t row, double t, double x, double y, double z)
{
    if (row < a.bt_max) {
        double4 *dst = reinterpret_cast<double4 *>(a.bt_rows + ((int64_t)slot * a.bt_max + row) * EMC_BTAPE_WIDTH);
        *dst = make_double4(t, x, y, z);
    }
}

/* A shared-memory address the compiler cannot rebuild: ptxas otherwise REMATERIALISES the address of a lane record
 * (S2R CgaCtaId, S2R tid, shift, add: six instructions and two long-latency special-register reads) twice per
 * derivative instead of holding it in one register (profiles/, round 2).  The round trip through the 32-bit shared
 * window address keeps the address space known, so the accesses stay LDS/STS. */
template <class T>
__device__ __forceinline__ T *opaque_shared(T *p)
{
#ifndef EMC_NO_OPAQUE
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));
    return reinterpret_cast<T *>(__cvta_shared_to_generic((size_t)a));
#else
    return p;
#endif
}

/* base state + RK4 accumulator of this thread's lane record (a column of the [14][BLOCK] pair block at the start of emc_dyn) */
template <int BLOCK>
struct SharedStore {
    static constexpr bool kShared = true;      /* a copy of the object addresses the same state (lane hand-back helpers) */
    Pair *col;
    __device__ __forceinline__ SharedStore() : col(opaque_shared(reinterpret_cast<Pair *>(emc_dyn) + threadIdx.x)) {}
    __device__ __forceinline__ Pair s2(int p) const { return col[p * BLOCK]; }
    __device__ __forceinline__ void set_s2(int p, Pair v) { col[p * BLOCK] = v; }
    __device__ __forceinline__ Pair acc2(int p) const { return col[(7 + p) * BLOCK]; }
    __device__ __forceinline__ void set_acc2(int p, Pair v) { col[(7 + p) * BLOCK] = v; }
    __device__ __forceinline__ double s(int i) const { return reinterpret_cast<const double *>(col + (i >> 1) * BLOCK)[i & 1]; }
    __device__ __forceinline__ void set_s(int i, double v) { reinterpret_cast<double *>(col + (i >> 1) * BLOCK)[i & 1] = v; }
    __device__ __forceinline__ void adopt(int src) {                      /* take over another record's column */
#pragma unroll
        for (int p = 0; p < 14; ++p) col[p * BLOCK] = reinterpret_cast<const Pair *>(emc_dyn)[p * BLOCK + src];
    }
};
struct RegStoreLane : RegStore { static constexpr bool kShared = false; __device__ __forceinline__ void adopt(int) {} };

/* Tail compaction (north star: "retired with warp ballot/compaction").  Once the work queue is empty the warps of a
 * block thin out: a warp that still flies ONE trajectory costs its scheduler as many issue slots as a full one, and an
 * SM whose 12 warps each hold a straggler runs every one of them at the loaded step latency.  All per-trajectory state
 * of the shared-memory variants lives in the lane RECORD (hot/cold bookkeeping, sample constants, brackets, state
 * store column); a thread only holds {active, sample index}.  So a sparse warp can hand its trajectories to the
 * block's collector warp by posting their record numbers on a board and exit; the collector's idle lanes adopt them.  `reserved` bounds collector lanes in use + posted: a donation is made only if it fits, so a posted
 * trajectory is adopted at the collector's next iteration.  The arithmetic of a trajectory does not depend on the lane
 * that integrates it: outputs are bit-identical with and without compaction (tests/).
 * (First version: records addressed through a slot number, no copy — the indirection cost 7 % of the step latency
 * everywhere.  Now the adopting lane COPIES the 536 bytes of the record into its own slot, once per hand-over, and the
 * steady-state code is the plain thread-indexed one.) */
template <int BLOCK>
struct CompactBoard {
    int reserved;                 /* collector lanes active or promised; starts at 32 (closed) until the collector drains */
    int posted;                   /* slots written to `slots` (reservation cursor) */
    int ready;                    /* slots whose records are visible (release counter) */
    int exited;                   /* warps other than the collector that have left the loop */
    short slots[BLOCK];
};

/* hand a trajectory to the strict continuation: copy its state and bookkeeping into the next park record and publish it */
template <class Store, class CA>
__device__ __forceinline__ void park_lane(const KernelArgs &a, int64_t idx, const Store &st, TrackHot &K, const CA &C)
{
    ParkRec &P = a.park[atomicAdd(a.park_count, 1ull)];
    store_get(st, P.s);
    K.replay = 0;
    P.K = K;
    for (int f = 0; f < TC_DCOUNT; ++f) P.C.d[f] = C.getd(f);
    for (int f = 0; f < TI_ICOUNT; ++f) P.C.i[f] = C.geti(f);
    P.C.i[TI_SAMPLE] = (int32_t)idx;
    __threadfence();
    *reinterpret_cast<volatile int32_t *>(&P.epoch) = a.epoch;       /* publish */
}

/* ---- lane hand-back -------------------------------------------------------------------------------------------------
 * A batch larger than the resident lanes is flown in waves, and the launch ends with its longest trajectory: T = start of
 * that trajectory + its steps x the single-trajectory step latency.  Which sample flies longest is not known in advance,
 * but after EMC_YIELD_STEP stored states it shows in the attitude rate: a flight whose amplitude max |omega| is still
 * GROWING (it has risen by a factor of ~1.8 or more since step EMC_YIELD_STEP / 2) is diverging and ends soon; one whose
 * amplitude has settled flies on, and the longest flights of a batch are all of that kind (measured on the oracle, 14 000
 * samples of the headline workload, states 1 000 against 500: ratio 1.00 .. 1.24 for every flight beyond 3 200 steps,
 * median 8.8 and 10th percentile 1.9 for the flights of 2 000 .. 3 200 steps, 11 for the shorter ones; states 900 against
 * 450: at most 1.5 for the long flights, 87 % of the others at 1.8 or more; launch -> landing flights: 1.0, 90th
 * percentile 1.3 — they never hand back).  So while unstarted samples are waiting, a flight that reaches that state with a growing
 * amplitude puts its state aside (one record in global memory) and its lane starts a fresh sample: every sample is
 * started within about two hand-back periods of the launch, and the long flights never wait.  The records of a warp form
 * that warp's own list (head and tail live in shared memory, positions come from ballots: no atomics, no waiting —
 * one list for all warps was measured first: 1 776 warps claiming from one head with compare-and-swap cost more than the
 * hand-back returns); a warp hands back no more than its share of the unstarted samples and resumes its records, oldest
 * first, on the lanes that fall idle once the sample queue is empty.  Brackets are caches and the state is stored exactly,
 * so the outputs are bit-identical to the undisturbed flight. */
#ifndef EMC_YIELD_STEP
#define EMC_YIELD_STEP 900          /* measured on the eight per-rank batches of the headline workload: 800 / 900 / 1 000 states give 43.3 /
                                     * 43.3 / 44.0 ms on average; at 800 one long flight of sixteen already reads x1.9, at 900 none above x1.5 */
#endif
#define EMC_YIELD_QUORUM 16         /* lanes of a warp that must want to hand back in the same iteration */
#define EMC_YIELD_GROWTH 7          /* omega_code units: a rise by a factor of 1.7 .. 1.9 or more */

/* max |omega| on a logarithmic scale that fits a byte: exponent and three mantissa bits from 2^-20 rad/s (0) up to
 * 2^12 rad/s (255); NaN -> 255 */
__device__ __forceinline__ int omega_code(double om)
{
    const int c = (__double2hiint(fabs(om)) >> 17) - ((1023 - 20) << 3);
    return c < 0 ? 0 : (c > 255 ? 255 : c);
}

template <class Store, class CA>
__device__ __forceinline__ void yield_lane(ParkRec &P, int64_t idx, const Store &st, const TrackHot &K, const CA &C)
{
    store_get(st, P.s);
    P.K = K;
    for (int f = 0; f < TC_DCOUNT; ++f) P.C.d[f] = C.getd(f);
    for (int f = 0; f < TI_ICOUNT; ++f) P.C.i[f] = C.geti(f);
    P.C.i[TI_SAMPLE] = (int32_t)idx;
}

/* The two refill-path pieces of the hand-back are kept OUT of line: inside the persistent loop they would only be code the
 * hot path has to jump across and registers its allocation has to respect (measured: +2 % on the steady-state rate with
 * both in line).  Values go in and come back by value so that nothing of the loop's state has to live in memory. */
struct YieldOut { unsigned act, handed; int active; };

/* the hand-back point of a warp whose lanes started together: the lanes of `ymask` (flights with a growing attitude
 * oscillation) write their records at consecutive positions of the warp's list and become idle */
template <int NW, class Store, class LANES>
__device__ __noinline__ YieldOut yield_write(const KernelArgs a, LANES lanes, Store st, bool active, int64_t idx, unsigned act, unsigned ymask, int *tail_p)
{
    const unsigned lane = threadIdx.x & 31u;
    YieldOut r = { act, 0u, active ? 1 : 0 };
    /* a warp hands back no more flights than its share of the samples that no lane could start at once: the fresh samples
     * then spread evenly over the warps, and so do the records that wait for them to finish */
    const int64_t all_warps = (int64_t)gridDim.x * NW;
    const int my_share = (a.n > all_warps * 32) ? (int)((a.n - all_warps * 32 + all_warps - 1) / all_warps) : 0;
    const int tail = *tail_p;
    const int want = __popc(ymask);
    int allow = my_share - tail;
    allow = allow < 0 ? 0 : (allow > a.resume_cap - tail ? a.resume_cap - tail : allow);
    /* only a warp whose flights are diverging TOGETHER hands back: its records are resumed as soon as those neighbours have
     * ended, which is soon.  A lone growing flight among settled ones would wait for a lane of its warp for as long as the
     * settled flights last (measured on launch -> landing flights: 0.4 % of them show a growing amplitude, and each
     * handed-back one ended a third of a flight late). */
    if (want < EMC_YIELD_QUORUM || *reinterpret_cast<volatile unsigned long long *>(a.queue) >= (unsigned long long)a.n) allow = 0;
    allow = want < allow ? want : allow;
    const int rank = __popc(ymask & ((1u << lane) - 1u));
    const bool mine = ((ymask >> lane) & 1u) && rank < allow;
    if (mine) {
        auto &RC = lanes.rec(threadIdx.x);
        ParkRec *const my_recs = a.resume + (size_t)(blockIdx.x * NW + (threadIdx.x >> 5)) * (size_t)a.resume_cap;
        yield_lane(my_recs[tail + rank], idx, st, RC.K, lanes.cold(threadIdx.x));
        r.active = 0;
        DBG_MARK(a, idx, 1);
    }
    __syncwarp();
    if (lane == 0) *tail_p = tail + allow;
    __syncwarp();
    r.handed = __ballot_sync(0xffffffffu, mine);
    r.act = act & ~r.handed;
    return r;
}

struct ResumeOut { unsigned act; int active; long long idx; };

/* idle lanes (mask `idle`) resume the oldest handed-back flights of this warp */
template <int NW, class Store, class LANES>
__device__ __noinline__ ResumeOut yield_resume(const KernelArgs a, LANES lanes, Store st, bool active, int64_t idx, unsigned idle, int *head_p, int tail)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned FULL = 0xffffffffu;
    ResumeOut r = { 0u, active ? 1 : 0, (long long)idx };
    const int head = *head_p;
    int k = tail - head;
    const int nidle = __popc(idle);
    const ParkRec *const my_recs = a.resume + (size_t)(blockIdx.x * NW + (threadIdx.x >> 5)) * (size_t)a.resume_cap;
    k = k < nidle ? k : nidle;
    const int rank = __popc(idle & ((1u << lane) - 1u));
    if (((idle >> lane) & 1u) && rank < k) {
        const ParkRec &P = my_recs[head + rank];
        /* L2 reads (the record was written by a lane of this warp, before a __syncwarp) straight into the lane's slot */
        idx = __ldcg(&P.C.i[TI_SAMPLE]);
        auto &RC = lanes.rec(threadIdx.x);
        const auto C = lanes.cold(threadIdx.x);
        load_sample(c_model, a.scalars + idx, a.ld, a.wind ? a.wind + idx * a.wind_stride : nullptr, RC.S);
        {
            State s;
            double *sp = reinterpret_cast<double *>(&s);
            const double *src = reinterpret_cast<const double *>(&P.s);
            for (int c = 0; c < 14; ++c) sp[c] = __ldcg(src + c);
            store_put(st, s);
        }
        {
            static_assert(sizeof(TrackHot) % 8 == 0, "TrackHot is copied word by word");
            union { TrackHot k; unsigned long long w[sizeof(TrackHot) / 8]; } u;
            const unsigned long long *src = reinterpret_cast<const unsigned long long *>(&P.K);
            for (int w = 0; w < (int)(sizeof(TrackHot) / 8); ++w) u.w[w] = __ldcg(src + w);
            RC.K = u.k;
        }
        for (int f = 0; f < TC_DCOUNT; ++f) C.setd(f, __ldcg(&P.C.d[f]));
        for (int f = 0; f < TI_ICOUNT; ++f) C.seti(f, __ldcg(&P.C.i[f]));
        wind_bracket_reset(RC.WB);
        r.active = 1; r.idx = idx;
        DBG_MARK(a, idx, 2);
    }
    __syncwarp();
    if (lane == 0) *head_p = head + k;
    __syncwarp();
    r.act = __ballot_sync(FULL, r.active != 0);
    return r;
}

/* the persistent loop, generic over where the lane records live (REC: registers or shared memory) and the cold accessor */
template <int BLOCK, class Store, bool COMPACT, int MK, int WK, class LANES>
__device__ __forceinline__ void flight_loop(const KernelArgs &a, const DevTables &Tb, const double *alt, LANES lanes,
                                            CompactBoard<BLOCK> *board)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned FULL = 0xffffffffu;
    constexpr int NW = BLOCK / 32;

    bool active = false, drained = false;
    Store st;
    int64_t idx = -1;
    unsigned long long n_steps = 0, n_replay = 0, n_refill = 0, n_tape = 0, n_yield = 0;
    /* lane hand-back (above).  Its state is kept out of the registers that live through the step: head and tail of the
     * warp's list in shared memory, one iteration counter and one flag in registers, everything else is recomputed where it
     * is needed (the refill path) */
    constexpr bool YIELD_OK = !COMPACT && Store::kShared;      /* the kernel instances whose lane state lives in shared memory */
    const bool yield_on = YIELD_OK && a.yield_step > 0;
    __shared__ int yl_head_s[NW], yl_tail_s[NW];
    int &my_head = yl_head_s[threadIdx.x >> 5], &my_tail = yl_tail_s[threadIdx.x >> 5];
    if (lane == 0) { my_head = 0; my_tail = 0; }
    __syncwarp();
    bool have_recs = false;         /* my_head < my_tail, warp-uniform: the idle path of the tail does not read shared memory for it */
    const int thr = a.refill_threshold < 1 ? 1 : (a.refill_threshold > 32 ? 32 : a.refill_threshold);
    /* compaction roles: blocks that share an SM (ids differ by the SM count) pick collectors on different schedulers */
    const int warp = threadIdx.x >> 5;
    const bool compact = COMPACT;
    const bool collector = compact && (warp == (int)((blockIdx.x / (unsigned)a.sm_count) % NW));
    if (a.flight_started && lane == 0) atomicAdd(a.flight_started, 1ull);        /* before this warp's first claim on the queue */
    bool opened = false;          /* collector: `reserved` switched from "closed" to its own lane count */
    int taken = 0, last_cnt = 33; /* collector: board entries adopted; donor: active count at the last donation attempt */

    int it = 0;                     /* steps taken by the lanes that started with the warp */
    for (;;) {
        /* ---- retire/refill: ONE ballot per iteration in the steady state; idle lanes are ranked with
         * popc and fetch their sample indices with one atomic per warp ---- */
        unsigned act = __ballot_sync(FULL, active);
        /* the hand-back point (and the point half way to it) of the flights that started with this warp: one warp-uniform
         * test per iteration, nothing in the step section (yield_step = yield_half = -1 when the hand-back is off) */
        const bool trig = YIELD_OK && (it == a.yield_step || it == a.yield_half);
        if (act != FULL || trig) {
            unsigned handed = 0u;        /* lanes that hand their flight back in this iteration: they start fresh samples */
            if (trig && yield_on) {
                auto &RC = lanes.rec(threadIdx.x);
                const bool cohort = active && RC.K.n_steps == it;       /* started with the warp, a step in every iteration since */
                if (it != a.yield_step) {
                    if (cohort) RC.K.om_half = (uint8_t)omega_code(lanes.cold(threadIdx.x).getd(TC_MAX_OM));
                } else {
                    const bool growing = cohort && !RC.K.finishing &&
                                         omega_code(lanes.cold(threadIdx.x).getd(TC_MAX_OM)) - (int)RC.K.om_half >= EMC_YIELD_GROWTH;
                    const unsigned ymask = __ballot_sync(FULL, growing);
                    if (ymask && !drained) {
                        const YieldOut r = yield_write<NW>(a, lanes, st, active, idx, act, ymask, &my_tail);
                        act = r.act; handed = r.handed; active = r.active != 0; n_yield += (unsigned)__popc(r.handed & (1u << lane));
                        have_recs = my_head < my_tail;
                    }
                }
            }
            if (have_recs && (~act & ~(drained ? 0u : handed))) {
                /* a lane whose flight has ENDED resumes the oldest handed-back flight of this warp before it would start a
                 * fresh sample (a record never waits for the whole queue); the lanes that have just handed back start fresh
                 * samples — that is what they gave their flight up for — unless the queue is empty */
                const ResumeOut r = yield_resume<NW>(a, lanes, st, active, idx, ~act & ~(drained ? 0u : handed), &my_head, my_tail);
                act = r.act; active = r.active != 0; idx = r.idx;
                have_recs = my_head < my_tail;
            }
            if (!drained) {
                const unsigned idle = ~act;
                const int nidle = __popc(idle);
                if (nidle >= thr) {
                    const int leader = __ffs(idle) - 1;
                    unsigned long long base = 0;
                    if ((int)lane == leader) base = atomicAdd(a.queue, (unsigned long long)nidle);
                    base = __shfl_sync(FULL, base, leader);
                    if (base + (unsigned long long)nidle >= (unsigned long long)a.n) drained = true;
                    if (!active) {
                        const int64_t my = (int64_t)base + __popc(idle & ((1u << lane) - 1u));
                        if (my < a.n) {
                            idx = my;
                            auto &R = lanes.rec(threadIdx.x);
                            TrackHot &K = R.K; Sample &S = R.S; WindBracket &WB = R.WB;
                            const auto C = lanes.cold(threadIdx.x);
                            load_sample(c_model, a.scalars + idx, a.ld, a.wind ? a.wind + idx * a.wind_stride : nullptr, S);
                            double t_rail;
                            State s;
                            load_flight_state(S, a.scalars + idx, a.ld, a.out + idx, a.old, s, t_rail);
                            store_put(st, s);
                            track_init(K, C, s, t_rail);
                            C.seti(TI_SAMPLE, (int32_t)idx);
                            wind_bracket_reset(WB);
                            if (!(K.t < c_model.max_time)) { K.term = EMC_TERM_MAX_TIME; K.finishing = true; }
                            if (a.tape && a.tape_cap > 0) {
                                a.tape[0] = K.t;
                                const double *sp = reinterpret_cast<const double *>(&s);
                                for (int c = 0; c < 14; ++c) a.tape[1 + c] = sp[c];
                            }
                            if (a.bt_slot) {
                                const int32_t slot = a.bt_slot[idx];
                                C.seti(TI_BT_SLOT, slot); C.seti(TI_BT_NEXT, a.bt_stride);
                                if (slot >= 0) bt_write(a, slot, 0, 0.0, s.x, s.y, s.z);
                            }
                            active = true;
                            ++n_refill;
                            DBG_MARK(a, idx, 0);
                            /* the fast path assumes a regular sample (derivative<., ., REG>): anything else — a non-positive
                             * or non-finite mass, degenerate inertia constants — is flown by the strict continuation
                             * from its first state, with the reference's own guards */
                            if (!sample_regular(c_model, S)) { park_lane(a, idx, st, K, C); active = false; }
                        }
                    }
                    act = __ballot_sync(FULL, active);
                }
            }
            if (compact && drained) {
                const int cnt = __popc(act);
                if (collector) {
                    if (!opened) {                       /* open the board: from "closed" (32) down to the lanes in use */
                        if (lane == 0) atomicSub(&board->reserved, 32 - cnt);
                        opened = true;
                    } else if (cnt < last_cnt && lane == 0) {
                        atomicSub(&board->reserved, last_cnt - cnt);        /* own trajectories that retired */
                    }
                    int rdy = 0;
                    if (lane == 0) rdy = *reinterpret_cast<volatile int *>(&board->ready);
                    rdy = __shfl_sync(FULL, rdy, 0);
                    int k = rdy - taken;
                    if (k > 0) {                         /* adopt: the reservation guarantees that the idle lanes suffice */
                        const unsigned idle = ~act;
                        const int rank = __popc(idle & ((1u << lane) - 1u));
                        if (!active && rank < k) {
                            __threadfence_block();
                            const int src = board->slots[taken + rank];
                            lanes.adopt(threadIdx.x, src);       /* record + state-store column into this lane's slot */
                            st.adopt(src);
                            idx = lanes.cold(threadIdx.x).geti(TI_SAMPLE);
                            active = true;
                            atomicAdd(a.counters + 8, 1ull);
                        }
                        taken += k;
                        act = __ballot_sync(FULL, active);
                    }
                    last_cnt = __popc(act);
                    if (act == 0u) {
                        /* leave only after every other warp has left and everything it posted has been adopted */
                        int ex = 0, po = 0;
                        if (lane == 0) { ex = *reinterpret_cast<volatile int *>(&board->exited); po = *reinterpret_cast<volatile int *>(&board->posted); }
                        ex = __shfl_sync(FULL, ex, 0); po = __shfl_sync(FULL, po, 0);
                        if (ex == NW - 1 && po == taken) break;
                        __nanosleep(256);                /* wait for the other warps without taking their issue slots */
                        continue;
                    }
                } else if (cnt > 0 && cnt <= 16 && cnt < last_cnt) {
                    /* donor: try once per change of the active count */
                    last_cnt = cnt;
                    int ok = 0, pos = 0;
                    if (lane == 0) {
                        const int r = atomicAdd(&board->reserved, cnt);
                        if (r + cnt <= 32) { ok = 1; pos = atomicAdd(&board->posted, cnt); }
                        else atomicSub(&board->reserved, cnt);
                    }
                    ok = __shfl_sync(FULL, ok, 0); pos = __shfl_sync(FULL, pos, 0);
                    if (ok) {
                        if (active) board->slots[pos + __popc(act & ((1u << lane) - 1u))] = (short)threadIdx.x;
                        __threadfence_block();
                        __syncwarp();
                        if (lane == 0) atomicAdd(&board->ready, cnt);
                        active = false;
                        act = 0u;
                    }
                }
            }
            if (act == 0u) {
                if (drained && !have_recs) break;
                continue;
            }
        }
        if (active) {
            auto &R = lanes.rec(threadIdx.x);
            TrackHot &K = R.K; Sample &S = R.S; WindBracket &WB = R.WB;
            const auto C = lanes.cold(threadIdx.x);
            bool stepped; int64_t rep = 0;
            const bool retired = lane_advance<Store, decltype(C), MK, WK, true>(c_model, Tb, alt, S, WB, K, C, st, a.nan_ff != 0, stepped, rep, true);
            if (stepped) {
                ++n_steps;
                if (a.tape && (int64_t)K.n_steps < a.tape_cap) {
                    double *row = a.tape + (int64_t)K.n_steps * EMC_TAPE_WIDTH;
                    row[0] = K.t;
                    for (int c = 0; c < 14; ++c) row[1 + c] = st.s(c);
                }
                if (a.bt_slot) {                                            /* every bt_stride-th stored state */
                    const int32_t slot = C.geti(TI_BT_SLOT);
                    if (slot >= 0 && K.n_steps == C.geti(TI_BT_NEXT)) {
                        bt_write(a, slot, K.n_steps / a.bt_stride, K.t - K.t_rail, st.s(0), st.s(1), st.s(2));
                        C.seti(TI_BT_NEXT, K.n_steps + a.bt_stride);
                    }
                }
            }
            if (retired && K.replay == EMC_REPLAY_PARK) {
                /* blow-up under way: hand the trajectory to the strict continuation (emc_strict_kernel) */
                park_lane(a, idx, st, K, C);
                active = false;
                DBG_MARK(a, idx, 3);
            } else if (retired) {
                n_replay += (unsigned long long)rep;
                State s; store_get(st, s);
                write_flight_outputs(K, C, s, a.out + idx, a.iout + idx, a.old);
                if (a.tape_n) *a.tape_n = (int64_t)K.n_steps + 1 - rep;
                if (a.bt_slot) {
                    const int32_t slot = C.geti(TI_BT_SLOT);
                    if (slot >= 0) {
                        /* the last integrated state closes the trajectory (a fast-forwarded NaN tail is not recorded) */
                        const int32_t last = K.n_steps - (int32_t)rep;
                        int32_t rows = last / a.bt_stride + 1;
                        if (rep == 0 && last % a.bt_stride != 0) { bt_write(a, slot, rows, K.t - K.t_rail, s.x, s.y, s.z); ++rows; }
                        a.bt_count[slot] = rows;
                        n_tape += (unsigned long long)(rows < a.bt_max ? rows : a.bt_max);
                    }
                }
                active = false;
                DBG_MARK(a, idx, 3);
            }
        }
        ++it;
    }
    if (compact && !collector && lane == 0) atomicAdd(&board->exited, 1);
    for (int o = 16; o > 0; o >>= 1) {
        n_steps += __shfl_down_sync(FULL, n_steps, o);
        n_replay += __shfl_down_sync(FULL, n_replay, o);
        n_refill += __shfl_down_sync(FULL, n_refill, o);
        n_tape += __shfl_down_sync(FULL, n_tape, o);
        n_yield += __shfl_down_sync(FULL, n_yield, o);
    }
    if (lane == 0) {
        if (n_steps) atomicAdd(a.counters + 0, n_steps);
        if (n_replay) atomicAdd(a.counters + 1, n_replay);
        if (n_refill) atomicAdd(a.counters + 3, n_refill);
        if (n_tape) atomicAdd(a.counters + 7, n_tape);
        if (n_yield) atomicAdd(a.counters + 20, n_yield);
        if (a.flight_done) { __threadfence(); atomicAdd(a.flight_done, 1ull); }   /* after every park of this warp */
    }
}

/* where the lane records live */
template <class REC> struct SmemLanesFull {          /* hot + cold halves side by side in shared memory */
    REC *recs, *me;                                  /* me: this thread's own record (opaque_shared) */
    __device__ __forceinline__ REC &rec(int) const { return *me; }
    __device__ __forceinline__ ColdStruct cold(int) const { return ColdStruct(me->C); }
    __device__ __forceinline__ void adopt(int dst, int src) const {
        const double *s = reinterpret_cast<const double *>(&recs[src]);
        double *d = reinterpret_cast<double *>(&recs[dst]);
#pragma unroll
        for (int w = 0; w < (int)(sizeof(REC) / 8); ++w) d[w] = s[w];
    }
};
template <class REC> struct SmemLanesHot {           /* hot half in shared memory, cold half in global memory */
    REC *recs; double *gd; int32_t *gi; int64_t ld;
    __device__ __forceinline__ REC &rec(int slot) const { return recs[slot]; }
    __device__ __forceinline__ GlobalCold cold(int slot) const { return GlobalCold{ gd + slot, gi + slot, ld }; }
    __device__ __forceinline__ void adopt(int, int) const {}
};
struct RegLanes {                                    /* everything in this thread's registers */
    ColdLaneFull *one;
    __device__ __forceinline__ ColdLaneFull &rec(int) const { return *one; }
    __device__ __forceinline__ ColdStruct cold(int) const { return ColdStruct(one->C); }
    __device__ __forceinline__ void adopt(int, int) const {}
};

/* COLD: 0 everything in registers; 1 bookkeeping / sample constants / brackets in shared memory; 2 additionally the base
 *       state and the RK4 accumulator (SharedStore) — the default; 4 as 2 with the lane records addressed through a slot
 *       number and tail compaction (EMC_RUN_COMPACTION); 3 as 2 with the once-per-step half of the bookkeeping in global
 *       memory (16 warps per SM; measured slower, kept selectable).
 * MK / WK: the motor kind and the presence of a wind table compiled in (-1: read from the model) */
template <int BLOCK, int COLD, int MK, int WK>
__device__ __forceinline__ void flight_body(const KernelArgs &a)
{
    __shared__ DevTables Tb;
    if constexpr (COLD == 3) {
        /* dynamic shared memory: [28][BLOCK] state store | BLOCK hot lane records | wind altitude grid (static shared
         * memory stops at 48 KB) */
        constexpr int HOT_WORDS = sizeof(Padded<ColdLaneHot>) / 8;
        double *alt = emc_dyn + (28 + HOT_WORDS) * BLOCK;
        stage_tables(Tb, alt, a);
        const int64_t g0 = (int64_t)blockIdx.x * BLOCK;
        SmemLanesHot<Padded<ColdLaneHot>> lanes = { reinterpret_cast<Padded<ColdLaneHot> *>(emc_dyn + 28 * BLOCK),
                                                    a.gcold_d + g0, a.gcold_i + g0, a.gcold_ld };
        flight_loop<BLOCK, SharedStore<BLOCK>, false, MK, WK>(a, Tb, alt, lanes, (CompactBoard<BLOCK> *)nullptr);
    } else if constexpr (COLD == 2 || COLD == 4) {
        /* lane records: static shared memory while they fit under its 48 KB limit, else behind the state store in the
         * dynamic block (one large block per SM) */
        constexpr bool REC_STATIC = sizeof(Padded<ColdLaneFull>) * BLOCK <= 40960;
        __shared__ Padded<ColdLaneFull> sh_cold[REC_STATIC ? BLOCK : 1];
        __shared__ CompactBoard<BLOCK> board;
        constexpr int REC_WORDS = REC_STATIC ? 0 : (int)(sizeof(Padded<ColdLaneFull>) / 8) * BLOCK;
        Padded<ColdLaneFull> *recs = REC_STATIC ? sh_cold : reinterpret_cast<Padded<ColdLaneFull> *>(emc_dyn + 28 * BLOCK);
        double *alt = emc_dyn + 28 * BLOCK + REC_WORDS;
        if (threadIdx.x == 0) { board.reserved = 32; board.posted = 0; board.ready = 0; board.exited = 0; }
        stage_tables(Tb, alt, a);                    /* ends with __syncthreads() */
        SmemLanesFull<Padded<ColdLaneFull>> lanes = { recs, opaque_shared(recs + threadIdx.x) };
        if constexpr (COLD == 4) flight_loop<BLOCK, SharedStore<BLOCK>, true, MK, WK>(a, Tb, alt, lanes, &board);
        else flight_loop<BLOCK, SharedStore<BLOCK>, false, MK, WK>(a, Tb, alt, lanes, &board);
    } else if constexpr (COLD == 1) {
        __shared__ Padded<ColdLaneFull> sh_cold[BLOCK];
        stage_tables(Tb, emc_dyn, a);
        SmemLanesFull<Padded<ColdLaneFull>> lanes = { sh_cold, opaque_shared(sh_cold + threadIdx.x) };
        flight_loop<BLOCK, RegStoreLane, false, MK, WK>(a, Tb, emc_dyn, lanes, (CompactBoard<BLOCK> *)nullptr);
    } else {
        stage_tables(Tb, emc_dyn, a);
        ColdLaneFull CL;
        RegLanes lanes = { &CL };
        flight_loop<BLOCK, RegStoreLane, false, MK, WK>(a, Tb, emc_dyn, lanes, (CompactBoard<BLOCK> *)nullptr);
    }
}

#ifdef EMC_SLIM_MAXNREG      /* developer builds: a register cap that launch bounds cannot express (odd warp counts) */
#define EMC_FLIGHT_BOUNDS __maxnreg__(EMC_SLIM_MAXNREG)
#else
#define EMC_FLIGHT_BOUNDS __launch_bounds__(BLOCK, MINB)
#endif
template <int BLOCK, int MINB, int COLD, int MK = -1, int WK = -1>
__global__ void EMC_FLIGHT_BOUNDS emc_flight_kernel(KernelArgs a) { flight_body<BLOCK, COLD, MK, WK>(a); }

/* ------------------------------------------------------------------------------------------------ */
/* One thread per parked trajectory: the strict continuation (emc_strict.cuh) to the end of the flight, then the
 * summary.  Convergent in code (every lane is strict), ragged in length (a handful of steps each; NaN tails are
 * replayed in closed form). */
struct StrictTape {
    const KernelArgs *a; ColdStruct C;
    __device__ __forceinline__ void operator()(const TrackHot &K, const double *st) const
    {
        if (a->tape && (int64_t)K.n_steps < a->tape_cap) {
            double *row = a->tape + (int64_t)K.n_steps * EMC_TAPE_WIDTH;
            row[0] = K.t;
            for (int c = 0; c < 14; ++c) row[1 + c] = st[c];
        }
        if (a->bt_slot) {
            const int32_t slot = C.geti(TI_BT_SLOT);
            if (slot >= 0 && K.n_steps == C.geti(TI_BT_NEXT)) {
                bt_write(*a, slot, K.n_steps / a->bt_stride, K.t - K.t_rail, st[0], st[1], st[2]);
                C.seti(TI_BT_NEXT, K.n_steps + a->bt_stride);
            }
        }
    }
};

/* The strict continuation runs on the context's second stream WHILE the flight kernel flies (one 64-lane block of <= 168
 * registers and no shared memory per SM fits next to the three resident flight blocks): every thread draws a ticket,
 * waits until the record of that number has been published — or until none can be published any more —, claims it
 * (epoch -> -epoch), finishes the flight, draws again.
 *   - Termination does not depend on flight blocks that are not resident yet (they would wait for the consumers'
 *     resources in turn): once the sample queue is exhausted, parks can only come from warps that have already started.
 *   - A consumer never has to succeed: a thread that has waited too long (no flight warp running anywhere after 20 ms —
 *     the two grids did not become co-resident —, or 2 s in any case) simply leaves.  Whatever is still unclaimed when
 *     the flight kernel has finished is swept by a second, ordinary launch of the same code on the main stream
 *     (emc_strict_tail_kernel: usually nothing).  So the pair cannot deadlock and cannot lose a record. */
#ifndef EMC_STRICT_BLOCK
#define EMC_STRICT_BLOCK 64
#define EMC_STRICT_MINB 8        /* 128 registers: three flight blocks (<= 144 registers a thread) and a consumer block share an SM */
#endif
__device__ __forceinline__ bool no_more_parks(const KernelArgs &a)
{
    /* order matters: queue, then finished, then started (equality then proves that every warp started by the last read
     * had finished by the second, and a warp that starts later finds no sample) */
    if (*reinterpret_cast<volatile unsigned long long *>(a.queue) < (unsigned long long)a.n) return false;
    __threadfence();
    const unsigned long long fin = *reinterpret_cast<volatile unsigned long long *>(a.flight_done);
    __threadfence();
    const unsigned long long sta = *reinterpret_cast<volatile unsigned long long *>(a.flight_started);
    return sta == fin;
}

/* finish the parked flight of record P (already claimed by the caller) */
__device__ __noinline__ void finish_parked(const KernelArgs &a, ParkRec &P, unsigned long long &steps, unsigned long long &replays)
{
    /* tables and altitude grid are read where they are (constant bank, global memory): a consumer block must not take
     * shared memory from the flight blocks it runs beside */
    const DevTables &Tb = c_tables;
    const double *alt = a.wind_alt;
    /* The record is read with L2 loads into a private copy.  Records are not cache-line aligned: a neighbour on this SM
     * that read record p - 1 may have pulled the line holding the head of record p into L1 BEFORE that record was
     * written, and L1 is not coherent. */
    ParkRec R;
    {
        static_assert(sizeof(ParkRec) % 8 == 0, "ParkRec is copied word by word");
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(&P);
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(&R);
        for (int w = 0; w < (int)(sizeof(ParkRec) / 8); ++w) dst[w] = __ldcg(src + w);
    }
    const int64_t idx = R.C.i[TI_SAMPLE];
    if (idx < 0 || idx >= a.n) { atomicAdd(a.strict_err, 1ull); return; }      /* not a record of this run: refuse */
    Sample S;
    load_sample(c_model, a.scalars + idx, a.ld, a.wind ? a.wind + idx * a.wind_stride : nullptr, S);
    TrackHot K = R.K;
    const ColdStruct C(R.C);
    double st[14];
    memcpy(st, &R.s, sizeof st);
    int64_t ns = 0;
    const StrictTape tape = { &a, C };
    const int64_t rep = strict_fly(c_model, Tb, alt, S, K, C, st, a.nan_ff != 0, &ns, tape);
    State s;
    memcpy(&s, st, sizeof s);
    write_flight_outputs(K, C, s, a.out + idx, a.iout + idx, a.old);
    if (a.tape_n) *a.tape_n = (int64_t)K.n_steps + 1 - rep;
    if (a.bt_slot) {
        const int32_t slot = C.geti(TI_BT_SLOT);
        if (slot >= 0) {
            const int32_t last = K.n_steps - (int32_t)rep;
            int32_t rows = last / a.bt_stride + 1;
            if (rep == 0 && last % a.bt_stride != 0) { bt_write(a, slot, rows, K.t - K.t_rail, s.x, s.y, s.z); ++rows; }
            a.bt_count[slot] = rows;
            atomicAdd(a.counters + 7, (unsigned long long)(rows < a.bt_max ? rows : a.bt_max));
        }
    }
    steps += (unsigned long long)ns; replays += (unsigned long long)rep;
}

__device__ __forceinline__ void strict_counts(const KernelArgs &a, unsigned long long steps, unsigned long long replays)
{
    __syncwarp();
    for (int o = 16; o > 0; o >>= 1) {
        steps += __shfl_down_sync(0xffffffffu, steps, o);
        replays += __shfl_down_sync(0xffffffffu, replays, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (steps) atomicAdd(a.counters + 9, steps);
        if (replays) atomicAdd(a.counters + 1, replays);
    }
}

__global__ void __launch_bounds__(EMC_STRICT_BLOCK, EMC_STRICT_MINB) emc_strict_kernel(KernelArgs a)
{
    unsigned long long steps = 0, replays = 0;
    const long long t_start = clock64();
    for (;;) {
        const unsigned long long p = atomicAdd(a.park_next, 1ull);
        if (p >= (unsigned long long)a.n) break;                       /* more tickets than samples: nothing can follow */
        ParkRec &P = a.park[p];
        bool have = false;
        for (unsigned spin = 0;; ++spin) {
            if (*reinterpret_cast<volatile int32_t *>(&P.epoch) == a.epoch) { have = true; break; }
            /* the shared counters are read on every eighth poll only: thousands of waiting threads, three cache lines */
            if ((spin & 7u) == 7u) {
                if (no_more_parks(a)) {
                    __threadfence();                                   /* every park precedes its warp's count */
                    have = *reinterpret_cast<volatile int32_t *>(&P.epoch) == a.epoch;
                    break;
                }
                const long long waited = clock64() - t_start;
                const bool nobody = *reinterpret_cast<volatile unsigned long long *>(a.flight_started) == 0ull;
                if ((nobody && waited > 40000000ll) || waited > 4000000000ll) {       /* ~20 ms / ~2 s: leave it to the sweep */
                    atomicAdd(a.strict_err + 1, 1ull);                 /* counted (d_ctrl[15 + ...]), not an error */
                    break;
                }
            }
            __nanosleep(spin < 16 ? 1000 : 8000);
        }
        if (!have) break;
        __threadfence();
        *reinterpret_cast<volatile int32_t *>(&P.epoch) = -a.epoch;    /* claimed */
        finish_parked(a, P, steps, replays);
    }
    strict_counts(a, steps, replays);
}

/* the sweep after the flight kernel: every record that is published and not claimed */
__global__ void __launch_bounds__(EMC_STRICT_BLOCK, EMC_STRICT_MINB) emc_strict_tail_kernel(KernelArgs a)
{
    unsigned long long steps = 0, replays = 0;
    const unsigned long long n_parked = *a.park_count;
    for (unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_parked; p += (unsigned long long)gridDim.x * blockDim.x) {
        ParkRec &P = a.park[p];
        if (__ldcg(&P.epoch) != a.epoch) continue;
        finish_parked(a, P, steps, replays);
    }
    strict_counts(a, steps, replays);
}

/* ------------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(128) emc_derivative_kernel(KernelArgs a, const double *t, const double *state,
                                                             int32_t *chute, double *state_dot)
{
    extern __shared__ double alt[];
    __shared__ DevTables Tb;
    stage_tables(Tb, alt, a);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    Sample S;
    load_sample(c_model, a.scalars + i, a.ld, a.wind ? a.wind + i * a.wind_stride : nullptr, S);
    WindBracket WB; wind_bracket_reset(WB);
    State s, k; Diag dg;
    const double *sp = state + 14 * i;
    s.x = sp[0]; s.y = sp[1]; s.z = sp[2]; s.vx = sp[3]; s.vy = sp[4]; s.vz = sp[5];
    s.q0 = sp[6]; s.q1 = sp[7]; s.q2 = sp[8]; s.q3 = sp[9]; s.wx = sp[10]; s.wy = sp[11]; s.wz = sp[12]; s.pf = sp[13];
    bool ch = chute[i] != 0; double ct = 0.0;
    derivative(c_model, Tb, alt, S, WB, t[i], s, ch, ct, k, true, dg);
    chute[i] = ch ? 1 : 0;
    double *kp = state_dot + 14 * i;
    kp[0] = k.x; kp[1] = k.y; kp[2] = k.z; kp[3] = k.vx; kp[4] = k.vy; kp[5] = k.vz;
    kp[6] = k.q0; kp[7] = k.q1; kp[8] = k.q2; kp[9] = k.q3; kp[10] = k.wx; kp[11] = k.wy; kp[12] = k.wz; kp[13] = k.pf;
}

/* ------------------------------------------------------------------------------------------------ */
/* one thread per stored state of ONE flight: the derived series of _extract_results (simulator.py:511-552) */
__global__ void __launch_bounds__(128) emc_series_kernel(KernelArgs a, const double *tape, int64_t n_states, double *series)
{
    extern __shared__ double alt[];
    __shared__ DevTables Tb;
    stage_tables(Tb, alt, a);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_states) return;
    Sample S;
    load_sample(c_model, a.scalars, a.ld, a.wind, S);
    const double t_rail = tape[0];
    const double *row = tape + i * EMC_TAPE_WIDTH;
    series_state(c_model, Tb, alt, S, row, row[0] - t_rail, series + i, n_states);
}

/* ------------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(128) emc_component_kernel(int comp, int64_t n, const double *in, double *out)
{
    __shared__ DevTables Tb;
    {
        const double *src = reinterpret_cast<const double *>(&c_tables);
        double *dst = reinterpret_cast<double *>(&Tb);
        for (int i = threadIdx.x; i < (int)(sizeof(DevTables) / sizeof(double)); i += blockDim.x) dst[i] = src[i];
        __syncthreads();
    }
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) component_eval(c_model, Tb, comp, in + i, out + i, n);
}

/* ------------------------------------------------------------------------------------------------ */
/* batch tape: sample index -> row block (later entries of a duplicated index win; indices outside the batch are ignored) */
__global__ void emc_tape_map_kernel(const int64_t *list, int64_t n_sel, int64_t n, int32_t *slot)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_sel && list[k] >= 0 && list[k] < n) slot[list[k]] = (int32_t)k;
}

/* ------------------------------------------------------------------------------------------------ */
__global__ void emc_math_kernel(int op, int64_t n, const double *x, const double *y, double *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r;
    if (op == 0) r = fast_rcp(x[i]);
    else if (op == 1) r = fast_rsqrt(x[i]);
    else if (op == 2) r = fast_atan2(y[i], x[i]);
    else if (op == 3) r = fast_sqrt(x[i]);
    else if (op == 4) r = fast_exp(x[i]);
    else r = fast_log(x[i]);
    out[i] = r;
}

/* ------------------------------------------------------------------------------------------------ */
/* 8 independent DFMA chains per thread, all in registers: 2*8*iters flop per thread */
__global__ void __launch_bounds__(256) emc_dfma_kernel(double *sink, int iters, double a, double b)
{
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1e-3, x2 = x0 + 2e-3, x3 = x0 + 3e-3;
    double x4 = x0 + 4e-3, x5 = x0 + 5e-3, x6 = x0 + 6e-3, x7 = x0 + 7e-3;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    const double r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (r == 123456.789) sink[0] = r;     /* never true; keeps the chains alive */
}

/* one dependent DFMA chain in one warp: cycles per dependent FP64 FMA (clock64 around the chain) */
__global__ void emc_dfma_latency_kernel(double *sink, long long *cycles, int iters, double a, double b)
{
    double x = threadIdx.x * 1e-9;
    const long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < iters; ++i) x = fma(x, a, b);
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    if (x == 123456.789) sink[0] = x;
}

/* ================================================================================================
 *  C ABI
 * ============================================================================================== */
struct emc_ctx {
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr;      /* stream2: the strict consumer, concurrent with the flight kernel */
    int32_t epoch = 0;                                     /* run counter: publication mark of the park records */
    int64_t strict_left_early = 0;                         /* concurrent consumers that gave up waiting in the last run (their records went to the sweep) */
    cudaEvent_t ev[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
    bool has_model = false;
    emc_model model;              /* raw copy (wind_altitudes pointer is NOT valid after set_model) */
    DevModel dmodel;              /* host copies of what c_model / c_tables must hold while this context launches */
    DevTables dtables;
    double *d_wind_alt = nullptr;
    unsigned long long *d_ctrl = nullptr;   /* [0] queue head, [1..4] counters, [5] tape_n */
    /* staging buffers for the host-buffer entry points (grown on demand) */
    double *d_scalars = nullptr; size_t cap_scalars = 0;
    double *d_wind = nullptr; size_t cap_wind = 0;
    double *d_out = nullptr; size_t cap_out = 0;
    int32_t *d_iout = nullptr; size_t cap_iout = 0;
    double *d_tape = nullptr; size_t cap_tape = 0;
    unsigned char *d_scratch = nullptr; size_t cap_scratch = 0;
    double *d_partial = nullptr; size_t cap_partial = 0;
    double *d_summary = nullptr; size_t cap_summary = 0;
    double *d_disp = nullptr; size_t cap_disp = 0;    /* dispersion tables (shear/base wind/rho/innov) */
    double *d_draws = nullptr; size_t cap_draws = 0;  /* caller-supplied draws */
    int64_t staged_n = 0; int staged_knots = 0;
    int64_t last_n = 0;                       /* samples held by d_out/d_iout after the last host-buffer run */
    /* one-sample paths (tape, series, derivative/debug seams) have their own small output block, so they never touch the
     * resident outputs of the last batch */
    double *d_out1 = nullptr; int32_t *d_iout1 = nullptr;
    /* downsampled batch tape (emc_tape_request) */
    int32_t *d_bt_slot = nullptr; size_t cap_bt_slot = 0;
    double *d_bt_rows = nullptr; size_t cap_bt_rows = 0;
    int32_t *d_bt_count = nullptr; size_t cap_bt_count = 0;
    int64_t *d_bt_list = nullptr; size_t cap_bt_list = 0;
    unsigned char *d_gcold = nullptr; size_t cap_gcold = 0;     /* GlobalCold arrays of the 16-warp kernel */
    ParkRec *d_resume = nullptr; size_t cap_resume = 0;         /* flights that handed their lane back (emc_counters.yielded) */
#ifdef EMC_YIELD_DEBUG
    unsigned int *dbg = nullptr; int64_t dbg_n = 0;
#endif
    ParkRec *d_park = nullptr; size_t cap_park = 0;             /* trajectories parked for the strict continuation */
    int64_t bt_n_sel = 0; int32_t bt_stride = 0, bt_max = 0;
    bool bt_armed = false;                    /* a request waits for the next run */
    int64_t bt_have = 0;                      /* n_sel of the tape the last armed run left in d_bt_rows */
    emc_counters counters;
    std::string err;
};

static thread_local std::string g_create_err;

/* c_model / c_tables are per-device globals shared by every context on that device: remember which context uploaded
 * last and re-upload (after draining the device) when another one is about to launch. */
static std::atomic<int32_t> g_epoch{0};      /* run counter of the process: publication mark of the park records (never 0) */
static std::mutex g_owner_mu;
static emc_ctx *g_owner[64] = { nullptr };

static int fail(emc_ctx *c, int code, const std::string &msg)
{
    if (c) c->err = msg; else g_create_err = msg;
    return code;
}

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(ctx, EMC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
    } while (0)

EMC_EXPORT int emc_abi_version(void) { return EMC_ABI_VERSION; }

EMC_EXPORT const char *emc_last_error(const emc_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

EMC_EXPORT int emc_create(emc_ctx **out, int device)
{
    emc_ctx *ctx = nullptr;
    if (!out) return fail(nullptr, EMC_ERR_INVALID, "emc_create: ctx is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(nullptr, EMC_ERR_NO_DEVICE,
                    std::string("emc_create: no CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0") +
                        "); this engine has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(nullptr, EMC_ERR_INVALID, "emc_create: device index out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, EMC_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return fail(nullptr, EMC_ERR_NO_DEVICE,
                    std::string("emc_create: device '") + prop.name + "' is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                        "; libemc.so carries sm_100a code only");
    ctx = new (std::nothrow) emc_ctx();
    if (!ctx) return fail(nullptr, EMC_ERR_INVALID, "out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    memset(&ctx->counters, 0, sizeof ctx->counters);
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking);
    for (int i = 0; i < 5 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev[i]);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_ctrl, 32 * sizeof(unsigned long long));
    if (e != cudaSuccess) {
        std::string m = std::string("emc_create: ") + cudaGetErrorString(e);
        delete ctx;
        return fail(nullptr, EMC_ERR_CUDA, m);
    }
    *out = ctx;
    return EMC_OK;
}

EMC_EXPORT int emc_destroy(emc_ctx *ctx)
{
    if (!ctx) return EMC_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    {
        std::lock_guard<std::mutex> lk(g_owner_mu);
        if (ctx->device >= 0 && ctx->device < 64 && g_owner[ctx->device] == ctx) g_owner[ctx->device] = nullptr;
    }
    cudaFree(ctx->d_out1); cudaFree(ctx->d_iout1); cudaFree(ctx->d_bt_slot); cudaFree(ctx->d_bt_rows); cudaFree(ctx->d_bt_count); cudaFree(ctx->d_bt_list); cudaFree(ctx->d_gcold); cudaFree(ctx->d_park); cudaFree(ctx->d_resume);
    cudaFree(ctx->d_wind_alt); cudaFree(ctx->d_ctrl); cudaFree(ctx->d_scalars); cudaFree(ctx->d_wind);
    cudaFree(ctx->d_out); cudaFree(ctx->d_iout); cudaFree(ctx->d_tape); cudaFree(ctx->d_scratch); cudaFree(ctx->d_partial); cudaFree(ctx->d_summary); cudaFree(ctx->d_disp); cudaFree(ctx->d_draws);
    for (int i = 0; i < 5; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return EMC_OK;
}

/* Make this context's run constants the ones in the device's __constant__ bank.  Called before every launch that reads
 * c_model / c_tables; a no-op while the same context keeps launching. */
static int make_resident(emc_ctx *ctx)
{
    if (ctx->device < 0 || ctx->device >= 64) return fail(ctx, EMC_ERR_INVALID, "device index beyond the ownership table");
    std::lock_guard<std::mutex> lk(g_owner_mu);
    if (g_owner[ctx->device] == ctx) return EMC_OK;
    CK(cudaDeviceSynchronize());               /* kernels of the previous owner may still be reading its constants */
    CK(cudaMemcpyToSymbol(c_model, &ctx->dmodel, sizeof(DevModel)));
    CK(cudaMemcpyToSymbol(c_tables, &ctx->dtables, sizeof(DevTables)));
    g_owner[ctx->device] = ctx;
    return EMC_OK;
}

EMC_EXPORT int emc_set_model(emc_ctx *ctx, const emc_model *model)
{
    if (!ctx || !model) return fail(ctx, EMC_ERR_INVALID, "emc_set_model: NULL argument");
    if (const char *why = validate_model(*model)) return fail(ctx, EMC_ERR_INVALID, std::string("emc_set_model: ") + why);
    CK(cudaSetDevice(ctx->device));
    DevModel D; DevTables T;
    build_dev_model(*model, D, T);
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->dmodel = D; ctx->dtables = T;
    {
        std::lock_guard<std::mutex> lk(g_owner_mu);
        if (g_owner[ctx->device] == ctx) g_owner[ctx->device] = nullptr;     /* force the upload below */
    }
    ctx->has_model = false;
    if (int rc = make_resident(ctx)) return rc;
    cudaFree(ctx->d_wind_alt); ctx->d_wind_alt = nullptr;
    if (D.has_wind) {
        CK(cudaMalloc(&ctx->d_wind_alt, sizeof(double) * D.n_wind));
        CK(cudaMemcpy(ctx->d_wind_alt, model->wind_altitudes, sizeof(double) * D.n_wind, cudaMemcpyHostToDevice));
    }
    ctx->model = *model;
    ctx->model.wind_altitudes = nullptr;
    ctx->has_model = true;
    return EMC_OK;
}

template <typename T>
static cudaError_t grow(T **p, size_t *cap, size_t need)
{
    if (need <= *cap) return cudaSuccess;
    cudaFree(*p); *p = nullptr; *cap = 0;
    cudaError_t e = cudaMalloc(p, need * sizeof(T));
    if (e == cudaSuccess) *cap = need;
    return e;
}

static int check_run_args(emc_ctx *ctx, const emc_inputs *in, int64_t n, const emc_outputs *out)
{
    if (!ctx || !in || !out) return fail(ctx, EMC_ERR_INVALID, "NULL argument");
    if (!ctx->has_model) return fail(ctx, EMC_ERR_NO_MODEL, "emc_set_model has not been called");
    if (n < 0 || n > 0x7fffffffLL) return fail(ctx, EMC_ERR_INVALID, "n must be in [0, 2^31)");
    if (n > 0 && (!in->scalars || !out->out || !out->iout)) return fail(ctx, EMC_ERR_INVALID, "NULL buffer");
    if (in->ld < n || out->ld < n) return fail(ctx, EMC_ERR_INVALID, "leading dimension smaller than n");
    if (ctx->dmodel.has_wind && n > 0 && !in->wind) return fail(ctx, EMC_ERR_INVALID, "model has a wind grid but inputs.wind is NULL");
    if (ctx->dmodel.has_wind && in->wind_sample_stride != 0 && in->wind_sample_stride < (int64_t)ctx->dmodel.n_wind * 3)
        return fail(ctx, EMC_ERR_INVALID, "wind_sample_stride smaller than n_wind*3");
    return EMC_OK;
}

static cudaError_t launch_kernel(emc_ctx *ctx, void (*kern)(KernelArgs), int block, KernelArgs &a, size_t smem, int blocks_per_sm_req)
{
    int occ = 0;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, block, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    if (blocks_per_sm_req > 0 && blocks_per_sm_req < occ) occ = blocks_per_sm_req;
    int64_t grid = (int64_t)ctx->sm_count * occ;
    const int64_t need = (a.n + block - 1) / block;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    a.flight_warps = (int32_t)(grid * (block / 32));      /* diagnostic only: the consumer compares flight_started with flight_done */
    kern<<<(unsigned)grid, block, smem, ctx->stream>>>(a);
    return cudaGetLastError();
}

template <int BLOCK, int MINB, int COLD>
static cudaError_t launch_flight(emc_ctx *ctx, KernelArgs &a, size_t smem, int blocks_per_sm_req)
{
    return launch_kernel(ctx, emc_flight_kernel<BLOCK, MINB, COLD>, BLOCK, a, smem, blocks_per_sm_req);
}

/* the production instances: motor kind and wind-table presence compiled in */
template <int BLOCK, int MINB, int COLD>
static cudaError_t launch_flight_cfg(emc_ctx *ctx, KernelArgs &a, size_t smem, int blocks_per_sm_req)
{
    const bool solid = ctx->dmodel.motor_kind == EMC_MOTOR_SOLID, wind = ctx->dmodel.has_wind != 0;
    if (solid) return wind ? launch_kernel(ctx, emc_flight_kernel<BLOCK, MINB, COLD, 1, 1>, BLOCK, a, smem, blocks_per_sm_req)
                           : launch_kernel(ctx, emc_flight_kernel<BLOCK, MINB, COLD, 1, 0>, BLOCK, a, smem, blocks_per_sm_req);
    return wind ? launch_kernel(ctx, emc_flight_kernel<BLOCK, MINB, COLD, 0, 1>, BLOCK, a, smem, blocks_per_sm_req)
                : launch_kernel(ctx, emc_flight_kernel<BLOCK, MINB, COLD, 0, 0>, BLOCK, a, smem, blocks_per_sm_req);
}

/* all pointers in `a` are device pointers */
static int run_device(emc_ctx *ctx, KernelArgs a, const emc_run_opts *opts)
{
    emc_run_opts o = { 0, 0, 0, 1, 0, 0 };
    if (opts) o = *opts;
    a.refill_threshold = o.refill_threshold > 0 ? o.refill_threshold : 1;
    a.nan_ff = o.nan_fast_forward;
    a.sm_count = ctx->sm_count > 0 ? ctx->sm_count : 1;
    a.compact = (o.flags & EMC_RUN_COMPACTION) ? 1 : 0;
    a.park_on = 1;       /* EMC_RUN_NO_STRICT_TAIL is ignored since the fast path relies on the strict continuation for irregular samples */
    if (a.park_on && a.n > 0) {
        const size_t cap_before = ctx->cap_park;
        CK(grow(&ctx->d_park, &ctx->cap_park, (size_t)a.n));
        /* fresh device memory may be a recycled park buffer of ANOTHER context of this process: clear the publication marks
         * (the epochs themselves come from one process-wide counter, so a stale mark can never equal a later run's) */
        if (ctx->cap_park != cap_before) CK(cudaMemsetAsync(ctx->d_park, 0, sizeof(ParkRec) * ctx->cap_park, ctx->stream));
        a.park = ctx->d_park; a.park_count = ctx->d_ctrl + 11;
    }
    a.yield_step = -1; a.yield_half = -1; a.resume = nullptr;
    a.wind_alt = ctx->d_wind_alt;
    a.queue = ctx->d_ctrl; a.counters = ctx->d_ctrl + 1;
    if (a.tape) a.tape_n = reinterpret_cast<int64_t *>(ctx->d_ctrl + 5);
    if (!ctx->dmodel.has_wind) { a.wind = nullptr; a.wind_stride = 0; }
    const size_t smem = smem_bytes(ctx->dmodel.n_wind);
    if (int rc = make_resident(ctx)) return rc;
    CK(cudaMemsetAsync(ctx->d_ctrl, 0, 32 * sizeof(unsigned long long), ctx->stream));
    memset(&ctx->counters, 0, sizeof ctx->counters);
    if (a.n == 0) return EMC_OK;
    if (ctx->bt_armed && !a.tape) {
        /* consume the tape request: slot map of this batch (-1 everywhere, then the listed samples), cleared counts */
        ctx->bt_armed = false; ctx->bt_have = 0;
        CK(grow(&ctx->d_bt_slot, &ctx->cap_bt_slot, (size_t)a.n));
        CK(cudaMemsetAsync(ctx->d_bt_slot, 0xff, sizeof(int32_t) * (size_t)a.n, ctx->stream));
        CK(cudaMemsetAsync(ctx->d_bt_count, 0, sizeof(int32_t) * (size_t)ctx->bt_n_sel, ctx->stream));
        emc_tape_map_kernel<<<(unsigned)((ctx->bt_n_sel + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_bt_list, ctx->bt_n_sel, a.n, ctx->d_bt_slot);
        CK(cudaGetLastError());
        a.bt_slot = ctx->d_bt_slot; a.bt_rows = ctx->d_bt_rows; a.bt_count = ctx->d_bt_count;
        a.bt_stride = ctx->bt_stride; a.bt_max = ctx->bt_max;
        ctx->bt_have = ctx->bt_n_sel;
    }
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    {
        int64_t grid = (a.n + 127) / 128;
        const int64_t cap = (int64_t)ctx->sm_count * 16;
        if (grid > cap) grid = cap;
        emc_rail_kernel<<<(unsigned)grid, 128, smem, ctx->stream>>>(a);
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    if (a.park_on) {
        /* the strict consumer starts on the second stream as soon as the inputs and the rail outputs are in place */
        a.epoch = ctx->epoch = ++g_epoch;
        a.park_next = ctx->d_ctrl + 12; a.flight_done = ctx->d_ctrl + 13; a.strict_err = ctx->d_ctrl + 14; a.flight_started = ctx->d_ctrl + 18;      /* d_ctrl[15]: consumers that left early */
        CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev[1], 0));
    }
    /* default launch: 128 threads, 3 blocks/SM (159 registers, 12 warps/SM), cold lane state, base state and RK4
     * accumulator in shared memory */
    const int bt = o.block_threads > 0 ? o.block_threads : 128;
    const int bps = (o.block_threads > 0 || o.blocks_per_sm > 0) ? o.blocks_per_sm : 3;
    cudaError_t e;
    const bool cold = o.cold_state_in_smem >= 0;
    const bool store = o.cold_state_in_smem == 0 || o.cold_state_in_smem >= 2;   /* default: base state + RK4 accumulator in shared memory as well */
#ifdef EMC_SLIM
    const bool default_instance = true;
#else
    const bool default_instance = bt == 128 && bps == 3 && store;
#endif
    /* lane hand-back: the default kernel instance, and only where a launch has more samples than resident lanes (otherwise
     * every sample starts at once) and not so many that its last wave is a small part of it */
    {
        const int64_t resident = (int64_t)ctx->sm_count * 384;
        const char *ys = getenv("EMC_YIELD_STEP");
        const int step = ys ? atoi(ys) : EMC_YIELD_STEP;
        if (default_instance && !(o.flags & EMC_RUN_NO_YIELD) && !a.compact && !a.tape && step > 0 && a.n > resident && a.n <= 8 * resident) {
            /* one list per warp (at most 16 warps per SM at >= 128 registers): twice a warp's share of the batch, rounded up */
            const int64_t warps = (int64_t)ctx->sm_count * 16;
            const int32_t cap = (int32_t)(2 * ((a.n + (int64_t)ctx->sm_count * 12 - 1) / ((int64_t)ctx->sm_count * 12)) + 64);
            CK(grow(&ctx->d_resume, &ctx->cap_resume, (size_t)(warps * cap)));
            a.resume_cap = cap;
            a.resume = ctx->d_resume;
            a.yield_step = step; a.yield_half = step >> 1;
        }
#ifdef EMC_YIELD_DEBUG
        a.dbg = nullptr;
        if (getenv("EMC_YIELD_DEBUG") && a.n >= 50000 && a.n <= 200000) {
            static unsigned int *d_dbg = nullptr; static size_t cap = 0;
            CK(grow(&d_dbg, &cap, (size_t)a.n * 4));
            CK(cudaMemsetAsync(d_dbg, 0, sizeof(unsigned int) * 4 * (size_t)a.n, ctx->stream));
            a.dbg = d_dbg; ctx->dbg = d_dbg; ctx->dbg_n = a.n;
        } else ctx->dbg = nullptr;
#endif
    }
#ifdef EMC_SLIM     /* developer builds (tools/build_variant.sh): only the default instances, seconds instead of a minute */
    (void)bt; (void)bps; (void)cold; (void)store;
#ifndef EMC_SLIM_BLOCK
#define EMC_SLIM_BLOCK 128
#define EMC_SLIM_MINB 3
#endif
    e = launch_flight_cfg<EMC_SLIM_BLOCK, EMC_SLIM_MINB, 2>(ctx, a, smem + (28 * sizeof(double) + (sizeof(Padded<ColdLaneFull>) * EMC_SLIM_BLOCK <= 40960 ? 0 : sizeof(Padded<ColdLaneFull>))) * EMC_SLIM_BLOCK, EMC_SLIM_MINB);
#else
    if (bt == 256 && bps == 2) {
        /* 16 warps per SM: 2 blocks x 256 lanes at 128 registers; the once-per-step bookkeeping lives in global memory */
        const int64_t lanes = (int64_t)ctx->sm_count * 2048;
        CK(grow(&ctx->d_gcold, &ctx->cap_gcold, (size_t)lanes * (TC_DCOUNT * sizeof(double) + TI_ICOUNT * sizeof(int32_t))));
        a.gcold_d = reinterpret_cast<double *>(ctx->d_gcold);
        a.gcold_i = reinterpret_cast<int32_t *>(ctx->d_gcold + (size_t)lanes * TC_DCOUNT * sizeof(double));
        a.gcold_ld = lanes;
        e = launch_flight_cfg<256, 2, 3>(ctx, a, smem + (28 * sizeof(double) + sizeof(Padded<ColdLaneHot>)) * 256, bps);
    }
    else if (bt == 128 && bps == 3 && store && a.compact) e = launch_flight_cfg<128, 3, 4>(ctx, a, smem + 28 * 128 * sizeof(double), bps);
    else if (bt == 128 && bps == 3 && store) e = launch_flight_cfg<128, 3, 2>(ctx, a, smem + 28 * 128 * sizeof(double), bps);
    else if (bt == 128 && bps == 4 && store) e = launch_flight_cfg<128, 4, 2>(ctx, a, smem + 28 * 128 * sizeof(double), bps);
    else if (bt == 128 && bps == 3) e = cold ? launch_flight<128, 3, 1>(ctx, a, smem, bps) : launch_flight<128, 3, 0>(ctx, a, smem, bps);
    else if (bt == 128) e = launch_flight<128, 1, 0>(ctx, a, smem, bps);
    else if (bt == 64) e = launch_flight<64, 1, 0>(ctx, a, smem, bps);
    else return fail(ctx, EMC_ERR_INVALID, "block_threads must be 64, 128 or 256 (with 2 blocks per SM)");
#endif
    if (e != cudaSuccess) return fail(ctx, EMC_ERR_CUDA, std::string("flight kernel launch: ") + cudaGetErrorString(e));
    CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    ctx->counters.kernel_launches = 2;
    if (a.park_on) {
        /* one 64-lane consumer block per SM (registers: 3 x 128 x 136 + 64 x 168 <= 65 536; no shared memory): tickets,
         * not indices, so the grid is independent of the batch */
        int64_t grid = (a.n + EMC_STRICT_BLOCK - 1) / EMC_STRICT_BLOCK;
        if (grid > ctx->sm_count) grid = ctx->sm_count;
        /* same shared-memory carve-out as the flight kernel: an SM cannot change its carve-out while blocks are resident, so
         * a consumer block that arrived first with a small one would keep the flight blocks of that SM out until it leaves */
        static bool carve_set = false;
        if (!carve_set) {
            CK(cudaFuncSetAttribute(emc_strict_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            CK(cudaFuncSetAttribute(emc_strict_tail_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            carve_set = true;
        }
        emc_strict_kernel<<<(unsigned)grid, EMC_STRICT_BLOCK, 0, ctx->stream2>>>(a);
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev[3], ctx->stream2));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev[3], 0));
        /* the sweep: what the concurrent consumers left behind (normally nothing: one look at each record's mark) */
        emc_strict_tail_kernel<<<(unsigned)(2 * ctx->sm_count), EMC_STRICT_BLOCK, 0, ctx->stream>>>(a);
        CK(cudaGetLastError());
        ctx->counters.kernel_launches = 4;
    }
    CK(cudaEventRecord(ctx->ev[4], ctx->stream));
    return EMC_OK;
}

static int finish_counters(emc_ctx *ctx)
{
    unsigned long long h[32];
    CK(cudaMemcpyAsync(h, ctx->d_ctrl, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->counters.rk4_steps = (int64_t)h[1];
    ctx->counters.replay_steps = (int64_t)h[2];
    ctx->counters.rail_steps = (int64_t)h[3];
    ctx->counters.refills = (int64_t)h[4];
    ctx->counters.tape_rows = (int64_t)h[8];
    ctx->counters.handovers = (int64_t)h[9];
    ctx->counters.strict_steps = (int64_t)h[10];
    ctx->counters.parked = (int64_t)h[11];
    ctx->counters.yielded = (int64_t)h[21];
#ifdef EMC_YIELD_DEBUG
    if (ctx->dbg) {
        std::vector<unsigned int> v((size_t)ctx->dbg_n * 4);
        CK(cudaMemcpy(v.data(), ctx->dbg, v.size() * 4, cudaMemcpyDeviceToHost));
        unsigned int t0 = 0xffffffffu;
        for (int64_t i = 0; i < ctx->dbg_n; ++i) if (v[i * 4] && v[i * 4] < t0) t0 = v[i * 4];
        const char *nm[4] = { "start", "handback", "resume", "end" };
        for (int e = 0; e < 4; ++e) {
            std::vector<double> x;
            for (int64_t i = 0; i < ctx->dbg_n; ++i) if (v[i * 4 + e]) x.push_back((v[i * 4 + e] - t0) * 1.024e-3);
            std::sort(x.begin(), x.end());
            if (x.empty()) continue;
            fprintf(stderr, "[emc-dbg] %-8s n=%zu  min %.2f p10 %.2f p50 %.2f p90 %.2f p99 %.2f max %.2f ms\n", nm[e], x.size(), x[0], x[x.size() / 10], x[x.size() / 2],
                    x[x.size() * 9 / 10], x[x.size() * 99 / 100], x.back());
        }
        /* the five flights that end last: their sample, start, hand-back, resume, end */
        std::vector<std::pair<unsigned int, int64_t>> ends;
        for (int64_t i = 0; i < ctx->dbg_n; ++i) ends.push_back({ v[i * 4 + 3], i });
        std::sort(ends.begin(), ends.end());
        for (size_t k = ends.size() >= 5 ? ends.size() - 5 : 0; k < ends.size(); ++k) {
            const int64_t i = ends[k].second;
            fprintf(stderr, "[emc-dbg] late sample %lld: start %.2f handback %.2f resume %.2f end %.2f ms\n", (long long)i, (v[i * 4] - t0) * 1.024e-3,
                    v[i * 4 + 1] ? (v[i * 4 + 1] - t0) * 1.024e-3 : -1.0, v[i * 4 + 2] ? (v[i * 4 + 2] - t0) * 1.024e-3 : -1.0, (v[i * 4 + 3] - t0) * 1.024e-3);
        }
    }
#endif
    if (h[14]) return fail(ctx, EMC_ERR_CUDA, "strict continuation: " + std::to_string(h[14]) + " park records do not belong to this run (epoch " +
                           std::to_string(ctx->epoch) + ", parked " + std::to_string(h[11]) + ", tickets " + std::to_string(h[12]) + ")");
    ctx->strict_left_early = (int64_t)h[15];
    if (h[15] && getenv("EMC_DEBUG")) fprintf(stderr, "[emc] %llu strict consumers left early (their records were swept after the flight kernel)\n", h[15]);
    float ms = 0.f;
    if (ctx->counters.kernel_launches) {
        CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1])); ctx->counters.rail_ms = ms;
        CK(cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2])); ctx->counters.flight_ms = ms;
        CK(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[4])); ctx->counters.strict_ms = ms;
    }
    return EMC_OK;
}

EMC_EXPORT int emc_run_batch_device(emc_ctx *ctx, const emc_inputs *in, int64_t n, const emc_outputs *out,
                                    const emc_run_opts *opts)
{
    if (int rc = check_run_args(ctx, in, n, out)) return rc;
    CK(cudaSetDevice(ctx->device));
    KernelArgs a;
    memset(&a, 0, sizeof a);
    a.scalars = in->scalars; a.ld = in->ld; a.wind = in->wind; a.wind_stride = in->wind_sample_stride;
    a.out = out->out; a.iout = out->iout; a.old = out->ld; a.n = n;
    if (int rc = run_device(ctx, a, opts)) return rc;
    return finish_counters(ctx);
}

/* batch = true: the run owns the context's resident output block; false (tape, series, debug seams): outputs, if any, go
 * to the one-sample block so that the resident outputs of the last batch stay valid.  Either way the context's input
 * staging area is overwritten, which invalidates inputs staged by emc_generate_inputs. */
static int upload_inputs(emc_ctx *ctx, const emc_inputs *in, int64_t n, KernelArgs &a, bool batch = true)
{
    const size_t ns = (size_t)EMC_IN_COUNT * (size_t)n;
    ctx->staged_n = 0;
    CK(grow(&ctx->d_scalars, &ctx->cap_scalars, ns));
    /* compact the leading dimension to n on the way up (one contiguous copy when it already is n) */
    if (in->ld == n) CK(cudaMemcpyAsync(ctx->d_scalars, in->scalars, sizeof(double) * ns, cudaMemcpyHostToDevice, ctx->stream));
    else CK(cudaMemcpy2DAsync(ctx->d_scalars, sizeof(double) * n, in->scalars, sizeof(double) * in->ld, sizeof(double) * n,
                              EMC_IN_COUNT, cudaMemcpyHostToDevice, ctx->stream));
    a.scalars = ctx->d_scalars; a.ld = n;
    if (ctx->dmodel.has_wind) {
        const size_t row = (size_t)ctx->dmodel.n_wind * 3;
        if (in->wind_sample_stride == 0) {
            CK(grow(&ctx->d_wind, &ctx->cap_wind, row));
            CK(cudaMemcpyAsync(ctx->d_wind, in->wind, sizeof(double) * row, cudaMemcpyHostToDevice, ctx->stream));
            a.wind_stride = 0;
        } else {
            CK(grow(&ctx->d_wind, &ctx->cap_wind, row * (size_t)n));
            if (in->wind_sample_stride == (int64_t)row)
                CK(cudaMemcpyAsync(ctx->d_wind, in->wind, sizeof(double) * row * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
            else
                CK(cudaMemcpy2DAsync(ctx->d_wind, sizeof(double) * row, in->wind, sizeof(double) * in->wind_sample_stride,
                                     sizeof(double) * row, (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
            a.wind_stride = (int64_t)row;
        }
        a.wind = ctx->d_wind;
    }
    a.n = n;
    if (!batch) {
        if (n == 1) {
            if (!ctx->d_out1) CK(cudaMalloc(&ctx->d_out1, sizeof(double) * EMC_OUT_COUNT));
            if (!ctx->d_iout1) CK(cudaMalloc(&ctx->d_iout1, sizeof(int32_t) * EMC_IOUT_COUNT));
            a.out = ctx->d_out1; a.iout = ctx->d_iout1; a.old = 1;
        }
        return EMC_OK;
    }
    if ((size_t)EMC_OUT_COUNT * (size_t)n > ctx->cap_out || (size_t)EMC_IOUT_COUNT * (size_t)n > ctx->cap_iout) ctx->last_n = 0;
    CK(grow(&ctx->d_out, &ctx->cap_out, (size_t)EMC_OUT_COUNT * (size_t)n));
    CK(grow(&ctx->d_iout, &ctx->cap_iout, (size_t)EMC_IOUT_COUNT * (size_t)n));
    a.out = ctx->d_out; a.iout = ctx->d_iout; a.old = n;
    ctx->last_n = 0;                          /* set again by the caller once the run has completed */
    return EMC_OK;
}

static int download_outputs(emc_ctx *ctx, const emc_outputs *out, int64_t n, const double *d_out, const int32_t *d_iout)
{
    if (out->ld == n) {
        CK(cudaMemcpyAsync(out->out, d_out, sizeof(double) * (size_t)EMC_OUT_COUNT * n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(out->iout, d_iout, sizeof(int32_t) * (size_t)EMC_IOUT_COUNT * n, cudaMemcpyDeviceToHost, ctx->stream));
        return EMC_OK;
    }
    CK(cudaMemcpy2DAsync(out->out, sizeof(double) * out->ld, d_out, sizeof(double) * n, sizeof(double) * n,
                         EMC_OUT_COUNT, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpy2DAsync(out->iout, sizeof(int32_t) * out->ld, d_iout, sizeof(int32_t) * n, sizeof(int32_t) * n,
                         EMC_IOUT_COUNT, cudaMemcpyDeviceToHost, ctx->stream));
    return EMC_OK;
}

EMC_EXPORT int emc_run_batch(emc_ctx *ctx, const emc_inputs *in, int64_t n, const emc_outputs *out,
                             const emc_run_opts *opts)
{
    if (int rc = check_run_args(ctx, in, n, out)) return rc;
    if (n == 0) { memset(&ctx->counters, 0, sizeof ctx->counters); return EMC_OK; }
    CK(cudaSetDevice(ctx->device));
    KernelArgs a;
    memset(&a, 0, sizeof a);
    if (int rc = upload_inputs(ctx, in, n, a)) return rc;
    if (int rc = run_device(ctx, a, opts)) return rc;
    if (int rc = download_outputs(ctx, out, n, ctx->d_out, ctx->d_iout)) return rc;
    if (int rc = finish_counters(ctx)) return rc;
    ctx->last_n = n;
    return EMC_OK;
}

EMC_EXPORT int emc_run_tape(emc_ctx *ctx, const emc_inputs *in, const emc_outputs *out, double *tape, int64_t cap,
                            int64_t *n_states)
{
    if (int rc = check_run_args(ctx, in, 1, out)) return rc;
    if (!tape || cap < 1 || !n_states) return fail(ctx, EMC_ERR_INVALID, "emc_run_tape: tape/cap/n_states");
    CK(cudaSetDevice(ctx->device));
    KernelArgs a;
    memset(&a, 0, sizeof a);
    if (int rc = upload_inputs(ctx, in, 1, a, false)) return rc;
    CK(grow(&ctx->d_tape, &ctx->cap_tape, (size_t)cap * EMC_TAPE_WIDTH));
    a.tape = ctx->d_tape; a.tape_cap = cap;
    /* every state is integrated: no fast-forward on the tape path.  Default launch shape: the SAME kernel instance that
     * flies the batches, so that a tape is bit for bit the flight the batch flew (two instances of the same source may
     * differ in FMA contraction) */
    emc_run_opts o = { 1, 0, 0, 0, 0, 0 };
    if (int rc = run_device(ctx, a, &o)) return rc;
    if (int rc = download_outputs(ctx, out, 1, ctx->d_out1, ctx->d_iout1)) return rc;
    if (int rc = finish_counters(ctx)) return rc;
    unsigned long long h[16];
    CK(cudaMemcpy(h, ctx->d_ctrl, sizeof h, cudaMemcpyDeviceToHost));
    const int64_t ns = (int64_t)h[5];
    *n_states = ns;
    const int64_t rows = ns < cap ? ns : cap;
    CK(cudaMemcpy(tape, ctx->d_tape, sizeof(double) * (size_t)rows * EMC_TAPE_WIDTH, cudaMemcpyDeviceToHost));
    if (ns > cap) return fail(ctx, EMC_ERR_CAPACITY, "emc_run_tape: tape capacity too small; n_states holds the required rows");
    return EMC_OK;
}

EMC_EXPORT int emc_extract_series(emc_ctx *ctx, const emc_inputs *in, const double *tape, int64_t n_states, double *series)
{
    emc_outputs dummy = { (double *)1, (int32_t *)1, 1 };
    if (int rc = check_run_args(ctx, in, 1, &dummy)) return rc;
    if (!tape || !series || n_states < 1) return fail(ctx, EMC_ERR_INVALID, "emc_extract_series: tape/series/n_states");
    CK(cudaSetDevice(ctx->device));
    KernelArgs a;
    memset(&a, 0, sizeof a);
    if (int rc = upload_inputs(ctx, in, 1, a, false)) return rc;
    if (int rc = make_resident(ctx)) return rc;
    a.wind_alt = ctx->d_wind_alt;
    if (!ctx->dmodel.has_wind) { a.wind = nullptr; a.wind_stride = 0; }
    CK(grow(&ctx->d_tape, &ctx->cap_tape, (size_t)n_states * EMC_TAPE_WIDTH));
    const size_t ns = (size_t)EMC_SERIES_COUNT * (size_t)n_states;
    CK(grow(&ctx->d_scratch, &ctx->cap_scratch, ns * sizeof(double)));
    double *d_series = reinterpret_cast<double *>(ctx->d_scratch);
    CK(cudaMemcpyAsync(ctx->d_tape, tape, sizeof(double) * (size_t)n_states * EMC_TAPE_WIDTH, cudaMemcpyHostToDevice, ctx->stream));
    emc_series_kernel<<<(unsigned)((n_states + 127) / 128), 128, smem_bytes(ctx->dmodel.n_wind), ctx->stream>>>(a, ctx->d_tape, n_states, d_series);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(series, d_series, ns * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return EMC_OK;
}

EMC_EXPORT int emc_derivative_debug(emc_ctx *ctx, const emc_inputs *in, int64_t n, const double *t, const double *state,
                                    int32_t *chute, double *state_dot)
{
    emc_outputs dummy = { (double *)1, (int32_t *)1, n };
    if (int rc = check_run_args(ctx, in, n, &dummy)) return rc;
    if (!t || !state || !chute || !state_dot) return fail(ctx, EMC_ERR_INVALID, "emc_derivative_debug: NULL buffer");
    if (n == 0) return EMC_OK;
    CK(cudaSetDevice(ctx->device));
    KernelArgs a;
    memset(&a, 0, sizeof a);
    if (int rc = upload_inputs(ctx, in, n, a, false)) return rc;
    if (int rc = make_resident(ctx)) return rc;
    a.wind_alt = ctx->d_wind_alt;
    if (!ctx->dmodel.has_wind) { a.wind = nullptr; a.wind_stride = 0; }
    double *d_t = nullptr, *d_s = nullptr, *d_k = nullptr; int32_t *d_c = nullptr;
    cudaError_t e = cudaMalloc(&d_t, sizeof(double) * n);
    if (e == cudaSuccess) e = cudaMalloc(&d_s, sizeof(double) * 14 * n);
    if (e == cudaSuccess) e = cudaMalloc(&d_k, sizeof(double) * 14 * n);
    if (e == cudaSuccess) e = cudaMalloc(&d_c, sizeof(int32_t) * n);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_t, t, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_s, state, sizeof(double) * 14 * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_c, chute, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        emc_derivative_kernel<<<(unsigned)((n + 127) / 128), 128, smem_bytes(ctx->dmodel.n_wind), ctx->stream>>>(a, d_t, d_s, d_c, d_k);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(state_dot, d_k, sizeof(double) * 14 * n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(chute, d_c, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_t); cudaFree(d_s); cudaFree(d_k); cudaFree(d_c);
    if (e != cudaSuccess) return fail(ctx, EMC_ERR_CUDA, std::string("emc_derivative_debug: ") + cudaGetErrorString(e));
    return EMC_OK;
}

/* ---------------------------------------------------------------------------------------------------
 *  device-side dispersions
 * ------------------------------------------------------------------------------------------------- */
static int generate_core(emc_ctx *ctx, const emc_dispersion *d, uint64_t seed, int64_t first_index, int64_t n,
                         const double *gauss, int64_t n_gauss, const double *unif, int draws_on_device,
                         double *scalars_dev, int64_t ld, double *wind_dev);

EMC_EXPORT int emc_generate_inputs(emc_ctx *ctx, const emc_dispersion *d, uint64_t seed, int64_t first_index, int64_t n,
                                   const double *gauss, int64_t n_gauss, const double *unif,
                                   double *scalars_dev, int64_t ld, double *wind_dev)
{
    return generate_core(ctx, d, seed, first_index, n, gauss, n_gauss, unif, 0, scalars_dev, ld, wind_dev);
}

static int numpy_draws_device(emc_ctx *ctx, int64_t first_seed, int64_t n, int64_t G, double **g_dev, double **u_dev)
{
    CK(cudaSetDevice(ctx->device));
    CK(grow(&ctx->d_draws, &ctx->cap_draws, (size_t)n * (size_t)(G + 3)));
    *g_dev = ctx->d_draws; *u_dev = ctx->d_draws + (size_t)n * G;
    emc_numpy_draws_kernel<<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>(first_seed, n, G, *g_dev, *u_dev, *u_dev + 2 * (size_t)n);
    CK(cudaGetLastError());
    return EMC_OK;
}

EMC_EXPORT int emc_numpy_draws(emc_ctx *ctx, int64_t first_seed, int64_t n, int64_t n_gauss, double *gauss, double *unif, double *density)
{
    if (!ctx || !gauss || !unif || n < 0 || n_gauss < 14) return fail(ctx, EMC_ERR_INVALID, "emc_numpy_draws: bad argument (n_gauss >= 14)");
    if (n == 0) return EMC_OK;
    double *g = nullptr, *u = nullptr;
    if (int rc = numpy_draws_device(ctx, first_seed, n, n_gauss, &g, &u)) return rc;
    CK(cudaMemcpyAsync(gauss, g, sizeof(double) * (size_t)n * n_gauss, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(unif, u, sizeof(double) * 2 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (density) CK(cudaMemcpyAsync(density, u + 2 * (size_t)n, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return EMC_OK;
}

EMC_EXPORT int emc_generate_inputs_numpy(emc_ctx *ctx, const emc_dispersion *d, int64_t first_seed, int64_t n,
                                         double *scalars_dev, int64_t ld, double *wind_dev)
{
    if (!ctx || !d || n < 0) return fail(ctx, EMC_ERR_INVALID, "emc_generate_inputs_numpy: bad argument");
    if (n == 0) return generate_core(ctx, d, 0, first_seed, 0, nullptr, 0, nullptr, 0, scalars_dev, ld, wind_dev);
    const int64_t G = (3 * (int64_t)d->n_knots > 15) ? 3 * (int64_t)d->n_knots : 15;
    double *g = nullptr, *u = nullptr;
    if (int rc = numpy_draws_device(ctx, first_seed, n, G, &g, &u)) return rc;
    return generate_core(ctx, d, 0, first_seed, n, g, G, u, 1, scalars_dev, ld, wind_dev);
}

static int generate_core(emc_ctx *ctx, const emc_dispersion *d, uint64_t seed, int64_t first_index, int64_t n,
                         const double *gauss, int64_t n_gauss, const double *unif, int draws_on_device,
                         double *scalars_dev, int64_t ld, double *wind_dev)
{
    if (!ctx || !d || n < 0) return fail(ctx, EMC_ERR_INVALID, "emc_generate_inputs: bad argument");
    if (d->n_knots < 0 || d->n_knots > EMC_MAX_WIND_KNOTS) return fail(ctx, EMC_ERR_INVALID, "emc_generate_inputs: n_knots");
    if (d->n_knots > 0 && (!d->rho || !d->innov || (d->wind_mode == 0 ? !d->shear : !d->base_wind)))
        return fail(ctx, EMC_ERR_INVALID, "emc_generate_inputs: missing wind tables");
    const int64_t need_g = (3 * (int64_t)d->n_knots > 15) ? 3 * (int64_t)d->n_knots : 15;
    if (gauss && n_gauss < need_g) return fail(ctx, EMC_ERR_INVALID, "emc_generate_inputs: n_gauss too small (max(15, 3*n_knots))");
    if (scalars_dev && d->n_knots > 0 && !wind_dev)
        return fail(ctx, EMC_ERR_INVALID, "emc_generate_inputs: wind_dev is NULL but the dispersion has wind knots");
    if (n == 0) { if (!scalars_dev) { ctx->staged_n = 0; } return EMC_OK; }
    CK(cudaSetDevice(ctx->device));
    const int K = d->n_knots;
    CK(grow(&ctx->d_disp, &ctx->cap_disp, (size_t)(6 * (K > 0 ? K : 1))));
    DevDispersion D;
    memset(&D, 0, sizeof D);
    memcpy(D.base_pos, d->base_pos, sizeof(double) * 3 * 8);        /* base_* and sigma_* are contiguous in both structs */
    D.mass_sigma = d->mass_sigma; D.wind_speed_lo = d->wind_speed_lo; D.wind_speed_hi = d->wind_speed_hi;
    D.wind_dir_lo = d->wind_dir_lo; D.wind_dir_hi = d->wind_dir_hi; D.dry_mass = d->dry_mass; D.propellant_mass = d->propellant_mass;
    D.thrust_vacuum = d->thrust_vacuum; D.thrust_sea_level = d->thrust_sea_level; D.mass_flow_rate = d->mass_flow_rate;
    D.nozzle_exit_area = d->nozzle_exit_area; D.motor_propellant_mass = d->motor_propellant_mass; D.motor_burn_time = d->motor_burn_time;
    D.thrust_sigma = d->thrust_sigma; D.flow_sigma = d->flow_sigma; D.burn_sigma = d->burn_sigma;
    D.motor_kind = d->motor_kind; D.wind_mode = d->wind_mode; D.n_knots = K;
    if (K > 0) {
        double *t = ctx->d_disp;
        D.shear = t; D.base_wind = t + K; D.rho = t + 4 * K; D.innov = t + 5 * K;
        if (d->shear) CK(cudaMemcpyAsync(t, d->shear, sizeof(double) * K, cudaMemcpyHostToDevice, ctx->stream));
        if (d->base_wind) CK(cudaMemcpyAsync(t + K, d->base_wind, sizeof(double) * 3 * K, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(t + 4 * K, d->rho, sizeof(double) * K, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(t + 5 * K, d->innov, sizeof(double) * K, cudaMemcpyHostToDevice, ctx->stream));
        D.scale0 = d->innov[0];
    }
    const double *g_dev = nullptr, *u_dev = nullptr;
    if (draws_on_device) { g_dev = gauss; u_dev = unif; }
    else if (gauss || unif) {
        if (!gauss || !unif) return fail(ctx, EMC_ERR_INVALID, "emc_generate_inputs: gauss and unif must be given together");
        CK(grow(&ctx->d_draws, &ctx->cap_draws, (size_t)n * (size_t)(n_gauss + 2)));
        CK(cudaMemcpyAsync(ctx->d_draws, gauss, sizeof(double) * (size_t)n * n_gauss, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_draws + (size_t)n * n_gauss, unif, sizeof(double) * 2 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        g_dev = ctx->d_draws; u_dev = ctx->d_draws + (size_t)n * n_gauss;
    }
    if (!scalars_dev) {
        CK(grow(&ctx->d_scalars, &ctx->cap_scalars, (size_t)EMC_IN_COUNT * (size_t)n));
        if (K > 0) CK(grow(&ctx->d_wind, &ctx->cap_wind, (size_t)n * (size_t)K * 3));
        scalars_dev = ctx->d_scalars; ld = n; wind_dev = K > 0 ? ctx->d_wind : nullptr;
        ctx->staged_n = n; ctx->staged_knots = K;
    }
    if (ld < n) return fail(ctx, EMC_ERR_INVALID, "emc_generate_inputs: ld < n");
    emc_generate_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(D, seed, first_index, n, g_dev, n_gauss, u_dev, scalars_dev, ld, wind_dev);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return EMC_OK;
}

EMC_EXPORT int emc_run_batch_staged(emc_ctx *ctx, int64_t n, const emc_outputs *out, const emc_run_opts *opts)
{
    if (!ctx || !out) return fail(ctx, EMC_ERR_INVALID, "emc_run_batch_staged: NULL argument");
    if (!ctx->has_model) return fail(ctx, EMC_ERR_NO_MODEL, "emc_set_model has not been called");
    if (n < 0 || n > ctx->staged_n) return fail(ctx, EMC_ERR_INVALID, "emc_run_batch_staged: n exceeds the staged samples");
    if (ctx->dmodel.has_wind && ctx->staged_knots != ctx->dmodel.n_wind) return fail(ctx, EMC_ERR_INVALID, "emc_run_batch_staged: staged wind tables do not match the model's altitude grid");
    if (n == 0) return EMC_OK;
    if ((out->out || out->iout) && (!out->out || !out->iout || out->ld < n)) return fail(ctx, EMC_ERR_INVALID, "emc_run_batch_staged: output buffers");
    CK(cudaSetDevice(ctx->device));
    KernelArgs a;
    memset(&a, 0, sizeof a);
    a.scalars = ctx->d_scalars; a.ld = ctx->staged_n;
    a.wind = ctx->dmodel.has_wind ? ctx->d_wind : nullptr; a.wind_stride = (int64_t)ctx->staged_knots * 3;
    ctx->last_n = 0;
    CK(grow(&ctx->d_out, &ctx->cap_out, (size_t)EMC_OUT_COUNT * (size_t)n));
    CK(grow(&ctx->d_iout, &ctx->cap_iout, (size_t)EMC_IOUT_COUNT * (size_t)n));
    a.out = ctx->d_out; a.iout = ctx->d_iout; a.old = n; a.n = n;
    if (int rc = run_device(ctx, a, opts)) return rc;
    if (out->out && out->iout) { if (int rc = download_outputs(ctx, out, n, ctx->d_out, ctx->d_iout)) return rc; }
    if (int rc = finish_counters(ctx)) return rc;
    ctx->last_n = n;
    return EMC_OK;
}

EMC_EXPORT int emc_staged_inputs(emc_ctx *ctx, int64_t n, double *scalars_host, double *wind_host)
{
    if (!ctx || n < 0 || n > ctx->staged_n) return fail(ctx, EMC_ERR_INVALID, "emc_staged_inputs: bad argument");
    CK(cudaSetDevice(ctx->device));
    if (scalars_host && n > 0)
        CK(cudaMemcpy2DAsync(scalars_host, sizeof(double) * n, ctx->d_scalars, sizeof(double) * ctx->staged_n, sizeof(double) * n,
                             EMC_IN_COUNT, cudaMemcpyDeviceToHost, ctx->stream));
    if (wind_host && n > 0 && ctx->staged_knots > 0)
        CK(cudaMemcpyAsync(wind_host, ctx->d_wind, sizeof(double) * (size_t)n * ctx->staged_knots * 3, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return EMC_OK;
}

EMC_EXPORT int emc_philox_draws(emc_ctx *ctx, uint64_t seed, int64_t first_index, int64_t n, int64_t n_gauss, double *gauss, double *unif)
{
    if (!ctx || !gauss || !unif || n < 0 || n_gauss < 1) return fail(ctx, EMC_ERR_INVALID, "emc_philox_draws: bad argument");
    if (n == 0) return EMC_OK;
    CK(cudaSetDevice(ctx->device));
    CK(grow(&ctx->d_draws, &ctx->cap_draws, (size_t)n * (size_t)(n_gauss + 2)));
    double *g = ctx->d_draws, *u = g + (size_t)n * n_gauss;
    emc_philox_draws_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(seed, first_index, n, n_gauss, g, u);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(gauss, g, sizeof(double) * (size_t)n * n_gauss, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(unif, u, sizeof(double) * 2 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return EMC_OK;
}

/* ---------------------------------------------------------------------------------------------------
 *  device statistics
 * ------------------------------------------------------------------------------------------------- */
EMC_EXPORT int emc_scratch(emc_ctx *ctx, int64_t bytes, void **dev_ptr)
{
    if (!ctx || !dev_ptr || bytes < 0) return fail(ctx, EMC_ERR_INVALID, "emc_scratch: bad argument");
    CK(cudaSetDevice(ctx->device));
    CK(grow(&ctx->d_scratch, &ctx->cap_scratch, (size_t)bytes));
    *dev_ptr = ctx->d_scratch;
    return EMC_OK;
}

EMC_EXPORT int emc_copy_to_host(emc_ctx *ctx, void *host, const void *dev, int64_t bytes)
{
    if (!ctx || !host || !dev || bytes < 0) return fail(ctx, EMC_ERR_INVALID, "emc_copy_to_host: bad argument");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(host, dev, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return EMC_OK;
}

EMC_EXPORT int emc_copy_to_device(emc_ctx *ctx, void *dev, const void *host, int64_t bytes)
{
    if (!ctx || !host || !dev || bytes < 0) return fail(ctx, EMC_ERR_INVALID, "emc_copy_to_device: bad argument");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(dev, host, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return EMC_OK;
}

EMC_EXPORT int emc_resident_outputs(emc_ctx *ctx, double **out_dev, int64_t *ld)
{
    if (!ctx || !out_dev || !ld) return fail(ctx, EMC_ERR_INVALID, "emc_resident_outputs: NULL argument");
    if (!ctx->d_out || ctx->last_n <= 0) return fail(ctx, EMC_ERR_INVALID, "emc_resident_outputs: no resident outputs");
    *out_dev = ctx->d_out; *ld = ctx->last_n;
    return EMC_OK;
}

/* ---------------------------------------------------------------------------------------------------
 *  downsampled batch tape (reference monte_carlo.py:296-302 'trajectory')
 * ------------------------------------------------------------------------------------------------- */
EMC_EXPORT int emc_fetch_outputs(emc_ctx *ctx, int64_t n, const emc_outputs *out)
{
    if (!ctx || !out || !out->out || !out->iout || out->ld < n) return fail(ctx, EMC_ERR_INVALID, "emc_fetch_outputs: bad argument");
    if (!ctx->d_out || !ctx->d_iout || n <= 0 || n != ctx->last_n) return fail(ctx, EMC_ERR_INVALID, "emc_fetch_outputs: the context holds no resident outputs of that size");
    CK(cudaSetDevice(ctx->device));
    if (int rc = download_outputs(ctx, out, n, ctx->d_out, ctx->d_iout)) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return EMC_OK;
}

EMC_EXPORT int emc_tape_request(emc_ctx *ctx, const int64_t *samples, int64_t n_sel, int32_t stride, int32_t max_rows)
{
    if (!ctx) return EMC_ERR_INVALID;
    if (n_sel == 0 || !samples) { ctx->bt_armed = false; return EMC_OK; }     /* clears a pending request */
    if (n_sel < 0 || n_sel > 0x7fffffffLL || stride < 1 || max_rows < 2) return fail(ctx, EMC_ERR_INVALID, "emc_tape_request: n_sel / stride >= 1 / max_rows >= 2");
    CK(cudaSetDevice(ctx->device));
    CK(grow(&ctx->d_bt_list, &ctx->cap_bt_list, (size_t)n_sel));
    CK(grow(&ctx->d_bt_count, &ctx->cap_bt_count, (size_t)n_sel));
    CK(grow(&ctx->d_bt_rows, &ctx->cap_bt_rows, (size_t)n_sel * (size_t)max_rows * EMC_BTAPE_WIDTH));
    CK(cudaMemcpyAsync(ctx->d_bt_list, samples, sizeof(int64_t) * (size_t)n_sel, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->bt_n_sel = n_sel; ctx->bt_stride = stride; ctx->bt_max = max_rows;
    ctx->bt_armed = true; ctx->bt_have = 0;
    return EMC_OK;
}

EMC_EXPORT int emc_tape_fetch(emc_ctx *ctx, double *rows, int32_t *n_rows)
{
    if (!ctx || !rows || !n_rows) return fail(ctx, EMC_ERR_INVALID, "emc_tape_fetch: NULL argument");
    if (ctx->bt_have <= 0) return fail(ctx, EMC_ERR_INVALID, "emc_tape_fetch: no run has recorded a tape since the last emc_tape_request");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(rows, ctx->d_bt_rows, sizeof(double) * (size_t)ctx->bt_have * (size_t)ctx->bt_max * EMC_BTAPE_WIDTH, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(n_rows, ctx->d_bt_count, sizeof(int32_t) * (size_t)ctx->bt_have, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return EMC_OK;
}

EMC_EXPORT int emc_tape_resident(emc_ctx *ctx, double **rows_dev, int32_t **n_rows_dev, int64_t *n_sel, int32_t *max_rows)
{
    if (!ctx || !rows_dev || !n_rows_dev || !n_sel || !max_rows) return fail(ctx, EMC_ERR_INVALID, "emc_tape_resident: NULL argument");
    if (ctx->bt_have <= 0) return fail(ctx, EMC_ERR_INVALID, "emc_tape_resident: no recorded tape");
    *rows_dev = ctx->d_bt_rows; *n_rows_dev = ctx->d_bt_count; *n_sel = ctx->bt_have; *max_rows = ctx->bt_max;
    return EMC_OK;
}

EMC_EXPORT int emc_upload_outputs(emc_ctx *ctx, const double *out_host, int64_t ld, int64_t n)
{
    if (!ctx || !out_host || n < 0 || ld < n) return fail(ctx, EMC_ERR_INVALID, "emc_upload_outputs: bad argument");
    CK(cudaSetDevice(ctx->device));
    CK(grow(&ctx->d_out, &ctx->cap_out, (size_t)EMC_OUT_COUNT * (size_t)(n > 0 ? n : 1)));
    if (n > 0)
        CK(cudaMemcpy2DAsync(ctx->d_out, sizeof(double) * n, out_host, sizeof(double) * ld, sizeof(double) * n, EMC_OUT_COUNT,
                             cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->last_n = n;
    return EMC_OK;
}

static int stats_source(emc_ctx *ctx, const double *&out_dev, int64_t &ld, int64_t n)
{
    if (!ctx || n < 0) return fail(ctx, EMC_ERR_INVALID, "emc_stats: bad argument");
    if (n == 0 && !out_dev) {                 /* an empty shard of a multi-GPU job still takes part in the reductions */
        out_dev = reinterpret_cast<const double *>(ctx->d_ctrl); ld = 0;
        return EMC_OK;
    }
    if (!out_dev) {
        if (!ctx->d_out || ctx->last_n < n || ctx->last_n == 0) return fail(ctx, EMC_ERR_INVALID, "emc_stats: no resident outputs of a previous emc_run_batch");
        out_dev = ctx->d_out; ld = ctx->last_n;
    }
    if (ld < n) return fail(ctx, EMC_ERR_INVALID, "emc_stats: ld < n");
    return EMC_OK;
}

static int stats_grid(emc_ctx *ctx, int64_t n)
{
    int64_t g = (n + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 4;
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}

EMC_EXPORT int emc_stats_moments1(emc_ctx *ctx, const double *out_dev, int64_t ld, int64_t n, double *sum_dev,
                                  double *min_dev, double *max_dev)
{
    if (int rc = stats_source(ctx, out_dev, ld, n)) return rc;
    if (!sum_dev || !min_dev || !max_dev) return fail(ctx, EMC_ERR_INVALID, "emc_stats_moments1: NULL result block");
    CK(cudaSetDevice(ctx->device));
    const int g = stats_grid(ctx, n);
    const size_t need = (size_t)g * (ST_SUM_COUNT + 2 * ST_MM_COUNT);
    CK(grow(&ctx->d_partial, &ctx->cap_partial, need));
    double *ps = ctx->d_partial, *pmin = ps + (size_t)g * ST_SUM_COUNT, *pmax = pmin + (size_t)g * ST_MM_COUNT;
    emc_stats_moments1_kernel<<<g, 256, 0, ctx->stream>>>(out_dev, ld, n, ps, pmin, pmax);
    if (min_dev == sum_dev + ST_SUM_COUNT && max_dev == min_dev + ST_MM_COUNT) {     /* contiguous result block: one launch */
        emc_stats_finish3_kernel<<<1, 32, 0, ctx->stream>>>(ps, pmin, pmax, g, sum_dev);
    } else {
        emc_stats_finish_kernel<<<1, 32, 0, ctx->stream>>>(ps, g, ST_SUM_COUNT, 0, sum_dev);
        emc_stats_finish_kernel<<<1, 32, 0, ctx->stream>>>(pmin, g, ST_MM_COUNT, 1, min_dev);
        emc_stats_finish_kernel<<<1, 32, 0, ctx->stream>>>(pmax, g, ST_MM_COUNT, 2, max_dev);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return EMC_OK;
}

EMC_EXPORT int emc_stats_moments2(emc_ctx *ctx, const double *out_dev, int64_t ld, int64_t n, const double *center_dev,
                                  double *sum_dev)
{
    if (int rc = stats_source(ctx, out_dev, ld, n)) return rc;
    if (!center_dev || !sum_dev) return fail(ctx, EMC_ERR_INVALID, "emc_stats_moments2: NULL block");
    CK(cudaSetDevice(ctx->device));
    const int g = stats_grid(ctx, n);
    CK(grow(&ctx->d_partial, &ctx->cap_partial, (size_t)g * ST2_COUNT));
    emc_stats_moments2_kernel<<<g, 256, 0, ctx->stream>>>(out_dev, ld, n, center_dev, ctx->d_partial);
    emc_stats_finish_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_partial, g, ST2_COUNT, 0, sum_dev);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return EMC_OK;
}

/* One stage of the summary chain, enqueued on the context stream without synchronising (stage 14 copies the result and
 * synchronises).  Stages: 0 moments1 -> block sum|min|max (20 words); 1 plan + moments2 -> block s2 (6 words);
 * 2+2k digit histogram of pass k -> block hist (3*2*n_pct*2048 words); 3+2k digit decision of pass k; 14 end.
 * A multi-GPU caller all-reduces `block` on the same stream between stages; a single GPU just runs 0..14. */
static int summary_stage(emc_ctx *ctx, const double *out_dev, int64_t ld, int64_t n, const double *percentiles, int n_pct, int stage,
                         void **block_dev, int64_t *block_words, double *result)
{
    const int nt = 2 * n_pct, rows = 3 * nt;
    const size_t res_words = 32 + (size_t)rows, words = res_words + 2 * (size_t)rows + (size_t)rows * EMC_SELECT_BINS;
    CK(grow(&ctx->d_summary, &ctx->cap_summary, words));
    SummaryLayout L;
    L.res = ctx->d_summary; L.nt = nt;
    L.prefix = reinterpret_cast<unsigned long long *>(ctx->d_summary + res_words);
    L.rem = reinterpret_cast<long long *>(ctx->d_summary + res_words + rows);
    L.hist = reinterpret_cast<unsigned long long *>(ctx->d_summary + res_words + 2 * (size_t)rows);
    const int g = stats_grid(ctx, n);
    CK(grow(&ctx->d_partial, &ctx->cap_partial, (size_t)g * (ST_SUM_COUNT + 2 * ST_MM_COUNT)));
    double *ps = ctx->d_partial, *pmin = ps + (size_t)g * ST_SUM_COUNT, *pmax = pmin + (size_t)g * ST_MM_COUNT;
    cudaStream_t st = ctx->stream;
    if (block_dev) *block_dev = nullptr;
    if (block_words) *block_words = 0;
    static const int passes[6][2] = { { 55, 64 }, { 44, 55 }, { 33, 44 }, { 22, 33 }, { 11, 22 }, { 0, 11 } };
    if (stage == 0) {
        CK(cudaMemsetAsync(ctx->d_summary, 0, sizeof(double) * words, st));
        emc_stats_moments1_kernel<<<g, 256, 0, st>>>(out_dev, ld, n, ps, pmin, pmax);
        emc_stats_finish3_kernel<<<1, 32, 0, st>>>(ps, pmin, pmax, g, L.res);
        if (block_dev) *block_dev = L.res;
        if (block_words) *block_words = ST_SUM_COUNT + 2 * ST_MM_COUNT;
    } else if (stage == 1) {
        SummaryPct P;
        memset(&P, 0, sizeof P);
        P.n_pct = n_pct;
        for (int j = 0; j < n_pct; ++j) P.pct[j] = percentiles[j];
        emc_stats_plan_kernel<<<1, 32, 0, st>>>(L, P);
        emc_stats_moments2_kernel<<<g, 256, 0, st>>>(out_dev, ld, n, L.res + 26, ps);
        emc_stats_finish_kernel<<<1, 32, 0, st>>>(ps, g, ST2_COUNT, 0, L.res + 20);
        if (block_dev) *block_dev = L.res + 20;
        if (block_words) *block_words = ST2_COUNT;
    } else if (stage >= 2 && stage <= 13) {
        const int k = (stage - 2) / 2, shift = passes[k][0], pshift = passes[k][1];
        if ((stage & 1) == 0) {
            emc_stats_select_dev_kernel<<<g, 256, 0, st>>>(out_dev, ld, n, L, shift, pshift);
            if (block_dev) *block_dev = L.hist;
            if (block_words) *block_words = (int64_t)rows * EMC_SELECT_BINS;
        } else {
            emc_stats_select_finish_kernel<<<rows, 32, 0, st>>>(L, (pshift >= 64) ? 64 - shift : pshift - shift, k == 5);
        }
    } else if (stage == 14) {
        if (!result) return fail(ctx, EMC_ERR_INVALID, "emc_stats_summary: NULL result");
        CK(cudaMemcpyAsync(result, L.res, sizeof(double) * res_words, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    } else {
        return fail(ctx, EMC_ERR_INVALID, "emc_stats_summary_stage: stage must be 0..14");
    }
    CK(cudaGetLastError());
    return EMC_OK;
}

EMC_EXPORT int emc_stats_summary_stage(emc_ctx *ctx, const double *out_dev, int64_t ld, int64_t n, const double *percentiles, int n_pct,
                                       int stage, void **block_dev, int64_t *block_words, double *result)
{
    if (int rc = stats_source(ctx, out_dev, ld, n)) return rc;
    if (!percentiles || n_pct < 1 || n_pct > EMC_SUMMARY_MAX_PCT) return fail(ctx, EMC_ERR_INVALID, "emc_stats_summary_stage: bad argument (1 <= n_pct <= 8)");
    CK(cudaSetDevice(ctx->device));
    return summary_stage(ctx, out_dev, ld, n, percentiles, n_pct, stage, block_dev, block_words, result);
}

EMC_EXPORT int emc_stats_summary(emc_ctx *ctx, const double *out_dev, int64_t ld, int64_t n, const double *percentiles, int n_pct,
                                 double *result)
{
    if (int rc = stats_source(ctx, out_dev, ld, n)) return rc;
    if (!result || !percentiles || n_pct < 1 || n_pct > EMC_SUMMARY_MAX_PCT) return fail(ctx, EMC_ERR_INVALID, "emc_stats_summary: bad argument (1 <= n_pct <= 8)");
    CK(cudaSetDevice(ctx->device));
    for (int stage = 0; stage <= 14; ++stage)
        if (int rc = summary_stage(ctx, out_dev, ld, n, percentiles, n_pct, stage, nullptr, nullptr, result)) return rc;
    return EMC_OK;
}

/* the context's CUDA stream (cudaStream_t), for callers that enqueue their own work between stages (NCCL all-reduce) */
EMC_EXPORT int emc_stream(emc_ctx *ctx, void **stream)
{
    if (!ctx || !stream) return fail(ctx, EMC_ERR_INVALID, "emc_stream: NULL argument");
    *stream = (void *)ctx->stream;
    return EMC_OK;
}

EMC_EXPORT int emc_stats_select_hist(emc_ctx *ctx, const double *out_dev, int64_t ld, int64_t n, int field, int shift,
                                     int prefix_shift, const uint64_t *prefixes, int n_prefix, uint64_t *hist_dev)
{
    if (int rc = stats_source(ctx, out_dev, ld, n)) return rc;
    if (!hist_dev || !prefixes || n_prefix < 1 || n_prefix > EMC_SELECT_MAX_PREFIX || field < 0 || field > 2 || shift < 0 || shift > 63)
        return fail(ctx, EMC_ERR_INVALID, "emc_stats_select_hist: bad argument");
    CK(cudaSetDevice(ctx->device));
    SelectArgs a;
    memset(&a, 0, sizeof a);
    for (int u = 0; u < n_prefix; ++u) a.prefix[u] = prefixes[u];
    a.n_prefix = n_prefix; a.field = field; a.shift = shift; a.prefix_shift = prefix_shift;
    CK(cudaMemsetAsync(hist_dev, 0, sizeof(uint64_t) * (size_t)n_prefix * EMC_SELECT_BINS, ctx->stream));
    emc_stats_select_kernel<<<stats_grid(ctx, n), 256, 0, ctx->stream>>>(out_dev, ld, n, a, reinterpret_cast<unsigned long long *>(hist_dev));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return EMC_OK;
}

EMC_EXPORT int emc_stats_select_hist3(emc_ctx *ctx, const double *out_dev, int64_t ld, int64_t n, int shift, int prefix_shift,
                                      const uint64_t *prefixes /*[3][EMC_SELECT_MAX_PREFIX]*/, const int32_t *n_prefix /*[3]*/,
                                      uint64_t *hist_dev /*[n_prefix[0]+n_prefix[1]+n_prefix[2]][EMC_SELECT_BINS], metric-major*/)
{
    if (int rc = stats_source(ctx, out_dev, ld, n)) return rc;
    if (!hist_dev || !prefixes || !n_prefix || shift < 0 || shift > 63) return fail(ctx, EMC_ERR_INVALID, "emc_stats_select_hist3: bad argument");
    Select3Args a;
    memset(&a, 0, sizeof a);
    for (int f = 0; f < 3; ++f) {
        if (n_prefix[f] < 0 || n_prefix[f] > EMC_SELECT_MAX_PREFIX) return fail(ctx, EMC_ERR_INVALID, "emc_stats_select_hist3: n_prefix");
        a.n_prefix[f] = n_prefix[f];
        for (int u = 0; u < n_prefix[f]; ++u) a.prefix[f][u] = prefixes[f * EMC_SELECT_MAX_PREFIX + u];
    }
    a.row0[0] = 0; a.row0[1] = n_prefix[0]; a.row0[2] = n_prefix[0] + n_prefix[1];
    const size_t rows = (size_t)(n_prefix[0] + n_prefix[1] + n_prefix[2]);
    a.shift = shift; a.prefix_shift = prefix_shift;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(hist_dev, 0, sizeof(uint64_t) * rows * EMC_SELECT_BINS, ctx->stream));
    emc_stats_select3_kernel<<<stats_grid(ctx, n), 256, 0, ctx->stream>>>(out_dev, ld, n, a, reinterpret_cast<unsigned long long *>(hist_dev));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return EMC_OK;
}

EMC_EXPORT int emc_stats_linear_hist(emc_ctx *ctx, const double *out_dev, int64_t ld, int64_t n, int field, double lo, double hi,
                                     int nbins, uint64_t *hist_dev)
{
    if (int rc = stats_source(ctx, out_dev, ld, n)) return rc;
    if (!hist_dev || nbins < 1 || field < 0 || field > 4) return fail(ctx, EMC_ERR_INVALID, "emc_stats_linear_hist: bad argument");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(hist_dev, 0, sizeof(uint64_t) * (size_t)nbins, ctx->stream));
    emc_stats_linear_hist_kernel<<<stats_grid(ctx, n), 256, 0, ctx->stream>>>(out_dev, ld, n, field, lo, hi, nbins,
                                                                             reinterpret_cast<unsigned long long *>(hist_dev));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return EMC_OK;
}

EMC_EXPORT int emc_component_debug(emc_ctx *ctx, int component, int64_t n, const double *in, double *out)
{
    static const int n_in[4] = { 1, 3, 6, 5 }, n_out[4] = { 5, 5, 7, 1 };
    if (!ctx || !in || !out || n < 0 || component < 0 || component > 3) return fail(ctx, EMC_ERR_INVALID, "emc_component_debug: bad argument");
    if (!ctx->has_model) return fail(ctx, EMC_ERR_NO_MODEL, "emc_set_model has not been called");
    if (n == 0) return EMC_OK;
    CK(cudaSetDevice(ctx->device));
    if (int rc = make_resident(ctx)) return rc;
    const size_t ni = (size_t)n_in[component] * n, no = (size_t)n_out[component] * n;
    CK(grow(&ctx->d_scratch, &ctx->cap_scratch, (ni + no) * sizeof(double)));
    double *d_in = reinterpret_cast<double *>(ctx->d_scratch), *d_out = d_in + ni;
    CK(cudaMemcpyAsync(d_in, in, ni * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    emc_component_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(component, n, d_in, d_out);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d_out, no * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return EMC_OK;
}

EMC_EXPORT int emc_math_debug(emc_ctx *ctx, int op, int64_t n, const double *x, const double *y, double *out)
{
    if (!ctx || !x || !out || n < 0 || op < 0 || op > 5 || (op == 2 && !y)) return fail(ctx, EMC_ERR_INVALID, "emc_math_debug: bad argument");
    if (n == 0) return EMC_OK;
    CK(cudaSetDevice(ctx->device));
    double *d = nullptr;
    CK(cudaMalloc(&d, sizeof(double) * 3 * n));
    cudaError_t e = cudaMemcpyAsync(d, x, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && y) e = cudaMemcpyAsync(d + n, y, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        emc_math_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(op, n, d, d + n, d + 2 * n);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d + 2 * n, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, EMC_ERR_CUDA, std::string("emc_math_debug: ") + cudaGetErrorString(e));
    return EMC_OK;
}

EMC_EXPORT int emc_get_counters(const emc_ctx *ctx, emc_counters *c)
{
    if (!ctx || !c) return EMC_ERR_INVALID;
    *c = ctx->counters;
    return EMC_OK;
}

EMC_EXPORT int emc_fp64_latency(emc_ctx *ctx, double *cycles_per_dependent_fma)
{
    if (!ctx || !cycles_per_dependent_fma) return fail(ctx, EMC_ERR_INVALID, "emc_fp64_latency: NULL argument");
    CK(cudaSetDevice(ctx->device));
    double *sink = reinterpret_cast<double *>(ctx->d_ctrl + 6);
    long long *cyc = reinterpret_cast<long long *>(ctx->d_ctrl + 7);
    const int iters = 1 << 14;
    emc_dfma_latency_kernel<<<1, 32, 0, ctx->stream>>>(sink, cyc, iters, 0.999999, 1e-9);
    emc_dfma_latency_kernel<<<1, 32, 0, ctx->stream>>>(sink, cyc, iters, 0.999999, 1e-9);
    CK(cudaGetLastError());
    long long h = 0;
    CK(cudaMemcpyAsync(&h, cyc, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *cycles_per_dependent_fma = (double)h / iters;
    return EMC_OK;
}

EMC_EXPORT int emc_fp64_peak(emc_ctx *ctx, double *tflops, double *ms_out)
{
    if (!ctx || !tflops) return fail(ctx, EMC_ERR_INVALID, "emc_fp64_peak: NULL argument");
    CK(cudaSetDevice(ctx->device));
    double *sink = reinterpret_cast<double *>(ctx->d_ctrl + 6);
    const int iters = 1 << 16, threads = 256;
    const int blocks = ctx->sm_count * 8;
    emc_dfma_kernel<<<blocks, threads, 0, ctx->stream>>>(sink, 1024, 0.999999, 1e-9);   /* warm-up */
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(ctx->ev[0], ctx->stream));
        emc_dfma_kernel<<<blocks, threads, 0, ctx->stream>>>(sink, iters, 0.999999, 1e-9);
        CK(cudaEventRecord(ctx->ev[1], ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const double flop = 2.0 * 8.0 * (double)iters * (double)threads * (double)blocks;
    *tflops = flop / ((double)best * 1e-3) * 1e-12;
    if (ms_out) *ms_out = best;
    return EMC_OK;
}

/* ------------------------------------------------------------------------------------------------ */
#include "emc_group.inl"
