/*
 * emc.h — C ABI of libemc.so, the B200 batch engine behind the drop-in Python API.
 *
 * The reference (smcconoughey/erpl_monte_carlo_sim) has no FFI layer; its boundary for this
 * path is the Python call FlightSimulator.simulate_flight(initial_conditions, wind_profile,
 * altitude_profile) (rocket_simulation/simulator.py:127) made once per dispersed sample by
 * MonteCarloAnalyzer._run_single_simulation (rocket_simulation/monte_carlo.py:291-295).
 * The entry points below are what a ctypes binding placed at those two call sites binds
 * (see INTEGRATION.md for the stub).
 *
 * Conventions
 *  - plain C: pointers, sizes, doubles and 32/64-bit integers only; no exceptions cross the ABI.
 *  - every function returns 0 on success or a negative emc_status; emc_last_error() gives text.
 *  - the library copies what it needs; the caller keeps ownership of every buffer it passes.
 *  - calls on one context are serialised by the caller; a context owns one CUDA device.  Several contexts may share a
 *    device: the run constants live in per-device __constant__ memory and are re-uploaded whenever the context that
 *    launches is not the one that uploaded last (the device is synchronised at that switch), so contexts never see each
 *    other's model; calls on DIFFERENT contexts of one device must not overlap in time.
 *  - all arithmetic is IEEE-754 binary64 (the reference's Python floats / NumPy float64).
 *  - there is NO CPU implementation behind these symbols: without a CUDA device emc_create fails.
 */
#ifndef EMC_H
#define EMC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMC_ABI_VERSION 3

/* ---- limits of the run-constant tables (rocket.py:43-53, motor.py:31-41) ---- */
#define EMC_MAX_CD_KNOTS 16
#define EMC_MAX_CP_KNOTS 16
#define EMC_MAX_THRUST_KNOTS 32
#define EMC_MAX_WIND_KNOTS 1024

typedef enum emc_status {
    EMC_OK = 0,
    EMC_ERR_INVALID = -1,   /* bad argument (NULL, size, table not increasing, ...) */
    EMC_ERR_NO_DEVICE = -2, /* no CUDA device / wrong architecture: there is no CPU fallback */
    EMC_ERR_CUDA = -3,      /* a CUDA runtime call failed; see emc_last_error */
    EMC_ERR_NO_MODEL = -4,  /* emc_set_model has not been called */
    EMC_ERR_CAPACITY = -5   /* caller-provided buffer too small (tape) */
} emc_status;

enum { EMC_MOTOR_LIQUID = 0, EMC_MOTOR_SOLID = 1 };

/*
 * Run-constant model: every attribute the hot path reads from the reference's parameter objects.
 * Raw attribute values go in; constants derived from them (layer base pressures, interpolation
 * slopes, fin aspect ratio ...) are computed inside the library the way the reference computes them.
 */
typedef struct emc_model {
    /* Rocket (rocket.py:15-66); dry_mass / propellant_mass are per-sample inputs */
    double center_of_mass_dry;
    double Ixx_dry, Iyy_dry;          /* Izz_dry is never read by the path (rocket.py:127) */
    double diameter;                  /* rocket.py:122 propellant Ixx uses diameter/4 */
    double reference_area, reference_diameter;
    double fin_root_chord, fin_tip_chord, fin_span, fin_sweep_angle;
    double cp_location;               /* Barrowman result, rocket.py:56,68-103 */
    double parachute_area, parachute_cd, parachute_deployment_altitude;
    double power_off_drag_factor;
    int32_t n_cd;                     /* rocket.py:43-47 */
    int32_t n_cp;                     /* rocket.py:50-53 */
    double cd_mach[EMC_MAX_CD_KNOTS], cd0[EMC_MAX_CD_KNOTS], cda[EMC_MAX_CD_KNOTS];
    double cp_mach[EMC_MAX_CP_KNOTS], cp_shift[EMC_MAX_CP_KNOTS];
    /* Motor (motor.py): scalars are per-sample inputs; the Solid base curve is run-constant */
    int32_t motor_kind;               /* EMC_MOTOR_LIQUID | EMC_MOTOR_SOLID */
    int32_t n_thrust;                 /* Solid only, motor.py:31-41 */
    double thrust_time[EMC_MAX_THRUST_KNOTS], thrust_curve[EMC_MAX_THRUST_KNOTS];
    /* StandardAtmosphere (environment.py:13-24); sea_level_density/gamma are not read by the path */
    double sea_level_pressure, sea_level_temperature, temperature_lapse_rate;
    double gas_constant, gravity;
    double troposphere_height, stratosphere_height, stratosphere_temp;
    /* FlightSimulator knobs (simulator.py:19-37,42) */
    double max_time, dt_initial, pitch_damping, yaw_damping, rail_length;
    /* Wind table grid (environment.py:267-276): shared altitude grid, per-sample (or shared) values */
    int32_t has_wind;                 /* 0: wind_profile/altitude_profile were None (simulator.py:333) */
    int32_t n_wind;                   /* knots in the altitude grid (<= EMC_MAX_WIND_KNOTS) */
    const double *wind_altitudes;     /* [n_wind], host pointer, copied by emc_set_model */
    /* ABI 2 (appended): atmosphere.gamma, read only by get_properties' speed_of_sound (environment.py:96); the Mach number
     * of the flight path uses the literals 1.4 and 287.053 of utils.mach_number (utils.py:152-157) whatever the
     * atmosphere object holds.  0 = 1.4. */
    double gamma;
} emc_model;

/* ---- per-sample inputs: one field-major block  scalars[EMC_IN_COUNT][ld] ---- */
enum emc_in_field {
    EMC_IN_X = 0, EMC_IN_Y, EMC_IN_Z,         /* initial position (simulator.py:134) */
    EMC_IN_VX, EMC_IN_VY, EMC_IN_VZ,          /* initial velocity (simulator.py:137) */
    EMC_IN_Q0, EMC_IN_Q1, EMC_IN_Q2, EMC_IN_Q3, /* attitude quaternion [w,x,y,z] (utils.py:129-136) */
    EMC_IN_WX, EMC_IN_WY, EMC_IN_WZ,          /* body angular velocity (simulator.py:150) */
    EMC_IN_DRY_MASS, EMC_IN_PROP_MASS,        /* rocket.dry_mass / rocket.propellant_mass (monte_carlo.py:315-316) */
    EMC_IN_THRUST_A,                          /* Liquid: thrust_vacuum; Solid: multiplier on the base curve */
    EMC_IN_NOZZLE_AREA,                       /* motor.nozzle_exit_area */
    EMC_IN_MDOT,                              /* motor.mass_flow_rate */
    EMC_IN_BURN_TIME,                         /* motor.burn_time (monte_carlo.py:258-260) */
    EMC_IN_CD_SCALE,                          /* multiplier on Cd_data['cd0'] (1.0 = reference) */
    EMC_IN_COUNT
};

typedef struct emc_inputs {
    const double *scalars;       /* [EMC_IN_COUNT][ld], field-major */
    int64_t ld;                  /* leading dimension (>= n) */
    const double *wind;          /* [n][n_wind][3] (u,v,w) rows as numpy (N,3); NULL iff !has_wind */
    int64_t wind_sample_stride;  /* doubles between consecutive samples' tables; 0 = one shared table */
} emc_inputs;

/* ---- per-sample outputs: field-major blocks  out[EMC_OUT_COUNT][ld], iout[EMC_IOUT_COUNT][ld] ---- */
enum emc_out_field {
    EMC_OUT_RAIL_EXIT_TIME = 0,               /* simulator.py:104 */
    EMC_OUT_RAIL_EXIT_X, EMC_OUT_RAIL_EXIT_Y, EMC_OUT_RAIL_EXIT_Z,
    EMC_OUT_RAIL_EXIT_VX, EMC_OUT_RAIL_EXIT_VY, EMC_OUT_RAIL_EXIT_VZ,
    EMC_OUT_RAIL_EXIT_SPEED,                  /* simulator.py:107 */
    EMC_OUT_RAIL_EXIT_ROLL, EMC_OUT_RAIL_EXIT_PITCH, EMC_OUT_RAIL_EXIT_YAW, /* simulator.py:108 */
    EMC_OUT_RAIL_EXIT_AOA, EMC_OUT_RAIL_EXIT_SIDESLIP, /* simulator.py:121-122 */
    EMC_OUT_WIND_AT_EXIT_U, EMC_OUT_WIND_AT_EXIT_V, EMC_OUT_WIND_AT_EXIT_W, /* simulator.py:123 */
    EMC_OUT_APOGEE_ALTITUDE, EMC_OUT_APOGEE_TIME, /* simulator.py:488-490 (time since rail exit) */
    EMC_OUT_RANGE,                            /* simulator.py:493-494 */
    EMC_OUT_FLIGHT_TIME,                      /* simulator.py:582 */
    EMC_OUT_FINAL_X, EMC_OUT_FINAL_Y, EMC_OUT_FINAL_Z,     /* landing point */
    EMC_OUT_FINAL_VX, EMC_OUT_FINAL_VY, EMC_OUT_FINAL_VZ,
    /* maxima/minima over the stored states, NumPy max/min semantics (NaN wins) */
    EMC_OUT_MAX_MACH,                         /* simulator.py:530-532 */
    EMC_OUT_MAX_Q,                            /* simulator.py:541 */
    EMC_OUT_MAX_SPEED,                        /* simulator.py:476; analyze_outlier.py:20 */
    EMC_OUT_MAX_ABS_OMEGA,                    /* analyze_outlier.py:25: max |component| */
    EMC_OUT_MIN_STABILITY, EMC_OUT_MAX_STABILITY, /* simulator.py:549; analyze_outlier.py:24 */
    EMC_OUT_MAX_ABS_AOA,                      /* simulator.py:533 */
    EMC_OUT_BURNOUT_TIME,                     /* simulator.py:479-480: time[argmax(time > burn_time)] */
    EMC_OUT_CHUTE_TIME,                       /* time since ignition of the derivative call that latched simulator.py:366-369; NaN if never */
    EMC_OUT_COUNT
};

enum emc_iout_field {
    EMC_IOUT_N_STEPS = 0,     /* accepted RK4 steps = stored states - 1 (simulator.py:216-264) */
    EMC_IOUT_TERMINATION,     /* emc_termination */
    EMC_IOUT_APOGEE_INDEX,    /* simulator.py:488 np.argmax(altitudes) */
    EMC_IOUT_FIRST_NAN_STEP,  /* first stored-state index whose altitude is NaN, -1 if none */
    EMC_IOUT_RAIL_STEPS,      /* Euler steps on the rail (simulator.py:63-96) */
    EMC_IOUT_COUNT
};

typedef enum emc_termination {
    EMC_TERM_NONE = 0,
    EMC_TERM_GROUND = 1,      /* simulator.py:238-239 */
    EMC_TERM_ALTITUDE = 2,    /* simulator.py:242-244 */
    EMC_TERM_COAST = 3,       /* simulator.py:260-264 */
    EMC_TERM_MAX_TIME = 4     /* loop guard, simulator.py:216 */
} emc_termination;

typedef struct emc_outputs {
    double *out;     /* [EMC_OUT_COUNT][ld] */
    int32_t *iout;   /* [EMC_IOUT_COUNT][ld] */
    int64_t ld;
} emc_outputs;

/* Tape row layout for emc_run_tape: t (since ignition) followed by the 14 state components */
#define EMC_TAPE_WIDTH 15

typedef struct emc_run_opts {
    int32_t refill_threshold; /* idle lanes per warp before the warp refills from the work queue; 0 = default */
    int32_t block_threads;    /* 0 = default (128) */
    int32_t blocks_per_sm;    /* 0 = default (3 when block_threads is 0, else the occupancy limit) */
    int32_t nan_fast_forward; /* 1 (default when opts==NULL): replay t += dt only once the altitude is NaN for good */
    int32_t cold_state_in_smem; /* 0 (default) or 2: per-lane bookkeeping, base state and RK4 accumulator live in shared memory; 1: bookkeeping only; -1: registers */
    int32_t flags;            /* ABI 2: EMC_RUN_* bits */
} emc_run_opts;
#define EMC_RUN_NO_STRICT_TAIL 2   /* accepted and IGNORED (it used to finish every trajectory on the fast path).  A trajectory whose stored
                                     state shows the reference's blow-up in its last steps (|v| > 1e7 m/s or |omega| > 1000 rad/s) is always
                                     finished by the strict continuation, which follows the reference's arithmetic operation by operation (no
                                     FMA contraction, IEEE division / square root, libm), so that the inf / NaN pattern of the overflowing steps
                                     — step count, termination, first-NaN index — is the reference's; so is every IRREGULAR sample (dry mass
                                     not positive and finite, propellant mass negative or not finite), from its first state: the fast path
                                     drops the guards of simulator.py:315-318,431-436 that only such samples can trigger. */
#define EMC_RUN_NO_YIELD 4        /* ABI 3: switch off the lane hand-back described at emc_counters.yielded (A/B runs, scheduling-invariance tests) */
#define EMC_RUN_COMPACTION 1      /* tail compaction: once the work queue is empty, sparse warps hand their trajectories (lane records in
                                     shared memory, addressed by slot) to one collector warp per block and exit.  Bit-identical outputs.
                                     Off by default: measured on the B200 the slot indirection costs more than the compacted tail
                                     returns. */

typedef struct emc_ctx emc_ctx;

/* Kernel-side counters of the last emc_run_* call */
typedef struct emc_counters {
    int64_t rk4_steps;        /* accepted RK4 steps integrated (excludes NaN fast-forward replays) */
    int64_t replay_steps;     /* t += dt replays of all-NaN trajectories */
    int64_t rail_steps;       /* Euler steps on the rail */
    int64_t refills;          /* work-queue fetches */
    int64_t kernel_launches;  /* kernels launched by the call */
    double rail_ms, flight_ms;/* device time of the two kernels (CUDA events on the context stream) */
    int64_t tape_rows;        /* ABI 2: rows stored by an armed batch tape (emc_tape_request) */
    int64_t handovers;        /* ABI 2: trajectories handed to a collector warp by the tail compaction (EMC_RUN_COMPACTION) */
    int64_t parked;           /* ABI 2: trajectories finished by the strict continuation (blow-up under way; csrc/emc_strict.cuh) */
    int64_t strict_steps;     /* ABI 2: RK4 steps taken there (not included in rk4_steps) */
    double strict_ms;         /* ABI 2: device time between the end of the flight kernel and the end of the strict continuation (it runs
                               * concurrently on a second stream: normally the cost of the final sweep only) */
    int64_t yielded;          /* ABI 3: trajectories that gave their lane back once, after 900 stored states, while unstarted samples
                               * were waiting, and were resumed later (a batch larger than the resident lanes, up to 8 x as large: every
                               * sample is STARTED early, and the flights whose attitude oscillation has settled — the ones that fly
                               * longest — keep their lane, so the longest trajectory of the batch is not one that started in the last
                               * wave; only warps whose lanes started together and of which at least half show a growing oscillation hand
                               * back: launch -> landing flights never do).  Outputs are bit-identical with and without it
                               * (EMC_RUN_NO_YIELD).  DESIGN.md section 5. */
} emc_counters;

int emc_abi_version(void);
const char *emc_last_error(const emc_ctx *ctx);     /* ctx may be NULL: error of a failed emc_create */

int emc_create(emc_ctx **ctx, int device);          /* replaces FlightSimulator.__init__ state (simulator.py:12-40) */
int emc_destroy(emc_ctx *ctx);
int emc_set_model(emc_ctx *ctx, const emc_model *model);

/* Host-buffer entry: simulate n samples, launch -> termination (simulator.py:127-293 per sample). */
int emc_run_batch(emc_ctx *ctx, const emc_inputs *in, int64_t n, const emc_outputs *out,
                  const emc_run_opts *opts);

/* Device-buffer entry: same, all pointers are device pointers owned by the caller (resident inputs). */
int emc_run_batch_device(emc_ctx *ctx, const emc_inputs *in_dev, int64_t n, const emc_outputs *out_dev,
                         const emc_run_opts *opts);

/* One sample with every stored state written to tape[cap][EMC_TAPE_WIDTH] (simulator.py:212-231). */
int emc_run_tape(emc_ctx *ctx, const emc_inputs *in, const emc_outputs *out,
                 double *tape, int64_t cap, int64_t *n_states);

/* ---- downsampled trajectory tape of a batch (reference monte_carlo.py:296-302: every result of a Monte Carlo run
 *      carries 'trajectory' = {time, altitude, position}; plot_trajectory_cloud(_3d) :635-707 read it) -------------
 * Arms the NEXT emc_run_batch / emc_run_batch_device / emc_run_batch_staged of this context: for every listed sample the
 * flight kernel stores the stored states 0, stride, 2*stride, ... (state 0 = rail exit, simulator.py:212-214) and the last
 * integrated one as rows {t - t_rail, x, y, z} (the reference's shifted time axis, simulator.py:464) into HBM while it
 * flies — nothing is re-flown.  Rows past max_rows are counted but not stored; a trajectory that is fast-forwarded as
 * all-NaN (emc_run_opts.nan_fast_forward) stops recording at the fast-forward point.  The request is consumed by that
 * run; emc_tape_fetch then copies rows[n_sel][max_rows][EMC_BTAPE_WIDTH] and n_rows[n_sel] to the host. */
#define EMC_BTAPE_WIDTH 4
int emc_tape_request(emc_ctx *ctx, const int64_t *samples /*host [n_sel], indices into the next batch*/, int64_t n_sel,
                     int32_t stride, int32_t max_rows);
int emc_tape_fetch(emc_ctx *ctx, double *rows /*host [n_sel][max_rows][4]*/, int32_t *n_rows /*host [n_sel]*/);
/* device pointers of the tape of the last armed run (rows_dev [n_sel][max_rows][4], n_rows_dev [n_sel]) and its size */
int emc_tape_resident(emc_ctx *ctx, double **rows_dev, int32_t **n_rows_dev, int64_t *n_sel, int32_t *max_rows);

/* ---- device-side dispersion draws (reference monte_carlo.py:156-201,225-288; motor.py:95-125,171-186;
 *      environment.py:125-200,218-265) -------------------------------------------------------------------
 * Counter-based Philox4x32-10: sample i uses counter (i, j, stream) and the run seed as key.  The structure of the
 * reference's per-sample streams is kept (one normal stream consumed from its start by the parameter draw, the motor
 * and the wind generator); the bits are Philox/Box-Muller, so this mode matches the reference in
 * distribution.  With `gauss`/`unif` given, the same kernel perturbs with the caller's draws (host-seeded NumPy). */
typedef struct emc_dispersion {
    double base_pos[3], base_vel[3], base_att[3], base_omega[3];       /* base initial conditions (euler xyz) */
    double sigma_pos[3], sigma_vel[3], sigma_att[3], sigma_omega[3];   /* monte_carlo.py:36-39 */
    double mass_sigma;                                                 /* :40 */
    double wind_speed_lo, wind_speed_hi, wind_dir_lo, wind_dir_hi;     /* :45-46 */
    double dry_mass, propellant_mass;                                  /* nominal rocket, :315-316 */
    double thrust_vacuum, thrust_sea_level, mass_flow_rate;            /* nominal Liquid motor, motor.py:177-184 */
    double nozzle_exit_area, motor_propellant_mass, motor_burn_time;   /* nominal Solid: Ae (motor.py:123), own burn time */
    double thrust_sigma, flow_sigma, burn_sigma;                       /* motor.py:50-51,149-150 */
    int32_t motor_kind;                                                /* EMC_MOTOR_LIQUID | EMC_MOTOR_SOLID */
    int32_t wind_mode;                                                 /* 0: generate_stochastic_profile, 1: perturb_wind_profile + offset */
    int32_t n_knots, pad_;
    const double *shear;       /* [n_knots] (z/10)^0.14, mode 0 (environment.py:118-123)            host pointer */
    const double *base_wind;   /* [n_knots][3], mode 1                                             host pointer */
    const double *rho;         /* [n_knots] AR(1) correlation (environment.py:175-177)              host pointer */
    const double *innov;       /* [n_knots] innovation scale; innov[0] = surface turbulence scale   host pointer */
} emc_dispersion;

/* Fill scalars_dev[EMC_IN_COUNT][ld] and wind_dev[n][n_knots][3] (DEVICE pointers; both NULL = the context's own staging
 * buffers, to be flown with emc_run_batch_staged) for samples first_index .. first_index+n-1.
 * gauss/unif: HOST arrays [n][n_gauss] / [n][2] of standard normals / uniforms to use instead of Philox, or NULL. */
int emc_generate_inputs(emc_ctx *ctx, const emc_dispersion *d, uint64_t seed, int64_t first_index, int64_t n,
                        const double *gauss, int64_t n_gauss, const double *unif,
                        double *scalars_dev, int64_t ld, double *wind_dev);
/* The same, with the reference's OWN streams regenerated on the device: sample i uses MT19937 seeded with
 * first_seed + i (np.random.seed(i) / RandomState(i)) and NumPy's legacy polar Gaussian — the draws a host-seeded run
 * uploads, without the host loop.  scalars_dev/wind_dev as above (NULL = stage inside the context). */
int emc_generate_inputs_numpy(emc_ctx *ctx, const emc_dispersion *d, int64_t first_seed, int64_t n,
                              double *scalars_dev, int64_t ld, double *wind_dev);
/* the device-regenerated NumPy draws themselves: gauss[n][n_gauss] (RandomState(seed).standard_normal(n_gauss)) and
 * unif[n][2] (the two random_sample() values that follow the first 14 normals in the parameter stream) and, if not
 * NULL, density[n] (the normal drawn after them); host arrays */
int emc_numpy_draws(emc_ctx *ctx, int64_t first_seed, int64_t n, int64_t n_gauss, double *gauss, double *unif, double *density);

/* fly the n samples staged by emc_generate_inputs(..., NULL, 0, NULL); outputs to host buffers */
int emc_run_batch_staged(emc_ctx *ctx, int64_t n, const emc_outputs *out, const emc_run_opts *opts);
/* copy the staged inputs back (either pointer may be NULL): scalars[EMC_IN_COUNT][n], wind[n][n_knots][3] */
int emc_staged_inputs(emc_ctx *ctx, int64_t n, double *scalars_host, double *wind_host);
/* the Philox draws of samples first_index.. : gauss[n][n_gauss], unif[n][2] (host) */
int emc_philox_draws(emc_ctx *ctx, uint64_t seed, int64_t first_index, int64_t n, int64_t n_gauss, double *gauss, double *unif);

/* ---- per-state result series of one flight (reference simulator.py:496-552, _extract_results) ----------
 * series[EMC_SERIES_COUNT][n_states], field-major, from the tape of emc_run_tape (tape[i] = t, state[14]). */
enum emc_series_field {
    EMC_SER_MASS = 0, EMC_SER_IXX, EMC_SER_IYY, EMC_SER_IZZ,        /* simulator.py:515-520 */
    EMC_SER_CENTER_OF_MASS,                                         /* :516 */
    EMC_SER_EULER_ROLL, EMC_SER_EULER_PITCH, EMC_SER_EULER_YAW,     /* :512 */
    EMC_SER_THRUST,                                                 /* :543 (evaluated at the SHIFTED time, as the reference does) */
    EMC_SER_DRAG,                                                   /* :542 */
    EMC_SER_CD, EMC_SER_CL, EMC_SER_CM,                             /* :544-546 */
    EMC_SER_CP_DYNAMIC, EMC_SER_STABILITY_MARGIN,                   /* :548-549 */
    EMC_SER_AOA, EMC_SER_SIDESLIP,                                  /* :551-552 */
    EMC_SER_SPEED,                                                  /* :476 */
    EMC_SER_MACH, EMC_SER_QDYN,                                     /* :532, :541 (not result keys; used for max Mach / max Q) */
    EMC_SERIES_COUNT
};
int emc_extract_series(emc_ctx *ctx, const emc_inputs *in /*one sample*/, const double *tape /*[n_states][EMC_TAPE_WIDTH], host*/,
                       int64_t n_states, double *series /*[EMC_SERIES_COUNT][n_states], host*/);

/* Test seam: out[i][14] = _rocket_dynamics(t[i], state[i]) (simulator.py:295-460) for sample i.
 * chute[i] is the sticky parachute flag, read and written back (simulator.py:366-369). */
int emc_derivative_debug(emc_ctx *ctx, const emc_inputs *in, int64_t n, const double *t,
                         const double *state /*[n][14]*/, int32_t *chute /*[n]*/,
                         double *state_dot /*[n][14]*/);

int emc_get_counters(const emc_ctx *ctx, emc_counters *c);

/* ---- device-side Monte Carlo statistics (reference monte_carlo.py:337-473) --------------------------------
 * All buffers below are DEVICE pointers (HBM).  `out_dev` is a field-major [EMC_OUT_COUNT][ld] block; NULL means
 * "the outputs of this context's last emc_run_batch, still resident".  The SUM/MIN/MAX blocks are what a
 * multi-GPU job all-reduces (NCCL) between the passes; see erpl_monte_carlo_sim_b200/stats.py. */
enum { EMC_ST_N = 0, EMC_ST_VALID, EMC_ST_OUTLIER, EMC_ST_NONFINITE, EMC_ST_AP_HIGH, EMC_ST_AP_LOW, EMC_ST_RANGE,
       EMC_ST_TIME, EMC_ST_ENERGY, EMC_ST_SUM_AP, EMC_ST_SUM_RG, EMC_ST_SUM_FT, EMC_ST_SUM_X, EMC_ST_SUM_Y, EMC_ST_SUM_COUNT };
#define EMC_ST_MM_COUNT 3     /* apogee, range, flight_time */
#define EMC_ST2_COUNT 6       /* centred: apogee^2, range^2, time^2, xx, xy, yy (landing ellipse) */
#define EMC_SELECT_BINS 2048
#define EMC_SELECT_MAX_PREFIX 16

/* context-owned device scratch (valid until the next emc_scratch call with a larger size or emc_destroy) */
int emc_scratch(emc_ctx *ctx, int64_t bytes, void **dev_ptr);
int emc_copy_to_host(emc_ctx *ctx, void *host, const void *dev, int64_t bytes);
int emc_copy_to_device(emc_ctx *ctx, void *dev, const void *host, int64_t bytes);

/* device pointer and leading dimension of the outputs left resident by the last host-buffer run (for statistics over
 * sub-ranges: pass out_dev + first_sample with the same ld) */
int emc_resident_outputs(emc_ctx *ctx, double **out_dev, int64_t *ld);

/* download the resident outputs (both blocks) of the last batch run of n samples: what emc_run_batch_staged with NULL
 * output pointers left in HBM.  The reference returns every sample's result dict (monte_carlo.py:296-302); the package
 * builds them on access, so a statistics-only campaign never moves the 300 bytes per sample. */
int emc_fetch_outputs(emc_ctx *ctx, int64_t n, const emc_outputs *out);

/* make a host [EMC_OUT_COUNT][ld] block the context's resident outputs (for statistics over a run that was
 * executed in several chunks) */
int emc_upload_outputs(emc_ctx *ctx, const double *out_host, int64_t ld, int64_t n);

/* outlier classification (monte_carlo.py:348-390) + counts, sums, min, max over the valid samples */
int emc_stats_moments1(emc_ctx *ctx, const double *out_dev, int64_t ld, int64_t n,
                       double *sum_dev /*[EMC_ST_SUM_COUNT]*/, double *min_dev /*[3]*/, double *max_dev /*[3]*/);
/* centred second moments about center_dev[5] = mean apogee, range, flight_time, landing x, y (np.std two-pass) */
int emc_stats_moments2(emc_ctx *ctx, const double *out_dev, int64_t ld, int64_t n, const double *center_dev,
                       double *sum_dev /*[EMC_ST2_COUNT]*/);
/* one digit pass of the exact radix select behind np.percentile: field 0 apogee, 1 range, 2 flight_time;
 * hist_dev[u][digit] += 1 for every valid sample whose key >> prefix_shift == prefixes[u] (prefix_shift 64: all) */
int emc_stats_select_hist(emc_ctx *ctx, const double *out_dev, int64_t ld, int64_t n, int field, int shift,
                          int prefix_shift, const uint64_t *prefixes /*host, [n_prefix]*/, int n_prefix,
                          uint64_t *hist_dev /*[n_prefix][EMC_SELECT_BINS], zeroed by the call*/);

/* the same digit pass for all three metrics at once: prefixes[3][EMC_SELECT_MAX_PREFIX] (host), n_prefix[3],
 * hist_dev[n_prefix[0]+n_prefix[1]+n_prefix[2]][EMC_SELECT_BINS] (compact, metric-major) */
int emc_stats_select_hist3(emc_ctx *ctx, const double *out_dev, int64_t ld, int64_t n, int shift, int prefix_shift,
                           const uint64_t *prefixes, const int32_t *n_prefix, uint64_t *hist_dev);

/* The whole single-GPU summary in one stream-ordered chain (no host round trips between the passes): counts by reason,
 * sum / min / max, centred second moments and the np.percentile(method="linear") order statistics of apogee, range and
 * flight time over the valid samples.  result (host) = sum[14] | min[3] | max[3] | s2[6] | means[5] | 0 |
 * val[3][2*n_pct] (lo and hi order statistic of every percentile; NaN when there is no valid sample).
 * n_pct <= 8.  out_dev NULL = the context's resident outputs.  Multi-GPU jobs use the pass-by-pass entry points above,
 * whose blocks are all-reduced between passes. */
int emc_stats_summary(emc_ctx *ctx, const double *out_dev, int64_t ld, int64_t n, const double *percentiles, int n_pct,
                      double *result);
/* The same chain one stage at a time, for several GPUs: every stage is only ENQUEUED on the context stream; after the
 * stages that return a block (0: sum[14] | min[3] | max[3]; 1: s2[6]; 2, 4, .., 12: the digit histograms, uint64) the
 * caller all-reduces that block across ranks on the same stream (emc_stream) and goes on; stage 14 copies the result
 * (layout of emc_stats_summary, now over the whole job) and synchronises.  Stages 0..14 in order. */
int emc_stats_summary_stage(emc_ctx *ctx, const double *out_dev, int64_t ld, int64_t n, const double *percentiles, int n_pct,
                            int stage, void **block_dev, int64_t *block_words, double *result);
/* the context's cudaStream_t */
int emc_stream(emc_ctx *ctx, void **stream);

/* fixed-bin histogram over the valid samples; field 0 apogee, 1 range, 2 flight_time, 3 landing x, 4 landing y */
int emc_stats_linear_hist(emc_ctx *ctx, const double *out_dev, int64_t ld, int64_t n, int field, double lo, double hi,
                          int nbins, uint64_t *hist_dev /*[nbins], zeroed by the call*/);

/* ------------------------------------------------------------------------------------------------------------------
 * Several GPUs of one box from ONE host process, without torch (ABI 2).  The reference fans samples out over a process pool
 * (rocket_simulation/monte_carlo.py:63-83) and analyses the gathered list in the parent (:337-473); a group owns one
 * context per device, cuts [0, n) into contiguous shards (np.array_split's sizes), flies them concurrently (one host
 * thread per device inside the call) and reduces the statistics with NCCL all-reduces of the small blocks between the
 * stages of emc_stats_summary_stage, on the devices' streams.  NCCL (libnccl.so.2) is loaded with dlopen by the first
 * group of more than one device.  Python jobs under torchrun use one context per rank and torch.distributed instead
 * (erpl_monte_carlo_sim_b200/stats.py); both drive the same staged chain.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct emc_group emc_group;
int emc_group_create(emc_group **out, const int *devices, int n_dev);
int emc_group_destroy(emc_group *g);
const char *emc_group_last_error(const emc_group *g);            /* g NULL: the last emc_group_create failure of this thread */
int emc_group_size(const emc_group *g);
emc_ctx *emc_group_context(emc_group *g, int i);                 /* the i-th device's context (owned by the group) */
int emc_group_set_model(emc_group *g, const emc_model *model);
/* host buffers as in emc_run_batch, covering all n samples; every device's shard stays resident in its HBM */
int emc_group_run_batch(emc_group *g, const emc_inputs *in, int64_t n, const emc_outputs *out, const emc_run_opts *opts);
int emc_group_shard(const emc_group *g, int i, int64_t *first, int64_t *count);       /* shard of device i in the last run */
/* statistics of the WHOLE job over the resident shards; result laid out as in emc_stats_summary */
int emc_group_stats_summary(emc_group *g, const double *percentiles, int n_pct, double *result);
int emc_group_get_counters(const emc_group *g, emc_counters *c);  /* sums over the devices; the times are maxima */

/* Component evaluation on arrays (the model-evaluation helpers of the reference's parameter classes, run on the device;
 * also the component known-answer test seam).  in[k][n] / out[k][n] are HOST, field-major.
 *   EMC_COMP_ATMOSPHERE  in: altitude                          out: temperature, pressure, density, speed_of_sound, gravity
 *                        (environment.py:26-108)
 *   EMC_COMP_MASS        in: propellant_fraction, dry_mass, propellant_mass        out: mass, center_of_mass, Ixx, Iyy, Izz
 *                        (rocket.py:110-136)
 *   EMC_COMP_AERO        in: mach, alpha, beta, center_of_mass, power_on(0/1), cd_scale
 *                        out: cd, cl, cm, cy, cyaw, cp, cn                  (rocket.py:105-108,138-218)
 *   EMC_COMP_THRUST      in: time, ambient_pressure, thrust_a, nozzle_exit_area, burn_time      out: thrust   (motor.py:54-76,152-156) */
enum { EMC_COMP_ATMOSPHERE = 0, EMC_COMP_MASS = 1, EMC_COMP_AERO = 2, EMC_COMP_THRUST = 3 };
int emc_component_debug(emc_ctx *ctx, int component, int64_t n, const double *in, double *out);

/* Test seam: the engine's device math helpers on arrays.  op 0: 1/x, 1: 1/sqrt(x), 2: atan2(y, x),
 * 3: sqrt(x), 4: exp(x), 5: log(x) (as used by the derivative kernel: MUFU seed + Newton / minimax polynomials). */
int emc_math_debug(emc_ctx *ctx, int op, int64_t n, const double *x, const double *y, double *out);

/* Register-resident DFMA chain: measures this GPU's FP64 FMA peak (the roofline denominator). */
int emc_fp64_peak(emc_ctx *ctx, double *tflops, double *ms);
/* One dependent DFMA chain in one warp: SM cycles per dependent FP64 FMA (the latency a single trajectory sees). */
int emc_fp64_latency(emc_ctx *ctx, double *cycles_per_dependent_fma);

#ifdef __cplusplus
}
#endif
#endif /* EMC_H */
