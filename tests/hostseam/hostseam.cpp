/*
 * hostseam.cpp — TEST SEAM.  Compiles the engine's device physics (csrc/emc_physics.cuh, the very
 * code the CUDA kernels inline) with g++ so that parity against the oracle and the golden vectors
 * can be checked in a container without a GPU.  It is built into tests/hostseam/_build only, is
 * loaded only by tests/, and is neither part of libemc.so nor reachable from the Python package:
 * the product has no CPU path.
 */
#include <stdint.h>
#include <string.h>

#include "../../erpl_monte_carlo_sim_b200/csrc/emc_model_build.h"
#include "../../erpl_monte_carlo_sim_b200/csrc/emc_strict.cuh"

using namespace emc;

extern "C" {

__attribute__((visibility("default")))
int hs_derivative(const emc_model *m, const emc_inputs *in, int64_t n, const double *t, const double *state,
                  int32_t *chute, double *state_dot)
{
    if (validate_model(*m)) return -1;
    DevModel D; DevTables T;
    build_dev_model(*m, D, T);
    for (int64_t i = 0; i < n; ++i) {
        Sample S;
        load_sample(D, in->scalars + i, in->ld, in->wind ? in->wind + i * in->wind_sample_stride : nullptr, S);
        WindBracket WB; wind_bracket_reset(WB);
        State s, k; Diag dg;
        memcpy(&s, state + 14 * i, sizeof s);
        bool ch = chute[i] != 0; double ct = NAN;
        derivative(D, T, m->wind_altitudes, S, WB, t[i], s, ch, ct, k, true, dg);
        chute[i] = ch ? 1 : 0;
        memcpy(state_dot + 14 * i, &k, sizeof k);
    }
    return 0;
}

/* n flights, rail + RK4 loop, exactly the per-lane sequence of the flight kernel */
__attribute__((visibility("default")))
int hs_batch(const emc_model *m, const emc_inputs *in, int64_t n, const emc_outputs *o, int nan_fast_forward,
             double *tape, int64_t tape_cap, int64_t *n_states)
{
    if (validate_model(*m)) return -1;
    DevModel D; DevTables T;
    build_dev_model(*m, D, T);
    for (int64_t i = 0; i < n; ++i) {
        const double *col = in->scalars + i;
        double *out = o->out + i; int32_t *iout = o->iout + i;
        Sample S;
        load_sample(D, col, in->ld, in->wind ? in->wind + i * in->wind_sample_stride : nullptr, S);
        iout[EMC_IOUT_RAIL_STEPS * o->ld] = rail_phase(D, T, m->wind_altitudes, S, col, in->ld, out, o->ld);
        State s; double t_rail;
        load_flight_state(S, col, in->ld, out, o->ld, s, t_rail);
        Track TK; TrackHot &K = TK.h; const ColdStruct C(TK.c);
        track_init(K, C, s, t_rail);
        RegStore st; store_put(st, s);
        WindBracket WB; wind_bracket_reset(WB);
        int64_t ns = 1, replayed = 0;
        if (tape && tape_cap > 0) { tape[0] = K.t; memcpy(tape + 1, &s, sizeof s); }
        if (!(K.t < D.max_time)) { K.term = EMC_TERM_MAX_TIME; K.finishing = true; }
        for (;;) {
            bool stepped;
            bool retired = lane_advance(D, T, m->wind_altitudes, S, WB, K, C, st, nan_fast_forward != 0, stepped, replayed);
            store_get(st, s);
            if (stepped) {
                if (tape && ns < tape_cap) { tape[ns * EMC_TAPE_WIDTH] = K.t; memcpy(tape + ns * EMC_TAPE_WIDTH + 1, &s, sizeof s); }
                ++ns;
            }
            if (retired) break;
        }
        write_flight_outputs(K, C, s, out, iout, o->ld);
        if (n_states) *n_states = ns;
    }
    return 0;
}


/* strict_derivative on arrays (same layout as hs_derivative) */
__attribute__((visibility("default")))
int hs_strict_derivative(const emc_model *m, const emc_inputs *in, int64_t n, const double *t, const double *state,
                         int32_t *chute, double *state_dot)
{
    if (validate_model(*m)) return -1;
    DevModel D; DevTables T;
    build_dev_model(*m, D, T);
    for (int64_t i = 0; i < n; ++i) {
        Sample S;
        load_sample(D, in->scalars + i, in->ld, in->wind ? in->wind + i * in->wind_sample_stride : nullptr, S);
        WindBracket WB; wind_bracket_reset(WB);
        bool ch = chute[i] != 0; double ct = NAN;
        strict_derivative(D, T, m->wind_altitudes, S, WB, t[i], state + 14 * i, ch, ct, state_dot + 14 * i);
        chute[i] = ch ? 1 : 0;
    }
    return 0;
}


/* n flights continued by the STRICT code (emc_strict.cuh).  from_step < 0: strict from the rail-exit state found in o->out
 * (the rail_* fields must be filled, e.g. by the oracle).  from_step >= 0: the fast path flies until its trigger fires
 * (strict_trigger) or the flight ends, exactly as the flight kernel + emc_strict_kernel pair does. */
struct HostTape {
    double *rows; int64_t cap;
    void operator()(const TrackHot &K, const double *st) const
    {
        if (rows && K.n_steps < cap) { rows[(int64_t)K.n_steps * EMC_TAPE_WIDTH] = K.t; memcpy(rows + (int64_t)K.n_steps * EMC_TAPE_WIDTH + 1, st, 14 * sizeof(double)); }
    }
};

__attribute__((visibility("default")))
int hs_batch_strict(const emc_model *m, const emc_inputs *in, int64_t n, const emc_outputs *o, int nan_fast_forward, int from_step,
                    int64_t *strict_steps, double *tape, int64_t tape_cap)
{
    if (validate_model(*m)) return -1;
    DevModel D; DevTables T;
    build_dev_model(*m, D, T);
    for (int64_t i = 0; i < n; ++i) {
        const double *col = in->scalars + i;
        double *out = o->out + i; int32_t *iout = o->iout + i;
        Sample S;
        load_sample(D, col, in->ld, in->wind ? in->wind + i * in->wind_sample_stride : nullptr, S);
        if (from_step >= 0) iout[EMC_IOUT_RAIL_STEPS * o->ld] = rail_phase(D, T, m->wind_altitudes, S, col, in->ld, out, o->ld);
        State s; double t_rail;
        load_flight_state(S, col, in->ld, out, o->ld, s, t_rail);
        Track TK; TrackHot &K = TK.h; const ColdStruct C(TK.c);
        track_init(K, C, s, t_rail);
        if (!(K.t < D.max_time)) { K.term = EMC_TERM_MAX_TIME; K.finishing = true; }
        int64_t replayed = 0;
        bool finished = false;
        if (from_step >= 0) {
            RegStore st; store_put(st, s);
            WindBracket WB; wind_bracket_reset(WB);
            for (;;) {
                bool stepped;
                if (lane_advance(D, T, m->wind_altitudes, S, WB, K, C, st, nan_fast_forward != 0, stepped, replayed, true)) { finished = (K.replay != EMC_REPLAY_PARK); break; }
            }
            store_get(st, s);
        }
        if (!finished) {
            K.replay = 0;
            double st14[14]; memcpy(st14, &s, sizeof st14);
            int64_t steps = 0;
            const HostTape ht = { tape, tape_cap };
            replayed += strict_fly(D, T, m->wind_altitudes, S, K, C, st14, nan_fast_forward != 0, &steps, ht);
            memcpy(&s, st14, sizeof st14);
            if (strict_steps) strict_steps[i] = steps;
        } else if (strict_steps) strict_steps[i] = 0;
        write_flight_outputs(K, C, s, out, iout, o->ld);
    }
    return 0;
}


/* replay_time (closed form) and replay_time_loop (plain) on the same start state: out = {n, t_end, burnout_time, found} */
__attribute__((visibility("default")))
int hs_replay(double t0, double t_rail, double dt, double max_time, double burn_time, int closed_form, double *out)
{
    DevModel D; memset(&D, 0, sizeof D);
    D.dt = dt; D.max_time = max_time; D.dt_over_6 = dt / 6.0;
    Sample S; memset(&S, 0, sizeof S); S.burn_time = burn_time;
    State s; memset(&s, 0, sizeof s);
    Track TK; TrackHot &K = TK.h; const ColdStruct C(TK.c);
    track_init(K, C, s, t_rail);
    K.t = t0; K.replay = 1;
    int64_t n = closed_form ? replay_time(D, S, K, C, s) : replay_time_loop(D, S, K, C, (int64_t)1 << 40);
    out[0] = (double)n; out[1] = K.t; out[2] = C.getd(TC_BURNOUT_TIME); out[3] = K.burnout_found ? 1.0 : 0.0; out[4] = (double)K.n_steps;
    return 0;
}


__attribute__((visibility("default")))
int hs_series(const emc_model *m, const emc_inputs *in, const double *tape, int64_t n_states, double *series)
{
    if (validate_model(*m)) return -1;
    DevModel D; DevTables T;
    build_dev_model(*m, D, T);
    Sample S;
    load_sample(D, in->scalars, in->ld, in->wind, S);
    for (int64_t i = 0; i < n_states; ++i)
        series_state(D, T, m->wind_altitudes, S, tape + i * EMC_TAPE_WIDTH, tape[i * EMC_TAPE_WIDTH] - tape[0], series + i, n_states);
    return 0;
}


__attribute__((visibility("default")))
int hs_component(const emc_model *m, int comp, int64_t n, const double *in, double *out)
{
    if (validate_model(*m)) return -1;
    DevModel D; DevTables T;
    build_dev_model(*m, D, T);
    for (int64_t i = 0; i < n; ++i) component_eval(D, T, comp, in + i, out + i, n);
    return 0;
}

}
