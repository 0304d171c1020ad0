/*
 * emc_group — several B200s of one box behind the C ABI, without torch (included at the end of emc_engine.cu).
 *
 * The reference fans its samples out over a process pool (rocket_simulation/monte_carlo.py:63-83) and analyses the
 * gathered list in the parent (:337-473).  Here one host process owns one emc_ctx per device: emc_group_run_batch cuts
 * the sample range into contiguous shards, one host thread per device runs the ordinary emc_run_batch on its shard
 * (uploads, rail / flight / strict kernels and downloads of different devices overlap), and emc_group_stats_summary runs
 * the staged statistics chain (emc_stats_summary_stage) on every device with the small blocks all-reduced by NCCL between
 * stages, stream-ordered, over NVLink.  There is no data-path collective: the only exchange is the statistics.
 * NCCL is loaded with dlopen when the first group of more than one device is created, so libemc.so itself has no link
 * dependency on it (a one-device group never needs it).
 */
#include <dlfcn.h>
#include <thread>
#include <vector>

namespace {
typedef void *nccl_comm;
struct NcclApi {
    void *lib = nullptr;
    int (*CommInitAll)(nccl_comm *, int, const int *) = nullptr;
    int (*CommDestroy)(nccl_comm) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
enum { NCCL_SUM = 0, NCCL_MAX = 2, NCCL_MIN = 3, NCCL_UINT64 = 5, NCCL_FLOAT64 = 8 };      /* nccl.h: ncclRedOp_t, ncclDataType_t */

bool load_nccl(NcclApi &A, std::string &err)
{
    if (A.lib) return true;
    const char *names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char *nm : names) { A.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (A.lib) break; }
    if (!A.lib) { err = std::string("NCCL not found (dlopen libnccl.so.2): ") + dlerror(); return false; }
    bool ok = true;
    auto sym = [&](const char *n) { void *p = dlsym(A.lib, n); if (!p) { ok = false; err = std::string("NCCL symbol missing: ") + n; } return p; };
    A.CommInitAll = reinterpret_cast<decltype(A.CommInitAll)>(sym("ncclCommInitAll"));
    A.CommDestroy = reinterpret_cast<decltype(A.CommDestroy)>(sym("ncclCommDestroy"));
    A.AllReduce = reinterpret_cast<decltype(A.AllReduce)>(sym("ncclAllReduce"));
    A.GroupStart = reinterpret_cast<decltype(A.GroupStart)>(sym("ncclGroupStart"));
    A.GroupEnd = reinterpret_cast<decltype(A.GroupEnd)>(sym("ncclGroupEnd"));
    A.GetErrorString = reinterpret_cast<decltype(A.GetErrorString)>(sym("ncclGetErrorString"));
    if (!ok) { dlclose(A.lib); A.lib = nullptr; }
    return ok;
}
NcclApi g_nccl;
}  // namespace

struct emc_group {
    std::vector<emc_ctx *> ctx;
    std::vector<nccl_comm> comm;
    std::vector<int64_t> lo, hi;          /* shard of the last run */
    std::string err;
};

static thread_local std::string g_group_create_err;
static int gfail(emc_group *g, int code, const std::string &msg)
{
    if (g) g->err = msg; else g_group_create_err = msg;
    return code;
}

EMC_EXPORT const char *emc_group_last_error(const emc_group *g) { return g ? g->err.c_str() : g_group_create_err.c_str(); }

EMC_EXPORT int emc_group_destroy(emc_group *g)
{
    if (!g) return EMC_OK;
    for (nccl_comm c : g->comm) if (c && g_nccl.CommDestroy) g_nccl.CommDestroy(c);
    for (emc_ctx *c : g->ctx) if (c) emc_destroy(c);
    delete g;
    return EMC_OK;
}

EMC_EXPORT int emc_group_create(emc_group **out, const int *devices, int n_dev)
{
    if (!out || !devices || n_dev < 1 || n_dev > 64) return gfail(nullptr, EMC_ERR_INVALID, "emc_group_create: bad argument");
    *out = nullptr;
    emc_group *g = new emc_group;
    for (int i = 0; i < n_dev; ++i) {
        emc_ctx *c = nullptr;
        const int rc = emc_create(&c, devices[i]);
        if (rc != EMC_OK) { const std::string m = std::string("device ") + std::to_string(devices[i]) + ": " + emc_last_error(nullptr); emc_group_destroy(g); return gfail(nullptr, rc, m); }
        g->ctx.push_back(c);
    }
    if (n_dev > 1) {
        std::string e;
        if (!load_nccl(g_nccl, e)) { emc_group_destroy(g); return gfail(nullptr, EMC_ERR_CUDA, e); }
        g->comm.assign((size_t)n_dev, nullptr);
        const int rc = g_nccl.CommInitAll(g->comm.data(), n_dev, devices);
        if (rc != 0) { const std::string m = std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(rc); g->comm.clear(); emc_group_destroy(g); return gfail(nullptr, EMC_ERR_CUDA, m); }
    }
    g->lo.assign((size_t)n_dev, 0); g->hi.assign((size_t)n_dev, 0);
    *out = g;
    return EMC_OK;
}

EMC_EXPORT int emc_group_size(const emc_group *g) { return g ? (int)g->ctx.size() : 0; }
EMC_EXPORT emc_ctx *emc_group_context(emc_group *g, int i) { return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[(size_t)i] : nullptr; }

EMC_EXPORT int emc_group_set_model(emc_group *g, const emc_model *model)
{
    if (!g || !model) return gfail(g, EMC_ERR_INVALID, "emc_group_set_model: NULL argument");
    for (size_t i = 0; i < g->ctx.size(); ++i) {
        const int rc = emc_set_model(g->ctx[i], model);
        if (rc != EMC_OK) return gfail(g, rc, std::string("device ") + std::to_string(g->ctx[i]->device) + ": " + emc_last_error(g->ctx[i]));
    }
    return EMC_OK;
}

/* samples [0, n) -> contiguous shards, the first (n mod D) devices take one more (the split of np.array_split) */
EMC_EXPORT int emc_group_run_batch(emc_group *g, const emc_inputs *in, int64_t n, const emc_outputs *out, const emc_run_opts *opts)
{
    if (!g || !in || !out || n < 0) return gfail(g, EMC_ERR_INVALID, "emc_group_run_batch: bad argument");
    const int64_t D = (int64_t)g->ctx.size();
    std::vector<int> rc((size_t)D, EMC_OK);
    std::vector<std::thread> th;
    for (int64_t i = 0; i < D; ++i) {
        g->lo[(size_t)i] = i * (n / D) + (i < n % D ? i : n % D);
        g->hi[(size_t)i] = g->lo[(size_t)i] + n / D + (i < n % D ? 1 : 0);
    }
    auto work = [&](int64_t i) {
        const int64_t lo = g->lo[(size_t)i], m = g->hi[(size_t)i] - lo;
        emc_inputs si = *in;
        si.scalars = in->scalars + lo;                                   /* field-major: same leading dimension */
        if (in->wind && in->wind_sample_stride > 0) si.wind = in->wind + lo * in->wind_sample_stride;
        emc_outputs so = *out;
        if (out->out) so.out = out->out + lo;
        if (out->iout) so.iout = out->iout + lo;
        rc[(size_t)i] = emc_run_batch(g->ctx[(size_t)i], &si, m, &so, opts);
    };
    for (int64_t i = 1; i < D; ++i) th.emplace_back(work, i);
    work(0);
    for (std::thread &t : th) t.join();
    for (int64_t i = 0; i < D; ++i)
        if (rc[(size_t)i] != EMC_OK) return gfail(g, rc[(size_t)i], std::string("device ") + std::to_string(g->ctx[(size_t)i]->device) + ": " + emc_last_error(g->ctx[(size_t)i]));
    return EMC_OK;
}

EMC_EXPORT int emc_group_shard(const emc_group *g, int i, int64_t *first, int64_t *count)
{
    if (!g || i < 0 || i >= (int)g->ctx.size() || !first || !count) return EMC_ERR_INVALID;
    *first = g->lo[(size_t)i]; *count = g->hi[(size_t)i] - g->lo[(size_t)i];
    return EMC_OK;
}

/* statistics of the whole job over the outputs every device holds from the last emc_group_run_batch; result as
 * emc_stats_summary.  Every stage is enqueued on each device's stream and the blocks it returns are all-reduced in place
 * on those streams: no host synchronisation before the last stage. */
EMC_EXPORT int emc_group_stats_summary(emc_group *g, const double *percentiles, int n_pct, double *result)
{
    if (!g || !percentiles || !result) return gfail(g, EMC_ERR_INVALID, "emc_group_stats_summary: NULL argument");
    const size_t D = g->ctx.size();
    if (D == 1) {
        const int rc = emc_stats_summary(g->ctx[0], nullptr, 0, g->hi[0] - g->lo[0], percentiles, n_pct, result);
        return rc == EMC_OK ? rc : gfail(g, rc, emc_last_error(g->ctx[0]));
    }
    std::vector<void *> blk(D, nullptr);
    for (int stage = 0; stage <= 14; ++stage) {
        int64_t words = 0;
        for (size_t i = 0; i < D; ++i) {
            int64_t w = 0;
            const int rc = emc_stats_summary_stage(g->ctx[i], nullptr, 0, g->hi[i] - g->lo[i], percentiles, n_pct, stage, &blk[i], &w, result);
            if (rc != EMC_OK) return gfail(g, rc, std::string("device ") + std::to_string(g->ctx[i]->device) + ": " + emc_last_error(g->ctx[i]));
            words = w;
        }
        if (!blk[0] || words <= 0) continue;
        struct Part { int64_t off, cnt; int type, op; };
        std::vector<Part> parts;
        if (stage == 0) parts = { { 0, ST_SUM_COUNT, NCCL_FLOAT64, NCCL_SUM }, { ST_SUM_COUNT, ST_MM_COUNT, NCCL_FLOAT64, NCCL_MIN },
                                  { ST_SUM_COUNT + ST_MM_COUNT, ST_MM_COUNT, NCCL_FLOAT64, NCCL_MAX } };
        else if (stage == 1) parts = { { 0, words, NCCL_FLOAT64, NCCL_SUM } };
        else parts = { { 0, words, NCCL_UINT64, NCCL_SUM } };
        int rc = g_nccl.GroupStart();
        for (const Part &p : parts)
            for (size_t i = 0; i < D && rc == 0; ++i) {
                char *b = static_cast<char *>(blk[i]) + 8 * p.off;
                rc = g_nccl.AllReduce(b, b, (size_t)p.cnt, p.type, p.op, g->comm[i], g->ctx[i]->stream);
            }
        const int rc2 = g_nccl.GroupEnd();
        if (rc != 0 || rc2 != 0) return gfail(g, EMC_ERR_CUDA, std::string("ncclAllReduce: ") + g_nccl.GetErrorString(rc != 0 ? rc : rc2));
    }
    return EMC_OK;
}

/* counters of the last run: sums over the devices, times as the maximum */
EMC_EXPORT int emc_group_get_counters(const emc_group *g, emc_counters *c)
{
    if (!g || !c) return EMC_ERR_INVALID;
    memset(c, 0, sizeof *c);
    for (emc_ctx *x : g->ctx) {
        const emc_counters &k = x->counters;
        c->rk4_steps += k.rk4_steps; c->replay_steps += k.replay_steps; c->rail_steps += k.rail_steps; c->refills += k.refills;
        c->kernel_launches += k.kernel_launches; c->tape_rows += k.tape_rows; c->handovers += k.handovers;
        c->parked += k.parked; c->strict_steps += k.strict_steps; c->yielded += k.yielded;
        if (k.rail_ms > c->rail_ms) c->rail_ms = k.rail_ms;
        if (k.flight_ms > c->flight_ms) c->flight_ms = k.flight_ms;
        if (k.strict_ms > c->strict_ms) c->strict_ms = k.strict_ms;
    }
    return EMC_OK;
}
