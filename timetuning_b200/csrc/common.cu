// Error state, launch counter, parameter validation, ABI info.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace timet {

static thread_local char g_error[512] = "";
static int64_t g_launches = 0;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int64_t &launch_counter() { return g_launches; }

int num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

static EnvCfg g_env;
static bool g_env_loaded = false;

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

void env_reload() {
    EnvCfg e;
    e.tc_pflags = env_int("TIMET_TC_PFLAGS", 0);
    e.tc_flags = env_int("TIMET_TC_FLAGS", 0);
    e.tc_stages = env_int("TIMET_TC_STAGES", 0);
    e.tc_nbuf = env_int("TIMET_TC_NBUF", 0);
    e.tc_clip_group = env_int("TIMET_TC_CLIP_GROUP", 0);
    e.tc_persist = env_int("TIMET_TC_PERSIST", 1) != 0;
    e.tc_dyn = env_int("TIMET_TC_DYN", 1) != 0;
    e.tc_trace = env_int("TIMET_TC_TRACE", 0) == 1;
    e.sk_streaming = env_int("TIMET_SK_STREAMING", 0) == 1;
    e.sk_no_dual = env_int("TIMET_SK_DUAL", 1) == 0;     // 0: the two calls of timet_sinkhorn_pair one after the other
    e.sk_ll = env_int("TIMET_SK_LL", 1) != 0 ? 1 : 0;
    e.sk_ustride = env_int("TIMET_SK_USTRIDE", 0);
    e.gather_batch = env_int("TIMET_GATHER_BATCH", 0);
    e.gather_l1 = env_int("TIMET_GATHER_L1", 1) != 0;
    e.fin_staged = env_int("TIMET_FIN_STAGED", 1) != 0;
    e.sc_stages = env_int("TIMET_SC_STAGES", 0);
    const char *to = getenv("TIMET_P2P_TIMEOUT_S");
    e.p2p_timeout_s = (to && atof(to) > 0.0) ? atof(to) : 600.0;    // NCCL-like patience: rank skew of minutes is legal
    g_env = e;
    g_env_loaded = true;
}

const EnvCfg &env_cfg() {
    if (!g_env_loaded) env_reload();
    return g_env;
}

int ff_validate(const timet_ff_params *p) {
    TIMET_CHECK_ARG(p != nullptr, "ff: params is NULL");
    TIMET_CHECK_ARG(p->n_clips >= 1, "ff: n_clips=%d must be >= 1", p->n_clips);
    TIMET_CHECK_ARG(p->n_frames >= 2, "ff: n_frames=%d must be >= 2", p->n_frames);
    TIMET_CHECK_ARG(p->grid_h >= 1 && p->grid_w >= 1, "ff: bad grid %dx%d", p->grid_h, p->grid_w);
    TIMET_CHECK_ARG((int64_t)p->grid_h * p->grid_w <= 65536, "ff: grid %dx%d too large", p->grid_h, p->grid_w);
    TIMET_CHECK_ARG(p->dim >= 1 && p->dim <= 4096, "ff: dim=%d out of range", p->dim);
    TIMET_CHECK_ARG(p->n_channels >= 1, "ff: n_channels=%d must be >= 1", p->n_channels);
    // queue.Queue(0) is unbounded and the reference then blocks on get() (mask_propagation.py:460,488)
    TIMET_CHECK_ARG(p->n_last_frames >= 1, "ff: n_last_frames=%d must be >= 1", p->n_last_frames);
    TIMET_CHECK_ARG(p->radius >= 0, "ff: radius=%d must be >= 0", p->radius);
    TIMET_CHECK_ARG(p->topk >= 1 && p->topk <= 16, "ff: topk=%d must be in 1..16", p->topk);
    TIMET_CHECK_ARG(p->t_begin >= 1 && p->t_begin < p->n_frames, "ff: t_begin=%d must be in 1..n_frames-1", p->t_begin);
    TIMET_CHECK_ARG(p->temperature > 0.f, "ff: temperature must be > 0");
    TIMET_CHECK_ARG((int64_t)p->n_frames * p->grid_h * p->grid_w < (1ll << 31), "ff: n_frames*N overflows int32 keys");
    return TIMET_OK;
}

}  // namespace timet

extern "C" {

const char *timet_last_error(void) { return timet::g_error; }
int timet_abi_version(void) { return TIMET_ABI_VERSION; }
int64_t timet_launch_count(void) { return timet::g_launches; }
int timet_debug_reload_env(void) { timet::env_reload(); return TIMET_OK; }

}
