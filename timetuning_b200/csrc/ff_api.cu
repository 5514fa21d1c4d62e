// C ABI of the Feature-Forwarding pipeline (include/timet_b200.h).
#include "common.cuh"

namespace timet {

extern thread_local cudaEvent_t g_ev_nominate_begin, g_ev_nominate_end;

__global__ void ff_export_kernel(const float *sel_w, const int32_t *sel_k, const int32_t *sel_cnt, int64_t q0, int N,
                                 int kw, float *w_out, int32_t *k_out, int32_t *c_out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < N * kw) {
        w_out[idx] = sel_w[q0 * kw + idx];
        k_out[idx] = sel_k[q0 * kw + idx];
    }
    if (idx < N) c_out[idx] = sel_cnt[q0 + idx];
}

static int check_ws(const timet_ff_params *p, const void *ws, size_t bytes, FFLayout *L) {
    int rc = ff_validate(p);
    if (rc != TIMET_OK) return rc;
    TIMET_CHECK_ARG(ws != nullptr, "ff: workspace is NULL");
    TIMET_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 1023) == 0, "ff: workspace must be 1024-byte aligned");
    *L = ff_layout(*p);
    if (bytes < L->total) {
        set_error("ff: workspace %zu < %zu bytes", bytes, L->total);
        return TIMET_ERR_WORKSPACE;
    }
    return TIMET_OK;
}

}  // namespace timet

using namespace timet;

extern "C" {

size_t timet_ff_workspace_bytes(const timet_ff_params *p) {
    if (ff_validate(p) != TIMET_OK) return 0;
    return ff_layout(*p).total;
}

int timet_ff_slots(const timet_ff_params *p) {
    if (ff_validate(p) != TIMET_OK) return TIMET_ERR_INVALID;
    return ff_layout(*p).kw;
}

int timet_ff_tc_supported(const timet_ff_params *p) {
    if (ff_validate(p) != TIMET_OK) return 0;
    return ff_tc_supported(*p) ? 1 : 0;
}

double timet_ff_tc_executed_flops(const timet_ff_params *p) {
    if (ff_validate(p) != TIMET_OK) return 0.0;
    return ff_tc_executed_flops(*p);
}

int timet_ff_tc_plan(const timet_ff_params *p, int32_t *plan) {
    if (ff_validate(p) != TIMET_OK) return TIMET_ERR_INVALID;
    TIMET_CHECK_ARG(plan != nullptr, "ff_tc_plan: plan is NULL");
    return ff_tc_plan(*p, plan);
}

int timet_ff_prepare(const timet_ff_params *p, const float *feats, void *workspace, size_t workspace_bytes,
                     timet_stream_t stream) {
    FFLayout L;
    int rc = check_ws(p, workspace, workspace_bytes, &L);
    if (rc != TIMET_OK) return rc;
    TIMET_CHECK_ARG(feats != nullptr, "ff_prepare: feats is NULL");
    return ff_prepare_launch(*p, L, feats, (char *)workspace, (cudaStream_t)stream);
}

int timet_ff_select(const timet_ff_params *p, int engine, const float *feats, void *workspace, size_t workspace_bytes,
                    timet_stream_t stream) {
    FFLayout L;
    int rc = check_ws(p, workspace, workspace_bytes, &L);
    if (rc != TIMET_OK) return rc;
    TIMET_CHECK_ARG(feats != nullptr, "ff_select: feats is NULL (the exact re-evaluation reads the feature rows in place)");
    TIMET_CHECK_ARG(engine == TIMET_FF_EXACT || engine == TIMET_FF_TC || engine == TIMET_FF_AUTO, "ff_select: bad engine %d", engine);
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)workspace;
    // diagnostics counters + the header of the re-do region (re-do count, work counter, wide-pool fill) are adjacent: one memset
    TIMET_CUDA(cudaMemsetAsync(ws + L.off_stats, 0, (L.off_redo - L.off_stats) + 256, st));
    if (engine == TIMET_FF_AUTO) engine = ff_tc_supported(*p) ? TIMET_FF_TC : TIMET_FF_EXACT;
    if (engine == TIMET_FF_TC) {
        if (!ff_tc_supported(*p)) {
            set_error("ff_select: tensor-core engine does not support this shape (grid %dx%d dim %d radius %d n_last %d)",
                      p->grid_h, p->grid_w, p->dim, p->radius, p->n_last_frames);
            return TIMET_ERR_UNSUPPORTED;
        }
        return ff_select_tc_launch(*p, L, feats, ws, st);
    }
    return ff_select_exact_launch(*p, L, feats, ws, st);
}

int timet_ff_gather(const timet_ff_params *p, float *labels, int64_t *hard, const void *workspace,
                    size_t workspace_bytes, timet_stream_t stream) {
    FFLayout L;
    int rc = check_ws(p, workspace, workspace_bytes, &L);
    if (rc != TIMET_OK) return rc;
    TIMET_CHECK_ARG(labels != nullptr, "ff_gather: labels is NULL");
    return ff_gather_launch(*p, L, labels, hard, (const char *)workspace, (cudaStream_t)stream);
}

int timet_ff_propagate(const timet_ff_params *p, int engine, const float *feats, float *labels, int64_t *hard,
                       void *workspace, size_t workspace_bytes, timet_stream_t stream) {
    int rc = timet_ff_prepare(p, feats, workspace, workspace_bytes, stream);
    if (rc != TIMET_OK) return rc;
    rc = timet_ff_select(p, engine, feats, workspace, workspace_bytes, stream);
    if (rc != TIMET_OK) return rc;
    return timet_ff_gather(p, labels, hard, workspace, workspace_bytes, stream);
}

int timet_ff_stats(const timet_ff_params *p, const void *workspace, size_t workspace_bytes, int64_t *out,
                   timet_stream_t stream) {
    FFLayout L;
    int rc = check_ws(p, workspace, workspace_bytes, &L);
    if (rc != TIMET_OK) return rc;
    TIMET_CHECK_ARG(out != nullptr, "ff_stats: out is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    TIMET_CUDA(cudaMemcpyAsync(out, (const char *)workspace + L.off_stats, 8 * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    const int64_t q = L.queries;
    TIMET_CUDA(cudaMemcpyAsync(out, &q, sizeof(int64_t), cudaMemcpyHostToDevice, st));
    TIMET_CUDA(cudaStreamSynchronize(st));   // `q` lives on this stack frame
    return TIMET_OK;
}

int timet_ff_export_selection(const timet_ff_params *p, const void *workspace, size_t workspace_bytes, int clip,
                              int t, float *weights, int32_t *keys, int32_t *counts, timet_stream_t stream) {
    FFLayout L;
    int rc = check_ws(p, workspace, workspace_bytes, &L);
    if (rc != TIMET_OK) return rc;
    TIMET_CHECK_ARG(weights && keys && counts, "ff_export_selection: NULL output");
    TIMET_CHECK_ARG(clip >= 0 && clip < p->n_clips && t >= p->t_begin && t < p->n_frames, "ff_export_selection: bad (clip=%d, t=%d)", clip, t);
    const char *ws = (const char *)workspace;
    const int64_t q0 = ((int64_t)clip * L.nT + (t - p->t_begin)) * L.N;
    const int n = L.N * L.kw;
    ff_export_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float *>(ws + L.off_sel_w), reinterpret_cast<const int32_t *>(ws + L.off_sel_k),
        reinterpret_cast<const int32_t *>(ws + L.off_sel_cnt), q0, L.N, L.kw, weights, keys, counts);
    TIMET_LAUNCHED();
    return TIMET_OK;
}


int timet_ff_export_wide(const timet_ff_params *p, const void *workspace, size_t workspace_bytes, int64_t offset, int n,
                         float *weights, int32_t *keys, timet_stream_t stream) {
    FFLayout L;
    int rc = check_ws(p, workspace, workspace_bytes, &L);
    if (rc != TIMET_OK) return rc;
    TIMET_CHECK_ARG(weights && keys, "ff_export_wide: NULL output");
    TIMET_CHECK_ARG(offset >= 0 && n >= 0 && offset + n <= L.wide_cap, "ff_export_wide: [%lld, +%d) outside the pool of %lld entries",
                    (long long)offset, n, (long long)L.wide_cap);
    const char *ws = (const char *)workspace;
    cudaStream_t st = (cudaStream_t)stream;
    TIMET_CUDA(cudaMemcpyAsync(weights, ws + L.off_wide_w + (size_t)offset * sizeof(float), (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TIMET_CUDA(cudaMemcpyAsync(keys, ws + L.off_wide_k + (size_t)offset * sizeof(int32_t), (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    return TIMET_OK;
}

int timet_debug_tc_tile(const timet_ff_params *p, void *workspace, size_t workspace_bytes, int64_t tile_id, float *dump,
                        timet_stream_t stream) {
    FFLayout L;
    int rc = check_ws(p, workspace, workspace_bytes, &L);
    if (rc != TIMET_OK) return rc;
    TIMET_CHECK_ARG(dump != nullptr, "debug_tc_tile: dump is NULL");
    return ff_tc_debug_tile(*p, L, (char *)workspace, tile_id, dump, (cudaStream_t)stream);
}


int timet_debug_tc_trace(const timet_ff_params *p, const void *workspace, size_t workspace_bytes, uint64_t *out, int n_ctas,
                         timet_stream_t stream) {
    FFLayout L;
    int rc = check_ws(p, workspace, workspace_bytes, &L);
    if (rc != TIMET_OK) return rc;
    TIMET_CHECK_ARG(out != nullptr, "debug_tc_trace: out is NULL");
    return ff_tc_debug_trace(*p, L, (const char *)workspace, (unsigned long long *)out, n_ctas, (cudaStream_t)stream);
}


int timet_ff_select_timed(const timet_ff_params *p, int engine, const float *feats, void *workspace, size_t workspace_bytes,
                          timet_stream_t stream, void *event_before_nominate, void *event_after_nominate) {
    g_ev_nominate_begin = (cudaEvent_t)event_before_nominate;
    g_ev_nominate_end = (cudaEvent_t)event_after_nominate;
    const int rc = timet_ff_select(p, engine, feats, workspace, workspace_bytes, stream);
    g_ev_nominate_begin = nullptr;
    g_ev_nominate_end = nullptr;
    return rc;
}

}
