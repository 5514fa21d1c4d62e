// Shared host/device helpers of libtimet_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/timet_b200.h"

namespace timet {

// ------------------------------------------------------------------ errors
void set_error(const char *fmt, ...);
int64_t &launch_counter();

#define TIMET_CHECK_ARG(cond, ...)                 \
    do {                                           \
        if (!(cond)) {                             \
            ::timet::set_error(__VA_ARGS__);       \
            return TIMET_ERR_INVALID;              \
        }                                          \
    } while (0)

#define TIMET_CUDA(call)                                                                        \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            ::timet::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return TIMET_ERR_CUDA;                                                              \
        }                                                                                       \
    } while (0)

// count + check a kernel launch
#define TIMET_LAUNCHED()                                                                        \
    do {                                                                                        \
        ::timet::launch_counter()++;                                                            \
        cudaError_t e__ = cudaGetLastError();                                                   \
        if (e__ != cudaSuccess) {                                                               \
            ::timet::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
            return TIMET_ERR_CUDA;                                                              \
        }                                                                                       \
    } while (0)

int num_sms();

// Experiment / debug switches, read from the environment ONCE (first use) -- never on the launch path.
// timet_debug_reload_env() re-reads them (tests flip switches inside one process).
struct EnvCfg {
    int tc_pflags, tc_flags, tc_stages, tc_nbuf, tc_clip_group;   // 0 = default
    bool tc_persist, tc_dyn, tc_trace;
    bool sk_streaming, sk_no_dual;
    int sk_ll;                                                     // 1 (default): tagged-word NVLink exchange; 0: data + flag
    int sk_ustride;                                                // 0 = default
    int sc_stages;                                                 // cosine-scores GEMM ring depth (2 = two CTAs per SM, default; 4)
    int gather_batch;                                              // gather: label rows in flight per batch (0 = by topk)
    bool fin_staged;                                               // finalize: rows staged in shared memory with cp.async (default, dim <= 384)
    bool gather_l1;                                                // gather: label rows through L1 where that is safe (default)
    double p2p_timeout_s;                                          // peer-exchange / marginal wait time-out (seconds)
};
const EnvCfg &env_cfg();
void env_reload();

// ------------------------------------------------------------------ peer-memory exchange buffer (one per rank)
constexpr int P2P_MAX_RANKS = 16;
constexpr int P2P_MAX_K = 1024;
struct P2PBuf {
    unsigned long long flag[P2P_MAX_RANKS];          // flag[src] = number of exchanges src has published here
    unsigned long long pad[16];
    float slot[3][P2P_MAX_RANKS][P2P_MAX_K];         // slot[e % 3][src] = src's K-vector of exchange e
    // low-latency form: one 8-byte word per value = (float bits, exchange number + 1).  The tag travels WITH the data in one
    // single-copy-atomic store, so no separate flag + release fence (one NVLink latency per exchange instead of two).
    unsigned long long ll[3][P2P_MAX_RANKS][P2P_MAX_K];
};
// *epoch points at the two per-channel exchange counters (every rank's buffer is P2PBuf[2])
bool comm_p2p_info(timet_comm_t comm, void ***peers_dev, int *rank, int *ws, unsigned long long **epoch);

// ------------------------------------------------------------------ FF workspace layout
constexpr int FF_CAND_CAP = 32;     // in-kernel candidate list capacity per (query, epilogue group)
constexpr int FF_CAND_STORE = 16;   // candidates published per query after the merged final compaction
constexpr int FF_TRACE_CTAS = 4096;  // debug timeline (env TIMET_TC_TRACE): first CTAs x 8 globaltimer stamps
constexpr int FF_TRACE_SLOTS = 8;
// header of the redo region (256 bytes, zeroed by timet_ff_select): byte offsets of its counters
constexpr int FF_HDR_REDO_COUNT = 0;     // u32: queries queued for the exact re-do
constexpr int FF_HDR_NEXT_ITEM = 128;    // u32: work counter of the persistent tensor-core kernel
constexpr int FF_HDR_WIDE_USED = 192;    // u64: entries allocated in the wide-row pool

struct FFLayout {
    int N, Dp, nT, kw;               // patches, padded dim (multiple of 64), target frames, slots per query
    int64_t rows;                    // n_clips * n_frames * N feature rows
    int64_t queries;                 // n_clips * nT * N
    size_t off_xpad, off_inv, off_fn16, off_sel_w, off_sel_k, off_sel_cnt, off_cand, off_cand_meta, off_stats, off_redo, off_trace;
    size_t off_wide_w, off_wide_k;   // overflow pool for rows with more than kw kept entries (exact tie sets)
    int64_t wide_cap;                // pool capacity in entries
    size_t total;
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static inline FFLayout ff_layout(const timet_ff_params &p) {
    FFLayout L;
    L.N = p.grid_h * p.grid_w;
    L.Dp = (p.dim + 63) / 64 * 64;
    L.nT = p.n_frames - p.t_begin;
    L.kw = p.topk <= 8 ? 16 : 32;
    L.rows = (int64_t)p.n_clips * p.n_frames * L.N;
    L.queries = (int64_t)p.n_clips * L.nT * L.N;
    size_t o = 0;
    // +256 rows of slack: TMA boxes of the last query/key tile may run past the last row
    // fp32 operands of the exact re-evaluation: the CALLER's feature rows are read in place (no normalised fp32 copy);
    // one inverse norm per row.  Only when dim % 4 != 0 (rows not float4-addressable) a zero-padded copy is kept.
    L.off_xpad = o; o = align_up(o + ((p.dim & 3) ? (size_t)L.rows * L.Dp * sizeof(float) : 0), 1024);
    L.off_inv = o; o = align_up(o + (size_t)L.rows * sizeof(float), 1024);
    L.off_fn16 = o; o = align_up(o + (size_t)(L.rows + 256) * L.Dp * sizeof(__half), 1024);
    L.off_sel_w = o; o = align_up(o + (size_t)L.queries * L.kw * sizeof(float), 1024);
    L.off_sel_k = o; o = align_up(o + (size_t)L.queries * L.kw * sizeof(int32_t), 1024);
    L.off_sel_cnt = o; o = align_up(o + (size_t)L.queries * sizeof(int32_t), 1024);
    L.off_cand = o; o = align_up(o + (size_t)L.queries * FF_CAND_STORE * sizeof(uint32_t), 1024);
    L.off_cand_meta = o; o = align_up(o + (size_t)L.queries * sizeof(uint32_t), 1024);
    L.off_stats = o; o = align_up(o + 8 * sizeof(int64_t), 1024);
    L.off_redo = o; o = align_up(o + 256 + (size_t)L.queries * sizeof(int32_t), 1024);   // count header + query ids
    L.off_trace = o; o = align_up(o + (size_t)FF_TRACE_CTAS * FF_TRACE_SLOTS * sizeof(unsigned long long), 1024);
    // wide rows: the reference keeps EVERY key tied with the k-th affinity (mask_propagation.py:432-436); rows with more
    // than kw survivors go to this pool (variable length, allocated with one atomic per row)
    L.wide_cap = L.queries * 4 < 65536 ? 65536 : (L.queries * 4 > (int64_t)(1 << 24) ? (int64_t)(1 << 24) : L.queries * 4);
    L.off_wide_w = o; o = align_up(o + (size_t)L.wide_cap * sizeof(float), 1024);
    L.off_wide_k = o; o = align_up(o + (size_t)L.wide_cap * sizeof(int32_t), 1024);
    L.total = o;
    return L;
}

int ff_validate(const timet_ff_params *p);

// fp32 view of the feature rows for the exact (canonical) similarity: row r = x + r * ld, its first n4 float4 are
// meaningful; inv[r] = 1 / max(||row r||, 1e-12)
struct FFSrc {
    const float *x;
    const float *inv;
    int ld, n4;
};
static inline FFSrc ff_src(const timet_ff_params &p, const FFLayout &L, const float *feats, const char *ws) {
    FFSrc S;
    S.inv = reinterpret_cast<const float *>(ws + L.off_inv);
    if (p.dim & 3) { S.x = reinterpret_cast<const float *>(ws + L.off_xpad); S.ld = L.Dp; S.n4 = L.Dp >> 2; }
    else { S.x = feats; S.ld = p.dim; S.n4 = p.dim >> 2; }
    return S;
}

// stage launchers (defined in the per-stage .cu files)
int ff_prepare_launch(const timet_ff_params &p, const FFLayout &L, const float *feats, char *ws, cudaStream_t st);
int ff_select_exact_launch(const timet_ff_params &p, const FFLayout &L, const float *feats, char *ws, cudaStream_t st);
int ff_select_tc_launch(const timet_ff_params &p, const FFLayout &L, const float *feats, char *ws, cudaStream_t st);
int ff_gather_launch(const timet_ff_params &p, const FFLayout &L, float *labels, int64_t *hard, const char *ws,
                     cudaStream_t st);
bool ff_tc_supported(const timet_ff_params &p);
double ff_tc_executed_flops(const timet_ff_params &p);
int ff_tc_plan(const timet_ff_params &p, int32_t *out);
int ff_tc_debug_trace(const timet_ff_params &p, const FFLayout &L, const char *ws, unsigned long long *out, int n_ctas, cudaStream_t st);
int ff_tc_debug_tile(const timet_ff_params &p, const FFLayout &L, char *ws, int64_t tile_id, float *dump, cudaStream_t st);

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Canonical fp32 similarity of two feature rows (n4 float4 elements each, rows 16-byte aligned).
// DEFINITION (every engine reproduces it bit for bit, so the tensor-core engine and the exact scan agree):
//   sim(q, k) = (dot(x_q, x_k) * inv_q) * inv_k        inv = 1 / max(||x||_2, 1e-12)   (F.normalize, :418-419)
// on the UN-normalised rows -- no normalised fp32 copy of the features exists (it would cost a 4*D-byte write and
// re-read per row); the reference normalises first and multiplies after, which differs by a few ulp.  dot:
//   partial[l], l = 0..31 : one fmaf chain over the float4 elements i = l, l+32, l+64, ... in that order,
//                           components x, y, z, w in that order;
//   result                : xor-butterfly tree  v[l] += v[l ^ 16]; v[l] += v[l ^ 8]; ... ; v[l] += v[l ^ 1].
// dot_canonical_warp: the 32 lanes of a warp each own one partial (coalesced 512-byte row segments).
// dot_canonical_seq : one thread emulates all 32 partials and the same tree (lane-per-key scans).
__device__ __forceinline__ float fma4_chain(float acc, const float4 x, const float4 y) {
    acc = fmaf(x.x, y.x, acc);
    acc = fmaf(x.y, y.y, acc);
    acc = fmaf(x.z, y.z, acc);
    return fmaf(x.w, y.w, acc);
}
__device__ __forceinline__ float dot_canonical_warp(const float4 *__restrict__ q, const float4 *__restrict__ k, int n4, int lane) {
    float acc = 0.f;
    for (int i = lane; i < n4; i += 32) acc = fma4_chain(acc, q[i], __ldg(k + i));
    return warp_sum(acc);
}
__device__ __forceinline__ float dot_canonical_seq(const float4 *__restrict__ q, const float4 *__restrict__ k, int n4) {
    float acc[32];
#pragma unroll
    for (int l = 0; l < 32; ++l) acc[l] = 0.f;
    for (int base = 0; base < n4; base += 32) {
#pragma unroll
        for (int l = 0; l < 32; ++l)
            if (base + l < n4) acc[l] = fma4_chain(acc[l], q[base + l], __ldg(k + base + l));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int l = 0; l < o; ++l) acc[l] += acc[l + o];
    }
    return acc[0];
}

__device__ __forceinline__ float sim_from_dot(float dot, float inv_q, float inv_k) { return __fmul_rn(__fmul_rn(dot, inv_q), inv_k); }

// exp(sim / T) exactly as mask_propagation.py:422 evaluates it in float32
__device__ __forceinline__ float affinity_from_sim(float sim, float temperature) {
    return expf(__fdiv_rn(sim, temperature));
}

// First context frame after frame 0 for target t (mask_propagation.py:482-493): the FIFO holds
// the n_last most recent targets, so contexts are {0} U {max(1, t - n_last) .. t-1}.
__device__ __host__ __forceinline__ int ctx_lo(int t, int n_last) { return (t - n_last > 1) ? t - n_last : 1; }
__device__ __host__ __forceinline__ int ctx_count(int t, int n_last) { return 1 + (t - ctx_lo(t, n_last)); }
__device__ __host__ __forceinline__ int ctx_frame(int t, int n_last, int ci) {
    return ci == 0 ? 0 : ctx_lo(t, n_last) + ci - 1;
}
#endif

}  // namespace timet
