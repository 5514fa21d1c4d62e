/*
 * timet_b200.h — C ABI of the B200-native Feature-Forwarding + Sinkhorn-Knopp library.
 *
 * This is the drop-in boundary for the ONE hot path of SMSD75/Timetuning that this
 * repository replaces (BASELINE.json north_star).  The reference has no
 * FFI layer: its boundary is a set of Python callables.  Each entry point
 * below names the reference callable it stands behind (file:line in /root/reference);
 * timetuning_b200/ops.py holds the Python shims with the reference signatures and
 * INTEGRATION.md shows the binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain C: pointers + sizes only, no torch / C++ types, no exceptions across the ABI;
 *   - every function returns 0 on success, <0 on error; timet_last_error() (thread-local)
 *     holds the message;
 *   - all data pointers are DEVICE pointers on the current CUDA device, allocated and owned
 *     by the caller (inputs, outputs and workspace; sizes from the *_workspace_bytes query);
 *   - every launch goes to the cudaStream_t passed as `stream` (timet_stream_t == cudaStream_t);
 *     no call synchronises the device;
 *   - there is NO CPU fallback: without a CUDA device every compute entry returns an error.
 */
#ifndef TIMET_B200_H
#define TIMET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TIMET_ABI_VERSION 2

typedef void *timet_stream_t; /* cudaStream_t */
typedef void *timet_comm_t;   /* opaque: owns an ncclComm_t */

/* ------------------------------------------------------------------ errors / info */
#define TIMET_OK 0
#define TIMET_ERR_INVALID (-1)     /* bad argument                                  */
#define TIMET_ERR_CUDA (-2)        /* CUDA runtime / driver error                   */
#define TIMET_ERR_UNSUPPORTED (-3) /* shape outside what a kernel variant supports  */
#define TIMET_ERR_NCCL (-4)        /* NCCL error or libnccl not loadable            */
#define TIMET_ERR_WORKSPACE (-5)   /* workspace too small                           */

const char *timet_last_error(void);
int timet_abi_version(void);
/* number of kernels this library has launched in this process (bench.py "gpu_launches") */
int64_t timet_launch_count(void);
/* The experiment / debug switches (environment variables TIMET_*) are read once, at first use; this
 * re-reads them (tests flip switches inside one process).  Not needed in production. */
int timet_debug_reload_env(void);

/* ------------------------------------------------------------------ Sinkhorn-Knopp
 * Stands behind  my_utils.sinkhorn(Q, nmb_iters, world_size)      my_utils.py:246-274
 * and            TimeT.find_optimal_assignment(scores, eps, iters) time_tuning.py:157-168.
 *
 * `in` is row-major [B, K] (B local samples, K prototypes) — the physical layout of the
 * K x B transposed view the reference passes (time_tuning.py:164).
 *   input_kind TIMET_SK_EXP    : in = exp(scores/eps) (what sinkhorn() receives); epsilon ignored
 *   input_kind TIMET_SK_SCORES : in = cosine scores; exp(in/epsilon) is fused into every pass
 * q_out is row-major [B, K] float32, rows sum to 1 (my_utils.py:274).
 * world_size > 1: `comm` must come from timet_comm_init; the K-vector of prototype marginals
 * is all-reduced once per iteration on `stream` (my_utils.py:259-272); the sample marginal
 * uses B * world_size (my_utils.py:257), i.e. equal B on every rank like the reference.
 */
#define TIMET_SK_EXP 0
#define TIMET_SK_SCORES 1
size_t timet_sinkhorn_workspace_bytes(int64_t B, int K);
int timet_sinkhorn(const float *in, int64_t B, int K, int input_kind, float epsilon, int iters,
                   int world_size, timet_comm_t comm, float *q_out, void *workspace,
                   size_t workspace_bytes, timet_stream_t stream);

/* Extended form.  opts may be NULL (= timet_sinkhorn).
 *   out_block_rows / out_block_stride: row j of Q is written at q_out + (j / out_block_rows) * out_block_stride +
 *     (j % out_block_rows) * K floats (out_block_rows <= 0: contiguous).  With out_block_rows = N and out_block_stride =
 *     n_frames * N * K the assignment of clip b lands in frame 0 of the channel-last label tensor of
 *     timet_ff_propagate -- TimeT.make_seg_maps' reshape + the first_seg copy (time_tuning.py:144-147) disappear.
 *   share_sm: run the resident kernel with 512 instead of 1024 threads per CTA so that kernels launched on another
 *     stream (the HBM/L2-bound Feature-Forwarding stages) can co-reside on the SMs; the call is latency-bound. */
typedef struct timet_sinkhorn_opts {
    int64_t out_block_rows;
    int64_t out_block_stride;
    int32_t share_sm;
    int32_t reserved;
} timet_sinkhorn_opts;
int timet_sinkhorn_ex(const float *in, int64_t B, int K, int input_kind, float epsilon, int iters, int world_size,
                      timet_comm_t comm, float *q_out, const timet_sinkhorn_opts *opts, void *workspace,
                      size_t workspace_bytes, timet_stream_t stream);
/* Two problems of the same shape (B, K, kind, eps, iters) in ONE launch: the source and the target assignment of a
 * training step (time_tuning.py:268,275).  A call is bound by the latency of its dependent grid-wide reductions, so the
 * two problems run SIDE BY SIDE on half of the SMs each (own reduction counters, own NVLink exchange channel) and overlap
 * completely; rows that do not fit half of the shared memory are re-read from L2.  Same maths as timet_sinkhorn_ex on
 * another row partition (equal within fp32 summation order, bit-reproducible).  workspace: 2 x
 * timet_sinkhorn_workspace_bytes(B, K).  Falls back to two sequential calls where it does not apply. */
int timet_sinkhorn_pair(const float *in0, const float *in1, int64_t B, int K, int input_kind, float epsilon, int iters,
                        int world_size, timet_comm_t comm, float *q0, const timet_sinkhorn_opts *opts0, float *q1,
                        const timet_sinkhorn_opts *opts1, void *workspace, size_t workspace_bytes, timet_stream_t stream);
/* How timet_sinkhorn_pair runs this shape: 1 = DUAL (one launch, the two problems side by side on half of the SMs each:
 * the default), 0 = two timet_sinkhorn_ex calls one after the other (TIMET_SK_DUAL=0, K % 4 != 0, K > 512) */
int timet_sinkhorn_pair_mode(int64_t B, int K);
/* How a call of this shape runs: 1 = ONE resident kernel (all rows of exp(S/eps) fit the SMs' shared memory);
 * 2 = ONE hybrid kernel (as many rows resident as fit, the rest re-read and re-exponentiated every iteration -- e.g.
 * BASELINE configs[2] at 2 / 4 GPUs); 0 = one streaming launch per pass (K % 4 != 0, K > 512, TIMET_SK_STREAMING=1). */
int timet_sinkhorn_resident(int64_t B, int K);

/* ------------------------------------------------------------------ cosine scores
 * No-grad branch of TimeT.get_feature_prototype_similarity (time_tuning.py:130-141):
 *   scores[B, K] = F.normalize(x[B, dh], dim=-1) @ prototypes[K, dh]^T      (float32 in / out)
 * computed on the tensor cores with an fp16 hi/lo split (3 products, <= 2^-22 relative error), because the
 * scores feed exp(s / eps).  The student branch that needs autograd stays in PyTorch. */
size_t timet_cosine_scores_workspace_bytes(int64_t B, int K, int dh);
int timet_cosine_scores(const float *x, const float *prototypes, int64_t B, int K, int dh, float *scores_out,
                        void *workspace, size_t workspace_bytes, timet_stream_t stream);
/* Several feature blocks of rows_each rows against the same prototypes in ONE GEMM launch (get_loss scores the source
 * and the target frame, time_tuning.py:268,275): scores_out is [n_x * rows_each, K], block i at row i * rows_each;
 * n_x <= 4; workspace sized for B = n_x * rows_each. */
int timet_cosine_scores_multi(const float *const *x_list, int n_x, int64_t rows_each, const float *prototypes, int K, int dh,
                              float *scores_out, void *workspace, size_t workspace_bytes, timet_stream_t stream);

/* ------------------------------------------------------------------ Feature-Forwarding
 * Stands behind  label_propagation(...)  mask_propagation.py:396-445   (one target frame)
 *                propagate_labels(...)   mask_propagation.py:448-496   (frame loop + FIFO)
 *                TimeT.make_seg_maps     time_tuning.py:143-154 and the per-clip loop :277-296.
 *
 * Layout in HBM (all row-major, contiguous):
 *   feats   float32 [n_clips, n_frames, N, dim]         N = grid_h * grid_w, un-normalised
 *   labels  float32 [n_clips, n_frames, N, n_channels]  channel-last soft labels; the caller
 *           fills frame 0 (first_seg, == Sinkhorn Q rows in training); frames
 *           t_begin..n_frames-1 are written by the library
 *   hard    int64   [n_clips, N]  argmax over channels of the LAST frame (time_tuning.py:296);
 *           may be NULL
 * Context frames of target t: frame 0 plus the n_last_frames previous targets
 * (mask_propagation.py:460,482-493).  Frames < t_begin are contexts only (their labels are
 * caller-provided): t_begin = 1 reproduces propagate_labels; t_begin = n_frames-1 with
 * n_last_frames >= n_frames-2 reproduces one label_propagation call on an explicit context list.
 */
typedef struct timet_ff_params {
    int32_t n_clips;
    int32_t n_frames;
    int32_t grid_h, grid_w;   /* patch grid; the reference only supports h == w (:407) */
    int32_t dim;              /* feature dim D */
    int32_t n_channels;       /* C label channels */
    int32_t n_last_frames;    /* FIFO depth (reference default 7) */
    int32_t radius;           /* size_mask_neighborhood; 0 = no spatial restriction (:423) */
    int32_t topk;             /* 1..16 */
    int32_t t_begin;          /* first target frame, >= 1 */
    float temperature;        /* 0.1 in the reference (:422) */
    int32_t reserved;
} timet_ff_params;

/* selection engines for the affinity / top-k stage */
#define TIMET_FF_EXACT 0 /* fp32 CUDA-core scan of every in-window key                         */
#define TIMET_FF_TC 1    /* tcgen05 fp16 tensor-core nomination + exact fp32 re-evaluation      */
#define TIMET_FF_AUTO 2  /* TC when the shape is supported by the tensor-core kernel, else EXACT */

size_t timet_ff_workspace_bytes(const timet_ff_params *p);
/* 1 if the tensor-core engine supports this shape (radius 1..15, n_last_frames <= 7, ...) */
int timet_ff_tc_supported(const timet_ff_params *p);
/* FLOPs the tensor-core kernel issues for this problem (key tiles inside the window band only, M padded to the
 * 128-row MMA): the numerator of the tensor-pipe roofline in bench.py.  0 if the TC engine does not support the shape.
 * For comparison, the dense definition of SURVEY.md §8d is 2 * N^2 * dim * sum_t ctx(t) per clip. */
double timet_ff_tc_executed_flops(const timet_ff_params *p);
/* What timet_ff_select(TIMET_FF_TC / AUTO) will run for this problem (host-side query, no device work; honours the
 * TIMET_TC_* experiment switches).  plan[8] = { kernel: 0 exact engine only, 1 per-item tcgen05 kernel, 2 persistent
 * tcgen05 kernel; column-blocked query tiles (0/1); candidate-list slots per (query, group); key-ring stages; bytes of the
 * resident query tile (0: streamed); key-tile columns; grid rows per query tile; dynamic shared memory in bytes }.
 * The quantities behind the kernel's roofline in DESIGN.md §4.3; mask_propagation.py:418-436 is what they implement. */
int timet_ff_tc_plan(const timet_ff_params *p, int32_t *plan);

/* stage 1: one pass over the feature rows (F.normalize, :418-419): inverse norms + the fp16 tensor-core operand go to
 * the workspace.  No normalised fp32 copy is made: stage 2 reads `feats` IN PLACE for the exact fp32 similarity, so the
 * same buffer must be passed to timet_ff_select and stay unchanged until that has run.  feats must be 16-byte aligned. */
int timet_ff_prepare(const timet_ff_params *p, const float *feats, void *workspace, size_t workspace_bytes,
                     timet_stream_t stream);
/* stage 2: per (clip, target frame, query): window mask, exp(sim/T), global top-k over all
 * contexts with ties kept, normalised weights (:422-436) -> sparse (weight, key) lists in the workspace */
int timet_ff_select(const timet_ff_params *p, int engine, const float *feats, void *workspace, size_t workspace_bytes,
                    timet_stream_t stream);
/* timet_ff_select with two caller-created cudaEvent_t (may be NULL) recorded on `stream` immediately before and
 * after the tensor-core nomination kernel: lets a caller time the dominant kernel alone without a profiler. */
int timet_ff_select_timed(const timet_ff_params *p, int engine, const float *feats, void *workspace, size_t workspace_bytes,
                          timet_stream_t stream, void *event_before_nominate, void *event_after_nominate);
/* stage 3: frame-sequential weighted gather of the context labels (:439-444) and argmax */
int timet_ff_gather(const timet_ff_params *p, float *labels, int64_t *hard, const void *workspace,
                    size_t workspace_bytes, timet_stream_t stream);
/* stages 1-3 */
int timet_ff_propagate(const timet_ff_params *p, int engine, const float *feats, float *labels, int64_t *hard,
                       void *workspace, size_t workspace_bytes, timet_stream_t stream);

/* Diagnostics of the last timet_ff_select on this workspace (device -> 8 x int64 at `out`, device ptr):
 * [0] queries, [1] selected (weight) entries, [2] queries with more than topk entries (exact ties),
 * [3] TC candidates nominated, [4] queries re-done by the exact scan (candidate-list overflow),
 * [5] queries whose tie set was TRUNCATED to the kw slots because the wide-row pool was exhausted (0 unless more
 *     than ~4 entries per query on average are needed; the Python shims raise when it is non-zero),
 * [6] wide rows (queries with more than kw survivors, kept completely in the pool), [7] reserved.
 *
 * Ties.  The reference keeps EVERY key whose affinity equals the k-th largest (mask_propagation.py:432-436), so a
 * query can have more than topk weights -- clips with repeated frames (data_loader.py:621-623) produce exact ties
 * across contexts.  Up to kw = timet_ff_slots(p) survivors live in the regular slots; larger sets are stored as
 * variable-length "wide rows": counts[i] = -n and keys[i * kw] = offset of the row's n entries in the pool. */
int timet_ff_stats(const timet_ff_params *p, const void *workspace, size_t workspace_bytes, int64_t *out,
                   timet_stream_t stream);
/* Copy the sparse selection of (clip, t) out of the workspace for inspection / tests:
 * weights float32 [N, kw], keys int32 [N, kw] (frame * N + patch, -1 = unused), counts int32 [N];
 * kw = timet_ff_slots(p). */
int timet_ff_slots(const timet_ff_params *p);
int timet_ff_export_selection(const timet_ff_params *p, const void *workspace, size_t workspace_bytes, int clip,
                              int t, float *weights, int32_t *keys, int32_t *counts, timet_stream_t stream);

/* Copy n entries of the wide-row pool starting at `offset` (see above): weights float32 [n], keys int32 [n]. */
int timet_ff_export_wide(const timet_ff_params *p, const void *workspace, size_t workspace_bytes, int64_t offset, int n,
                         float *weights, int32_t *keys, timet_stream_t stream);

/* Test hook of the tensor-core engine: run ONE query tile (tile_id in the kernel's launch order) after
 * timet_ff_prepare and dump its raw fp32 TMEM accumulators, float32 [n_key_tiles, 128, 256]. */
int timet_debug_tc_tile(const timet_ff_params *p, void *workspace, size_t workspace_bytes, int64_t tile_id,
                        float *dump, timet_stream_t stream);

/* Test hook: with env TIMET_TC_TRACE=1 the tensor-core kernel stamps %globaltimer at 8 points of every CTA's
 * life; this copies [n_ctas, 8] uint64 ns (device -> device) for timeline analysis (profiles/tc_trace.py). */
int timet_debug_tc_trace(const timet_ff_params *p, const void *workspace, size_t workspace_bytes, uint64_t *out,
                         int n_ctas, timet_stream_t stream);

/* ------------------------------------------------------------------ small routines
 * restrict_neighborhood(h, w, s)  mask_propagation.py:377-391 -> float32 [h*w, h*w] of 0/1
 * norm_mask(mask)                 mask_propagation.py:363-374 -> per-channel min-max, [C, HW];
 *                                 dtype_bytes 4 (float32) or 8 (float64)                         */
int timet_restrict_neighborhood(int h, int w, int radius, float *mask_out, timet_stream_t stream);
int timet_norm_mask(const void *mask, void *out, int n_channels, int64_t hw, int dtype_bytes,
                    timet_stream_t stream);

/* Eval tail of the DAVIS/YTVOS propagation (mask_propagation.py:822-824):
 * F.interpolate(maps, size=(out_h, out_w), mode="bilinear", align_corners=False) followed by max over channels,
 * fused.  labels: channel-last float32 frames [n_frames][h*w, C] with `frame_stride` floats between frames (the
 * layout timet_ff_propagate writes); out: int64 [n_frames, out_h, out_w]. */
int timet_upsample_argmax(const float *labels, int n_frames, int h, int w, int n_channels, int out_h, int out_w,
                          int64_t frame_stride, int64_t *out, timet_stream_t stream);

/* ------------------------------------------------------------------ multi-GPU plumbing
 * One process per GPU (time_tuning.py:516-521,717).  Rank 0 calls timet_comm_unique_id and
 * broadcasts the 128 bytes with whatever it has (torch.distributed in timetuning_b200/dist.py);
 * every rank then calls timet_comm_init.  libnccl.so.2 is dlopen'ed on first use. */
#define TIMET_UNIQUE_ID_BYTES 128
int timet_comm_unique_id(void *id_out);
int timet_comm_init(const void *id, int rank, int world_size, timet_comm_t *comm_out);
int timet_comm_destroy(timet_comm_t comm);
/* Optional NVLink peer-memory path for the Sinkhorn marginals (single node): every rank exports a small exchange
 * buffer through CUDA IPC (timet_comm_p2p_handle -> 64 bytes), the host side all-gathers the handles, and
 * timet_comm_p2p_connect maps the peers' buffers.  With it, timet_sinkhorn(world_size > 1) runs as ONE resident
 * kernel per call that exchanges the K-vector by direct peer stores + flags (no NCCL launch per iteration). */
#define TIMET_IPC_HANDLE_BYTES 64
int timet_comm_p2p_handle(timet_comm_t comm, void *handle_out);
int timet_comm_p2p_connect(timet_comm_t comm, const void *all_handles /* world_size x 64 bytes, rank order */);
int timet_comm_p2p_disable(timet_comm_t comm); /* back to the NCCL path (all ranks must agree) */
/* sum-all-reduce of n float32 in place on `stream` (exposed for tests of the plumbing) */
int timet_comm_allreduce_f32(timet_comm_t comm, float *buf, int64_t n, timet_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TIMET_B200_H */
