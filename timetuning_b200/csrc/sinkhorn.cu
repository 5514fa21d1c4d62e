// Sinkhorn-Knopp balanced assignment, scaling-vector form.
//
// Reference: my_utils.sinkhorn (/root/reference/my_utils.py:246-274) called from
// TimeT.find_optimal_assignment (/root/reference/time_tuning.py:157-168).
//
//   E = exp(S/eps)  [B, K] row-major (B samples, K prototypes)      r = 1/K, c = 1/(B*ws)
//   pass 0      : R_i = sum_j E_ji                                  (ws>1: all-reduce R)
//   pass p=1..n-1: a_i = r/R_i ; s_j = sum_i a_i E_ji ; b_j = c/s_j ; R_i = sum_j E_ji b_j   (all-reduce R)
//   pass n      : a_i = r/R_i ; s_j = sum_i a_i E_ji ; Q_ji = a_i E_ji / s_j
//
// One streaming read of the input per pass, one write of Q: (iters+2)*B*K*4 bytes in total
// instead of the reference's ~66 elementwise passes.  E is never stored: in SCORES mode
// exp(S/eps) is recomputed in registers every pass (bit-identical each time).
// The column marginals are reduced deterministically: per-CTA partials in a fixed order, the
// last CTA to finish (atomic ticket) folds them in CTA order -> bit-reproducible runs.
#include <stdlib.h>

#include "common.cuh"

namespace timet {

constexpr int SK_THREADS = 512;
constexpr int SK_WARPS = SK_THREADS / 32;
constexpr int SK_MAX_V4 = 4;   // float4 chunks per lane -> K <= 512 on the vector path

// Row j of Q goes to q_out + (j / block_rows) * block_stride + (j % block_rows) * K  (block_rows <= 0: contiguous).
// Lets a caller have the assignment written straight into frame 0 of the channel-last label tensor
// [clip, frame, N, K] (block_rows = N, block_stride = n_frames * N * K): no copy between Sinkhorn and Feature-Forwarding.
__device__ __forceinline__ float *sk_out_row(float *q_out, int64_t row, int K, int64_t block_rows, int64_t block_stride) {
    if (block_rows <= 0) return q_out + row * K;
    const int64_t blk = row / block_rows;
    return q_out + blk * block_stride + (row - blk * block_rows) * K;
}

struct SkArgs {
    const float *in;
    float *q_out;
    int64_t out_block_rows, out_block_stride;
    float *partials;      // [grid, K]
    float *R;             // [K] marginals produced by this pass
    const float *R_prev;  // [K] marginals of the previous pass (nullptr in pass 0)
    unsigned int *ticket;
    int64_t B;
    int K;
    float inv_eps;        // SCORES mode: 1/eps
    float r, c;
    int scores_mode;
};

// Fixed-order, coalesced fold of per-CTA marginal partials [n_parts, K] -> out[K].
// Column i is split over P = min(NT / K, 8) threads: thread (i, j) adds parts g = j, j+P, ... in order
// (consecutive threads read consecutive columns: full 128-byte lines), then j = 0..P-1 are added in order.
// out[i] = r / sum if r != 0 (the scaling a_i), else the sum itself.  `scratch` holds >= 8*K floats.
template <int NT>
__device__ __forceinline__ void fold_partials(const float *partials, unsigned int n_parts, int K, float *scratch,
                                              float *out, float r) {
    int P = NT / K;
    P = P > 8 ? 8 : (P < 1 ? 1 : P);
    __syncthreads();
    for (int idx = threadIdx.x; idx < K * P; idx += NT) {
        const int j = idx / K, i = idx - j * K;
        float t = 0.f;
#pragma unroll 4
        for (unsigned int g = j; g < n_parts; g += P) t += __ldcg(partials + (size_t)g * K + i);
        scratch[j * K + i] = t;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += NT) {
        float t = scratch[i];
        for (int j = 1; j < P; ++j) t += scratch[j * K + i];
        out[i] = (r != 0.f) ? __fdiv_rn(r, t) : t;
    }
    __syncthreads();
}

// MODE 0: first pass (column sums only); 1: middle pass; 2: final pass (writes Q)
template <int MODE, int NV4>
__global__ void __launch_bounds__(SK_THREADS) sk_pass_vec(SkArgs A) {
    extern __shared__ float smem[];
    float *a_s = smem;                 // [K] scaling a_i = r / R_prev_i
    float *red = smem + A.K;           // [SK_WARPS, K] per-warp marginal partials
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = A.K, K4 = K >> 2;

    if (MODE != 0) {
        for (int i = threadIdx.x; i < K; i += SK_THREADS) a_s[i] = __fdiv_rn(A.r, A.R_prev[i]);
        __syncthreads();
    }
    float4 av[NV4];
    float4 acc[NV4];
#pragma unroll
    for (int v = 0; v < NV4; ++v) {
        acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int i4 = lane + 32 * v;
        av[v] = (MODE != 0 && i4 < K4) ? reinterpret_cast<const float4 *>(a_s)[i4] : make_float4(0.f, 0.f, 0.f, 0.f);
    }

    const int64_t warps_total = (int64_t)gridDim.x * SK_WARPS;
    for (int64_t row = (int64_t)blockIdx.x * SK_WARPS + warp; row < A.B; row += warps_total) {
        const float4 *src = reinterpret_cast<const float4 *>(A.in + row * K);
        float4 e[NV4];
#pragma unroll
        for (int v = 0; v < NV4; ++v) {
            const int i4 = lane + 32 * v;
            e[v] = (i4 < K4) ? __ldcs(src + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (A.scores_mode) {
#pragma unroll
            for (int v = 0; v < NV4; ++v) {
                const int i4 = lane + 32 * v;
                if (i4 < K4) {
                    e[v].x = expf(e[v].x * A.inv_eps);
                    e[v].y = expf(e[v].y * A.inv_eps);
                    e[v].z = expf(e[v].z * A.inv_eps);
                    e[v].w = expf(e[v].w * A.inv_eps);
                }
            }
        }
        if (MODE == 0) {
#pragma unroll
            for (int v = 0; v < NV4; ++v) {
                acc[v].x += e[v].x; acc[v].y += e[v].y; acc[v].z += e[v].z; acc[v].w += e[v].w;
            }
        } else {
            float4 p[NV4];                                  // a_i * E_ji
            float s = 0.f;
#pragma unroll
            for (int v = 0; v < NV4; ++v) {
                p[v] = make_float4(e[v].x * av[v].x, e[v].y * av[v].y, e[v].z * av[v].z, e[v].w * av[v].w);
                s += (p[v].x + p[v].y) + (p[v].z + p[v].w);
            }
            s = warp_sum(s);                                // s_j = sum_i a_i E_ji
            if (MODE == 1) {
                const float b = __fdiv_rn(A.c, s);          // b_j = c / s_j
#pragma unroll
                for (int v = 0; v < NV4; ++v) {             // R_i += E_ji * b_j
                    acc[v].x = fmaf(e[v].x, b, acc[v].x); acc[v].y = fmaf(e[v].y, b, acc[v].y);
                    acc[v].z = fmaf(e[v].z, b, acc[v].z); acc[v].w = fmaf(e[v].w, b, acc[v].w);
                }
            } else {
                const float inv = __fdiv_rn(1.f, s);
                float4 *dst = reinterpret_cast<float4 *>(sk_out_row(A.q_out, row, K, A.out_block_rows, A.out_block_stride));
#pragma unroll
                for (int v = 0; v < NV4; ++v) {
                    const int i4 = lane + 32 * v;
                    if (i4 < K4) __stcs(dst + i4, make_float4(p[v].x * inv, p[v].y * inv, p[v].z * inv, p[v].w * inv));
                }
            }
        }
    }
    if (MODE == 2) return;

    // ---- deterministic reduction of the prototype marginals
#pragma unroll
    for (int v = 0; v < NV4; ++v) {
        const int i4 = lane + 32 * v;
        if (i4 < K4) reinterpret_cast<float4 *>(red + warp * K)[i4] = acc[v];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += SK_THREADS) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < SK_WARPS; ++w) t += red[w * K + i];
        A.partials[(int64_t)blockIdx.x * K + i] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(A.ticket, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    fold_partials<SK_THREADS>(A.partials, gridDim.x, K, red, A.R, 0.f);
    if (threadIdx.x == 0) *A.ticket = 0;
}

// Generic-K scalar path (K not a multiple of 4, or K > 512): lane strides over columns.
template <int MODE>
__global__ void __launch_bounds__(SK_THREADS) sk_pass_scalar(SkArgs A) {
    extern __shared__ float smem[];
    float *a_s = smem;                 // [K]
    float *red = smem + A.K;           // [K] CTA marginal; the tail fold reuses red[0 .. 8K) as scratch
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = A.K;
    for (int i = threadIdx.x; i < K; i += SK_THREADS) {
        a_s[i] = (MODE != 0) ? __fdiv_rn(A.r, A.R_prev[i]) : 0.f;
        red[i] = 0.f;
    }
    __syncthreads();
    // each CTA owns a contiguous block of rows; warps take rows round-robin; the per-column
    // accumulation order inside a CTA is fixed by processing rows in warp-synchronous sweeps
    const int64_t rows_per_cta = (A.B + gridDim.x - 1) / gridDim.x;
    const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t row1 = (row0 + rows_per_cta < A.B) ? row0 + rows_per_cta : A.B;
    for (int64_t base = row0; base < row1; base += SK_WARPS) {
        const int64_t row = base + warp;
        const bool live = row < row1;
        float s = 0.f;
        if (live && MODE != 0) {
            for (int i = lane; i < K; i += 32) {
                float e = A.in[row * K + i];
                if (A.scores_mode) e = expf(e * A.inv_eps);
                s += e * a_s[i];
            }
        }
        s = warp_sum(s);
        const float b = (MODE == 1) ? __fdiv_rn(A.c, s) : 1.f;
        const float inv = (MODE == 2) ? __fdiv_rn(1.f, s) : 0.f;
        float *qrow = (MODE == 2 && live) ? sk_out_row(A.q_out, row, K, A.out_block_rows, A.out_block_stride) : nullptr;
        // sweep: warps add their row to the CTA marginal one after another (fixed order)
        for (int w = 0; w < SK_WARPS; ++w) {
            if (w == warp && live) {
                for (int i = lane; i < K; i += 32) {
                    float e = A.in[row * K + i];
                    if (A.scores_mode) e = expf(e * A.inv_eps);
                    if (MODE == 0) red[i] += e;
                    else if (MODE == 1) red[i] = fmaf(e, b, red[i]);
                    else qrow[i] = e * a_s[i] * inv;
                }
            }
            if (MODE != 2) __syncthreads();
        }
    }
    if (MODE == 2) return;
    for (int i = threadIdx.x; i < K; i += SK_THREADS) A.partials[(int64_t)blockIdx.x * K + i] = red[i];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(A.ticket, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    fold_partials<SK_THREADS>(A.partials, gridDim.x, K, red, A.R, 0.f);
    if (threadIdx.x == 0) *A.ticket = 0;
}

__global__ void sk_fill(float *p, int n, float v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---------------------------------------------------------------------------------------------
// Resident variant (single GPU): ONE cooperative launch for the whole call.  Each CTA keeps its
// block of rows of E = exp(S/eps) in shared memory across all iterations (config 2: 25 088 x 200
// fp32 = 20 MB over 148 SMs = 136 KB per CTA), so HBM traffic is the compulsory read of S and write
// of Q.  Per iteration: one sweep over shared memory, per-CTA marginal partials to global
// (double-buffered), a grid barrier, and a fixed-order fold of all partials by every CTA
// (bit-reproducible, identical on every CTA).
// Thread count of the resident kernel: 1024 (32 warps; the whole register file of the SM) by default; 512 when the
// caller wants to SHARE the SMs with other kernels running on another stream (half the register file and 1536 thread
// slots stay free: the HBM/L2-bound Feature-Forwarding kernels co-reside; the Sinkhorn call is latency-bound anyway).
constexpr int SKR_RED = 16;           // rows of the cross-warp reduction scratch (warps fold in SKR_WARPS / SKR_RED rounds)
// Every CTA adds its K marginal partials to K global fixed-point accumulators per iteration.  Packed, the 200 accumulators
// of config 2 share 13 cache lines and the 29 600 atomics of an iteration serialise in a handful of L2 slices; with one
// accumulator per 256 B (the granularity of the address -> L2 slice hash) they spread over all slices.
constexpr int SKR_USTRIDE = 32;
// An accumulator is (fixed-point sum << 16) | arrivals: one atomicAdd delivers a CTA's partial AND its arrival, and every
// CTA polls the K accumulators themselves until all CTAs are in -- no separate grid barrier (bar.sync + release fence +
// counter + re-read) inside the iteration loop.  The two buffers alternate by iteration parity and are never reset
// during a call: the marginal of iteration `it` is the difference to the value read two iterations earlier.
constexpr int SKR_CNT_BITS = 16;

struct SkResArgs {
    const float *in;
    float *q_out;
    int64_t out_block_rows, out_block_stride;
    float *partials;          // [2, grid, K]
    unsigned int *bar;        // monotonic grid-barrier counter (zeroed before launch)
    unsigned long long *ufix; // [2, K * ustride] fixed-point marginal accumulators with arrival counts (zeroed before launch)
    float ufix_scale, ufix_inv; // 2^sbits and 2^-sbits of the fixed-point part
    int ustride;              // u64 elements between two accumulators (SKR_USTRIDE: every accumulator in its own 256 B L2 block)
    void *const *peers;       // world_size > 1: every rank's P2PBuf[2] (NVLink peer memory), else nullptr
    int rank, ws;
    int chan;                 // exchange channel (P2PBuf index): two problems running side by side use one each
    int ll;                   // 1: low-latency tagged-word exchange (default), 0: data + flag with release / acquire
    int g_first, g_size;      // CTAs [g_first, g_first + g_size) of the launch work on this problem (0 / 0 = the whole grid)
    unsigned long long epoch0; // exchanges completed before this call (same on every rank)
    unsigned long long timeout_ns; // patience of the in-kernel waits (env TIMET_P2P_TIMEOUT_S, default 10 min)
    int64_t B;
    int K, iters, rows_per_cta, scores_mode;
    float inv_eps, r, c;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        // release at gpu scope: cumulative over everything the CTA wrote before the bar.sync above
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

// Deterministic cross-warp sum of per-lane column partials: warps fold into SKR_RED scratch rows in
// SKR_WARPS / SKR_RED ordered rounds; afterwards column i = sum of red[0..SKR_RED) [i].
template <int NV4, int SKR_WARPS>
__device__ __forceinline__ void skr_fold_warps(float *red, const float4 (&acc)[NV4], int K, int K4, int warp, int lane) {
    for (int round = 0; round < SKR_WARPS / SKR_RED; ++round) {
        if ((warp / SKR_RED) == round) {
            float4 *dst = reinterpret_cast<float4 *>(red + (warp % SKR_RED) * K);
#pragma unroll
            for (int v = 0; v < NV4; ++v) {
                const int i4 = lane + 32 * v;
                if (i4 < K4) {
                    float4 a = acc[v];
                    if (round > 0) { const float4 o = dst[i4]; a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w; }
                    dst[i4] = a;
                }
            }
        }
        __syncthreads();
    }
}

// Cross-GPU sum of a K-vector through NVLink peer memory (world_size > 1): CTA 0 stores this rank's vector into
// slot[e % 3][rank] of EVERY rank's exchange buffer, then publishes flag[rank] = e + 1 with release semantics at
// system scope; every CTA of every rank waits until all world_size flags of its own buffer reached e + 1 and adds
// the world_size vectors in rank order (bit-identical on all ranks).  `vec` (shared memory, [K]) holds the local
// vector on entry and the global sum on exit.  A slot is reused after 3 exchanges; a peer can only be one exchange
// ahead, so it is never overwritten while still being read.
// Low-latency exchange (default): every value is sent as ONE 8-byte word (float bits | tag = exchange number in the high
// half).  The leader CTA of the problem stores the rank's K words into every rank's buffer.  An 8-byte store is
// single-copy atomic, so data and "flag" arrive together: no fence, no second round trip.  (32-bit non-zero tags; a slot
// is reused every 3 exchanges, so a stale word can never carry the awaited tag.)
// Receive: every CTA of every rank polls the K x world_size words of its own buffer, ONE WORD PER THREAD (all of them in
// flight together: one L2 round trip for a whole node instead of one per rank, and no registers held across the sweep);
// arrivals are parked in `scratch` ([world_size, K] floats of shared memory) and summed per column in rank order.
template <int SKR_THREADS>
__device__ __forceinline__ void skr_exchange_ll(const SkResArgs &A, unsigned long long e, float *vec, float *scratch) {
    const int K = A.K;
    P2PBuf *own = reinterpret_cast<P2PBuf *>(A.peers[A.rank]) + A.chan;
    const int slot = (int)(e % 3ull);
    const unsigned long long tag = ((e % 0xFFFFFFFFull) + 1ull) << 32;      // never 0: a zero-initialised buffer matches nothing
    if ((int)blockIdx.x == A.g_first) {
        for (int idx = threadIdx.x; idx < K * A.ws; idx += SKR_THREADS) {
            const int p = idx / K, i = idx - p * K;
            unsigned long long *dst = &(reinterpret_cast<P2PBuf *>(A.peers[p]) + A.chan)->ll[slot][A.rank][i];
            const unsigned long long word = tag | (unsigned long long)__float_as_uint(vec[i]);
            asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(word) : "memory");
        }
    }
    for (int idx = threadIdx.x; idx < K * A.ws; idx += SKR_THREADS) {
        const int r = idx / K, i = idx - r * K;
        const unsigned long long *src = &own->ll[slot][r][i];
        unsigned long long v, t0 = 0ull;
        unsigned int spins = 0;
        do {
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory");
            if ((v & 0xFFFFFFFF00000000ull) == tag) break;
            if ((++spins & 0x3FFu) == 0u) {
                const unsigned long long now = globaltimer_ns();
                if (t0 == 0ull) t0 = now;
                else if (now - t0 > A.timeout_ns) {
                    printf("timet: sinkhorn peer exchange timed out after %llu s (rank %d waiting for rank %d, exchange %llu)\n",
                           A.timeout_ns / 1000000000ull, A.rank, r, e);
                    __trap();
                }
                __nanosleep(100);
            }
        } while (true);
        scratch[idx] = __uint_as_float((unsigned int)(v & 0xFFFFFFFFull));      // scratch[r * K + i]
    }
    __syncthreads();                                   // all arrivals parked; vec is rewritten below
    for (int i = threadIdx.x; i < K; i += SKR_THREADS) {
        float t = 0.f;
        for (int r = 0; r < A.ws; ++r) t += scratch[r * K + i];     // rank order: identical bits on every rank
        vec[i] = t;
    }
    __syncthreads();
}

template <int SKR_THREADS>
__device__ __forceinline__ void skr_exchange_flag(const SkResArgs &A, unsigned long long e, float *vec) {
    const int K = A.K;
    P2PBuf *own = reinterpret_cast<P2PBuf *>(A.peers[A.rank]) + A.chan;
    const int slot = (int)(e % 3ull);
    if ((int)blockIdx.x == A.g_first) {
        for (int p = 0; p < A.ws; ++p) {
            float *dst = (reinterpret_cast<P2PBuf *>(A.peers[p]) + A.chan)->slot[slot][A.rank];
            for (int i = threadIdx.x; i < K; i += SKR_THREADS) dst[i] = vec[i];
        }
        __syncthreads();                       // bar.sync + the release below (system scope) cover every thread's stores
        if ((int)threadIdx.x < A.ws) {
            unsigned long long *f = &(reinterpret_cast<P2PBuf *>(A.peers[threadIdx.x]) + A.chan)->flag[A.rank];
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(e + 1ull) : "memory");
        }
    }
    if ((int)threadIdx.x < A.ws) {
        const unsigned long long *f = &own->flag[threadIdx.x];
        // A peer may legitimately arrive late (lazy library build, data loader stall, checkpointing): wait like a
        // blocking NCCL collective would, by wall clock (%globaltimer), not by a spin count.
        unsigned long long v, t0 = 0ull;
        unsigned int spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= e + 1ull) break;
            if ((++spins & 0xFFFu) == 0u) {
                const unsigned long long now = globaltimer_ns();
                if (t0 == 0ull) t0 = now;
                else if (now - t0 > A.timeout_ns) {
                    printf("timet: sinkhorn peer exchange timed out after %llu s (rank %d waiting for rank %d, exchange %llu)\n",
                           A.timeout_ns / 1000000000ull, A.rank, (int)threadIdx.x, e);
                    __trap();
                }
                __nanosleep(200);
            }
        } while (true);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += SKR_THREADS) {
        float t = 0.f;
        for (int r = 0; r < A.ws; ++r) t += __ldcv(&own->slot[slot][r][i]);
        vec[i] = t;
    }
    __syncthreads();
}

// `scratch`: [world_size, K] floats of shared memory that do not overlap `vec` (the low-latency form parks arrivals there)
template <int SKR_THREADS>
__device__ __forceinline__ void skr_exchange(const SkResArgs &A, unsigned long long e, float *vec, float *scratch) {
    if (A.ll) skr_exchange_ll<SKR_THREADS>(A, e, vec, scratch);
    else skr_exchange_flag<SKR_THREADS>(A, e, vec);
}

// Packed fp32x2 arithmetic (sm_100: FMUL2 / FADD2 / FFMA2 issue one instruction for two IEEE fp32 operations): the sweep is
// issue-bound, so the multiply / accumulate work of a row is done on float2 halves of each float4.
__device__ __forceinline__ float2 lo2(const float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4 v) { return make_float2(v.z, v.w); }

// One sweep over `nrows` rows (warp `warp` takes rows warp, warp + WARPS, ...; two rows in flight):
//   s_j = sum_i a_i E_ji;   not last: acc_i += a_i E_ji * c / s_j   (marginals of the next Q);   last: Q_ji = a_i E_ji / s_j
// RESIDENT: E read from shared memory; else re-read from global memory and re-exponentiated (bit-identical values).
// Row sum order (canonical for every kernel of this file): per float4 ((x + z) + (y + w)), float4 chunks in lane order,
// then the xor butterfly.
template <int NV4, int WARPS, bool RESIDENT, bool ZERO = true>
__device__ __forceinline__ void skp_sweep(const SkResArgs &A, const float4 *E, int64_t row0, int nrows, const float *a_s,
                                          float4 (&acc)[NV4], bool last, int warp, int lane) {
    const int K = A.K, K4 = K >> 2;
    float4 av[NV4];
#pragma unroll
    for (int v = 0; v < NV4; ++v) {
        const int i4 = lane + 32 * v;
        av[v] = (i4 < K4) ? reinterpret_cast<const float4 *>(a_s)[i4] : make_float4(0.f, 0.f, 0.f, 0.f);
        if (ZERO) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    constexpr int RPW = (NV4 <= 2) ? 2 : 1;           // wider rows (K > 256) would spill with two rows in registers
    for (int rl = warp; rl < nrows; rl += RPW * WARPS) {
        const int rl2 = rl + WARPS;
        const bool two = RPW == 2 && rl2 < nrows;
        float2 pl[NV4], ph[NV4], ql[NV4], qh[NV4];
        float s = 0.f, s2 = 0.f;
#pragma unroll
        for (int v = 0; v < NV4; ++v) {
            const int i4 = lane + 32 * v;
            float4 e = make_float4(0.f, 0.f, 0.f, 0.f), f = make_float4(0.f, 0.f, 0.f, 0.f);
            if (RESIDENT) {
                if (i4 < K4) e = E[(size_t)rl * K4 + i4];
                if (two && i4 < K4) f = E[(size_t)rl2 * K4 + i4];
            } else {
                if (i4 < K4) e = __ldg(reinterpret_cast<const float4 *>(A.in + (row0 + rl) * K) + i4);
                if (two && i4 < K4) f = __ldg(reinterpret_cast<const float4 *>(A.in + (row0 + rl2) * K) + i4);
                if (A.scores_mode) {
                    if (i4 < K4) {
                        e.x = expf(e.x * A.inv_eps); e.y = expf(e.y * A.inv_eps); e.z = expf(e.z * A.inv_eps); e.w = expf(e.w * A.inv_eps);
                    }
                    if (two && i4 < K4) {
                        f.x = expf(f.x * A.inv_eps); f.y = expf(f.y * A.inv_eps); f.z = expf(f.z * A.inv_eps); f.w = expf(f.w * A.inv_eps);
                    }
                }
            }
            pl[v] = __fmul2_rn(lo2(e), lo2(av[v])); ph[v] = __fmul2_rn(hi2(e), hi2(av[v]));
            ql[v] = __fmul2_rn(lo2(f), lo2(av[v])); qh[v] = __fmul2_rn(hi2(f), hi2(av[v]));
            const float2 t = __fadd2_rn(pl[v], ph[v]), t2 = __fadd2_rn(ql[v], qh[v]);
            s += t.x + t.y;
            s2 += t2.x + t2.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (!last) {
            const float b = __fdiv_rn(A.c, s);
            const float b2 = two ? __fdiv_rn(A.c, s2) : 0.f;
            const float2 bb = make_float2(b, b), bb2 = make_float2(b2, b2);
#pragma unroll
            for (int v = 0; v < NV4; ++v) {
                float2 al = __ffma2_rn(pl[v], bb, lo2(acc[v])), ah = __ffma2_rn(ph[v], bb, hi2(acc[v]));
                if (two) { al = __ffma2_rn(ql[v], bb2, al); ah = __ffma2_rn(qh[v], bb2, ah); }
                acc[v] = make_float4(al.x, al.y, ah.x, ah.y);
            }
        } else {
            const float inv = __fdiv_rn(1.f, s);
            const float inv2 = two ? __fdiv_rn(1.f, s2) : 0.f;
            const float2 ii = make_float2(inv, inv), ii2 = make_float2(inv2, inv2);
            float4 *dst = reinterpret_cast<float4 *>(sk_out_row(A.q_out, row0 + rl, K, A.out_block_rows, A.out_block_stride));
            float4 *dst2 = reinterpret_cast<float4 *>(sk_out_row(A.q_out, row0 + (two ? rl2 : rl), K, A.out_block_rows, A.out_block_stride));
#pragma unroll
            for (int v = 0; v < NV4; ++v) {
                const int i4 = lane + 32 * v;
                if (i4 < K4) {
                    const float2 a = __fmul2_rn(pl[v], ii), bq = __fmul2_rn(ph[v], ii);
                    __stcs(dst + i4, make_float4(a.x, a.y, bq.x, bq.y));
                }
                if (two && i4 < K4) {
                    const float2 a = __fmul2_rn(ql[v], ii2), bq = __fmul2_rn(qh[v], ii2);
                    __stcs(dst2 + i4, make_float4(a.x, a.y, bq.x, bq.y));
                }
            }
        }
    }
}

template <int NV4, int SKR_THREADS>
__global__ void __launch_bounds__(SKR_THREADS, SKR_THREADS == 512 ? 2 : 1) sk_resident(SkResArgs A) {   // <= 64 registers either way
    constexpr int SKR_WARPS = SKR_THREADS / 32;
    extern __shared__ float4 smem4[];
    const int K = A.K, K4 = K >> 2;
    float4 *E = smem4;                                                   // [rows_per_cta, K4]
    float *a_s = reinterpret_cast<float *>(smem4 + (size_t)A.rows_per_cta * K4);    // [K]
    float *red = a_s + K;                                                // [SKR_RED, K]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * A.rows_per_cta;
    const int nrows = (int)max((int64_t)0, min((int64_t)A.rows_per_cta, A.B - row0));
    unsigned int epoch = 0;

    float4 acc[NV4];
#pragma unroll
    for (int v = 0; v < NV4; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);

    // ---- pass 0: load, exponentiate, keep in shared memory, column sums
    for (int rl = warp; rl < nrows; rl += SKR_WARPS) {
        const float4 *src = reinterpret_cast<const float4 *>(A.in + (row0 + rl) * K);
#pragma unroll
        for (int v = 0; v < NV4; ++v) {
            const int i4 = lane + 32 * v;
            if (i4 < K4) {
                float4 e = __ldcs(src + i4);
                if (A.scores_mode) {
                    e.x = expf(e.x * A.inv_eps); e.y = expf(e.y * A.inv_eps);
                    e.z = expf(e.z * A.inv_eps); e.w = expf(e.w * A.inv_eps);
                }
                E[(size_t)rl * K4 + i4] = e;
                acc[v].x += e.x; acc[v].y += e.y; acc[v].z += e.z; acc[v].w += e.w;
            }
        }
    }

    // ---- R^(0) = column sums of E: per-CTA float partials, grid barrier, fixed-order fold -> a_i = r / R_i
    skr_fold_warps<NV4, SKR_WARPS>(red, acc, K, K4, warp, lane);
    for (int i = threadIdx.x; i < K; i += SKR_THREADS) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < SKR_RED; ++w) t += red[w * K + i];
        A.partials[(size_t)blockIdx.x * K + i] = t;
    }
    grid_barrier(A.bar, (++epoch) * gridDim.x);
    fold_partials<SKR_THREADS>(A.partials, gridDim.x, K, red, a_s, 0.f);      // a_s = local column sums
    unsigned long long xch = A.epoch0;
    if (A.ws > 1) skr_exchange<SKR_THREADS>(A, xch++, a_s, red);
    for (int i = threadIdx.x; i < K; i += SKR_THREADS) a_s[i] = __fdiv_rn(A.r, a_s[i]);
    __syncthreads();

    unsigned long long prev0 = 0ull, prev1 = 0ull;                    // accumulator values read at the previous even / odd iteration
    for (int it = 0; it < A.iters; ++it) {
        const bool last = (it == A.iters - 1);
        // ---- sweep the resident rows (skp_sweep): s_j = sum_i a_i E_ji ; then either u_i += a_i E_ji * c/s_j or write Q
        skp_sweep<NV4, SKR_WARPS, true>(A, E, row0, nrows, a_s, acc, last, warp, lane);
        if (last) break;
        // ---- marginals of Q itself, u_i = sum_j a_i E_ji b_j (they sum to 1 over i, so the fixed point never
        // overflows): integer atomics are associative -> the grid-wide sum is bit-reproducible without a fold.
        // The add carries the arrival (+1 in the low 16 bits); each column's thread polls its own accumulator.
        skr_fold_warps<NV4, SKR_WARPS>(red, acc, K, K4, warp, lane);
        unsigned long long *ufix = A.ufix + (size_t)(it & 1) * K * A.ustride;
        const unsigned long long want = (unsigned long long)((it >> 1) + 1) * gridDim.x;     // arrivals after this iteration
        float u_mine = 0.f;
        const int i = threadIdx.x;                                    // K <= 512 <= SKR_THREADS: one column per thread
        if (i < K) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < SKR_RED; ++w) t += red[w * K + i];
            unsigned long long *acc_i = ufix + (size_t)i * A.ustride;
            atomicAdd(acc_i, ((unsigned long long)__float2ll_rn(t * A.ufix_scale) << SKR_CNT_BITS) + 1ull);
            // with world_size > 1 another CTA of this grid may itself be waiting for a late peer: same patience
            unsigned long long v, t0 = 0ull;
            unsigned int spins = 0;
            do {
                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(acc_i) : "memory");
                if ((v & ((1ull << SKR_CNT_BITS) - 1ull)) >= want) break;
                if ((++spins & 0xFFFFu) == 0u) {
                    const unsigned long long now = globaltimer_ns();
                    if (t0 == 0ull) t0 = now;
                    else if (now - t0 > A.timeout_ns) {
                        printf("timet: sinkhorn marginal wait timed out (block %d column %d iteration %d)\n", (int)blockIdx.x, i, it);
                        __trap();
                    }
                }
            } while (true);
            const unsigned long long cur = v >> SKR_CNT_BITS;
            const unsigned long long prev = (it & 1) ? prev1 : prev0;
            u_mine = (float)((double)(long long)(cur - prev) * (double)A.ufix_inv);            // local u_i
            if (it & 1) prev1 = cur; else prev0 = cur;
        }
        __syncthreads();                                               // everyone is done with the fold scratch
        if (i < K) red[i] = u_mine;
        __syncthreads();
        if (A.ws > 1) skr_exchange<SKR_THREADS>(A, xch++, red, red + K);                     // u_i summed over ranks (my_utils.py:270-272)
        if (i < K) a_s[i] = a_s[i] * __fdiv_rn(A.r, red[i]);           // Q *= r / u  (my_utils.py:268)
        __syncthreads();
    }
}

// Arguments of a launch that works on one or two problems (sk_hybrid).
struct SkPairArgs {
    SkResArgs a[2];           // per-problem pointers / accumulators (shape fields equal); a[0] is the resident one
};

// P: fold the CTA's marginal partials and add them (with the arrival) to the grid-wide fixed-point accumulators
template <int NV4, int WARPS>
__device__ __forceinline__ void skp_post(const SkResArgs &A, int it, float *red, const float4 (&acc)[NV4], int warp, int lane) {
    const int K = A.K, K4 = K >> 2;
    skr_fold_warps<NV4, WARPS>(red, acc, K, K4, warp, lane);
    const int i = threadIdx.x;
    if (i < K) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < SKR_RED; ++w) t += red[w * K + i];
        unsigned long long *acc_i = A.ufix + (size_t)(it & 1) * K * A.ustride + (size_t)i * A.ustride;
        atomicAdd(acc_i, ((unsigned long long)__float2ll_rn(t * A.ufix_scale) << SKR_CNT_BITS) + 1ull);
    }
    __syncthreads();                                       // the fold scratch is free again
}

// W: wait until every CTA's contribution of iteration `it` is in, (multi-GPU: exchange the K-vector with the peers,)
// and apply Q *= r / u to the scaling vector (my_utils.py:268)
template <int THREADS>
__device__ __forceinline__ void skp_wait(const SkResArgs &A, int it, unsigned long long &prev0, unsigned long long &prev1,
                                         unsigned long long &xch, float *a_s, float *u_s, float *red) {
    const int K = A.K;
    const int i = threadIdx.x;
    if (i < K) {
        unsigned long long *acc_i = A.ufix + (size_t)(it & 1) * K * A.ustride + (size_t)i * A.ustride;
        const unsigned long long want = (unsigned long long)((it >> 1) + 1) * (unsigned)(A.g_size ? A.g_size : (int)gridDim.x);
        unsigned long long v, t0 = 0ull;
        unsigned int spins = 0;
        do {
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(acc_i) : "memory");
            if ((v & ((1ull << SKR_CNT_BITS) - 1ull)) >= want) break;
            if ((++spins & 0xFFFFu) == 0u) {
                const unsigned long long now = globaltimer_ns();
                if (t0 == 0ull) t0 = now;
                else if (now - t0 > A.timeout_ns) {
                    printf("timet: sinkhorn (pair) marginal wait timed out (block %d column %d iteration %d)\n", (int)blockIdx.x, i, it);
                    __trap();
                }
            }
        } while (true);
        const unsigned long long cur = v >> SKR_CNT_BITS;
        const unsigned long long prev = (it & 1) ? prev1 : prev0;
        u_s[i] = (float)((double)(long long)(cur - prev) * (double)A.ufix_inv);
        if (it & 1) prev1 = cur; else prev0 = cur;
    }
    __syncthreads();
    if (A.ws > 1) skr_exchange<THREADS>(A, xch++, u_s, red);
    if (i < K) a_s[i] = a_s[i] * __fdiv_rn(A.r, u_s[i]);
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// HYBRID variant: ONE cooperative launch for a call whose rows do NOT fit shared memory (BASELINE configs[2] at 2 / 4
// GPUs: 128 / 64 clips = 100 352 / 50 176 rows per rank).  Every CTA keeps as many of its rows of exp(S/eps) resident as
// fit and re-reads the rest every iteration (from L2 while the score matrix fits there, else from HBM), recomputing exp --
// bit-identical values.  Same reductions and the same in-kernel NVLink exchange as sk_resident; replaces the one-launch-
// per-pass streaming path (+ one ncclAllReduce launch per pass at world_size > 1: 514 us per call at 8 ranks).
template <int NV4>
__global__ void __launch_bounds__(1024, 1) sk_hybrid(SkPairArgs P, int nprob, int res_rows) {
    constexpr int THREADS = 1024, WARPS = THREADS / 32;
    extern __shared__ float4 smem4[];
    // nprob = 2: the two problems run SIDE BY SIDE, each on half of the grid (DUAL mode).  A Sinkhorn call is bound by the
    // latency of its ~10 dependent grid-wide reductions, not by the sweep, so a problem loses little on half of the SMs
    // -- and the two calls of a training step (time_tuning.py:268,275) then overlap completely instead of queueing.
    const int G = (int)gridDim.x / nprob;
    const int which = (int)blockIdx.x / G;
    const SkResArgs &A = P.a[which];
    const int lb = (int)blockIdx.x - which * G;
    const int K = A.K, K4 = K >> 2;
    float4 *E = smem4;                                                              // [res_rows, K4]
    float *a_s = reinterpret_cast<float *>(smem4 + (size_t)res_rows * K4);          // [K]
    float *u_s = a_s + K;                                                           // [K]
    float *red = u_s + K;                                                           // [SKR_RED, K]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row0 = (int64_t)lb * A.rows_per_cta;
    const int nrows = (int)max((int64_t)0, min((int64_t)A.rows_per_cta, A.B - row0));
    const int nres = min(nrows, res_rows);

    float4 acc[NV4];
#pragma unroll
    for (int v = 0; v < NV4; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int rl = warp; rl < nrows; rl += WARPS) {
        const float4 *src = reinterpret_cast<const float4 *>(A.in + (row0 + rl) * K);
#pragma unroll
        for (int v = 0; v < NV4; ++v) {
            const int i4 = lane + 32 * v;
            if (i4 < K4) {
                float4 e = __ldg(src + i4);
                if (A.scores_mode) {
                    e.x = expf(e.x * A.inv_eps); e.y = expf(e.y * A.inv_eps);
                    e.z = expf(e.z * A.inv_eps); e.w = expf(e.w * A.inv_eps);
                }
                if (rl < nres) E[(size_t)rl * K4 + i4] = e;
                acc[v].x += e.x; acc[v].y += e.y; acc[v].z += e.z; acc[v].w += e.w;
            }
        }
    }
    skr_fold_warps<NV4, WARPS>(red, acc, K, K4, warp, lane);
    for (int i = threadIdx.x; i < K; i += THREADS) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < SKR_RED; ++w) t += red[w * K + i];
        A.partials[(size_t)lb * K + i] = t;
    }
    grid_barrier(A.bar, (unsigned)G);                   // the CTAs of this problem only (own counter)
    fold_partials<THREADS>(A.partials, (unsigned)G, K, red, a_s, 0.f);
    unsigned long long xch = A.epoch0;
    if (A.ws > 1) skr_exchange<THREADS>(A, xch++, a_s, red);
    for (int i = threadIdx.x; i < K; i += THREADS) a_s[i] = __fdiv_rn(A.r, a_s[i]);
    __syncthreads();

    unsigned long long p0 = 0ull, p1 = 0ull;
    for (int it = 0; it < A.iters; ++it) {
        const bool last = (it == A.iters - 1);
        skp_sweep<NV4, WARPS, true, true>(A, E, row0, nres, a_s, acc, last, warp, lane);
        skp_sweep<NV4, WARPS, false, false>(A, nullptr, row0 + nres, nrows - nres, a_s, acc, last, warp, lane);
        if (last) break;
        skp_post<NV4, WARPS>(A, it, red, acc, warp, lane);
        skp_wait<THREADS>(A, it, p0, p1, xch, a_s, u_s, red);
    }
}

static bool sk_hybrid_plan(int64_t B, int K, int *grid, int *rows_per_cta, int *res_rows, size_t *smem, int g = 0) {
    if (K % 4 != 0 || K > 128 * SK_MAX_V4) return false;
    if (g <= 0) g = num_sms();
    const int64_t rpc = (B + g - 1) / g;
    const size_t fixed = (size_t)2 * K * 4 + (size_t)SKR_RED * K * 4;
    const size_t room = (size_t)226 * 1024 - fixed;
    int64_t rr = (int64_t)(room / ((size_t)K * 4));
    if (rr > rpc) rr = rpc;
    if (rr < 0) rr = 0;
    *grid = (int)((B + rpc - 1) / rpc);
    *rows_per_cta = (int)rpc;
    *res_rows = (int)rr;
    *smem = fixed + (size_t)rr * K * 4;
    return rpc < (1ll << 30);
}

template <int NV4>
static int sk_hybrid_launch(SkPairArgs &P, int nprob, int res_rows, int grid, size_t smem, cudaStream_t st) {
    TIMET_CUDA(cudaFuncSetAttribute(sk_hybrid<NV4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void *args[] = {&P, &nprob, &res_rows};
    TIMET_CUDA(cudaLaunchCooperativeKernel((const void *)sk_hybrid<NV4>, dim3(grid), dim3(1024), args, smem, st));
    launch_counter()++;
    return TIMET_OK;
}

static bool sk_resident_plan(int64_t B, int K, int *grid, int *rows_per_cta, size_t *smem) {
    if (K % 4 != 0 || K > 128 * SK_MAX_V4) return false;
    const int g = num_sms();
    const int64_t rpc = (B + g - 1) / g;
    const size_t need = (size_t)rpc * K * 4 + (size_t)K * 4 + (size_t)SKR_RED * K * 4;
    if (need > 226 * 1024) return false;
    *grid = (int)((B + rpc - 1) / rpc);
    *rows_per_cta = (int)rpc;
    *smem = need;
    return true;
}

template <int NV4, int NT>
static int sk_resident_launch_nt(SkResArgs &A, int grid, size_t smem, cudaStream_t st) {
    TIMET_CUDA(cudaFuncSetAttribute(sk_resident<NV4, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void *args[] = {&A};
    TIMET_CUDA(cudaLaunchCooperativeKernel((const void *)sk_resident<NV4, NT>, dim3(grid), dim3(NT), args, smem, st));
    launch_counter()++;
    return TIMET_OK;
}
template <int NV4>
static int sk_resident_launch(SkResArgs &A, int grid, size_t smem, cudaStream_t st, bool share_sm) {
    return share_sm ? sk_resident_launch_nt<NV4, 512>(A, grid, smem, st) : sk_resident_launch_nt<NV4, 1024>(A, grid, smem, st);
}

static int sk_grid(int64_t B) {
    const int64_t want = (B + SK_WARPS - 1) / SK_WARPS;
    const int64_t cap = (int64_t)num_sms() * 2;        // 2 x 512-thread CTAs per SM
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

template <int MODE>
static int sk_launch(const SkArgs &A, int grid, cudaStream_t st) {
    const int K = A.K;
    const bool vec = (K % 4 == 0) && (K <= 128 * SK_MAX_V4) && ((reinterpret_cast<uintptr_t>(A.in) & 15) == 0) &&
                     (A.q_out == nullptr || (reinterpret_cast<uintptr_t>(A.q_out) & 15) == 0) && (A.out_block_stride % 4) == 0;
    if (vec) {
        const size_t smem = (size_t)(K + SK_WARPS * K) * sizeof(float);
        const int nv4 = (K / 4 + 31) / 32;
        switch (nv4) {
            case 1: sk_pass_vec<MODE, 1><<<grid, SK_THREADS, smem, st>>>(A); break;
            case 2: sk_pass_vec<MODE, 2><<<grid, SK_THREADS, smem, st>>>(A); break;
            case 3: sk_pass_vec<MODE, 3><<<grid, SK_THREADS, smem, st>>>(A); break;
            default: sk_pass_vec<MODE, 4><<<grid, SK_THREADS, smem, st>>>(A); break;
        }
    } else {
        const size_t smem = (size_t)9 * K * sizeof(float);   // a_s[K] | red[K] (+ 8*K fold scratch starting at red)
        if (smem > 48 * 1024) {
            set_error("sinkhorn: K=%d too large for the scalar path", K);
            return TIMET_ERR_UNSUPPORTED;
        }
        sk_pass_scalar<MODE><<<grid, SK_THREADS, smem, st>>>(A);
    }
    TIMET_LAUNCHED();
    return TIMET_OK;
}

int comm_allreduce_f32(timet_comm_t comm, float *buf, int64_t n, cudaStream_t st);

}  // namespace timet

using namespace timet;

extern "C" {

// workspace: partials [grid_max, K] | R ping [K] | R pong [K] | ticket
size_t timet_sinkhorn_workspace_bytes(int64_t B, int K) {
    (void)B;
    if (K < 1) return 0;
    const size_t grid_max = 2 * 160;   // >= 2 * SM count on every B200 SKU
    return align_up((grid_max + 2) * (size_t)K * sizeof(float) + 512 + 2 * (size_t)K * SKR_USTRIDE * sizeof(unsigned long long) + 64, 256);
}

int timet_sinkhorn(const float *in, int64_t B, int K, int input_kind, float epsilon, int iters, int world_size,
                   timet_comm_t comm, float *q_out, void *workspace, size_t workspace_bytes, timet_stream_t stream) {
    return timet_sinkhorn_ex(in, B, K, input_kind, epsilon, iters, world_size, comm, q_out, nullptr, workspace, workspace_bytes, stream);
}

int timet_sinkhorn_resident(int64_t B, int K) {
    int g, rpc, res;
    size_t smem;
    if (B < 1 || K < 1 || env_cfg().sk_streaming) return 0;
    if (sk_resident_plan(B, K, &g, &rpc, &smem) && g <= 160) return 1;
    if (sk_hybrid_plan(B, K, &g, &rpc, &res, &smem) && g <= 160) return 2;
    return 0;
}

int timet_sinkhorn_pair_mode(int64_t B, int K) {
    const EnvCfg &E = env_cfg();
    int g, rpc, res;
    size_t smem;
    if (B < 1 || K < 1 || E.sk_streaming) return 0;
    if (!E.sk_no_dual && sk_hybrid_plan(B, K, &g, &rpc, &res, &smem, num_sms() / 2) && g <= 160) return 1;
    return 0;
}

int timet_sinkhorn_ex(const float *in, int64_t B, int K, int input_kind, float epsilon, int iters, int world_size,
                      timet_comm_t comm, float *q_out, const timet_sinkhorn_opts *opts, void *workspace, size_t workspace_bytes,
                      timet_stream_t stream) {
    TIMET_CHECK_ARG(in && q_out && workspace, "sinkhorn: NULL pointer");
    const int64_t ob_rows = opts ? opts->out_block_rows : 0, ob_stride = opts ? opts->out_block_stride : 0;
    const bool share_sm = opts && opts->share_sm;
    TIMET_CHECK_ARG(ob_rows <= 0 || (ob_stride >= ob_rows * K && B % ob_rows == 0),
                    "sinkhorn: output blocks of %lld rows with stride %lld do not tile B=%lld rows of K=%d", (long long)ob_rows,
                    (long long)ob_stride, (long long)B, K);
    TIMET_CHECK_ARG(B >= 1 && K >= 1, "sinkhorn: bad shape B=%lld K=%d", (long long)B, K);
    TIMET_CHECK_ARG(iters >= 0, "sinkhorn: iters=%d must be >= 0", iters);
    TIMET_CHECK_ARG(input_kind == TIMET_SK_EXP || input_kind == TIMET_SK_SCORES, "sinkhorn: bad input_kind %d", input_kind);
    TIMET_CHECK_ARG(input_kind == TIMET_SK_EXP || epsilon > 0.f, "sinkhorn: epsilon must be > 0");
    TIMET_CHECK_ARG(world_size >= 1, "sinkhorn: world_size=%d must be >= 1", world_size);
    TIMET_CHECK_ARG(world_size == 1 || comm != nullptr, "sinkhorn: world_size=%d needs a communicator", world_size);
    if (workspace_bytes < timet_sinkhorn_workspace_bytes(B, K)) {
        set_error("sinkhorn: workspace %zu < %zu bytes", workspace_bytes, timet_sinkhorn_workspace_bytes(B, K));
        return TIMET_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    {   // single GPU, >= 1 iteration, rows fit in shared memory: one cooperative launch
        int rgrid, rpc;
        size_t rsmem;
        const EnvCfg &E = env_cfg();
        void **peers = nullptr;
        int prank = 0, pws = 1;
        unsigned long long *pepoch = nullptr;
        const bool p2p = world_size > 1 && comm_p2p_info(comm, &peers, &prank, &pws, &pepoch) && pws == world_size && K <= P2P_MAX_K;
        if ((world_size == 1 || p2p) && iters >= 1 && !E.sk_streaming && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
            (reinterpret_cast<uintptr_t>(q_out) & 15) == 0 && (ob_stride % 4) == 0 && sk_resident_plan(B, K, &rgrid, &rpc, &rsmem) && rgrid <= 160 &&
            (int64_t)(iters / 2 + 1) * rgrid < (1 << SKR_CNT_BITS)) {      // arrival counts of a call fit their 16 bits
            float *partials = (float *)workspace;
            unsigned int *bar = (unsigned int *)(partials + (size_t)322 * K);
            // bar (64 B slot) followed by the 8-byte aligned fixed-point buffers; one memset clears both
            const size_t bar_off = (size_t)322 * K * sizeof(float);
            const size_t ufix_off = align_up(bar_off + 64, 256);
            const int ustride = (E.sk_ustride >= 1 && E.sk_ustride <= SKR_USTRIDE) ? E.sk_ustride : SKR_USTRIDE;   // 1 = packed accumulators (for comparison)
            TIMET_CUDA(cudaMemsetAsync((char *)workspace + bar_off, 0, ufix_off - bar_off + 2 * (size_t)K * ustride * sizeof(unsigned long long), st));
            SkResArgs R;
            R.chan = 0; R.g_first = 0; R.g_size = 0; R.ll = (E.sk_ll && world_size <= 15) ? 1 : 0;
            R.ustride = ustride;
            // 48 data bits: the marginals of one iteration sum to <= 1 and a buffer accumulates ceil(iters / 2) of them
            int head = 1;
            while ((1 << head) < iters / 2 + 2) ++head;
            R.ufix_scale = ldexpf(1.0f, 47 - head);
            R.ufix_inv = ldexpf(1.0f, head - 47);
            R.ufix = (unsigned long long *)((char *)workspace + ufix_off);
            R.out_block_rows = ob_rows; R.out_block_stride = ob_stride;
            R.in = in; R.q_out = q_out; R.partials = partials; R.bar = bar; R.B = B; R.K = K; R.iters = iters;
            R.rows_per_cta = rpc; R.scores_mode = (input_kind == TIMET_SK_SCORES);
            R.inv_eps = (input_kind == TIMET_SK_SCORES) ? 1.0f / epsilon : 0.f;
            R.r = 1.0f / (float)K; R.c = 1.0f / ((float)B * (float)world_size);
            R.peers = p2p ? peers : nullptr; R.rank = prank; R.ws = p2p ? pws : 1;
            R.epoch0 = p2p ? *pepoch : 0ull;
            R.timeout_ns = (unsigned long long)(E.p2p_timeout_s * 1e9);
            if (p2p) *pepoch += (unsigned long long)iters;             // pass 0 + (iters - 1) iterations exchange a vector
            switch ((K / 4 + 31) / 32) {
                case 1: return sk_resident_launch<1>(R, rgrid, rsmem, st, share_sm);
                case 2: return sk_resident_launch<2>(R, rgrid, rsmem, st, share_sm);
                case 3: return sk_resident_launch<3>(R, rgrid, rsmem, st, share_sm);
                default: return sk_resident_launch<4>(R, rgrid, rsmem, st, share_sm);
            }
        }
        // rows do not fit shared memory: hybrid kernel (resident part + re-read part), still one cooperative launch
        int hgrid, hrpc, hres;
        size_t hsmem;
        if ((world_size == 1 || p2p) && iters >= 1 && !E.sk_streaming && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
            (reinterpret_cast<uintptr_t>(q_out) & 15) == 0 && (ob_stride % 4) == 0 && !sk_resident_plan(B, K, &rgrid, &rpc, &rsmem) &&
            sk_hybrid_plan(B, K, &hgrid, &hrpc, &hres, &hsmem) && hgrid <= 160 && (int64_t)(iters / 2 + 1) * hgrid < (1 << SKR_CNT_BITS)) {
            float *partials = (float *)workspace;
            const size_t bar_off = (size_t)322 * K * sizeof(float);
            const size_t ufix_off = align_up(bar_off + 64, 256);
            const int ustride = (E.sk_ustride >= 1 && E.sk_ustride <= SKR_USTRIDE) ? E.sk_ustride : SKR_USTRIDE;
            TIMET_CUDA(cudaMemsetAsync((char *)workspace + bar_off, 0, ufix_off - bar_off + 2 * (size_t)K * ustride * sizeof(unsigned long long), st));
            SkPairArgs HP;
            SkResArgs &R = HP.a[0];
            R.chan = 0; R.g_first = 0; R.g_size = hgrid; R.ll = (E.sk_ll && world_size <= 15) ? 1 : 0;
            R.ustride = ustride;
            int head = 1;
            while ((1 << head) < iters / 2 + 2) ++head;
            R.ufix_scale = ldexpf(1.0f, 47 - head);
            R.ufix_inv = ldexpf(1.0f, head - 47);
            R.ufix = (unsigned long long *)((char *)workspace + ufix_off);
            R.out_block_rows = ob_rows; R.out_block_stride = ob_stride;
            R.in = in; R.q_out = q_out; R.partials = partials; R.bar = (unsigned int *)((char *)workspace + bar_off);
            R.B = B; R.K = K; R.iters = iters; R.rows_per_cta = hrpc; R.scores_mode = (input_kind == TIMET_SK_SCORES);
            R.inv_eps = (input_kind == TIMET_SK_SCORES) ? 1.0f / epsilon : 0.f;
            R.r = 1.0f / (float)K; R.c = 1.0f / ((float)B * (float)world_size);
            R.peers = p2p ? peers : nullptr; R.rank = prank; R.ws = p2p ? pws : 1;
            R.epoch0 = p2p ? *pepoch : 0ull;
            R.timeout_ns = (unsigned long long)(E.p2p_timeout_s * 1e9);
            if (p2p) *pepoch += (unsigned long long)iters;
            switch ((K / 4 + 31) / 32) {
                case 1: return sk_hybrid_launch<1>(HP, 1, hres, hgrid, hsmem, st);
                case 2: return sk_hybrid_launch<2>(HP, 1, hres, hgrid, hsmem, st);
                case 3: return sk_hybrid_launch<3>(HP, 1, hres, hgrid, hsmem, st);
                default: return sk_hybrid_launch<4>(HP, 1, hres, hgrid, hsmem, st);
            }
        }
    }
    const int grid = sk_grid(B);
    TIMET_CHECK_ARG(grid <= 320, "sinkhorn: grid %d exceeds the workspace layout", grid);
    float *partials = (float *)workspace;
    float *Rbuf[2] = {partials + (size_t)320 * K, partials + (size_t)321 * K};
    unsigned int *ticket = (unsigned int *)(partials + (size_t)322 * K);
    TIMET_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned int), st));

    SkArgs A;
    A.in = in; A.q_out = q_out; A.partials = partials; A.ticket = ticket;
    A.out_block_rows = ob_rows; A.out_block_stride = ob_stride;
    A.B = B; A.K = K;
    A.inv_eps = (input_kind == TIMET_SK_SCORES) ? 1.0f / epsilon : 0.f;
    A.scores_mode = (input_kind == TIMET_SK_SCORES);
    A.r = 1.0f / (float)K;
    A.c = 1.0f / ((float)B * (float)world_size);

    int rc;
    if (iters == 0) {
        // no scaling iterations: Q = E / rowsum(E)  (my_utils.py:274 applied to the normalised input)
        // run the final pass with a_i = 1: feed R_prev = r so that r / R_prev = 1
        float *ones = Rbuf[0];
        sk_fill<<<(K + 255) / 256, 256, 0, st>>>(ones, K, A.r);
        TIMET_LAUNCHED();
        A.R_prev = ones; A.R = nullptr;
        return sk_launch<2>(A, grid, st);
    }
    // pass 0: plain column sums
    A.R_prev = nullptr; A.R = Rbuf[0];
    if ((rc = sk_launch<0>(A, grid, st)) != TIMET_OK) return rc;
    if (world_size > 1 && (rc = comm_allreduce_f32(comm, A.R, K, st)) != TIMET_OK) return rc;
    for (int it = 1; it < iters; ++it) {
        A.R_prev = Rbuf[(it - 1) & 1]; A.R = Rbuf[it & 1];
        if ((rc = sk_launch<1>(A, grid, st)) != TIMET_OK) return rc;
        if (world_size > 1 && (rc = comm_allreduce_f32(comm, A.R, K, st)) != TIMET_OK) return rc;
    }
    A.R_prev = Rbuf[(iters - 1) & 1]; A.R = nullptr;
    return sk_launch<2>(A, grid, st);
}


int timet_sinkhorn_pair(const float *in0, const float *in1, int64_t B, int K, int input_kind, float epsilon, int iters,
                        int world_size, timet_comm_t comm, float *q0, const timet_sinkhorn_opts *opts0, float *q1,
                        const timet_sinkhorn_opts *opts1, void *workspace, size_t workspace_bytes, timet_stream_t stream) {
    TIMET_CHECK_ARG(in0 && in1 && q0 && q1 && workspace, "sinkhorn_pair: NULL pointer");
    TIMET_CHECK_ARG(B >= 1 && K >= 1, "sinkhorn_pair: bad shape B=%lld K=%d", (long long)B, K);
    const size_t one = timet_sinkhorn_workspace_bytes(B, K);
    if (workspace_bytes < 2 * one) {
        set_error("sinkhorn_pair: workspace %zu < %zu bytes (2 x timet_sinkhorn_workspace_bytes)", workspace_bytes, 2 * one);
        return TIMET_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const float *in[2] = {in0, in1};
    float *qo[2] = {q0, q1};
    const timet_sinkhorn_opts *op[2] = {opts0, opts1};
    char *wsp[2] = {(char *)workspace, (char *)workspace + one};
    const EnvCfg &E = env_cfg();
    void **peers = nullptr;
    int prank = 0, pws = 1;
    unsigned long long *pepoch = nullptr;
    const bool p2p = world_size > 1 && comm != nullptr && comm_p2p_info(comm, &peers, &prank, &pws, &pepoch) && pws == world_size && K <= P2P_MAX_K;
    // ---- DUAL: the two problems side by side on half of the SMs each (sk_hybrid with nprob = 2), default
    {
        int dgrid, drpc, dres;
        size_t dsmem;
        const int half = num_sms() / 2;
        bool dual = (world_size == 1 || p2p) && iters >= 1 && !E.sk_streaming && !E.sk_no_dual && epsilon > 0.f &&
                    (input_kind == TIMET_SK_EXP || input_kind == TIMET_SK_SCORES) && half >= 1 &&
                    sk_hybrid_plan(B, K, &dgrid, &drpc, &dres, &dsmem, half) && dgrid <= 160 &&
                    (int64_t)(iters / 2 + 1) * dgrid < (1 << SKR_CNT_BITS);
        for (int c = 0; c < 2 && dual; ++c) {
            const int64_t obr = op[c] ? op[c]->out_block_rows : 0, obs = op[c] ? op[c]->out_block_stride : 0;
            dual = (reinterpret_cast<uintptr_t>(in[c]) & 15) == 0 && (reinterpret_cast<uintptr_t>(qo[c]) & 15) == 0 && (obs % 4) == 0 &&
                   (obr <= 0 || (obs >= obr * K && B % obr == 0));
        }
        if (dual) {
            SkPairArgs P;
            const size_t bar_off = (size_t)322 * K * sizeof(float);
            const size_t ufix_off = align_up(bar_off + 64, 256);
            const int ustride = (E.sk_ustride >= 1 && E.sk_ustride <= SKR_USTRIDE) ? E.sk_ustride : SKR_USTRIDE;
            int head = 1;
            while ((1 << head) < iters / 2 + 2) ++head;
            for (int c = 0; c < 2; ++c) {
                TIMET_CUDA(cudaMemsetAsync(wsp[c] + bar_off, 0, ufix_off - bar_off + 2 * (size_t)K * ustride * sizeof(unsigned long long), st));
                SkResArgs &R = P.a[c];
                R.chan = c; R.g_first = c * dgrid; R.g_size = dgrid; R.ll = (E.sk_ll && world_size <= 15) ? 1 : 0;
                R.ustride = ustride;
                R.ufix_scale = ldexpf(1.0f, 47 - head);
                R.ufix_inv = ldexpf(1.0f, head - 47);
                R.ufix = (unsigned long long *)(wsp[c] + ufix_off);
                R.out_block_rows = op[c] ? op[c]->out_block_rows : 0;
                R.out_block_stride = op[c] ? op[c]->out_block_stride : 0;
                R.in = in[c]; R.q_out = qo[c]; R.partials = (float *)wsp[c]; R.bar = (unsigned int *)(wsp[c] + bar_off);
                R.B = B; R.K = K; R.iters = iters; R.rows_per_cta = drpc; R.scores_mode = (input_kind == TIMET_SK_SCORES);
                R.inv_eps = (input_kind == TIMET_SK_SCORES) ? 1.0f / epsilon : 0.f;
                R.r = 1.0f / (float)K; R.c = 1.0f / ((float)B * (float)world_size);
                R.peers = p2p ? peers : nullptr; R.rank = prank; R.ws = p2p ? pws : 1;
                R.epoch0 = p2p ? pepoch[c] : 0ull;
                R.timeout_ns = (unsigned long long)(E.p2p_timeout_s * 1e9);
                if (p2p) pepoch[c] += (unsigned long long)iters;       // pass 0 + (iters - 1) iterations exchange a vector
            }
            switch ((K / 4 + 31) / 32) {
                case 1: return sk_hybrid_launch<1>(P, 2, dres, 2 * dgrid, dsmem, st);
                case 2: return sk_hybrid_launch<2>(P, 2, dres, 2 * dgrid, dsmem, st);
                case 3: return sk_hybrid_launch<3>(P, 2, dres, 2 * dgrid, dsmem, st);
                default: return sk_hybrid_launch<4>(P, 2, dres, 2 * dgrid, dsmem, st);
            }
        }
    }
    // not a dual launch: two independent calls (each validates its own arguments)
    int rc = timet_sinkhorn_ex(in0, B, K, input_kind, epsilon, iters, world_size, comm, q0, opts0, wsp[0], one, stream);
    if (rc != TIMET_OK) return rc;
    return timet_sinkhorn_ex(in1, B, K, input_kind, epsilon, iters, world_size, comm, q1, opts1, wsp[1], one, stream);
}

}
