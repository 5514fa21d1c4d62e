// FF stage 2, EXACT engine: fp32 CUDA-core scan of every in-window key of every context.
//
// Reference maths per target frame (/root/reference/mask_propagation.py:418-436):
//   sim = f^_t[i] . f^_c[j]; aff = exp(sim/0.1) * window(i,j); theta_i = k-th largest aff over all
//   contexts; w = aff if aff >= theta_i else 0; w /= sum(w).
// The reference materialises the [ctx*N, N] affinity in HBM and sweeps it ~12 times; here each
// query keeps a sorted list in the registers of one warp and nothing but the <= kw survivors is
// written.  This engine is (a) the general path for shapes the tensor-core engine does not
// cover, (b) the re-do path for queries whose tensor-core candidate list overflowed, and
// (c) the bit-exact yardstick for the tensor-core engine (same dot_canonical order).
#include "ff_select.cuh"

namespace timet {

constexpr int EX_WARPS = 8;

struct WidePool {
    float *w;              // [cap] normalised weights
    int32_t *k;            // [cap] keys (frame * N + patch)
    unsigned long long *used;   // entries allocated so far (zeroed by timet_ff_select); 64-bit: never wraps
    unsigned long long cap;
};

// Second scan of one query whose survivor set does not fit kw slots: every in-window key with aff >= theta is kept,
// exactly like the reference (mask_propagation.py:434-436: aff[aff < kth] = 0; aff /= aff.sum()).  Two passes over the
// window (count + sum, then emit in scan order = context-major, window row-major: deterministic); the row lives in the
// pool, its descriptor in the regular slots: sel_cnt = -n, sel_k[0] = pool offset.  Returns false if the pool is full.
__device__ __forceinline__ bool ex_wide_row(const timet_ff_params &p, int N, const FFSrc &S, float inv_q,
                                            const float4 *qs, int64_t clip_row0, int t, int r0, int c0, int wcols, int nwin,
                                            float theta, int lane, const WidePool &pool, float *__restrict__ w_out,
                                            int32_t *__restrict__ k_out, int32_t *__restrict__ cnt_out, int kw, int *n_out) {
    const int W = p.grid_w;
    const int nctx = ctx_count(t, p.n_last_frames);
    unsigned long long off = 0;
    float total = 0.f;
    int n = 0;
    for (int pass = 0; pass < 2; ++pass) {
        float part = 0.f;
        int emitted = 0;
        for (int ci = 0; ci < nctx; ++ci) {
            const int f = ctx_frame(t, p.n_last_frames, ci);
            const int64_t frow0 = clip_row0 + (int64_t)f * N;
            for (int m0 = 0; m0 < nwin; m0 += 32) {
                const int m = m0 + lane;
                float aff = -1.f;
                int32_t key = 0;
                if (m < nwin) {
                    const int wr = m / wcols;
                    const int j = (r0 + wr) * W + c0 + (m - wr * wcols);
                    const float dot = dot_canonical_seq(qs, reinterpret_cast<const float4 *>(S.x + (frow0 + j) * S.ld), S.n4);
                    const float sim = sim_from_dot(dot, inv_q, __ldg(S.inv + frow0 + j));
                    aff = affinity_from_sim(sim, p.temperature);
                    key = f * N + j;
                }
                const bool hit = aff >= theta;
                const unsigned mask = __ballot_sync(0xffffffffu, hit);
                if (pass == 0) {
                    if (hit) part += aff;
                } else if (hit) {
                    const unsigned long long slot = off + (unsigned)emitted + (unsigned)__popc(mask & ((1u << lane) - 1u));
                    pool.w[slot] = __fdiv_rn(aff, total);
                    pool.k[slot] = key;
                }
                emitted += __popc(mask);
            }
        }
        if (pass == 0) {
            n = emitted;
            total = warp_sum(part);
            if (lane == 0) off = atomicAdd(pool.used, (unsigned long long)n);
            off = __shfl_sync(0xffffffffu, off, 0);
            if (off + (unsigned long long)n > pool.cap) return false;
        }
    }
    if (lane < kw) {
        w_out[lane] = 0.f;
        k_out[lane] = (lane == 0) ? (int32_t)off : -1;
    }
    if (lane == 0) *cnt_out = -n;
    *n_out = n;
    return true;
}

__global__ void __launch_bounds__(EX_WARPS * 32)
ff_select_exact_kernel(timet_ff_params p, int N, FFSrc S, int nT, int kw,
                       float *__restrict__ sel_w, int32_t *__restrict__ sel_k, int32_t *__restrict__ sel_cnt,
                       unsigned long long *__restrict__ stats, const int32_t *__restrict__ qlist,
                       const unsigned int *__restrict__ qcount, int64_t n_queries, WidePool pool) {
    extern __shared__ float4 qsm[];                       // [EX_WARPS][n4]
    __shared__ unsigned long long s_stat[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 4) s_stat[threadIdx.x] = 0ull;
    __syncthreads();
    float4 *qs = qsm + (size_t)warp * S.n4;
    const int H = p.grid_h, W = p.grid_w;
    const int64_t total = qlist ? (int64_t)*qcount : n_queries;
    const int64_t nwarps = (int64_t)gridDim.x * EX_WARPS;
    if (qlist && blockIdx.x == 0 && threadIdx.x == 0) stats[4] = (unsigned long long)*qcount;   // queries re-done after a candidate-list overflow
    unsigned long long st_sel = 0, st_ties = 0, st_trunc = 0, st_wide = 0;

    for (int64_t it = (int64_t)blockIdx.x * EX_WARPS + warp; it < total; it += nwarps) {
        const int64_t qid = qlist ? (int64_t)qlist[it] : it;
        const int clip = (int)(qid / ((int64_t)nT * N));
        const int rem = (int)(qid - (int64_t)clip * nT * N);
        const int tt = rem / N, i = rem - tt * N;
        const int t = p.t_begin + tt;
        const int64_t clip_row0 = (int64_t)clip * p.n_frames * N;

        __syncwarp();
        const int64_t q_row = clip_row0 + (int64_t)t * N + i;
        const float4 *qrow = reinterpret_cast<const float4 *>(S.x + q_row * S.ld);
        for (int d = lane; d < S.n4; d += 32) qs[d] = qrow[d];
        const float inv_q = __ldg(S.inv + q_row);
        __syncwarp();

        const int qr = i / W, qc = i - qr * W;
        int r0 = 0, r1 = H - 1, c0 = 0, c1 = W - 1;
        if (p.radius > 0) {
            r0 = max(qr - p.radius, 0); r1 = min(qr + p.radius, H - 1);
            c0 = max(qc - p.radius, 0); c1 = min(qc + p.radius, W - 1);
        }
        const int wcols = c1 - c0 + 1, nwin = (r1 - r0 + 1) * wcols;

        TopList L;
        list_init(L);
        const int nctx = ctx_count(t, p.n_last_frames);
        for (int ci = 0; ci < nctx; ++ci) {
            const int f = ctx_frame(t, p.n_last_frames, ci);
            const int64_t frow0 = clip_row0 + (int64_t)f * N;
            for (int m0 = 0; m0 < nwin; m0 += 32) {
                const int m = m0 + lane;
                const bool valid = m < nwin;
                float aff = 0.f;
                int32_t key = 0;
                if (valid) {
                    const int wr = m / wcols;
                    const int j = (r0 + wr) * W + c0 + (m - wr * wcols);
                    const float dot = dot_canonical_seq(qs, reinterpret_cast<const float4 *>(S.x + (frow0 + j) * S.ld), S.n4);
                    const float sim = sim_from_dot(dot, inv_q, __ldg(S.inv + frow0 + j));
                    aff = affinity_from_sim(sim, p.temperature);
                    key = f * N + j;
                }
                list_offer(L, valid, aff, key, p.topk, lane);
            }
        }
        // survivors: everything >= the k-th value.  The sorted list holds the 32 largest entries, so the k-th value is
        // exact; the survivor SET is complete unless it fills all kw slots' worth (or all 32 list slots with entries
        // dropped behind them) -- then the query is rescanned into a variable-length row of the pool.
        const int m_list = list_kept(L, p.topk, lane);
        int m = m_list;
        bool done = false;
        if (m_list > kw || (m_list == 32 && L.dropped > 0)) {
            int n_wide = 0;
            done = ex_wide_row(p, N, S, inv_q, qs, clip_row0, t, r0, c0, wcols, nwin, L.kth, lane, pool, sel_w + qid * kw,
                               sel_k + qid * kw, sel_cnt + qid, kw, &n_wide);
            if (done) { m = n_wide; st_wide += 1; st_sel += (unsigned long long)n_wide; }
            else st_trunc += 1;                     // pool exhausted: the row below is truncated to kw entries
        }
        if (!done) {
            list_finish(L, p.topk, kw, lane, sel_w + qid * kw, sel_k + qid * kw, sel_cnt + qid);
            st_sel += (unsigned long long)(m < kw ? m : kw);
        }
        st_ties += (m > p.topk);
    }
    if (lane == 0) {
        atomicAdd(&s_stat[0], st_sel);
        atomicAdd(&s_stat[1], st_ties);
        atomicAdd(&s_stat[2], st_trunc);
        atomicAdd(&s_stat[3], st_wide);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_stat[0]) atomicAdd(&stats[1], s_stat[0]);
        if (s_stat[1]) atomicAdd(&stats[2], s_stat[1]);
        if (s_stat[2]) atomicAdd(&stats[5], s_stat[2]);
        if (s_stat[3]) atomicAdd(&stats[6], s_stat[3]);
    }
}

// qlist == nullptr: every query.  Otherwise the first *qcount entries of qlist (device).
int ff_select_exact_run(const timet_ff_params &p, const FFLayout &L, const float *feats, char *ws, const int32_t *qlist,
                        const unsigned int *qcount, int64_t max_items, cudaStream_t st) {
    const FFSrc S = ff_src(p, L, feats, ws);
    const size_t smem = (size_t)EX_WARPS * S.n4 * sizeof(float4);
    if (smem > 48 * 1024)
        TIMET_CUDA(cudaFuncSetAttribute(ff_select_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (max_items + EX_WARPS - 1) / EX_WARPS;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    WidePool pool;
    pool.w = reinterpret_cast<float *>(ws + L.off_wide_w);
    pool.k = reinterpret_cast<int32_t *>(ws + L.off_wide_k);
    pool.used = reinterpret_cast<unsigned long long *>(ws + L.off_redo + FF_HDR_WIDE_USED);
    pool.cap = (unsigned long long)L.wide_cap;
    ff_select_exact_kernel<<<(int)blocks, EX_WARPS * 32, smem, st>>>(
        p, L.N, S, L.nT, L.kw,
        reinterpret_cast<float *>(ws + L.off_sel_w), reinterpret_cast<int32_t *>(ws + L.off_sel_k),
        reinterpret_cast<int32_t *>(ws + L.off_sel_cnt), reinterpret_cast<unsigned long long *>(ws + L.off_stats),
        qlist, qcount, L.queries, pool);
    TIMET_LAUNCHED();
    return TIMET_OK;
}

int ff_select_exact_launch(const timet_ff_params &p, const FFLayout &L, const float *feats, char *ws, cudaStream_t st) {
    return ff_select_exact_run(p, L, feats, ws, nullptr, nullptr, L.queries, st);
}

}  // namespace timet
