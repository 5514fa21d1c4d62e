// FF stage 3: frame-sequential weighted gather of context labels, then argmax of the last frame.
//
// Reference: seg_tar = segs.double() @ aff.double()  — a float64 dense GEMM [C, ctx*N] x [ctx*N, N]
// whose right operand has <= k (+ties) non-zeros per column (/root/reference/mask_propagation.py:439-444),
// plus a GPU->CPU->GPU bounce of the labels per clip (:456, :439).  Here labels stay on the device,
// channel-last [clip, frame, N, C] float32, and each output row is the sum of <= kw gathered rows.
// HBM/L2-bound: writes N*C*4 bytes per (clip, frame); gathered reads mostly hit L2 (one frame of
// labels is N*C*4 = 627 KB at config 2).  Frames must be processed in order (frame t reads the
// propagated labels of frames < t), so there is one launch per target frame, all clips at once.
#include "common.cuh"

namespace timet {

// One thread-block CLUSTER per clip walks the target frames in order; a cluster barrier (release/acquire)
// between frames replaces a kernel launch per frame: frame t of a clip only depends on earlier frames of
// the SAME clip, whose labels (n_frames * N * C * 4 bytes, 5 MB at config 2) stay in L2 while the cluster works.
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_nctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}

constexpr int GA_THREADS = 512;
constexpr int GA_WARPS = GA_THREADS / 32;

template <typename V>
__device__ __forceinline__ V ga_zero();
template <>
__device__ __forceinline__ float ga_zero<float>() { return 0.f; }
template <>
__device__ __forceinline__ float4 ga_zero<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void ga_fma(float &a, float w, float l) { a = fmaf(w, l, a); }
__device__ __forceinline__ void ga_fma(float4 &a, float w, const float4 l) {
    a.x = fmaf(w, l.x, a.x); a.y = fmaf(w, l.y, a.y); a.z = fmaf(w, l.z, a.z); a.w = fmaf(w, l.w, a.w);
}

// acc += sum_m w[m] * labels_row(k[m]) for the n_c (<= 32) entries held by the lanes of the warp (lane m: entry m).
// GB gathered rows are in flight together (coherent L2 loads: earlier frames were written by other CTAs of this launch).
// L1: gathered rows may be served from L1.  Safe when no 128-byte line holds data of two frames (see the launcher): a row is
// written once, in the phase of its frame, and only read in later phases (after a cluster barrier / a kernel boundary), so
// an SM never caches a line before its last write.  Neighbouring queries gather neighbouring rows: without L1 every one of
// them goes to L2.
template <typename V, int GB, bool L1>
__device__ __forceinline__ void ga_rows(V &acc0, V &acc1, float w_c, int32_t k_c, int n_c, const float *clip_base, int C, int c,
                                        int c2, bool one, bool two) {
    for (int m0 = 0; m0 < n_c; m0 += GB) {
        float wm[GB];
        V l0[GB], l1[GB];
#pragma unroll
        for (int u = 0; u < GB; ++u) {
            const int m = min(m0 + u, 31);
            wm[u] = __shfl_sync(0xffffffffu, w_c, m);
            const int32_t km = __shfl_sync(0xffffffffu, k_c, m);
            const V *row = reinterpret_cast<const V *>(clip_base + (int64_t)km * C);
            const bool on = m0 + u < n_c;
            l0[u] = (on && one) ? (L1 ? __ldca(row + c) : __ldcg(row + c)) : ga_zero<V>();
            l1[u] = (on && two) ? (L1 ? __ldca(row + c2) : __ldcg(row + c2)) : ga_zero<V>();
        }
#pragma unroll
        for (int u = 0; u < GB; ++u) {
            if (m0 + u < n_c) { ga_fma(acc0, wm[u], l0[u]); ga_fma(acc1, wm[u], l1[u]); }
        }
    }
}

// Wide rows are rare (degenerate features): kept out of line so that they do not cost the hot path registers.
template <typename V>
__device__ __noinline__ void ga_rows_wide(V &acc0, V &acc1, const float *__restrict__ wide_w, const int32_t *__restrict__ wide_k,
                                          int wide_off, int n_row, const float *clip_base, int C, int c, int c2, bool one, bool two,
                                          int lane) {
    for (int base = 0; base < n_row; base += 32) {
        const bool in = base + lane < n_row;
        const float w_c = in ? __ldg(wide_w + wide_off + base + lane) : 0.f;
        const int32_t k_c = in ? __ldg(wide_k + wide_off + base + lane) : 0;
        ga_rows<V, 2, false>(acc0, acc1, w_c, k_c, min(32, n_row - base), clip_base, C, c, c2, one, two);
    }
}

// V = float4 (C % 4 == 0) or float.  One WARP per query: lane m holds (weight, key) m of the sparse row,
// broadcast by shuffle; every lane then owns channel vectors c = lane, lane+32, ... and issues the <= 4 gathered
// row loads of a batch back to back (coherent L2 loads: earlier frames were written by other CTAs of this
// launch).  The next query's sparse row is prefetched while the current one is gathered.
// CLUSTER = false: the same body for ONE target frame per launch with `vcs` plain CTAs per clip (few clips: a
// cluster of 8 CTAs per clip would leave most SMs idle; frames then cost one launch each).
template <typename V, bool CLUSTER, int GB, bool L1>
__global__ void __launch_bounds__(GA_THREADS, 2)
ff_gather_kernel(float *__restrict__ labels, int64_t *__restrict__ hard, const float *__restrict__ sel_w,
                 const int32_t *__restrict__ sel_k, const int32_t *__restrict__ sel_cnt, const float *__restrict__ wide_w,
                 const int32_t *__restrict__ wide_k, int n_frames, int N, int C, int nT, int kw, int t_begin, int t_first,
                 int t_last, int vcs) {
    const unsigned cs = CLUSTER ? cluster_nctarank() : (unsigned)vcs;
    const unsigned cr = CLUSTER ? cluster_ctarank() : (blockIdx.x % (unsigned)vcs);
    const int clip = blockIdx.x / cs;
    constexpr int VW = sizeof(V) / sizeof(float);
    const int CV = C / VW;
    const int lane = threadIdx.x & 31;
    const int wid = cr * GA_WARPS + (threadIdx.x >> 5), nw = cs * GA_WARPS;
    float *clip_base = labels + (int64_t)clip * n_frames * N * C;
    for (int t = t_first; t <= t_last; ++t) {
        const int64_t q0 = ((int64_t)clip * nT + (t - t_begin)) * N;
        int i = wid;
        int cnt = 0;
        float w_l = 0.f;
        int32_t k_l = 0;
        if (i < N) {
            cnt = __ldg(sel_cnt + q0 + i);
            if (lane < kw) { w_l = __ldg(sel_w + (q0 + i) * kw + lane); k_l = __ldg(sel_k + (q0 + i) * kw + lane); }
        }
        while (i < N) {
            const int inext = i + nw;
            int cnt_n = 0;
            float w_n = 0.f;
            int32_t k_n = 0;
            if (inext < N) {   // prefetch the next sparse row
                cnt_n = __ldg(sel_cnt + q0 + inext);
                if (lane < kw) { w_n = __ldg(sel_w + (q0 + inext) * kw + lane); k_n = __ldg(sel_k + (q0 + inext) * kw + lane); }
            }
            V *dst = reinterpret_cast<V *>(clip_base + ((int64_t)t * N + i) * C);
            for (int c0 = 0; c0 < CV; c0 += 64) {          // warp-uniform: the shuffles below need every lane
                const int c = c0 + lane, c2 = c + 32;
                const bool one = c < CV, two = c2 < CV;
                V acc0 = ga_zero<V>(), acc1 = ga_zero<V>();
                // every lane holds the same cnt; taking it from a warp collective tells the compiler so (no per-shuffle
                // reconvergence code around the loop below)
                const int cnt_u = __reduce_max_sync(0xffffffffu, cnt);
                ga_rows<V, GB, L1>(acc0, acc1, w_l, k_l, cnt_u, clip_base, C, c, c2, one, two);
                if (cnt_u < 0)   // wide row (more than kw survivors: exact tie sets): -cnt entries in the pool at offset sel_k[0]
                    ga_rows_wide<V>(acc0, acc1, wide_w, wide_k, __shfl_sync(0xffffffffu, k_l, 0), -cnt_u, clip_base, C, c, c2, one, two, lane);
                if (one) dst[c] = acc0;
                if (two) dst[c2] = acc1;
            }
            i = inext; cnt = cnt_n; w_l = w_n; k_l = k_n;
        }
        if (CLUSTER) cluster_barrier();
    }
    if (hard && t_last == n_frames - 1) {   // argmax over channels of the last frame, lowest index on ties (time_tuning.py:296)
        // the rows read below were stored by other lanes of this warp (float4 stores, scalar loads): order them
        if (!CLUSTER) { __threadfence_block(); __syncwarp(); }
        for (int i = wid; i < N; i += nw) {
            const float *row = clip_base + ((int64_t)(n_frames - 1) * N + i) * C;
            float best = -INFINITY;
            int bi = 0x7fffffff;
            for (int c = lane; c < C; c += 32) {
                const float v = __ldcg(row + c);
                if (v > best || (v == best && c < bi)) { best = v; bi = c; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) hard[(int64_t)clip * N + i] = (bi == 0x7fffffff) ? 0 : (int64_t)bi;
        }
    }
}

template <int GBSEL, bool L1>
static int ff_gather_launch_gb(const timet_ff_params &p, const FFLayout &L, float *labels, int64_t *hard, const char *ws,
                               cudaStream_t st) {
    const float *sel_w = reinterpret_cast<const float *>(ws + L.off_sel_w);
    const int32_t *sel_k = reinterpret_cast<const int32_t *>(ws + L.off_sel_k);
    const int32_t *sel_cnt = reinterpret_cast<const int32_t *>(ws + L.off_sel_cnt);
    const float *wide_w = reinterpret_cast<const float *>(ws + L.off_wide_w);
    const int32_t *wide_k = reinterpret_cast<const int32_t *>(ws + L.off_wide_k);
    const bool vec = (p.n_channels % 4 == 0) && ((reinterpret_cast<uintptr_t>(labels) & 15) == 0);
    const int nfr = p.n_frames, nch = p.n_channels, tb = p.t_begin;
    if ((int64_t)p.n_clips * 8 >= num_sms() / 2) {
        // cluster size: as many CTAs per clip as useful (<= 8, power of two) without exceeding ~2 CTAs per SM in total
        int cs = 8;
        while (cs > 1 && ((int64_t)p.n_clips * cs > 2 * (int64_t)num_sms() || (int64_t)L.N < (int64_t)cs * GA_WARPS)) cs >>= 1;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(p.n_clips * cs));
        cfg.blockDim = dim3(GA_THREADS);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (vec)
            TIMET_CUDA(cudaLaunchKernelEx(&cfg, ff_gather_kernel<float4, true, GBSEL, L1>, labels, hard, sel_w, sel_k, sel_cnt, wide_w, wide_k, nfr, L.N, nch,
                                          L.nT, L.kw, tb, tb, nfr - 1, cs));
        else
            TIMET_CUDA(cudaLaunchKernelEx(&cfg, ff_gather_kernel<float, true, 4, false>, labels, hard, sel_w, sel_k, sel_cnt, wide_w, wide_k, nfr, L.N, nch,
                                          L.nT, L.kw, tb, tb, nfr - 1, cs));
        TIMET_LAUNCHED();
        return TIMET_OK;
    }
    // few clips: one launch per target frame, every SM busy
    int vcs = (int)((2 * (int64_t)num_sms() + p.n_clips - 1) / p.n_clips);
    const int max_useful = (L.N + GA_WARPS - 1) / GA_WARPS;
    if (vcs > max_useful) vcs = max_useful;
    if (vcs < 1) vcs = 1;
    for (int t = tb; t < nfr; ++t) {
        if (vec)
            ff_gather_kernel<float4, false, GBSEL, L1><<<p.n_clips * vcs, GA_THREADS, 0, st>>>(labels, hard, sel_w, sel_k, sel_cnt, wide_w, wide_k, nfr, L.N, nch,
                                                                                  L.nT, L.kw, tb, t, t, vcs);
        else
            ff_gather_kernel<float, false, 4, false><<<p.n_clips * vcs, GA_THREADS, 0, st>>>(labels, hard, sel_w, sel_k, sel_cnt, wide_w, wide_k, nfr, L.N, nch,
                                                                                 L.nT, L.kw, tb, t, t, vcs);
        TIMET_LAUNCHED();
    }
    return TIMET_OK;
}

// Gathered rows in flight per batch (env TIMET_GATHER_BATCH).  Measured at BASELINE configs[1] (topk 5) on the final kernel
// (62 registers at 4): 4 -> 0.149 ms, 5 -> 0.141 ms (one round trip per query), 7 -> 0.146 ms; before the warp-uniform row
// loop (spills) the order was reversed (0.180 / 0.220 / 0.332).  Default: 5 for the training top-k of 5, else 4.
int ff_gather_launch(const timet_ff_params &p, const FFLayout &L, float *labels, int64_t *hard, const char *ws,
                     cudaStream_t st) {
    const int gb = env_cfg().gather_batch ? env_cfg().gather_batch : (p.topk == 5 ? 5 : 4);
    // L1-cached row loads only when a frame of labels is a whole number of 128-byte lines (no line shared by two frames)
    const bool l1 = env_cfg().gather_l1 && ((int64_t)L.N * p.n_channels * 4) % 128 == 0 && (reinterpret_cast<uintptr_t>(labels) & 127) == 0;
    if (l1) {
        if (gb == 5) return ff_gather_launch_gb<5, true>(p, L, labels, hard, ws, st);
        if (gb >= 7) return ff_gather_launch_gb<7, true>(p, L, labels, hard, ws, st);
        return ff_gather_launch_gb<4, true>(p, L, labels, hard, ws, st);
    }
    if (gb == 5) return ff_gather_launch_gb<5, false>(p, L, labels, hard, ws, st);
    if (gb >= 7) return ff_gather_launch_gb<7, false>(p, L, labels, hard, ws, st);
    return ff_gather_launch_gb<4, false>(p, L, labels, hard, ws, st);
}

}  // namespace timet
