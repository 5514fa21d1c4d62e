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

template <bool VEC>
__global__ void __launch_bounds__(GA_THREADS)
ff_gather_kernel(float *__restrict__ labels, int64_t *__restrict__ hard, const float *__restrict__ sel_w,
                 const int32_t *__restrict__ sel_k, const int32_t *__restrict__ sel_cnt, int n_frames, int N, int C,
                 int nT, int kw, int t_begin) {
    const unsigned cs = cluster_nctarank(), cr = cluster_ctarank();
    const int clip = blockIdx.x / cs;
    const int CV = VEC ? (C >> 2) : C;
    const int items = N * CV;
    float *clip_base = labels + (int64_t)clip * n_frames * N * C;
    for (int t = t_begin; t < n_frames; ++t) {
        const int64_t q0 = ((int64_t)clip * nT + (t - t_begin)) * N;
        for (int idx = cr * GA_THREADS + threadIdx.x; idx < items; idx += cs * GA_THREADS) {
            const int i = idx / CV, c = idx - i * CV;
            const int64_t q = q0 + i;
            const int cnt = __ldg(sel_cnt + q);
            const float *w = sel_w + q * kw;
            const int32_t *kk = sel_k + q * kw;
            if (VEC) {
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int m = 0; m < cnt; ++m) {
                    const float wm = __ldg(w + m);
                    // labels of earlier frames were written in this kernel by other CTAs: coherent (L2) loads
                    const float4 l = __ldcg(reinterpret_cast<const float4 *>(clip_base + (int64_t)__ldg(kk + m) * C) + c);
                    acc.x = fmaf(wm, l.x, acc.x); acc.y = fmaf(wm, l.y, acc.y);
                    acc.z = fmaf(wm, l.z, acc.z); acc.w = fmaf(wm, l.w, acc.w);
                }
                *(reinterpret_cast<float4 *>(clip_base + ((int64_t)t * N + i) * C) + c) = acc;
            } else {
                float acc = 0.f;
                for (int m = 0; m < cnt; ++m) acc = fmaf(__ldg(w + m), __ldcg(clip_base + (int64_t)__ldg(kk + m) * C + c), acc);
                clip_base[((int64_t)t * N + i) * C + c] = acc;
            }
        }
        cluster_barrier();
    }
    if (hard) {   // argmax over channels of the last frame, lowest index on ties (time_tuning.py:296)
        const int lane = threadIdx.x & 31;
        const int warps = cs * (GA_THREADS >> 5);
        for (int i = cr * (GA_THREADS >> 5) + (threadIdx.x >> 5); i < N; i += warps) {
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

int ff_gather_launch(const timet_ff_params &p, const FFLayout &L, float *labels, int64_t *hard, const char *ws,
                     cudaStream_t st) {
    const float *sel_w = reinterpret_cast<const float *>(ws + L.off_sel_w);
    const int32_t *sel_k = reinterpret_cast<const int32_t *>(ws + L.off_sel_k);
    const int32_t *sel_cnt = reinterpret_cast<const int32_t *>(ws + L.off_sel_cnt);
    const bool vec = (p.n_channels % 4 == 0) && ((reinterpret_cast<uintptr_t>(labels) & 15) == 0);
    const int CV = vec ? p.n_channels / 4 : p.n_channels;
    // cluster size: as many CTAs per clip as useful (<= 8, power of two) without exceeding ~2 CTAs per SM in total
    int cs = 8;
    while (cs > 1 && ((int64_t)p.n_clips * cs > 2 * (int64_t)num_sms() || (int64_t)L.N * CV < (int64_t)cs * GA_THREADS / 2)) cs >>= 1;
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
        TIMET_CUDA(cudaLaunchKernelEx(&cfg, ff_gather_kernel<true>, labels, hard, sel_w, sel_k, sel_cnt, (int)p.n_frames,
                                      L.N, (int)p.n_channels, L.nT, L.kw, (int)p.t_begin));
    else
        TIMET_CUDA(cudaLaunchKernelEx(&cfg, ff_gather_kernel<false>, labels, hard, sel_w, sel_k, sel_cnt, (int)p.n_frames,
                                      L.N, (int)p.n_channels, L.nT, L.kw, (int)p.t_begin));
    TIMET_LAUNCHED();
    return TIMET_OK;
}

}  // namespace timet
