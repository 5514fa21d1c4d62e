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

template <bool VEC>
__global__ void __launch_bounds__(256)
ff_gather_kernel(float *__restrict__ labels, const float *__restrict__ sel_w, const int32_t *__restrict__ sel_k,
                 const int32_t *__restrict__ sel_cnt, int n_clips, int n_frames, int N, int C, int nT, int kw,
                 int t, int t_begin) {
    const int CV = VEC ? (C >> 2) : C;
    const int64_t total = (int64_t)n_clips * N * CV;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % CV);
        const int64_t qi = idx / CV;
        const int i = (int)(qi % N);
        const int clip = (int)(qi / N);
        const int64_t q = ((int64_t)clip * nT + (t - t_begin)) * N + i;
        const int cnt = __ldg(sel_cnt + q);
        const float *w = sel_w + q * kw;
        const int32_t *kk = sel_k + q * kw;
        const float *clip_base = labels + (int64_t)clip * n_frames * N * C;
        if (VEC) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int m = 0; m < cnt; ++m) {
                const float wm = __ldg(w + m);
                const float4 l = *(reinterpret_cast<const float4 *>(clip_base + (int64_t)__ldg(kk + m) * C) + c);
                acc.x = fmaf(wm, l.x, acc.x); acc.y = fmaf(wm, l.y, acc.y);
                acc.z = fmaf(wm, l.z, acc.z); acc.w = fmaf(wm, l.w, acc.w);
            }
            *(reinterpret_cast<float4 *>(labels + (((int64_t)clip * n_frames + t) * N + i) * C) + c) = acc;
        } else {
            float acc = 0.f;
            for (int m = 0; m < cnt; ++m) acc = fmaf(__ldg(w + m), clip_base[(int64_t)__ldg(kk + m) * C + c], acc);
            labels[(((int64_t)clip * n_frames + t) * N + i) * C + c] = acc;
        }
    }
}

// argmax over channels of the last frame, lowest index on ties (time_tuning.py:296)
__global__ void __launch_bounds__(256)
ff_argmax_kernel(const float *__restrict__ labels, int64_t *__restrict__ hard, int n_clips, int n_frames, int N, int C) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t qi = warp; qi < (int64_t)n_clips * N; qi += nwarps) {
        const int clip = (int)(qi / N), i = (int)(qi % N);
        const float *row = labels + (((int64_t)clip * n_frames + (n_frames - 1)) * N + i) * C;
        float best = -INFINITY;
        int bi = 0x7fffffff;
        for (int c = lane; c < C; c += 32) {
            const float v = row[c];
            if (v > best || (v == best && c < bi)) { best = v; bi = c; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) hard[qi] = (bi == 0x7fffffff) ? 0 : (int64_t)bi;
    }
}

int ff_gather_launch(const timet_ff_params &p, const FFLayout &L, float *labels, int64_t *hard, const char *ws,
                     cudaStream_t st) {
    const float *sel_w = reinterpret_cast<const float *>(ws + L.off_sel_w);
    const int32_t *sel_k = reinterpret_cast<const int32_t *>(ws + L.off_sel_k);
    const int32_t *sel_cnt = reinterpret_cast<const int32_t *>(ws + L.off_sel_cnt);
    const bool vec = (p.n_channels % 4 == 0) && ((reinterpret_cast<uintptr_t>(labels) & 15) == 0);
    const int CV = vec ? p.n_channels / 4 : p.n_channels;
    const int64_t total = (int64_t)p.n_clips * L.N * CV;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    for (int t = p.t_begin; t < p.n_frames; ++t) {
        if (vec)
            ff_gather_kernel<true><<<(int)blocks, 256, 0, st>>>(labels, sel_w, sel_k, sel_cnt, p.n_clips, p.n_frames,
                                                                L.N, p.n_channels, L.nT, L.kw, t, p.t_begin);
        else
            ff_gather_kernel<false><<<(int)blocks, 256, 0, st>>>(labels, sel_w, sel_k, sel_cnt, p.n_clips, p.n_frames,
                                                                 L.N, p.n_channels, L.nT, L.kw, t, p.t_begin);
        TIMET_LAUNCHED();
    }
    if (hard) {
        int64_t ab = ((int64_t)p.n_clips * L.N + 7) / 8;
        if (ab > cap) ab = cap;
        ff_argmax_kernel<<<(int)ab, 256, 0, st>>>(labels, hard, p.n_clips, p.n_frames, L.N, p.n_channels);
        TIMET_LAUNCHED();
    }
    return TIMET_OK;
}

}  // namespace timet
