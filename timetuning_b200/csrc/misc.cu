// Small routines named by the north star: restrict_neighborhood and norm_mask.
#include "common.cuh"

namespace timet {

// mask[i*w + j, p*w + q] = 1 iff |i-p| <= s and |j-q| <= s   (/root/reference/mask_propagation.py:377-391;
// the reference fills it with a 4-deep Python loop, 0.57 s at 28x28 — SURVEY.md §8a4)
__global__ void restrict_neighborhood_kernel(float *out, int h, int w, int s) {
    const int64_t n = (int64_t)h * w;
    const int64_t total = n * n;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(idx / n), k = (int)(idx % n);
        const int dr = q / w - k / w, dc = q % w - k % w;
        out[idx] = (dr <= s && dr >= -s && dc <= s && dc >= -s) ? 1.f : 0.f;
    }
}

// Per channel: if max > 0: m -= min; m /= max(m)  else 0   (/root/reference/mask_propagation.py:363-374).
// A constant positive channel gives 0/0 = NaN exactly as the reference does.
template <typename T>
__global__ void __launch_bounds__(256) norm_mask_kernel(const T *in, T *out, int64_t hw) {
    __shared__ T smin[8], smax[8];
    const T *src = in + (int64_t)blockIdx.x * hw;
    T *dst = out + (int64_t)blockIdx.x * hw;
    T mn = src[0], mx = src[0];
    for (int64_t i = threadIdx.x; i < hw; i += blockDim.x) {
        const T v = src[i];
        mn = v < mn ? v : mn;
        mx = v > mx ? v : mx;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const T a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = mn; smax[threadIdx.x >> 5] = mx; }
    __syncthreads();
    for (int wv = 0; wv < 8; ++wv) {
        mn = smin[wv] < mn ? smin[wv] : mn;
        mx = smax[wv] > mx ? smax[wv] : mx;
    }
    const bool live = mx > (T)0;
    const T range = mx - mn;               // == max(m - min): subtraction is monotone
    for (int64_t i = threadIdx.x; i < hw; i += blockDim.x) dst[i] = live ? (src[i] - mn) / range : (T)0;
}

}  // namespace timet

using namespace timet;

extern "C" {

int timet_restrict_neighborhood(int h, int w, int radius, float *mask_out, timet_stream_t stream) {
    TIMET_CHECK_ARG(mask_out != nullptr, "restrict_neighborhood: NULL output");
    TIMET_CHECK_ARG(h >= 1 && w >= 1 && radius >= 0, "restrict_neighborhood: bad arguments h=%d w=%d s=%d", h, w, radius);
    const int64_t total = (int64_t)h * w * h * w;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    restrict_neighborhood_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(mask_out, h, w, radius);
    TIMET_LAUNCHED();
    return TIMET_OK;
}

int timet_norm_mask(const void *mask, void *out, int n_channels, int64_t hw, int dtype_bytes, timet_stream_t stream) {
    TIMET_CHECK_ARG(mask && out, "norm_mask: NULL pointer");
    TIMET_CHECK_ARG(n_channels >= 1 && hw >= 1, "norm_mask: bad shape C=%d HW=%lld", n_channels, (long long)hw);
    TIMET_CHECK_ARG(dtype_bytes == 4 || dtype_bytes == 8, "norm_mask: dtype_bytes must be 4 or 8");
    if (dtype_bytes == 4)
        norm_mask_kernel<float><<<n_channels, 256, 0, (cudaStream_t)stream>>>((const float *)mask, (float *)out, hw);
    else
        norm_mask_kernel<double><<<n_channels, 256, 0, (cudaStream_t)stream>>>((const double *)mask, (double *)out, hw);
    TIMET_LAUNCHED();
    return TIMET_OK;
}

}

// ------------------------------------------------------------------ eval tail
// mask_propagation.py:822-824: stack -> F.interpolate(size=(R, R), mode="bilinear", align_corners=False) -> max over
// channels.  Fused: one thread per output pixel interpolates the C channels of its 4 taps straight from the
// channel-last label frames and keeps the arg max (lowest channel on ties); the [T, C, R, R] tensor never exists.
namespace timet {

__global__ void __launch_bounds__(256)
upsample_argmax_kernel(const float *__restrict__ labels, int64_t *__restrict__ out, int T, int h, int w, int C, int R_h,
                       int R_w, int64_t frame_stride) {
    const int64_t total = (int64_t)T * R_h * R_w;
    const float sy = (float)h / (float)R_h, sx = (float)w / (float)R_w;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(idx % R_w);
        const int y = (int)((idx / R_w) % R_h);
        const int t = (int)(idx / ((int64_t)R_w * R_h));
        // PyTorch area_pixel_compute_source_index, align_corners = False: src = max((dst + 0.5) * scale - 0.5, 0)
        const float fy = fmaxf(((float)y + 0.5f) * sy - 0.5f, 0.f), fx = fmaxf(((float)x + 0.5f) * sx - 0.5f, 0.f);
        const int y0 = (int)fy, x0 = (int)fx;
        const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
        const float ly = fy - (float)y0, lx = fx - (float)x0;
        const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
        const float *f = labels + (int64_t)t * frame_stride;
        const float *p00 = f + (int64_t)(y0 * w + x0) * C, *p01 = f + (int64_t)(y0 * w + x1) * C;
        const float *p10 = f + (int64_t)(y1 * w + x0) * C, *p11 = f + (int64_t)(y1 * w + x1) * C;
        float best = -INFINITY;
        int bi = 0;
        for (int c = 0; c < C; ++c) {
            const float v = w00 * __ldg(p00 + c) + w01 * __ldg(p01 + c) + w10 * __ldg(p10 + c) + w11 * __ldg(p11 + c);
            if (v > best) { best = v; bi = c; }
        }
        out[idx] = bi;
    }
}

}  // namespace timet

extern "C" int timet_upsample_argmax(const float *labels, int n_frames, int h, int w, int n_channels, int out_h, int out_w,
                                     int64_t frame_stride, int64_t *out, timet_stream_t stream) {
    using namespace timet;
    TIMET_CHECK_ARG(labels && out, "upsample_argmax: NULL pointer");
    TIMET_CHECK_ARG(n_frames >= 1 && h >= 1 && w >= 1 && n_channels >= 1 && out_h >= 1 && out_w >= 1, "upsample_argmax: bad shape");
    const int64_t total = (int64_t)n_frames * out_h * out_w;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    upsample_argmax_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(labels, out, n_frames, h, w, n_channels, out_h, out_w,
                                                                         frame_stride);
    TIMET_LAUNCHED();
    return TIMET_OK;
}
