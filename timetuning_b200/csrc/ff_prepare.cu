// FF stage 1: one pass over the features per step.
//
// Reference: F.normalize(feat, dim=D, p=2) on the target and on ALL context frames, recomputed
// for every target frame (/root/reference/mask_propagation.py:418-419).  Here each row is visited once:
//   inv  [rows]      float32  1 / max(||x||_2, 1e-12)         -- scales the exact fp32 similarity (common.cuh)
//   fn16 [rows, Dp]  float16  x / max(||x||_2, 1e-12) rounded -- K-major operand of the tcgen05 nomination GEMM (TMA-loaded)
// Dp = dim rounded up to 64, zero padded.  The fp32 rows themselves are NOT copied: the exact re-evaluation reads the
// caller's tensor in place (only for dim % 4 != 0 a zero-padded fp32 copy is written so that rows are float4-addressable).
// HBM-bound: reads 4*D, writes 2*Dp + 4 bytes per row.  One warp per row, 128-bit accesses.
#include "common.cuh"

namespace timet {

__device__ __forceinline__ void prep_store(__half *d16, int i, float4 v, float denom) {
    v.x = __fdiv_rn(v.x, denom); v.y = __fdiv_rn(v.y, denom);
    v.z = __fdiv_rn(v.z, denom); v.w = __fdiv_rn(v.w, denom);
    const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t *>(&lo);
    pk.y = *reinterpret_cast<const uint32_t *>(&hi);
    reinterpret_cast<uint2 *>(d16)[i] = pk;
}

// Fast path: dim % 4 == 0 and dim <= 128 * KEEP.  The lane's slice of a row lives in KEEP float4 registers (read once),
// and TWO rows of a warp are in flight together: the kernel is bound by bytes in flight per SM (DRAM latency x
// bandwidth), so registers are spent on loads, not on a worst-case slice (the round-1 kernel kept 8 float4 for every dim:
// 80 registers, 24 resident warps, 55 % of the HBM peak).
template <int KEEP>
__global__ void __launch_bounds__(256) ff_prepare_vec_kernel(const float *__restrict__ feats, float *__restrict__ inv,
                                                             __half *__restrict__ fn16, int64_t rows, int dim, int Dp) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int n4 = dim >> 2, np4 = Dp >> 2;
    for (int64_t row = 2 * warp; row < rows; row += 2 * nwarps) {
        const bool two = row + 1 < rows;
        const float4 *s0 = reinterpret_cast<const float4 *>(feats + row * dim);
        const float4 *s1 = reinterpret_cast<const float4 *>(feats + (row + (two ? 1 : 0)) * dim);
        float4 a[KEEP], b[KEEP];
#pragma unroll
        for (int u = 0; u < KEEP; ++u) {
            const int i = lane + 32 * u;
            a[u] = (i < n4) ? __ldg(s0 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            b[u] = (i < n4) ? __ldg(s1 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int u = 0; u < KEEP; ++u) {
            sa += (a[u].x * a[u].x + a[u].y * a[u].y) + (a[u].z * a[u].z + a[u].w * a[u].w);
            sb += (b[u].x * b[u].x + b[u].y * b[u].y) + (b[u].z * b[u].z + b[u].w * b[u].w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sa += __shfl_xor_sync(0xffffffffu, sa, o);
            sb += __shfl_xor_sync(0xffffffffu, sb, o);
        }
        const float da = fmaxf(sqrtf(sa), 1e-12f), db = fmaxf(sqrtf(sb), 1e-12f);
        if (lane == 0) {
            inv[row] = __fdiv_rn(1.0f, da);
            if (two) inv[row + 1] = __fdiv_rn(1.0f, db);
        }
        __half *d0 = fn16 + row * Dp, *d1 = fn16 + (row + 1) * Dp;
#pragma unroll
        for (int u = 0; u < KEEP; ++u) {
            const int i = lane + 32 * u;
            if (i < np4) {                                  // a[u] / b[u] are zero past dim: the padding
                prep_store(d0, i, a[u], da);
                if (two) prep_store(d1, i, b[u], db);
            }
        }
    }
}

// General path (any dim): lane-strided loops; for dim % 4 != 0 also writes the float4-addressable fp32 copy.
__global__ void __launch_bounds__(256) ff_prepare_kernel(const float *__restrict__ feats, float *__restrict__ xpad,
                                                         float *__restrict__ inv, __half *__restrict__ fn16, int64_t rows,
                                                         int dim, int Dp) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const bool vec = (dim & 3) == 0;
    for (int64_t row = warp; row < rows; row += nwarps) {
        const float *src = feats + row * dim;
        float ss = 0.f;
        if (vec) {
            const float4 *s4 = reinterpret_cast<const float4 *>(src);
            for (int i = lane; i < (dim >> 2); i += 32) {
                const float4 v = s4[i];
                ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
            }
        } else {
            for (int i = lane; i < dim; i += 32) { const float v = src[i]; ss = fmaf(v, v, ss); }
        }
        ss = warp_sum(ss);
        const float denom = fmaxf(sqrtf(ss), 1e-12f);
        if (lane == 0) inv[row] = __fdiv_rn(1.0f, denom);
        __half *d16 = fn16 + row * Dp;
        for (int i = lane; i < (Dp >> 2); i += 32) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const int d = i << 2;
            if (vec) {
                if (d < dim) v = reinterpret_cast<const float4 *>(src)[i];
            } else {
                if (d + 0 < dim) v.x = src[d + 0];
                if (d + 1 < dim) v.y = src[d + 1];
                if (d + 2 < dim) v.z = src[d + 2];
                if (d + 3 < dim) v.w = src[d + 3];
                reinterpret_cast<float4 *>(xpad + row * Dp)[i] = v;     // float4-addressable copy of the raw row
            }
            prep_store(d16, i, v, denom);
        }
    }
}

template <int KEEP>
static void prep_launch_vec(const float *feats, float *inv, __half *fn16, int64_t rows, int dim, int Dp, cudaStream_t st) {
    int64_t blocks = (rows + 15) / 16;                      // 8 warps x 2 rows per CTA and sweep
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    ff_prepare_vec_kernel<KEEP><<<(int)blocks, 256, 0, st>>>(feats, inv, fn16, rows, dim, Dp);
}

int ff_prepare_launch(const timet_ff_params &p, const FFLayout &L, const float *feats, char *ws, cudaStream_t st) {
    float *xpad = reinterpret_cast<float *>(ws + L.off_xpad);
    float *inv = reinterpret_cast<float *>(ws + L.off_inv);
    __half *fn16 = reinterpret_cast<__half *>(ws + L.off_fn16);
    if (!(p.dim & 3) && (reinterpret_cast<uintptr_t>(feats) & 15)) {
        set_error("ff_prepare: feats must be 16-byte aligned (dim %% 4 == 0 rows are read in place as float4)");
        return TIMET_ERR_INVALID;
    }
    const int keep = (L.Dp / 4 + 31) / 32;                  // float4 per lane covering the padded row
    if (!(p.dim & 3) && keep <= 8) {
        switch (keep) {
            case 1: prep_launch_vec<1>(feats, inv, fn16, L.rows, p.dim, L.Dp, st); break;
            case 2: prep_launch_vec<2>(feats, inv, fn16, L.rows, p.dim, L.Dp, st); break;
            case 3: prep_launch_vec<3>(feats, inv, fn16, L.rows, p.dim, L.Dp, st); break;
            case 4: prep_launch_vec<4>(feats, inv, fn16, L.rows, p.dim, L.Dp, st); break;
            case 5: case 6: prep_launch_vec<6>(feats, inv, fn16, L.rows, p.dim, L.Dp, st); break;
            default: prep_launch_vec<8>(feats, inv, fn16, L.rows, p.dim, L.Dp, st); break;
        }
    } else {
        int64_t blocks = (L.rows + 7) / 8;
        const int64_t cap = (int64_t)num_sms() * 8;
        if (blocks > cap) blocks = cap;
        ff_prepare_kernel<<<(int)blocks, 256, 0, st>>>(feats, xpad, inv, fn16, L.rows, p.dim, L.Dp);
    }
    TIMET_LAUNCHED();
    // the 256 slack rows behind the last frame are read by out-of-range TMA boxes: keep them finite
    TIMET_CUDA(cudaMemsetAsync(fn16 + L.rows * L.Dp, 0, (size_t)256 * L.Dp * sizeof(__half), st));
    return TIMET_OK;
}

}  // namespace timet
