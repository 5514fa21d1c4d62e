// Cosine scores  S = normalize(x) @ prototypes^T  (no-grad branch of TimeT.get_feature_prototype_similarity,
// /root/reference/time_tuning.py:130-141: F.normalize(x, dim=-1) then torch.mm with the [K, dh] prototypes).
//
// SURVEY.md §8f item 2 ("next"): the reference issues an fp32 SIMT GEMM (cuBLAS picks a 32x64x16 SIMT kernel,
// 80 us at 25 088 x 256 x 200) plus three elementwise launches.  The scores feed exp(s / 0.05), so a 1e-3
// relative error (TF32 / bf16) is not acceptable; here the GEMM runs on the tcgen05 tensor cores with the
// fp16 SPLIT  x^ = hi + lo,  p = phi + plo  (hi = fp16(x^), lo = fp16(x^ - hi)):
//     x^ . p  ~=  hi.phi + hi.plo + lo.phi        (the dropped lo.plo term is <= 2^-22 relative)
// as ONE K = 3*dh accumulation in TMEM (fp32).  Measured deviation from the fp32 torch path: ~1e-7 abs.
//   scores_prep_kernel   : per row  x^ = x / max(||x||, 1e-12)  ->  [hi | lo]  fp16, 2*dhp wide (dhp = dh up to 64)
//   scores_gemm_kernel   : one CTA per 128 rows x (<= 256 prototypes): TMA ring (A 128x64 + B Npx64 per stage),
//                          tcgen05.mma M=128 N=Np K=16, epilogue tcgen05.ld -> fp32 global rows
#include "ff_tc_dev.cuh"

namespace timet {

constexpr int SC_THREADS = 192;       // warp 0 TMA, warp 1 MMA (+TMEM alloc), warps 2-5 epilogue
constexpr int SC_MAX_STAGES = 4;
// Ring depth 2 (85 KB per CTA at K = 200) lets TWO CTAs share an SM: one CTA's epilogue (TMEM -> shared -> global) and
// prologue run under the other's TMA / MMA main loop; depth 4 = one CTA per SM, everything in sequence (round 1).

constexpr int SC_MAX_INPUTS = 4;
struct ScInputs {
    const float *x[SC_MAX_INPUTS];     // n row blocks of rows_each rows; output rows are their concatenation
    int64_t rows_each;
};

__device__ __forceinline__ uint2 pack_half4(float a, float b, float c, float d) {
    const __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, d);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t *>(&lo);
    pk.y = *reinterpret_cast<const uint32_t *>(&hi);
    return pk;
}

// One warp per row: x^ = x / max(||x||, 1e-12) (normalize) or x itself -> [hi | lo] fp16, 2 * dhp wide; src == nullptr
// writes a zero row.  dh % 4 == 0: 128-bit loads, the lane's slice kept in registers (dh <= 512), 64-bit stores of 4 halfs.
__device__ __forceinline__ void scores_prep_row(const float *__restrict__ src, __half *__restrict__ dst, int dh, int dhp,
                                                bool normalize, int lane) {
    constexpr int KEEP = 4;
    if (src == nullptr) {
        for (int i = lane; i < (dhp >> 1); i += 32) reinterpret_cast<uint2 *>(dst)[i] = make_uint2(0u, 0u);   // 2 * dhp halfs
        return;
    }
    if ((dh & 3) == 0 && (dh >> 2) <= 32 * KEEP && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        float4 v[KEEP];
        float ss = 0.f;
#pragma unroll
        for (int u = 0; u < KEEP; ++u) {
            const int i = lane + 32 * u;
            v[u] = (i < (dh >> 2)) ? __ldg(reinterpret_cast<const float4 *>(src) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            ss = fmaf(v[u].x, v[u].x, ss); ss = fmaf(v[u].y, v[u].y, ss); ss = fmaf(v[u].z, v[u].z, ss); ss = fmaf(v[u].w, v[u].w, ss);
        }
        ss = warp_sum(ss);
        const float denom = normalize ? fmaxf(sqrtf(ss), 1e-12f) : 1.f;
#pragma unroll
        for (int u = 0; u < KEEP; ++u) {
            const int i = lane + 32 * u;
            if (i < (dhp >> 2)) {
                const float a = __fdiv_rn(v[u].x, denom), b = __fdiv_rn(v[u].y, denom), c = __fdiv_rn(v[u].z, denom), d = __fdiv_rn(v[u].w, denom);
                const uint2 hi = pack_half4(a, b, c, d);
                const __half2 h01 = *reinterpret_cast<const __half2 *>(&hi.x), h23 = *reinterpret_cast<const __half2 *>(&hi.y);
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                reinterpret_cast<uint2 *>(dst)[i] = hi;
                reinterpret_cast<uint2 *>(dst + dhp)[i] = pack_half4(a - f01.x, b - f01.y, c - f23.x, d - f23.y);
            }
        }
        return;
    }
    float ss = 0.f;
    for (int i = lane; i < dh; i += 32) { const float v = src[i]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    const float denom = normalize ? fmaxf(sqrtf(ss), 1e-12f) : 1.f;
    for (int i = lane; i < dhp; i += 32) {
        const float v = (i < dh) ? __fdiv_rn(src[i], denom) : 0.f;
        const __half hi = __float2half_rn(v);
        dst[i] = hi;
        dst[dhp + i] = __float2half_rn(v - __half2float(hi));
    }
}

// ONE launch prepares both GEMM operands: logical rows [0, B) = feature rows (normalised) -> a2; [B, B + K) = the prototypes
// as given (time_tuning.py:138,140) -> b2; [B + K, B + Kp) = zero padding of b2 up to the MMA's N tile.
__global__ void __launch_bounds__(256)
scores_prep_kernel(ScInputs in, const float *__restrict__ prototypes, __half *__restrict__ a2, __half *__restrict__ b2, int64_t B,
                   int K, int Kp, int dh, int dhp) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = warp; row < B + Kp; row += nwarps) {
        if (row < B) {
            const int64_t blk = row / in.rows_each;
            scores_prep_row(in.x[blk] + (row - blk * in.rows_each) * dh, a2 + row * 2 * dhp, dh, dhp, true, lane);
        } else {
            const int64_t k = row - B;
            scores_prep_row(k < K ? prototypes + k * dh : nullptr, b2 + k * 2 * dhp, dh, dhp, false, lane);
        }
    }
}

struct __align__(8) ScCtl {
    uint64_t full[SC_MAX_STAGES], empty[SC_MAX_STAGES], tmem_full;
    uint32_t tmem_base;
};

template <int SC_STAGES>
__global__ void __launch_bounds__(SC_THREADS, SC_STAGES <= 2 ? 2 : 1)
scores_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   float *__restrict__ out, int64_t rows, int K, int Np, int dhp) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = 128u * 128u, b_bytes = (uint32_t)Np * 128u, stage_bytes = a_bytes + b_bytes;
    ScCtl *ctl = reinterpret_cast<ScCtl *>(smem + (size_t)SC_STAGES * stage_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128;
    const int n0 = blockIdx.y * 256;
    const int nkc = dhp / 64;               // 64-wide chunks per segment; 3 segments: (hi,phi) (hi,plo) (lo,phi)
    const int nsteps = 3 * nkc;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_a);
        ptx::prefetch_tensormap(&map_b);
        for (int s = 0; s < SC_STAGES; ++s) { ptx::mbar_init(&ctl->full[s], 1); ptx::mbar_init(&ctl->empty[s], 1); }
        ptx::mbar_init(&ctl->tmem_full, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc<256>(&ctl->tmem_base);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = ctl->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int step = 0; step < nsteps; ++step) {
                const int seg = step / nkc, kc = step - seg * nkc;
                const int a_col = (seg == 2 ? dhp : 0) + kc * 64;      // lo only in the third segment
                const int b_col = (seg == 1 ? dhp : 0) + kc * 64;      // plo only in the second segment
                ptx::mbar_wait(&ctl->empty[stage], phase ^ 1u);
                ptx::mbar_expect_tx(&ctl->full[stage], stage_bytes);
                uint8_t *st = smem + (size_t)stage * stage_bytes;
                ptx::tma_load_2d(st, &map_a, a_col, m0, &ctl->full[stage]);
                ptx::tma_load_2d(st + a_bytes, &map_b, b_col, n0, &ctl->full[stage]);
                if (++stage == SC_STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = ptx::umma_idesc_f16(128, Np);
            const uint64_t d0 = ptx::umma_desc_sw128(ptx::smem_u32(smem));
            uint32_t stage = 0, phase = 0;
            for (int step = 0; step < nsteps; ++step) {
                ptx::mbar_wait(&ctl->full[stage], phase);
                ptx::tc_fence_after();
                const uint64_t da = d0 + (uint64_t)((stage * stage_bytes) >> 4);
                const uint64_t db = da + (uint64_t)(a_bytes >> 4);
                ptx::umma_f16(tmem_base, da, db, idesc, step != 0);
                ptx::umma_f16(tmem_base, da + 2, db + 2, idesc, true);
                ptx::umma_f16(tmem_base, da + 4, db + 4, idesc, true);
                ptx::umma_f16(tmem_base, da + 6, db + 6, idesc, true);
                ptx::umma_commit(&ctl->empty[stage]);
                if (++stage == SC_STAGES) { stage = 0; phase ^= 1u; }
            }
            ptx::umma_commit(&ctl->tmem_full);
        }
    } else {
        // epilogue: warp w reads TMEM lanes 32*(w%4) .. +31 (rows m0 + lane index), 32 columns at a time
        const int q = warp & 3;
        if (lane == 0) ptx::mbar_wait(&ctl->tmem_full, 0);
        __syncwarp();
        ptx::tc_fence_after();
        const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16);
        // TMEM -> registers -> a 32 x 33 shared-memory tile per warp -> coalesced 128-byte row segments
        float *tile = reinterpret_cast<float *>(ctl + 1) + (size_t)q * 32 * 33;
        for (int c0 = 0; c0 < Np; c0 += 32) {
            uint32_t r[32];
            ptx::tmem_ld_32x32(t_acc + (uint32_t)c0, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) tile[lane * 33 + e] = __uint_as_float(r[e]);      // row = lane (conflict-free: stride 33)
            __syncwarp();
            const int col = n0 + c0 + lane;
            for (int rr = 0; rr < 32; ++rr) {
                const int64_t grow = (int64_t)m0 + q * 32 + rr;
                if (grow < rows && col < K) out[grow * K + col] = tile[rr * 33 + lane];
            }
            __syncwarp();
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<256>(tmem_base);
    }
}

static size_t scores_ws_bytes(int64_t B, int K, int dh) {
    const int dhp = (dh + 63) / 64 * 64;
    const int64_t Kp = (K + 255) / 256 * 256;
    return align_up((size_t)(B + 128) * 2 * dhp * sizeof(__half), 1024) + align_up((size_t)Kp * 2 * dhp * sizeof(__half), 1024);
}

}  // namespace timet

using namespace timet;

extern "C" {

size_t timet_cosine_scores_workspace_bytes(int64_t B, int K, int dh) {
    if (B < 1 || K < 1 || dh < 1) return 0;
    return scores_ws_bytes(B, K, dh);
}

int timet_cosine_scores(const float *x, const float *prototypes, int64_t B, int K, int dh, float *scores_out,
                        void *workspace, size_t workspace_bytes, timet_stream_t stream) {
    const float *xs[1] = {x};
    return timet_cosine_scores_multi(xs, 1, B, prototypes, K, dh, scores_out, workspace, workspace_bytes, stream);
}

int timet_cosine_scores_multi(const float *const *x_list, int n_x, int64_t rows_each, const float *prototypes, int K, int dh,
                              float *scores_out, void *workspace, size_t workspace_bytes, timet_stream_t stream) {
    TIMET_CHECK_ARG(x_list && prototypes && scores_out && workspace, "cosine_scores: NULL pointer");
    TIMET_CHECK_ARG(n_x >= 1 && n_x <= SC_MAX_INPUTS, "cosine_scores: n_x=%d must be in 1..%d", n_x, SC_MAX_INPUTS);
    ScInputs in;
    in.rows_each = rows_each;
    for (int i = 0; i < SC_MAX_INPUTS; ++i) in.x[i] = (i < n_x) ? x_list[i] : nullptr;
    for (int i = 0; i < n_x; ++i) TIMET_CHECK_ARG(x_list[i] != nullptr, "cosine_scores: input %d is NULL", i);
    const int64_t B = rows_each * n_x;
    TIMET_CHECK_ARG(rows_each >= 1 && K >= 1 && dh >= 1 && dh <= 4096, "cosine_scores: bad shape B=%lld K=%d dh=%d", (long long)B, K, dh);
    TIMET_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "cosine_scores: workspace must be 1024-byte aligned");
    if (workspace_bytes < scores_ws_bytes(B, K, dh)) {
        set_error("cosine_scores: workspace %zu < %zu bytes", workspace_bytes, scores_ws_bytes(B, K, dh));
        return TIMET_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int dhp = (dh + 63) / 64 * 64;
    const int64_t Kp = (K + 255) / 256 * 256;
    __half *a2 = reinterpret_cast<__half *>(workspace);
    __half *b2 = reinterpret_cast<__half *>((char *)workspace + align_up((size_t)(B + 128) * 2 * dhp * sizeof(__half), 1024));
    int64_t pb = (B + Kp + 7) / 8;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (pb > cap) pb = cap;
    scores_prep_kernel<<<(int)pb, 256, 0, st>>>(in, prototypes, a2, b2, B, K, (int)Kp, dh, dhp);
    TIMET_LAUNCHED();

    const int n_tiles = (int)((K + 255) / 256);
    const int Np = (n_tiles > 1) ? 256 : (K + 15) / 16 * 16;
    CUtensorMap map_a, map_b;
    int rc;
    {   // [rows, 2*dhp] fp16 tensors, 64-wide boxes, SWIZZLE_128B (same encoder as the FF kernel, different pitch)
        // tc_make_map assumes a Dp-wide row; the row here is 2*dhp wide
        if ((rc = tc_make_map(&map_a, a2, B + 128, 2 * dhp, 128)) != TIMET_OK) return rc;
        if ((rc = tc_make_map(&map_b, b2, Kp, 2 * dhp, Np)) != TIMET_OK) return rc;
    }
    const int stages = env_cfg().sc_stages == 4 ? 4 : 2;
    const size_t smem = 1024 + (size_t)stages * (128 * 128 + (size_t)Np * 128) + sizeof(ScCtl) + 4 * 32 * 33 * sizeof(float) + 64;
    dim3 grid((unsigned)((B + 127) / 128), (unsigned)n_tiles);
    if (stages == 4) {
        TIMET_CUDA(cudaFuncSetAttribute(scores_gemm_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        scores_gemm_kernel<4><<<grid, SC_THREADS, smem, st>>>(map_a, map_b, scores_out, B, K, Np, dhp);
    } else {
        TIMET_CUDA(cudaFuncSetAttribute(scores_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        scores_gemm_kernel<2><<<grid, SC_THREADS, smem, st>>>(map_a, map_b, scores_out, B, K, Np, dhp);
    }
    TIMET_LAUNCHED();
    return TIMET_OK;
}

}
