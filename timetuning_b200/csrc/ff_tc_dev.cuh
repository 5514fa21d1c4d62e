// Device-side building blocks shared by the per-item (ff_tc.cu) and the persistent (ff_tc3.cu) forms of the FF
// tensor-core kernel: geometry, shared-memory control block, packed candidate lists, predicated offer,
// warp-synchronous compaction.
#pragma once
#include "ff_select.cuh"
#include "ptx_sm100.cuh"

namespace timet {

constexpr int TC_GROUPS = 4;                    // epilogue warpgroups (4 warps each)
constexpr int TC_THREADS = 128 + TC_GROUPS * 128;
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_MAX_NKC = 6;                  // resident query tile: Dp <= 384
constexpr int TC_MAX_NKC_STREAM = 32;          // streamed query tile (A chunk re-loaded with every B chunk): Dp <= 2048
constexpr int TC_CLIP_GROUP = 8;                // clips whose tiles are launched together (L2 locality)
constexpr int TC_CAP = FF_CAND_CAP;            // 32 candidates per (query, epilogue group)
constexpr float FF_TC_DELTA = 1.05e-3f;        // bound on |sim~ - sim|: fp16 RN of both unit vectors (2^-10) + fp32 accumulation
constexpr float FF_TC_SLACK = 2.0f * FF_TC_DELTA + 3.1e-5f;   // + 2 x fixed-point quantisation (2^-17) with margin
constexpr float TC_FIX_BIAS = 66.0f;           // sim~ + 2 in [1,3] lands in [64,128): ulp = 2^-17 -> 19-bit fixed point in the mantissa

struct TcGeom {
    int H, W, N, Dp, NKC;
    int a_resident;           // 1: query tile A stays in smem (Dp <= 384); 0: its 64-wide chunks ride in the ring stages
    int QR, tiles_per_frame;
    int RPC, NT, qrows;
    int n_clips, n_frames, nT, t_begin, n_last, radius, topk;
    int nbuf, buf_cols, nstages;   // TMEM accumulator buffers (4 x 128 or 2 x 256 columns), B ring depth
    int clip_group;           // clips whose tiles are launched together (L2 locality); env TIMET_TC_CLIP_GROUP
    int flags;                // debug (env TIMET_TC_FLAGS): 1 = epilogue releases tiles unscanned, 2 = scan but never append
    int colblk;               // persistent kernel: query tile rows arranged as 4 column blocks of 4 x 8 queries (ff_tc3.cu)
    int ncl;                  // ... valid columns of the last column block (W - 24): it is stored as 4 x ncl queries
    int a_chunk_bytes;        // bytes between the 64-wide K chunks of the resident query tile (16384, or less: see ff_tc3.cu)
    int64_t total_tiles;
};

struct __align__(8) TcSmemCtl {
    uint64_t full[TC_MAX_STAGES], empty[TC_MAX_STAGES], a_full, tmem_full[4], tmem_empty[4];
    uint32_t tmem_base;
    uint32_t thr_sh[128];     // per-query nomination threshold shared by the groups: float bits of (thr + 4), atomicMax
    uint32_t xchg[3][128];    // groups 1..3 -> group 0: cnt | lost << 16 per query
};

__device__ __forceinline__ uint32_t thr_enc(float thr) { return __float_as_uint(fmaxf(thr, -3.0f) + 4.0f); }
__device__ __forceinline__ float thr_dec(uint32_t v) { return __uint_as_float(v) - 4.0f; }

__device__ __forceinline__ float tc_decode(uint32_t entry) { return (float)(entry >> 13) * (1.0f / 131072.0f) - 2.0f; }

// Candidate lists live in shared memory: slot s of thread qi of group g at list + s * 512 B (conflict-free).
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
constexpr uint32_t TC_SLOT_STRIDE = 128u * 4u;     // bytes between consecutive slots of one thread's list

// One epilogue step, predicated (no branch): if the key is inside the window (bit BIT of wmask) and
// sim~ > thr, append the packed candidate ((sim~ + 2) in 19-bit fixed point << 13 | code) and advance.
template <uint32_t BIT>
__device__ __forceinline__ void tc_offer(uint32_t &slot_addr, float v, float thr, uint32_t wmask, uint32_t code) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 u;\n\t.reg .f32 t;\n\t"
        "and.b32 u, %3, %5;\n\t"
        "setp.ne.b32 p, u, 0;\n\t"
        "setp.gt.and.f32 p, %1, %2, p;\n\t"
        "@p add.rn.f32 t, %1, 0f42840000;\n\t"     // + 66.0f
        "@p mov.b32 u, t;\n\t"
        "@p mad.lo.u32 u, u, 8192, %4;\n\t"
        "@p st.shared.u32 [%0], u;\n\t"
        "@p add.u32 %0, %0, 512;\n\t}"
        : "+r"(slot_addr)
        : "f"(v), "f"(thr), "r"(wmask), "r"(code), "n"(BIT)
        : "memory");
}

// Four epilogue steps (elements E .. E+3 of a 16-column TMEM block) in ONE asm statement: ptxas may interleave the four
// window-test -> compare -> append chains, which a sequence of volatile single-step statements forbids.
template <uint32_t E>
__device__ __forceinline__ void tc_offer4(uint32_t &slot_addr, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3, float thr,
                                          uint32_t wmask, uint32_t code0) {
    // the packed candidate is computed unconditionally (FADD + IMAD on the FMA pipe); only the store and the slot advance
    // are predicated.  Predicated writes to the temporaries would make ptxas carry their old values around (SEL / MOV).
    asm volatile(
        "{\n\t.reg .pred p0, p1, p2, p3;\n\t.reg .b32 m0, m1, m2, m3, x0, x1, x2, x3, y0, y1, y2, y3;\n\t.reg .f32 t0, t1, t2, t3;\n\t"
        "and.b32 m0, %6, %8;\n\t"
        "and.b32 m1, %6, %9;\n\t"
        "and.b32 m2, %6, %10;\n\t"
        "and.b32 m3, %6, %11;\n\t"
        "setp.ne.b32 p0, m0, 0;\n\t"
        "setp.ne.b32 p1, m1, 0;\n\t"
        "setp.ne.b32 p2, m2, 0;\n\t"
        "setp.ne.b32 p3, m3, 0;\n\t"
        "setp.gt.and.f32 p0, %1, %5, p0;\n\t"
        "setp.gt.and.f32 p1, %2, %5, p1;\n\t"
        "setp.gt.and.f32 p2, %3, %5, p2;\n\t"
        "setp.gt.and.f32 p3, %4, %5, p3;\n\t"
        "add.rn.f32 t0, %1, 0f42840000;\n\t"     // + 66.0f
        "add.rn.f32 t1, %2, 0f42840000;\n\t"
        "add.rn.f32 t2, %3, 0f42840000;\n\t"
        "add.rn.f32 t3, %4, 0f42840000;\n\t"
        "mov.b32 x0, t0;\n\t"
        "mov.b32 x1, t1;\n\t"
        "mov.b32 x2, t2;\n\t"
        "mov.b32 x3, t3;\n\t"
        "mad.lo.u32 y0, x0, 8192, %7;\n\t"
        "mad.lo.u32 y1, x1, 8192, %12;\n\t"
        "mad.lo.u32 y2, x2, 8192, %13;\n\t"
        "mad.lo.u32 y3, x3, 8192, %14;\n\t"
        "@p0 st.shared.u32 [%0], y0;\n\t"
        "@p0 add.u32 %0, %0, 512;\n\t"
        "@p1 st.shared.u32 [%0], y1;\n\t"
        "@p1 add.u32 %0, %0, 512;\n\t"
        "@p2 st.shared.u32 [%0], y2;\n\t"
        "@p2 add.u32 %0, %0, 512;\n\t"
        "@p3 st.shared.u32 [%0], y3;\n\t"
        "@p3 add.u32 %0, %0, 512;\n\t}"
        : "+r"(slot_addr)
        : "f"(__uint_as_float(v0)), "f"(__uint_as_float(v1)), "f"(__uint_as_float(v2)), "f"(__uint_as_float(v3)), "f"(thr),
          "r"(wmask), "r"(code0 + E), "n"(1u << E), "n"(2u << E), "n"(4u << E), "n"(8u << E), "r"(code0 + E + 1),
          "r"(code0 + E + 2), "r"(code0 + E + 3)
        : "memory");
}

// Drop entries below thr.  Warp-synchronous (loop bound = warp max of cnt).
__device__ __forceinline__ void tc_filter(uint32_t list, int &cnt, float thr) {
    int maxcnt = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxcnt = max(maxcnt, __shfl_xor_sync(0xffffffffu, maxcnt, o));
    // keep entries whose quantised value is >= thr (quantisation is already inside FF_TC_SLACK)
    const float lim = (thr + 2.0f) * 131072.0f;
    const uint32_t enc = (lim <= 0.f) ? 0u : ((uint32_t)lim << 13);
    uint32_t dst = list;
#pragma unroll 4
    for (int s = 0; s < maxcnt; ++s) {
        const uint32_t e = (s < cnt) ? lds_u32(list + s * TC_SLOT_STRIDE) : 0u;
        if (s < cnt && e >= enc) { sts_u32(dst, e); dst += TC_SLOT_STRIDE; }
    }
    cnt = (int)((dst - list) / TC_SLOT_STRIDE);
}

// Raise thr from the list content and drop entries that can no longer be among the top-k.
// One pass over the list keeps the KK largest packed entries in sorted registers (max/min chain), so the
// k-th largest (k <= KK) is read off directly.  Warp-synchronous; loop bounds are warp-uniform.
template <int KK>
__device__ __forceinline__ uint32_t tc_kth_largest(uint32_t list, int cnt, int maxcnt, int k) {
    uint32_t top[KK];
#pragma unroll
    for (int j = 0; j < KK; ++j) top[j] = 0u;
#pragma unroll 4
    for (int s = 0; s < maxcnt; ++s) {
        uint32_t e = (s < cnt) ? lds_u32(list + s * TC_SLOT_STRIDE) : 0u;
#pragma unroll
        for (int j = 0; j < KK; ++j) {
            const uint32_t hi = max(e, top[j]);
            e = min(e, top[j]);
            top[j] = hi;
        }
    }
    uint32_t kth = top[0];
#pragma unroll
    for (int j = 1; j < KK; ++j) kth = (k - 1 == j) ? top[j] : kth;
    return kth;
}

// Afterwards cnt <= keep_max (entries beyond that are dropped and the query is flagged for the exact re-do).
__device__ __forceinline__ void tc_compact(uint32_t list, int &cnt, float &thr, int &lost, int k, int keep_max) {
    int maxcnt = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxcnt = max(maxcnt, __shfl_xor_sync(0xffffffffu, maxcnt, o));
    // the chain costs 2 instructions per (entry, register): the usual k = 5 gets a 5-deep chain
    const uint32_t kth = (k <= 5) ? tc_kth_largest<5>(list, cnt, maxcnt, k) : tc_kth_largest<8>(list, cnt, maxcnt, k);
    if (kth != 0u) thr = fmaxf(thr, tc_decode(kth) - FF_TC_SLACK);
    tc_filter(list, cnt, thr);
    if (cnt > keep_max) { cnt = keep_max; lost = 1; }
}

// host helpers defined in ff_tc.cu
bool tc_geometry(const timet_ff_params &p, const FFLayout &L, TcGeom *G);
size_t tc_smem_bytes(const TcGeom &G);
int tc_make_map(CUtensorMap *m, const void *base, int64_t rows, int Dp, int box_rows);
int tc_make_map_colblk(CUtensorMap *m, const void *base, int64_t grid_rows, int Dp, int grid_w, int box_cols);

}  // namespace timet
