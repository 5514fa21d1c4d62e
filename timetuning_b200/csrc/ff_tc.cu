// FF stage 2, TC engine: tcgen05 / TMEM / TMA affinity GEMM with the neighbourhood window and the
// top-k NOMINATION fused into the TMEM epilogue; the N x N affinity never exists in memory.
//
// Reference maths (/root/reference/mask_propagation.py:418-436): sim = f^_t f^_c^T for every context c,
// aff = exp(sim/0.1) * window, theta_i = k-th largest over all contexts, keep aff >= theta_i, normalise.
//
// Precision contract.  The GEMM runs on fp16-rounded unit vectors with fp32 accumulation, so every
// approximate similarity satisfies |sim~ - sim| <= delta (delta = 2^-10 + accumulation slack, see
// FF_TC_DELTA).  The tensor cores only NOMINATE: a key is appended to the query's candidate list
// when sim~ > thr, where thr is always (k-th largest sim~ seen so far) - 2*delta, a provable lower
// bound of (final theta_i - delta).  Hence every key of the true top-k (ties included) is in the
// list.  ff_finalize then re-evaluates the <= 16 candidates with the canonical fp32 dot product
// (common.cuh) and applies the reference's selection exactly.  A query whose list overflowed is
// re-done by the exact engine.  The result is bit-identical to TIMET_FF_EXACT by construction.
//
// One CTA (640 threads) per (clip, target frame, query tile of QR grid rows <= 128 queries):
//   warp 0      TMA producer: query tile A (resident for Dp <= 384: Dp/64 swizzled 64-wide chunks loaded once;
//               streamed with the key chunks for larger Dp), then a ring of key chunks B (NT = RPC*W keys x 64)
//               for every context x key-row chunk that intersects the tile's window band (tiles outside the band
//               are never loaded or multiplied)
//   warp 1      MMA issuer: tcgen05.mma cta_group::1 kind::f16, M=128, N=NT (<=256), K=16, fp32 accumulators in
//               TMEM (2 x 256 or 4 x 128 columns); descriptors advanced incrementally, no divisions on this path
//   warp 2      TMEM allocator          warp 3   initialises the shared per-query thresholds
//   warps 4-19  four epilogue groups of 4 warps (thread = query row = TMEM lane).  2 buffers: groups g, g+2 share
//               the key tiles of buffer g & 1 and split their key rows by parity; 4 buffers: one group per buffer.
//               tcgen05.ld 32x32b.x16; window test by index arithmetic; predicated append of packed
//               (sim~, ctx, drow, dcol) candidates to a per-(query, group) list in shared memory; warp-synchronous
//               single-pass compaction raises the nomination threshold, which the groups of a query share through
//               an atomicMax word in shared memory; at the end the four lists are filtered with the final
//               threshold, merged and <= 16 candidates per query are published.
// (A CTA-pair cta_group::2 variant was measured slower twice and removed in round 2: profiles/r2_experiments.md.)
#include <stdlib.h>

#include "ff_tc_dev.cuh"

namespace timet {


template <bool DUMP>
__global__ void __launch_bounds__(TC_THREADS, 1)
ff_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, TcGeom G,
             uint32_t *__restrict__ cand, uint32_t *__restrict__ cand_meta, int64_t tile_override, float *__restrict__ dump, unsigned long long *__restrict__ trace) {
    extern __shared__ uint8_t smem_raw[];
    // carve: [A: NKC x 16 KB][B: nstages x NT*128][lists: 4 x 32 x 128 u32][ctl]; 1024-aligned for SWIZZLE_128B
    // (streamed-A variant for Dp > 384: no resident A, every ring stage is [A chunk 16 KB | B chunk])
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    const uint32_t b_stage_bytes = (uint32_t)G.NT * 128u;
    const uint32_t a_in_stage = G.a_resident ? 0u : 16384u;
    const uint32_t stage_bytes = b_stage_bytes + a_in_stage;
    uint8_t *sB = sA + (G.a_resident ? (size_t)G.NKC * 16384 : 0);   // ring base
    uint32_t *sList = reinterpret_cast<uint32_t *>(sB + (size_t)G.nstages * stage_bytes);
    TcSmemCtl *ctl = reinterpret_cast<TcSmemCtl *>(sList + TC_GROUPS * TC_CAP * 128);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // debug timeline: stamp s of this CTA (globaltimer ns); trace == nullptr in production
    auto stamp = [&](int s) {
        if (trace && blockIdx.x < FF_TRACE_CTAS) {
            unsigned long long tns;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
            trace[(size_t)blockIdx.x * FF_TRACE_SLOTS + s] = tns;
        }
    };
    if (threadIdx.x == 0) stamp(0);

    // ---- which tile: heavy (late) target frames first
    const int64_t tile_id = (tile_override >= 0) ? tile_override : (int64_t)blockIdx.x;
    // launch order: groups of TC_CLIP_GROUP clips (their frames stay L2-resident while the group runs),
    // inside a group the heavy (late) target frames first
    const int per_clip = G.nT * G.tiles_per_frame;
    const int grp = (int)(tile_id / ((int64_t)G.clip_group * per_clip));
    const int grp_clips = min(G.clip_group, G.n_clips - grp * G.clip_group);
    const int in_grp = (int)(tile_id - (int64_t)grp * G.clip_group * per_clip);
    const int per_t = grp_clips * G.tiles_per_frame;
    const int tdesc = in_grp / per_t;
    const int rem = in_grp - tdesc * per_t;
    const int t = G.n_frames - 1 - tdesc;
    const int clip = grp * G.clip_group + rem / G.tiles_per_frame, qt = rem % G.tiles_per_frame;
    const int qr0 = qt * G.QR;
    const int qr1 = min(G.H - 1, qr0 + G.QR - 1);
    const int nq = (qr1 - qr0 + 1) * G.W;
    const int kr_lo = max(0, qr0 - G.radius), kr_hi = min(G.H - 1, qr1 + G.radius);
    const int nchunks = (kr_hi - kr_lo + G.RPC) / G.RPC;
    const int nctx = ctx_count(t, G.n_last);
    const int ntiles = nctx * nchunks;
    const int64_t clip_row0 = (int64_t)clip * G.n_frames * G.N;
    const int q_row0 = (int)(clip_row0 + (int64_t)t * G.N + qr0 * G.W);

    // ---- one-time setup
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_a);
        ptx::prefetch_tensormap(&map_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < G.nstages; ++s) { ptx::mbar_init(&ctl->full[s], 1); ptx::mbar_init(&ctl->empty[s], 1); }
        ptx::mbar_init(&ctl->a_full, 1);
        // every warp that reads a buffer arrives once per tile: 4 warps (one group) or 8 (two groups splitting the rows)
        for (int b = 0; b < G.nbuf; ++b) { ptx::mbar_init(&ctl->tmem_full[b], 1); ptx::mbar_init(&ctl->tmem_empty[b], G.nbuf == 4 ? 4 : 8); }
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc<512>(&ctl->tmem_base);
    if (warp == 3) for (int i = lane; i < 128; i += 32) ctl->thr_sh[i] = thr_enc(-INFINITY);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = ctl->tmem_base;
    if (threadIdx.x == 0) stamp(1);

    if (warp == 0) {
        // =========================== TMA producer ===========================
        if (lane == 0) {
            if (G.a_resident) {
                ptx::mbar_expect_tx(&ctl->a_full, (uint32_t)G.NKC * 16384u);
                for (int kc = 0; kc < G.NKC; ++kc) ptx::tma_load_2d(sA + kc * 16384, &map_a, kc * 64, q_row0, &ctl->a_full);
            }
            uint32_t stage = 0, phase = 0;                     // ring position without divisions
            for (int ci = 0; ci < nctx; ++ci) {
                const int f = ctx_frame(t, G.n_last, ci);
                for (int ch = 0; ch < nchunks; ++ch) {
                    const int k_row0 = (int)(clip_row0 + (int64_t)f * G.N + (kr_lo + ch * G.RPC) * G.W);
                    for (int kc = 0; kc < G.NKC; ++kc) {
                        ptx::mbar_wait(&ctl->empty[stage], phase ^ 1u);
                        ptx::mbar_expect_tx(&ctl->full[stage], stage_bytes);
                        uint8_t *st = sB + (size_t)stage * stage_bytes;
                        if (!G.a_resident) ptx::tma_load_2d(st, &map_a, kc * 64, q_row0, &ctl->full[stage]);
                        ptx::tma_load_2d(st + a_in_stage, &map_b, kc * 64, k_row0, &ctl->full[stage]);
                        if (++stage == (uint32_t)G.nstages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            if (G.a_resident) ptx::mbar_wait(&ctl->a_full, 0);
            ptx::tc_fence_after();
            stamp(2);
            // the issuing thread is on the critical path of every MMA: descriptors are built once and advanced by
            // adding to the 14-bit start-address field; the ring position is tracked without divisions
            const uint64_t da0 = ptx::umma_desc_sw128(ptx::smem_u32(sA)), db0 = ptx::umma_desc_sw128(ptx::smem_u32(sB));
            const uint32_t stage_step = stage_bytes >> 4, a_step = a_in_stage >> 4;
            uint32_t stage = 0, phase = 0, buf = 0, use = 0;
            int ch = 0;
            for (int tile = 0; tile < ntiles; ++tile) {
                const int rc = min(G.RPC, kr_hi + 1 - (kr_lo + ch * G.RPC));
                const int n_mma = min(G.NT, (rc + G.qrows - 1) / G.qrows * G.qrows * G.W);
                const uint32_t idesc = ptx::umma_idesc_f16(128, n_mma);
                ptx::mbar_wait(&ctl->tmem_empty[buf], (use & 1u) ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * (uint32_t)G.buf_cols;
                uint64_t da = da0;
                for (int kc = 0; kc < G.NKC; ++kc, da += 16384 >> 4) {
                    ptx::mbar_wait(&ctl->full[stage], phase);
                    ptx::tc_fence_after();
                    const uint64_t ds = db0 + (uint64_t)(stage * stage_step);
                    const uint64_t db = ds + a_step;
                    const uint64_t dq = G.a_resident ? da : ds;              // streamed A: the chunk sits in front of B
                    ptx::umma_f16(d_tmem, dq, db, idesc, kc != 0);
                    ptx::umma_f16(d_tmem, dq + 2, db + 2, idesc, true);      // +32 B per K = 16 step
                    ptx::umma_f16(d_tmem, dq + 4, db + 4, idesc, true);
                    ptx::umma_f16(d_tmem, dq + 6, db + 6, idesc, true);
                    ptx::umma_commit(&ctl->empty[stage]);          // frees the smem stage when the MMAs retire
                    if (++stage == (uint32_t)G.nstages) { stage = 0; phase ^= 1u; }
                }
                ptx::umma_commit(&ctl->tmem_full[buf]);            // accumulator ready for its epilogue group
                if (++buf == (uint32_t)G.nbuf) { buf = 0; ++use; }
                if (++ch == nchunks) ch = 0;
            }
            stamp(3);
        }
    } else if (warp >= 4) {
        // =========================== epilogue groups ===========================
        // 4 buffers: group g owns key tiles g, g+4, ...   2 buffers: groups g and g+2 share the tiles of buffer
        // g & 1 and split their key rows by parity.
        const int g = (warp - 4) >> 2;
        const int buf = (G.nbuf == 4) ? g : (g & 1);
        const int row_par = (G.nbuf == 4) ? -1 : (g >> 1);
        const int qi = ((warp & 3) << 5) + lane;                    // TMEM lane == query row of the tile
        const uint32_t lane_base = (uint32_t)((warp & 3) << 5) << 16;
        const uint32_t list = ptx::smem_u32(sList + (size_t)g * TC_CAP * 128 + qi);   // slot s at list + s * 512 B
        constexpr int half = TC_CAP / 2;                              // max entries kept across a compaction
        const bool valid = qi < nq;
        const int qrow = qr0 + qi / G.W, qcol = qi % G.W;
        const int r_lo = qrow - G.radius, r_hi = qrow + G.radius;
        const int c_lo = qcol - G.radius;
        const int c_lo_cl = max(c_lo, 0), c_hi_cl = min(qcol + G.radius, G.W - 1);
        float thr = (G.flags & 2) ? INFINITY : -INFINITY;
        int cnt = 0, lost = 0;

        int ci = 0, ch = buf;                                       // tile = ci * nchunks + ch, advanced without divisions
        while (ch >= nchunks) { ch -= nchunks; ++ci; }
        uint32_t use = 0;
        for (int tile = buf; tile < ntiles; tile += G.nbuf, ++use) {
            const int kr_start = kr_lo + ch * G.RPC;
            const int rc = min(G.RPC, kr_hi + 1 - kr_start);
            if (lane == 0) ptx::mbar_wait(&ctl->tmem_full[buf], use & 1u);   // one poller per warp: the tensor core needs the smem bandwidth
            __syncwarp();
            ptx::tc_fence_after();
            const uint32_t t_acc = tmem_base + (uint32_t)(buf * G.buf_cols) + lane_base;
            thr = fmaxf(thr, thr_dec(ctl->thr_sh[qi]));               // what the other groups have established

            if (DUMP && row_par <= 0) {
                for (int c0 = 0; c0 < G.buf_cols; c0 += 16) {
                    uint32_t r[16];
                    ptx::tmem_ld_32x16(t_acc + c0, r);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) dump[((size_t)tile * 128 + qi) * 256 + c0 + e] = __uint_as_float(r[e]);
                }
            }

            for (int rr = 0; rr < ((G.flags & 1) ? 0 : rc); ++rr) {
                if (row_par >= 0 && (rr & 1) != row_par) continue;
                const int kr = kr_start + rr;
                const bool row_ok = valid && kr >= r_lo && kr <= r_hi;
                if (!__any_sync(0xffffffffu, row_ok)) continue;
                const int code_row = (ci << 10) | ((kr - r_lo) << 5);
                for (int cb = 0; cb < G.W; cb += 16) {
                    int col0 = rr * G.W + cb;
                    const int shift = max(0, col0 + 16 - G.buf_cols);   // keep the 16-column load inside the buffer
                    col0 -= shift;
                    const int cb_eff = cb - shift;                      // register e holds key column cb_eff + e
                    uint32_t r[16];
                    ptx::tmem_ld_32x16(t_acc + (uint32_t)col0, r);
                    // window columns of this thread inside the block -> bit mask over e
                    const int lo = max(c_lo_cl - cb_eff, shift), hi = min(c_hi_cl - cb_eff, 15);
                    uint32_t wmask = 0u;
                    if (row_ok && hi >= lo) wmask = (0xFFFFu >> (15 - hi)) & (0xFFFFu << lo) & 0xFFFFu;
                    const uint32_t code0 = (uint32_t)(code_row + (cb_eff - c_lo));
                    uint32_t slot = list + (uint32_t)cnt * TC_SLOT_STRIDE;
                    ptx::tmem_ld_wait();
#define TC_OFFER(E) tc_offer<(1u << (E))>(slot, __uint_as_float(r[E]), thr, wmask, code0 + (E));
                    // invariant: cnt <= 16 before every block of 16 offers -> the 32 slots cannot overflow.
                    // (Screening pairs by their max behind a branch was measured 25 % SLOWER: divergence beats the
                    // saved predicated instructions, so every element takes the predicated path.)
                    TC_OFFER(0) TC_OFFER(1) TC_OFFER(2) TC_OFFER(3) TC_OFFER(4) TC_OFFER(5) TC_OFFER(6) TC_OFFER(7)
                    TC_OFFER(8) TC_OFFER(9) TC_OFFER(10) TC_OFFER(11) TC_OFFER(12) TC_OFFER(13) TC_OFFER(14) TC_OFFER(15)
#undef TC_OFFER
                    cnt = (int)((slot - list) / TC_SLOT_STRIDE);
                    if (__any_sync(0xffffffffu, cnt > half)) {
                        const float before = thr;
                        tc_compact(list, cnt, thr, lost, G.topk, half);
                        if (thr > before) atomicMax(&ctl->thr_sh[qi], thr_enc(thr));
                    }
                }
            }
            // release the accumulator buffer to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&ctl->tmem_empty[buf]);
            ch += G.nbuf;
            while (ch >= nchunks) { ch -= nchunks; ++ci; }
        }

        if (warp == 4 && lane == 0) stamp(4);
        // ---- final: agree on the per-query threshold, filter, merge the 4 lists, publish <= FF_CAND_STORE candidates
        if (!DUMP) {
            {
                const float before = thr;
                tc_compact(list, cnt, thr, lost, G.topk, half);
                if (thr > before) atomicMax(&ctl->thr_sh[qi], thr_enc(thr));
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");            // the 16 epilogue warps
            thr = fmaxf(thr, thr_dec(ctl->thr_sh[qi]));
            tc_filter(list, cnt, thr);
            if (g > 0) ctl->xchg[g - 1][qi] = (uint32_t)cnt | (lost ? 0x10000u : 0u);
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (warp == 4 && lane == 0) stamp(5);
            if (g == 0) {
                for (int og = 1; og < TC_GROUPS; ++og) {
                    const uint32_t other = ctl->xchg[og - 1][qi];
                    const int ocnt = (int)(other & 0xFFFFu);
                    lost |= (int)(other >> 16);
                    const uint32_t olist = list + (uint32_t)og * TC_CAP * 128u * 4u;
                    for (int s = 0; s < ocnt; ++s) {
                        if (cnt < TC_CAP) { sts_u32(list + (uint32_t)cnt * TC_SLOT_STRIDE, lds_u32(olist + s * TC_SLOT_STRIDE)); ++cnt; }
                        else lost = 1;
                    }
                }
                tc_compact(list, cnt, thr, lost, G.topk, FF_CAND_STORE);   // joint k-th - slack
                if (valid && tile_override < 0) {
                    const int64_t q = ((int64_t)clip * G.nT + (t - G.t_begin)) * G.N + qr0 * G.W + qi;
                    uint32_t *dst = cand + q * FF_CAND_STORE;
#pragma unroll
                    for (int s4 = 0; s4 < FF_CAND_STORE; s4 += 4) {
                        if (s4 < cnt) {
                            uint4 v;
                            v.x = lds_u32(list + (s4 + 0) * TC_SLOT_STRIDE); v.y = lds_u32(list + (s4 + 1) * TC_SLOT_STRIDE);
                            v.z = lds_u32(list + (s4 + 2) * TC_SLOT_STRIDE); v.w = lds_u32(list + (s4 + 3) * TC_SLOT_STRIDE);
                            *reinterpret_cast<uint4 *>(dst + s4) = v;
                        }
                    }
                    cand_meta[q] = (uint32_t)cnt | (lost ? 0x10000u : 0u);
                }
            }
        }
    }

    // ---- teardown
    if (warp == 4 && lane == 0) stamp(6);
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) stamp(7);
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------ finalize: exact re-evaluation
// One warp per query: canonical fp32 similarity of every nominated key, exp, reference selection with ties.
//
// The kernel is bound by dependent memory round trips per query, not by bandwidth (ncu: L2 -> SM traffic well below the
// limit; halving the resident warps halved the speed), so the per-query chain is kept as short as possible:
//   * the candidate list (meta word + <= 16 packed entries) of the warp's NEXT query is fetched while the current one
//     is evaluated;
//   * the query row is not staged through shared memory: it is loaded from global memory in the same batch as the key
//     rows (and hits L1 when a second batch needs it again);
//   * BATCH candidate rows are in flight together; the usual k = 5 needs ONE batch with BATCH = 6.
// Per-candidate chains keep their i = lane, lane + 32, ... order (common.cuh), so the bits do not depend on BATCH.
constexpr int FIN_WARPS = 8;
constexpr int FIN_QPB = 64;                    // consecutive queries per CTA (neighbouring queries nominate neighbouring keys: L1 reuse)
constexpr int FIN_BATCH = 4;                   // candidate rows in flight together (6 / 8 measured slower: registers, occupancy)

// Canonical xor-butterfly of FOUR accumulators at once ("transposed" reduction): after the xor-16 step a lane keeps two of
// the four values, after the xor-8 step one; steps 4, 2, 1 finish it.  7 shuffles instead of 20, and every partial sum is
// the same  v[l] + v[l ^ o]  in the same order as warp_sum (fp add is commutative), so the bits are those of common.cuh's
// definition.  Returns the sum of acc[u] in the lanes with bit 4 == (u >> 1) and bit 3 == (u & 1).
__device__ __forceinline__ float fin_reduce4(const float (&acc)[4], int lane) {
    const bool hi4 = (lane & 16) != 0, hi3 = (lane & 8) != 0;
    float k0 = hi4 ? acc[2] : acc[0], k1 = hi4 ? acc[3] : acc[1];
    const float s0 = hi4 ? acc[0] : acc[2], s1 = hi4 ? acc[1] : acc[3];
    k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    float v = hi3 ? k1 : k0;
    const float sv = hi3 ? k0 : k1;
    v += __shfl_xor_sync(0xffffffffu, sv, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// STAGED (rows of at most FIN_STAGE_N4 float4, i.e. dim <= 384): the query row and up to FIN_SB key rows of a query are
// copied to the warp's slice of shared memory with cp.async -- every 16-byte piece of all six rows is in flight at once and
// costs no register -- and the dot products then read shared memory.  The register form below can keep only part of a batch
// in flight inside the 64 registers that four CTAs per SM allow (ptxas interleaves loads and fma chains: several dependent
// L2 round trips per query); staged, a query with k = 5 candidates costs ONE round trip, at three CTAs per SM.  Same
// per-lane fma chains, same butterfly: identical bits.
constexpr int FIN_SB = 5;                      // key rows staged together
constexpr int FIN_STAGE_N4 = 96;
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

template <bool STAGED>
__global__ void __launch_bounds__(FIN_WARPS * 32, STAGED ? 3 : 4)
ff_finalize_kernel(timet_ff_params p, int N, FFSrc S, int nT, int kw, uint32_t w_magic,
                   const uint32_t *__restrict__ cand, const uint32_t *__restrict__ cand_meta,
                   float *__restrict__ sel_w, int32_t *__restrict__ sel_k, int32_t *__restrict__ sel_cnt,
                   unsigned long long *__restrict__ stats, int32_t *__restrict__ redo_list,
                   unsigned int *__restrict__ redo_count, uint32_t n_queries) {
    __shared__ unsigned long long s_stat[4];
    extern __shared__ float4 fin_stage[];          // STAGED: [warp][1 + FIN_SB][n4]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 4) s_stat[threadIdx.x] = 0ull;
    __syncthreads();
    const int W = p.grid_w;
    const int n4 = S.n4;
    unsigned long long st_sel = 0, st_ties = 0, st_cand = 0;
    // index arithmetic without per-query divisions: (frame, patch) of the CTA's first query once, then increments
    const uint32_t q_first = blockIdx.x * (uint32_t)FIN_QPB;
    const uint32_t q_end = min(n_queries, q_first + (uint32_t)FIN_QPB);
    const uint32_t ft0 = q_first / (uint32_t)N;              // clip * nT + tt
    const uint32_t clip0 = ft0 / (uint32_t)nT;
    int i = (int)(q_first - ft0 * (uint32_t)N) + warp;
    int tt = (int)(ft0 - clip0 * (uint32_t)nT), clip = (int)clip0;
    uint32_t qid = q_first + (uint32_t)warp;
    uint32_t m_next = 0u, c_next = 0u;
    if (qid < q_end) {
        m_next = __ldg(cand_meta + qid);
        if (lane < FF_CAND_STORE) c_next = __ldg(cand + (size_t)qid * FF_CAND_STORE + lane);
    }
    const int src_lane = ((lane & 2) << 3) | ((lane & 1) << 3);     // where fin_reduce4 leaves candidate (lane & 3)
    for (; qid < q_end; qid += FIN_WARPS, i += FIN_WARPS) {
        while (i >= N) { i -= N; if (++tt == nT) { tt = 0; ++clip; } }
        // every lane loaded the same meta word; taking it from a warp collective tells the compiler that the candidate
        // loops below are warp-uniform (no reconvergence code around their shuffles)
        const uint32_t m0 = __reduce_max_sync(0xffffffffu, m_next), c_cur = c_next;
        if (qid + FIN_WARPS < q_end) {                 // next query's list: in flight during this query's evaluation
            m_next = __ldg(cand_meta + qid + FIN_WARPS);
            if (lane < FF_CAND_STORE) c_next = __ldg(cand + (size_t)(qid + FIN_WARPS) * FF_CAND_STORE + lane);
        }
        const int nc = (int)(m0 & 0xFFFFu);
        if ((m0 & 0x10000u) != 0u) {
            if (lane == 0) redo_list[atomicAdd(redo_count, 1u)] = (int32_t)qid;
            continue;
        }
        const int t = p.t_begin + tt;
        const int64_t clip_row0 = (int64_t)clip * p.n_frames * N;
        const int64_t q_row = clip_row0 + (int64_t)t * N + i;
        const float4 *qrow = reinterpret_cast<const float4 *>(S.x + q_row * S.ld);
        const int qr = (W == 1) ? i : (int)__umulhi((uint32_t)i, w_magic), qc = i - qr * W;   // i / W, i % W (exact for i < 65536)
        // lane j owns candidate j (nc <= FF_CAND_STORE)
        const bool has = lane < nc;
        int32_t key = 0x7fffffff, krow = (int32_t)q_row;
        if (has) {
            const uint32_t code = c_cur & 0x1FFFu;
            const int ci = (int)(code >> 10), wr = (int)((code >> 5) & 31u), wc = (int)(code & 31u);
            const int f = ctx_frame(t, p.n_last_frames, ci);
            key = f * N + (qr - p.radius + wr) * W + (qc - p.radius + wc);
            krow = (int32_t)(clip_row0 + key);
        }
        const float inv_q = __ldg(S.inv + q_row);
        const float inv_k = __ldg(S.inv + krow);
        // warp-cooperative canonical dot, four candidates at a time: the query row and the key rows of a batch are in
        // flight together
        float my_dot = 0.f;
        if (STAGED) {
            float4 *wq = fin_stage + (size_t)warp * (1 + FIN_SB) * n4;      // row 0: query, rows 1..FIN_SB: keys
            __syncwarp();                                                  // the previous query's reads are done
            for (int e = lane; e < n4; e += 32) cp_async16(wq + e, qrow + e);
            for (int c0 = 0; c0 < nc; c0 += FIN_SB) {
                const int nb = min(FIN_SB, nc - c0);
                for (int u = 0; u < nb; ++u) {
                    const int32_t row = __shfl_sync(0xffffffffu, krow, c0 + u);
                    const float4 *kp = reinterpret_cast<const float4 *>(S.x + (int64_t)row * S.ld);
                    float4 *dst = wq + (size_t)(1 + u) * n4;
                    for (int e = lane; e < n4; e += 32) cp_async16(dst + e, kp + e);
                }
                cp_async_wait_all();
                __syncwarp();
                for (int u0 = 0; u0 < nb; u0 += 4) {
                    float acc[4] = {0.f, 0.f, 0.f, 0.f};
                    const float4 *k0 = wq + (size_t)(1 + u0) * n4;
                    const int nu = min(4, nb - u0);                       // warp-uniform
#pragma unroll 3
                    for (int e = lane; e < n4; e += 32) {
                        const float4 x = wq[e];
                        acc[0] = fma4_chain(acc[0], x, k0[e]);
                        if (nu > 1) acc[1] = fma4_chain(acc[1], x, k0[n4 + e]);
                        if (nu > 2) acc[2] = fma4_chain(acc[2], x, k0[2 * n4 + e]);
                        if (nu > 3) acc[3] = fma4_chain(acc[3], x, k0[3 * n4 + e]);
                    }
                    const float red = fin_reduce4(acc, lane);
                    const int u_me = lane - (c0 + u0);                    // candidate `lane` is accumulator u_me of this sub-batch
                    const float mine = __shfl_sync(0xffffffffu, red, ((u_me & 2) << 3) | ((u_me & 1) << 3));
                    if (u_me >= 0 && u_me < nu) my_dot = mine;
                }
                __syncwarp();                                              // before the next batch overwrites the key rows
            }
        }
        for (int c0 = 0; !STAGED && c0 < nc; c0 += FIN_BATCH) {
            const float4 *kp[FIN_BATCH];
#pragma unroll
            for (int u = 0; u < FIN_BATCH; ++u) {
                const int32_t row = __shfl_sync(0xffffffffu, krow, (c0 + u < nc) ? c0 + u : c0);
                kp[u] = reinterpret_cast<const float4 *>(S.x + (int64_t)row * S.ld);
            }
            float acc[FIN_BATCH] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 3
            for (int e = lane; e < n4; e += 32) {
                const float4 x = __ldg(qrow + e);
                float4 y[FIN_BATCH];
#pragma unroll
                for (int u = 0; u < FIN_BATCH; ++u) y[u] = __ldg(kp[u] + e);
#pragma unroll
                for (int u = 0; u < FIN_BATCH; ++u) acc[u] = fma4_chain(acc[u], x, y[u]);
            }
            const float red = fin_reduce4(acc, lane);
            const float mine = __shfl_sync(0xffffffffu, red, src_lane);
            if ((lane >> 2) == (c0 >> 2)) my_dot = mine;
        }
        // reference selection (mask_propagation.py:432-436) on the <= 16 exact affinities: canonical order (affinity desc,
        // key asc) by counting, k-th value, everything >= it kept (ties), normalised
        const float my_aff = has ? affinity_from_sim(sim_from_dot(my_dot, inv_q, inv_k), p.temperature) : -1.f;
        int rank = 0;
        for (int j = 0; j < nc; ++j) {
            const float a = __shfl_sync(0xffffffffu, my_aff, j);
            const int32_t kk = __shfl_sync(0xffffffffu, key, j);
            rank += (a > my_aff || (a == my_aff && kk < key)) ? 1 : 0;
        }
        // lane r takes the candidate of rank r: the sorted layout of the exact engine's list (slot = lane), so that the
        // normalising sum below adds the same values in the same butterfly positions -> identical bits
        int src = 0;
        for (int j = 0; j < nc; ++j)
            if (__shfl_sync(0xffffffffu, rank, j) == lane) src = j;
        const float s_aff = __shfl_sync(0xffffffffu, my_aff, src);
        const int32_t s_key = __shfl_sync(0xffffffffu, key, src);
        const bool in_list = lane < nc;
        float kth = -1.f;                               // fewer than k candidates (tiny windows): all kept
        if (nc >= p.topk) kth = __shfl_sync(0xffffffffu, s_aff, p.topk - 1);
        const bool keep = in_list && s_aff >= kth;
        const int m = __popc(__ballot_sync(0xffffffffu, keep));                 // m <= nc <= kw
        const float sum = warp_sum(keep ? s_aff : 0.f);
        if (lane < kw) {
            sel_w[(size_t)qid * kw + lane] = keep ? __fdiv_rn(s_aff, sum) : 0.f;
            sel_k[(size_t)qid * kw + lane] = keep ? s_key : -1;
        }
        if (lane == 0) sel_cnt[qid] = m;
        st_sel += (unsigned long long)m;
        st_ties += (m > p.topk);
        st_cand += (unsigned long long)nc;
    }
    if (lane == 0) {
        atomicAdd(&s_stat[0], st_sel);
        atomicAdd(&s_stat[1], st_ties);
        atomicAdd(&s_stat[3], st_cand);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_stat[0]) atomicAdd(&stats[1], s_stat[0]);
        if (s_stat[1]) atomicAdd(&stats[2], s_stat[1]);
        if (s_stat[3]) atomicAdd(&stats[3], s_stat[3]);
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)sym;
    }
    return fn;
}

int tc_make_map(CUtensorMap *m, const void *base, int64_t rows, int Dp, int box_rows) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return TIMET_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)Dp, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)Dp * sizeof(__half)};
    const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld Dp=%d box_rows=%d)", (int)r, (long long)rows, Dp, box_rows);
        return TIMET_ERR_CUDA;
    }
    return TIMET_OK;
}

// Query-tile map of the column-blocked persistent kernel (ff_tc3.cu): the same fp16 rows viewed as
// {Dp, grid_w patch columns, grid_rows = clips x frames x grid_h}; one box {64, 8, 4} fetches a 4 x 8 block of the patch
// grid (8 consecutive columns of 4 consecutive grid rows) as 32 consecutive tile rows; box_cols < 8 for the last, narrower
// column block.  Columns >= grid_w and grid rows past the end are out of bounds and arrive as zeros.
int tc_make_map_colblk(CUtensorMap *m, const void *base, int64_t grid_rows, int Dp, int grid_w, int box_cols) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return TIMET_ERR_CUDA;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)Dp, (cuuint64_t)grid_w, (cuuint64_t)grid_rows};
    const cuuint64_t strides[2] = {(cuuint64_t)Dp * sizeof(__half), (cuuint64_t)grid_w * Dp * sizeof(__half)};
    const cuuint32_t box[3] = {64u, (cuuint32_t)box_cols, 4u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (column-blocked query tile) failed with CUresult %d (grid_rows=%lld Dp=%d W=%d)", (int)r,
                  (long long)grid_rows, Dp, grid_w);
        return TIMET_ERR_CUDA;
    }
    return TIMET_OK;
}

size_t tc_smem_bytes(const TcGeom &G) {
    const size_t a_res = G.a_resident ? (size_t)G.NKC * 16384 : 0, a_stage = G.a_resident ? 0 : 16384;
    return 1024 + a_res + (size_t)G.nstages * ((size_t)G.NT * 128 + a_stage) + (size_t)TC_GROUPS * TC_CAP * 128 * 4 + sizeof(TcSmemCtl) + 64;
}

bool tc_geometry(const timet_ff_params &p, const FFLayout &L, TcGeom *G) {
    const int W = p.grid_w, H = p.grid_h;
    if (p.radius < 1 || p.radius > 15) return false;
    if (p.n_last_frames > 7 || p.topk > 8) return false;   // 3-bit context slot in the packed candidate; k + ties must fit 16 slots
    if (W > 128 || L.Dp > 64 * TC_MAX_NKC_STREAM) return false;
    G->a_resident = (L.Dp <= 64 * TC_MAX_NKC) ? 1 : 0;
    int g = W, b = 16;
    while (b) { const int tmp = g % b; g = b; b = tmp; }   // gcd(W, 16)
    const int qrows = 16 / g;
    if (qrows * W > 256) return false;
    // prefer 4 accumulator buffers of 128 TMEM columns (4 independent epilogue groups); else 2 x 256
    // 4 accumulator buffers of 128 TMEM columns (one epilogue group each) or 2 x 256 (two groups per buffer,
    // splitting the key rows).  Bigger key tiles amortise the per-tile pipeline overhead better, so 2 x 256 is the
    // default; TIMET_TC_NBUF=4 selects the other layout for experiments.
    const EnvCfg &E = env_cfg();
    int RPC = (256 / W) / qrows * qrows;
    G->nbuf = 2;
    if (E.tc_nbuf == 4 && (128 / W) / qrows >= 1) { RPC = (128 / W) / qrows * qrows; G->nbuf = 4; }
    G->buf_cols = 512 / G->nbuf;
    G->H = H; G->W = W; G->N = L.N; G->Dp = L.Dp; G->NKC = L.Dp / 64;
    G->QR = (128 / W) < H ? (128 / W) : H;
    G->tiles_per_frame = (H + G->QR - 1) / G->QR;
    G->RPC = RPC; G->NT = RPC * W; G->qrows = qrows;
    G->n_clips = p.n_clips; G->n_frames = p.n_frames; G->nT = L.nT; G->t_begin = p.t_begin;
    G->n_last = p.n_last_frames; G->radius = p.radius; G->topk = p.topk;
    G->flags = E.tc_flags;
    G->colblk = 0; G->ncl = 8; G->a_chunk_bytes = 16384;
    G->clip_group = E.tc_clip_group >= 1 ? E.tc_clip_group : TC_CLIP_GROUP;
    G->total_tiles = (int64_t)p.n_clips * L.nT * G->tiles_per_frame;
    G->nstages = TC_MAX_STAGES;                                  // as deep a B ring as shared memory allows
    if (E.tc_stages >= 2 && E.tc_stages <= TC_MAX_STAGES) G->nstages = E.tc_stages;
    while (tc_smem_bytes(*G) > 227 * 1024 && G->nstages > 2) G->nstages--;
    return tc_smem_bytes(*G) <= 227 * 1024;
}


// FLOPs the tensor-core kernel ISSUES for this problem (what the tensor pipe executes): every key tile inside a
// query tile's window band, M padded to the 128-row MMA, N = the tile's key columns, K = Dp.  0 if unsupported.
double ff_tc_executed_flops(const timet_ff_params &p) {
    TcGeom G;
    const FFLayout L = ff_layout(p);
    if (!tc_geometry(p, L, &G)) return 0.0;
    double total = 0.0;
    for (int t = p.t_begin; t < p.n_frames; ++t) {
        const int nctx = ctx_count(t, p.n_last_frames);
        for (int qt = 0; qt < G.tiles_per_frame; ++qt) {
            const int qr0 = qt * G.QR, qr1 = (G.H - 1 < qr0 + G.QR - 1) ? G.H - 1 : qr0 + G.QR - 1;
            const int kr_lo = (qr0 - G.radius > 0) ? qr0 - G.radius : 0;
            const int kr_hi = (G.H - 1 < qr1 + G.radius) ? G.H - 1 : qr1 + G.radius;
            const int nchunks = (kr_hi - kr_lo + G.RPC) / G.RPC;
            for (int ch = 0; ch < nchunks; ++ch) {
                int rc = kr_hi + 1 - (kr_lo + ch * G.RPC);
                if (rc > G.RPC) rc = G.RPC;
                int n_mma = (rc + G.qrows - 1) / G.qrows * G.qrows * G.W;
                if (n_mma > G.NT) n_mma = G.NT;
                total += (double)nctx * 2.0 * 128.0 * n_mma * G.Dp;
            }
        }
    }
    return total * p.n_clips;
}

bool ff_tc_supported(const timet_ff_params &p) {
    TcGeom G;
    const FFLayout L = ff_layout(p);
    return tc_geometry(p, L, &G);
}

int ff_select_exact_run(const timet_ff_params &p, const FFLayout &L, const float *feats, char *ws, const int32_t *qlist,
                        const unsigned int *qcount, int64_t max_items, cudaStream_t st);

int ff_select_tc_persist_launch(const timet_ff_params &p, const FFLayout &L, char *ws, cudaStream_t st);

// optional CUDA events recorded on the stream right before / after the nomination (tensor-core) kernel, so a
// caller can time the dominant kernel alone (bench.py roofline); set through timet_ff_select_timed
thread_local cudaEvent_t g_ev_nominate_begin = nullptr, g_ev_nominate_end = nullptr;

int ff_select_tc_launch(const timet_ff_params &p, const FFLayout &L, const float *feats, char *ws, cudaStream_t st) {
    TcGeom G;
    if (!tc_geometry(p, L, &G)) {
        set_error("tensor-core engine: unsupported shape");
        return TIMET_ERR_UNSUPPORTED;
    }
    int rc;
    uint32_t *cand = reinterpret_cast<uint32_t *>(ws + L.off_cand);
    uint32_t *meta = reinterpret_cast<uint32_t *>(ws + L.off_cand_meta);
    if (g_ev_nominate_begin) TIMET_CUDA(cudaEventRecord(g_ev_nominate_begin, st));
    const EnvCfg &E = env_cfg();
    rc = TIMET_ERR_UNSUPPORTED;
    if (rc == TIMET_ERR_UNSUPPORTED) {
        // persistent kernel (ff_tc3.cu): one CTA per SM walks the work items; TIMET_TC_PERSIST=0 selects the per-tile kernel
        const bool debug = E.tc_trace || E.tc_flags != 0;
        if (E.tc_persist && !debug) {
            rc = ff_select_tc_persist_launch(p, L, ws, st);
            if (rc != TIMET_OK && rc != TIMET_ERR_UNSUPPORTED) return rc;
        }
    }
    if (rc == TIMET_ERR_UNSUPPORTED) {
        const __half *fn16 = reinterpret_cast<const __half *>(ws + L.off_fn16);
        CUtensorMap map_a, map_b;
        if ((rc = tc_make_map(&map_a, fn16, L.rows + 256, L.Dp, 128)) != TIMET_OK) return rc;
        if ((rc = tc_make_map(&map_b, fn16, L.rows + 256, L.Dp, G.NT)) != TIMET_OK) return rc;
        const size_t smem = tc_smem_bytes(G);
        TIMET_CUDA(cudaFuncSetAttribute(ff_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        unsigned long long *trace = E.tc_trace ? reinterpret_cast<unsigned long long *>(ws + L.off_trace) : nullptr;
        ff_tc_kernel<false><<<(unsigned)G.total_tiles, TC_THREADS, smem, st>>>(map_a, map_b, G, cand, meta, -1, nullptr, trace);
        TIMET_LAUNCHED();
    }

    if (g_ev_nominate_end) TIMET_CUDA(cudaEventRecord(g_ev_nominate_end, st));
    unsigned int *redo_count = reinterpret_cast<unsigned int *>(ws + L.off_redo);
    int32_t *redo_list = reinterpret_cast<int32_t *>(ws + L.off_redo + 256);
    const FFSrc S = ff_src(p, L, feats, ws);
    const int64_t blocks = (L.queries + FIN_QPB - 1) / FIN_QPB;
    TIMET_CHECK_ARG(L.queries < (1ll << 31), "ff_select: %lld queries exceed the 32-bit query index of the tensor-core engine", (long long)L.queries);
    const uint32_t w_magic = (uint32_t)((0x100000000ull + (uint64_t)p.grid_w - 1) / (uint64_t)p.grid_w);   // ceil(2^32 / W)
    const bool staged = E.fin_staged && S.n4 <= FIN_STAGE_N4;
    const size_t fin_smem = staged ? (size_t)FIN_WARPS * (1 + FIN_SB) * S.n4 * sizeof(float4) : 0;
    auto fin = staged ? ff_finalize_kernel<true> : ff_finalize_kernel<false>;
    if (fin_smem > 48 * 1024) TIMET_CUDA(cudaFuncSetAttribute(fin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem));
    fin<<<(unsigned)blocks, FIN_WARPS * 32, fin_smem, st>>>(
        p, L.N, S, L.nT, L.kw, w_magic, cand, meta, reinterpret_cast<float *>(ws + L.off_sel_w),
        reinterpret_cast<int32_t *>(ws + L.off_sel_k), reinterpret_cast<int32_t *>(ws + L.off_sel_cnt),
        reinterpret_cast<unsigned long long *>(ws + L.off_stats), redo_list, redo_count, (uint32_t)L.queries);
    TIMET_LAUNCHED();
    // overflowed queries: exact scan (device-side count; a fixed small grid loops over the list)
    if ((rc = ff_select_exact_run(p, L, feats, ws, redo_list, redo_count, (int64_t)num_sms() * 2 * 8, st)) != TIMET_OK) return rc;   // two CTAs per SM: the list is usually empty (grid-stride loop otherwise)
    return TIMET_OK;
}

// debug: run ONE query tile and dump its raw fp32 accumulators [ntiles, 128, 256] (tests only)
int ff_tc_debug_tile(const timet_ff_params &p, const FFLayout &L, char *ws, int64_t tile_id, float *dump, cudaStream_t st) {
    TcGeom G;
    if (!tc_geometry(p, L, &G)) {
        set_error("tensor-core engine: unsupported shape");
        return TIMET_ERR_UNSUPPORTED;
    }
    TIMET_CHECK_ARG(tile_id >= 0 && tile_id < G.total_tiles, "debug tile %lld out of range", (long long)tile_id);
    const __half *fn16 = reinterpret_cast<const __half *>(ws + L.off_fn16);
    CUtensorMap map_a, map_b;
    int rc;
    if ((rc = tc_make_map(&map_a, fn16, L.rows + 256, L.Dp, 128)) != TIMET_OK) return rc;
    if ((rc = tc_make_map(&map_b, fn16, L.rows + 256, L.Dp, G.NT)) != TIMET_OK) return rc;
    const size_t smem = tc_smem_bytes(G);
    TIMET_CUDA(cudaFuncSetAttribute(ff_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ff_tc_kernel<true><<<1, TC_THREADS, smem, st>>>(map_a, map_b, G, nullptr, nullptr, tile_id, dump, nullptr);
    TIMET_LAUNCHED();
    return TIMET_OK;
}

int ff_tc_debug_trace(const timet_ff_params &p, const FFLayout &L, const char *ws, unsigned long long *out, int n_ctas, cudaStream_t st) {
    (void)p;
    TIMET_CHECK_ARG(n_ctas >= 1 && n_ctas <= FF_TRACE_CTAS, "debug trace: n_ctas out of range");
    TIMET_CUDA(cudaMemcpyAsync(out, ws + L.off_trace, (size_t)n_ctas * FF_TRACE_SLOTS * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    return TIMET_OK;
}

}  // namespace timet
