// FF stage 2, TC engine, CTA-PAIR variant: tcgen05.mma.cta_group::2 (M = 256 over two SMs).
//
// Why: with one CTA per query tile every key tile B is pulled from L2 and read from shared memory for only
// 128 query rows; the 1-CTA kernel is bound by exactly that traffic (shared memory: A + B operand reads + TMA
// writes = 162 B/clk of the 128 B/clk an SM has; L2->SM: 2.2 GB per step).  Target frames t and t-1 of the same
// (clip, query tile) need the SAME key tiles (their context sets differ by one frame), so they run as a CTA pair:
// each CTA keeps its own 128-query A tile resident, loads HALF of every key tile, and one tcgen05.mma.cta_group::2
// (issued by the leader) multiplies both query tiles with the whole key tile.  B bytes per query row halve, in L2
// and in shared memory.
//
// Protocol (identical smem layout in both CTAs; rank 0 = leader = the later target frame):
//   producers (warp 0 of each CTA)  wait own empty[s]; TMA own B half -> own smem, complete_tx on the LEADER's full[s]
//   MMA issuer (leader warp 1)       wait leader tmem_empty[b] (arrivals from both CTAs' epilogues), wait full[s],
//                                    4 x tcgen05.mma.cta_group::2, commit.multicast -> empty[s] in both CTAs,
//                                    after the tile commit.multicast -> tmem_full[b] in both CTAs
//   epilogues (warps 4-19, both)     as in ff_tc.cu, each CTA on its own TMEM half / own target frame; tiles of a
//                                    context frame the CTA's target does not use are released unscanned;
//                                    release = remote mbarrier arrive on the leader's tmem_empty[b]
// Selection semantics, candidate encoding, finalize: unchanged (ff_tc.cu / ff_tc_dev.cuh).
#include <stdlib.h>

#include "ff_tc_dev.cuh"

namespace timet {

namespace p2 {

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(ptx::smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    uint32_t spins = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (++spins > (1u << 24)) {
            printf("timet: pair-kernel mbarrier wait timed out (block %d thread %d bar %p parity %u)\n", (int)blockIdx.x,
                   (int)threadIdx.x, (void *)bar, parity);
            __trap();
        }
    }
}
// 2-SM TMA load: tile lands in the executing CTA's smem, transaction bytes complete on `bar_cluster_addr`
// (the leader CTA's mbarrier, a shared::cluster address)
__device__ __forceinline__ void tma_load_2d_2sm(void *smem_dst, const CUtensorMap *m, int c0, int c1, uint32_t bar_cluster_addr) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc2(uint32_t *smem_result) {   // one full warp in EACH CTA, same warp id
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(smem_result)), "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// D[tmem, both CTAs] (+)= A[smem, 128 rows per CTA] * B[smem, N/2 rows per CTA]^T ; issued by ONE thread of the leader
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// arrive on the mbarrier at this smem offset in every CTA of `mask` once all prior MMAs of this thread retired
__device__ __forceinline__ void umma_commit_2sm(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(ptx::smem_u32(bar)), "h"(mask)
                 : "memory");
}

}  // namespace p2

struct PairGeom {
    TcGeom g;
    int pairs_per_cq;          // ceil(nT / 2) target pairs per (clip, query tile)
    int64_t total_pairs;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
ff_tc_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, PairGeom PG,
                  uint32_t *__restrict__ cand, uint32_t *__restrict__ cand_meta, unsigned long long *__restrict__ trace) {
    const TcGeom &G = PG.g;
    extern __shared__ uint8_t smem_raw[];
    // identical carve in both CTAs: [A: NKC x 16 KB][B halves: nstages x (NT/2)*128][lists][ctl]
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    const uint32_t half_rows = (uint32_t)G.NT / 2;
    const uint32_t b_stage_bytes = half_rows * 128u;
    uint8_t *sB = sA + (size_t)G.NKC * 16384;
    uint32_t *sList = reinterpret_cast<uint32_t *>(sB + (size_t)G.nstages * b_stage_bytes);
    TcSmemCtl *ctl = reinterpret_cast<TcSmemCtl *>(sList + TC_GROUPS * TC_CAP * 128);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto stamp = [&](int sidx) {
        if (trace && blockIdx.x < FF_TRACE_CTAS) {
            unsigned long long tns;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
            trace[(size_t)blockIdx.x * FF_TRACE_SLOTS + sidx] = tns;
        }
    };
    if (threadIdx.x == 0) stamp(0);
    const uint32_t rank = p2::cluster_ctarank();
    const bool leader = rank == 0;

    // ---- which pair: clip groups (L2 locality), heavy pairs first
    const int64_t pair_id = (int64_t)(blockIdx.x >> 1);
    const int per_clip = PG.pairs_per_cq * G.tiles_per_frame;
    const int grp = (int)(pair_id / ((int64_t)G.clip_group * per_clip));
    const int grp_clips = min(G.clip_group, G.n_clips - grp * G.clip_group);
    const int in_grp = (int)(pair_id - (int64_t)grp * G.clip_group * per_clip);
    const int per_p = grp_clips * G.tiles_per_frame;
    const int pidx = in_grp / per_p;
    const int rem = in_grp - pidx * per_p;
    const int clip = grp * G.clip_group + rem / G.tiles_per_frame, qt = rem % G.tiles_per_frame;
    const int t_hi = G.n_frames - 1 - 2 * pidx;
    const int t_lo = (t_hi - 1 >= G.t_begin) ? t_hi - 1 : t_hi;   // odd count: the last target pairs with itself
    const int t_own = leader ? t_hi : t_lo;

    const int qr0 = qt * G.QR;
    const int qr1 = min(G.H - 1, qr0 + G.QR - 1);
    const int nq = (qr1 - qr0 + 1) * G.W;
    const int kr_lo = max(0, qr0 - G.radius), kr_hi = min(G.H - 1, qr1 + G.radius);
    const int nchunks = (kr_hi - kr_lo + G.RPC) / G.RPC;
    // union of the two context sets: frame 0, then frames lo_u .. t_hi - 1
    const int lo_u = ctx_lo(t_lo, G.n_last);
    const int nunion = 1 + (t_hi - lo_u);
    const int ntiles = nunion * nchunks;
    const int own_lo = ctx_lo(t_own, G.n_last);
    const int64_t clip_row0 = (int64_t)clip * G.n_frames * G.N;
    const int q_row0 = (int)(clip_row0 + (int64_t)t_own * G.N + qr0 * G.W);

    // ---- one-time setup
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_a);
        ptx::prefetch_tensormap(&map_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < G.nstages; ++s) { ptx::mbar_init(&ctl->full[s], 1); ptx::mbar_init(&ctl->empty[s], 1); }
        ptx::mbar_init(&ctl->a_full, 1);
        // tmem_empty lives on the leader: every epilogue warp of BOTH CTAs that reads the buffer arrives once per tile
        for (int b = 0; b < G.nbuf; ++b) { ptx::mbar_init(&ctl->tmem_full[b], 1); ptx::mbar_init(&ctl->tmem_empty[b], 2 * (G.nbuf == 4 ? 4 : 8)); }
        ptx::fence_barrier_init();
    }
    if (warp == 2) p2::tmem_alloc2<512>(&ctl->tmem_base);
    if (warp == 3) for (int i = lane; i < 128; i += 32) ctl->thr_sh[i] = thr_enc(-INFINITY);
    ptx::tc_fence_before();
    __syncthreads();
    p2::cluster_sync();                      // both CTAs' barriers are initialised before any remote signal
    ptx::tc_fence_after();
    const uint32_t tmem_base = ctl->tmem_base;
    if (threadIdx.x == 0) stamp(1);

    if (warp == 0) {
        // =========================== TMA producer (both CTAs) ===========================
        if (lane == 0) {
            const uint32_t a_full_leader = p2::mapa(ptx::smem_u32(&ctl->a_full), 0);
            if (leader) ptx::mbar_expect_tx(&ctl->a_full, 2u * (uint32_t)G.NKC * 16384u);
            for (int kc = 0; kc < G.NKC; ++kc) p2::tma_load_2d_2sm(sA + kc * 16384, &map_a, kc * 64, q_row0, a_full_leader);
            const uint32_t full0_leader = p2::mapa(ptx::smem_u32(&ctl->full[0]), 0);
            uint32_t stage = 0, phase = 0;
            for (int u = 0; u < nunion; ++u) {
                const int f = (u == 0) ? 0 : lo_u + u - 1;
                for (int ch = 0; ch < nchunks; ++ch) {
                    const int rc = min(G.RPC, kr_hi + 1 - (kr_lo + ch * G.RPC));
                    const int n_mma = min(G.NT, (rc + G.qrows - 1) / G.qrows * G.qrows * G.W);
                    // the N columns of a cta_group::2 MMA are split at N/2: leader supplies keys [0, N/2), peer [N/2, N)
                    const int k_row0 = (int)(clip_row0 + (int64_t)f * G.N + (kr_lo + ch * G.RPC) * G.W) + (leader ? 0 : n_mma / 2);
                    for (int kc = 0; kc < G.NKC; ++kc) {
                        p2::mbar_wait_cluster(&ctl->empty[stage], phase ^ 1u);
                        if (leader) ptx::mbar_expect_tx(&ctl->full[stage], 2u * b_stage_bytes);
                        p2::tma_load_2d_2sm(sB + (size_t)stage * b_stage_bytes, &map_b, kc * 64, k_row0, full0_leader + stage * 8u);
                        if (++stage == (uint32_t)G.nstages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer (leader only) ===========================
        if (leader && lane == 0) {
            p2::mbar_wait_cluster(&ctl->a_full, 0);
            ptx::tc_fence_after();
            stamp(2);
            const uint64_t da0 = ptx::umma_desc_sw128(ptx::smem_u32(sA)), db0 = ptx::umma_desc_sw128(ptx::smem_u32(sB));
            const uint32_t stage_step = b_stage_bytes >> 4;
            uint32_t stage = 0, phase = 0, buf = 0, use = 0;
            int ch = 0;
            for (int tile = 0; tile < ntiles; ++tile) {
                const int rc = min(G.RPC, kr_hi + 1 - (kr_lo + ch * G.RPC));
                const int n_mma = min(G.NT, (rc + G.qrows - 1) / G.qrows * G.qrows * G.W);
                const uint32_t idesc = ptx::umma_idesc_f16(256, n_mma);
                p2::mbar_wait_cluster(&ctl->tmem_empty[buf], (use & 1u) ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * (uint32_t)G.buf_cols;
                uint64_t da = da0;
                for (int kc = 0; kc < G.NKC; ++kc, da += 16384 >> 4) {
                    p2::mbar_wait_cluster(&ctl->full[stage], phase);
                    ptx::tc_fence_after();
                    const uint64_t db = db0 + (uint64_t)(stage * stage_step);
                    p2::umma_f16_2sm(d_tmem, da, db, idesc, kc != 0);
                    p2::umma_f16_2sm(d_tmem, da + 2, db + 2, idesc, true);
                    p2::umma_f16_2sm(d_tmem, da + 4, db + 4, idesc, true);
                    p2::umma_f16_2sm(d_tmem, da + 6, db + 6, idesc, true);
                    p2::umma_commit_2sm(&ctl->empty[stage], 0b11);      // frees the stage in both CTAs
                    if (++stage == (uint32_t)G.nstages) { stage = 0; phase ^= 1u; }
                }
                p2::umma_commit_2sm(&ctl->tmem_full[buf], 0b11);        // accumulators ready in both CTAs
                if (++buf == (uint32_t)G.nbuf) { buf = 0; ++use; }
                if (++ch == nchunks) ch = 0;
            }
            stamp(3);
        }
    } else if (warp >= 4) {
        // =========================== epilogue groups (both CTAs, own target frame) ===========================
        const int g = (warp - 4) >> 2;
        const int buf = (G.nbuf == 4) ? g : (g & 1);
        const int row_par = (G.nbuf == 4) ? -1 : (g >> 1);
        const int qi = ((warp & 3) << 5) + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) << 5) << 16;
        const uint32_t list = ptx::smem_u32(sList + (size_t)g * TC_CAP * 128 + qi);
        constexpr int half = TC_CAP / 2;
        const bool valid = qi < nq;
        const int qrow = qr0 + qi / G.W, qcol = qi % G.W;
        const int r_lo = qrow - G.radius, r_hi = qrow + G.radius;
        const int c_lo = qcol - G.radius;
        const int c_lo_cl = max(c_lo, 0), c_hi_cl = min(qcol + G.radius, G.W - 1);
        float thr = (G.flags & 2) ? INFINITY : -INFINITY;
        int cnt = 0, lost = 0;

        for (int tile = buf; tile < ntiles; tile += G.nbuf) {
            const int u = tile / nchunks, ch = tile - u * nchunks;
            const int f = (u == 0) ? 0 : lo_u + u - 1;
            const bool mine = (f == 0) || (f >= own_lo && f <= t_own - 1);     // context frame of MY target?
            const int ci = (f == 0) ? 0 : f - own_lo + 1;
            const int kr_start = kr_lo + ch * G.RPC;
            const int rc = min(G.RPC, kr_hi + 1 - kr_start);
            const uint32_t use = (uint32_t)(tile / G.nbuf);
            if (lane == 0) p2::mbar_wait_cluster(&ctl->tmem_full[buf], use & 1u);
            __syncwarp();
            ptx::tc_fence_after();
            const uint32_t t_acc = tmem_base + (uint32_t)(buf * G.buf_cols) + lane_base;
            thr = fmaxf(thr, thr_dec(ctl->thr_sh[qi]));

            for (int rr = 0; rr < ((mine && !(G.flags & 1)) ? rc : 0); ++rr) {
                if (row_par >= 0 && (rr & 1) != row_par) continue;
                const int kr = kr_start + rr;
                const bool row_ok = valid && kr >= r_lo && kr <= r_hi;
                if (!__any_sync(0xffffffffu, row_ok)) continue;
                const int code_row = (ci << 10) | ((kr - r_lo) << 5);
                for (int cb = 0; cb < G.W; cb += 16) {
                    int col0 = rr * G.W + cb;
                    const int shift = max(0, col0 + 16 - G.buf_cols);
                    col0 -= shift;
                    const int cb_eff = cb - shift;
                    uint32_t r[16];
                    ptx::tmem_ld_32x16(t_acc + (uint32_t)col0, r);
                    const int lo = max(c_lo_cl - cb_eff, shift), hi = min(c_hi_cl - cb_eff, 15);
                    uint32_t wmask = 0u;
                    if (row_ok && hi >= lo) wmask = (0xFFFFu >> (15 - hi)) & (0xFFFFu << lo) & 0xFFFFu;
                    const uint32_t code0 = (uint32_t)(code_row + (cb_eff - c_lo));
                    uint32_t slot = list + (uint32_t)cnt * TC_SLOT_STRIDE;
                    ptx::tmem_ld_wait();
#define TC_OFFER(E) tc_offer<(1u << (E))>(slot, __uint_as_float(r[E]), thr, wmask, code0 + (E));
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
            // release the accumulator buffer: arrive on the LEADER's tmem_empty (remote for the peer CTA)
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) p2::mbar_arrive_cluster(p2::mapa(ptx::smem_u32(&ctl->tmem_empty[buf]), 0));
        }

        if (warp == 4 && lane == 0) { stamp(4); if (!leader) { stamp(2); stamp(3); } }
        // ---- final: agree on the per-query threshold, filter, merge the 4 lists, publish
        {
            const float before = thr;
            tc_compact(list, cnt, thr, lost, G.topk, half);
            if (thr > before) atomicMax(&ctl->thr_sh[qi], thr_enc(thr));
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
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
            tc_compact(list, cnt, thr, lost, G.topk, FF_CAND_STORE);
            if (valid) {
                const int64_t q = ((int64_t)clip * G.nT + (t_own - G.t_begin)) * G.N + qr0 * G.W + qi;
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

    if (warp == 4 && lane == 0) stamp(6);
    // ---- teardown: neither CTA may leave (or free TMEM) while the other can still touch its smem / barriers
    ptx::tc_fence_before();
    __syncthreads();
    p2::cluster_sync();
    if (warp == 2) {
        ptx::tc_fence_after();
        p2::tmem_dealloc2<512>(tmem_base);
    }
    if (threadIdx.x == 0) stamp(7);
}

static size_t pair_smem_bytes(const TcGeom &G) {
    return 1024 + (size_t)G.NKC * 16384 + (size_t)G.nstages * (G.NT / 2) * 128 + (size_t)TC_GROUPS * TC_CAP * 128 * 4 +
           sizeof(TcSmemCtl) + 64;
}

// CTA-pair launch; returns TIMET_ERR_UNSUPPORTED if the shape does not qualify (caller falls back to the 1-CTA kernel)
int ff_select_tc_pair_launch(const timet_ff_params &p, const FFLayout &L, char *ws, cudaStream_t st) {
    PairGeom PG;
    if (!tc_geometry(p, L, &PG.g)) return TIMET_ERR_UNSUPPORTED;
    TcGeom &G = PG.g;
    if ((G.NT / 2) % 8 != 0 || !G.a_resident) return TIMET_ERR_UNSUPPORTED;
    G.nstages = TC_MAX_STAGES;
    const EnvCfg &E = env_cfg();
    if (E.tc_stages >= 2 && E.tc_stages <= TC_MAX_STAGES) G.nstages = E.tc_stages;
    while (pair_smem_bytes(G) > 227 * 1024 && G.nstages > 2) G.nstages--;
    if (pair_smem_bytes(G) > 227 * 1024) return TIMET_ERR_UNSUPPORTED;
    PG.pairs_per_cq = (L.nT + 1) / 2;
    PG.total_pairs = (int64_t)p.n_clips * G.tiles_per_frame * PG.pairs_per_cq;

    const __half *fn16 = reinterpret_cast<const __half *>(ws + L.off_fn16);
    CUtensorMap map_a, map_b;
    int rc;
    if ((rc = tc_make_map(&map_a, fn16, L.rows + 256, L.Dp, 128)) != TIMET_OK) return rc;
    if ((rc = tc_make_map(&map_b, fn16, L.rows + 256, L.Dp, G.NT / 2)) != TIMET_OK) return rc;
    const size_t smem = pair_smem_bytes(G);
    TIMET_CUDA(cudaFuncSetAttribute(ff_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint32_t *cand = reinterpret_cast<uint32_t *>(ws + L.off_cand);
    uint32_t *meta = reinterpret_cast<uint32_t *>(ws + L.off_cand_meta);
    unsigned long long *trace = E.tc_trace ? reinterpret_cast<unsigned long long *>(ws + L.off_trace) : nullptr;
    ff_tc_pair_kernel<<<(unsigned)(2 * PG.total_pairs), TC_THREADS, smem, st>>>(map_a, map_b, PG, cand, meta, trace);
    TIMET_LAUNCHED();
    return TIMET_OK;
}

}  // namespace timet
