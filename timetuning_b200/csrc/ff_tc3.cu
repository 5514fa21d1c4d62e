// FF stage 2, TC engine, PERSISTENT variant of the 1-CTA kernel in ff_tc.cu (same maths, same candidate format).
//
// Why: with one CTA per (clip, target frame, query tile) about 10 us of every ~50 us CTA are not spent on MMAs:
// barrier/TMEM set-up, the load of the query tile, the drain of the last key tiles through the epilogue, the merge
// of the candidate lists and the publish (profiles/tc_trace.py); shared memory allows only one CTA per SM, so nothing
// overlaps them.  Here one CTA per SM walks a static sequence of work items (see item_at) and the three roles run
// decoupled: while the epilogue groups drain / merge / publish item i, the producer already loads the query tile of
// item i+1 (as soon as the last MMA of item i has retired: a_free barrier) and the MMA warp fills the TMEM buffers
// with its first key tiles.  TMEM buffer and smem ring phases simply continue across items.
//
// Hazards handled explicitly:
//   * query tile A is single-buffered        -> a_free mbarrier, committed after the last MMA of the item
//   * per-query shared threshold             -> two copies, alternating per item; group 1 resets the idle copy
//   * candidate lists of groups 1-3          -> they PUSH their entries into list 0 between two named barriers
//                                               (group 0 never reads another group's list while that group has
//                                               moved on to the next item)
//
// Issue path.  The producer and the MMA warp run their loops with ALL 32 lanes on warp-uniform values (the item sequence
// is a function of blockIdx and kernel parameters only) and elect one lane per TMA / MMA / commit.  With a divergent
// `if (lane == 0)` around the loops ptxas cannot keep descriptors in uniform registers and wraps every UTCHMMA / UTMALDG
// in an ELECT + 7 x R2UR "waterfall" (about 100 scalar instructions per 64-wide K chunk, more than the 504 cycles the
// four MMAs of the chunk take): the tensor pipe then idles on instruction issue (MMA/TMA side alone 0.31 -> 0.245 ms at
// BASELINE configs[1]).  Warp ids: epilogue 0-15 (TMEM lane quadrant = warp & 3), producer 16, MMA 17; giving the issue warps
// the highest ids (which the SMSP arbiter favours) made no measurable difference.
//
// TIMET_TC_PFLAGS (attribution switches, profiles/tc_kernel_time.py): 1 release tiles unscanned, 2 scan without appends,
// 4 TMEM loads only, 8 oldest-first context order, 32 nanosleep back-off while polling tmem_full, 64 shared threshold read
// once per tile instead of once per key row, 128 each TMEM buffer scanned by its own two groups only, 256 raster query tiles
// (no column blocks), 512 epilogue waits for a key tile with a suspend-time hint, 1024 no final merge / publish, 2048 full lists are
// emptied instead of compacted (1024 / 2048: timing attribution only, wrong results), 8192 no compaction while waiting for a key tile,
// 16384 padding rows inside every K chunk of A + 32-slot lists (two-stage key ring at 28 x 28 x 384), 32768 key rows always
// top-down, bits 20+ entries appended before a waiting warp compacts (default 6).
//
// Shared-memory budget at BASELINE configs[1] (28 x 28, D = 384), 227 KB per CTA: query tile 84 KB (6 K chunks of 112 rows),
// key ring 3 x 28 KB, candidate lists 4 groups x 24 slots x 128 queries x 4 B = 48 KB, control block 3.4 KB.  The ring is
// latency-bound (bytes in flight per SM against ~1 us of TMA latency), so whatever fits goes to ring stages.
#include <stdlib.h>

#include "ff_tc_dev.cuh"

namespace timet {

struct __align__(8) PsCtl {
    uint64_t full[TC_MAX_STAGES], empty[TC_MAX_STAGES], a_full, a_free, tmem_full[4], tmem_empty[4];
    uint64_t item_bar[4];      // DYN: item_id[k & 3] is valid (the producer fetches from a global counter, everyone follows)
    int32_t item_id[4];
    uint32_t tmem_base;
    uint32_t thr_sh[2][128];
    uint32_t xchg[TC_GROUPS][128];
};

struct Item {
    int t, clip, qr0, qr1, nq, kr_lo, kr_hi, nchunks, nctx, ntiles, q_row0;
    int64_t clip_row0;
};

__device__ __forceinline__ Item item_geom(const TcGeom &G, int64_t id) {
    Item I;
    const int per_clip = G.nT * G.tiles_per_frame;
    const int grp = (int)(id / ((int64_t)G.clip_group * per_clip));
    const int grp_clips = min(G.clip_group, G.n_clips - grp * G.clip_group);
    const int in_grp = (int)(id - (int64_t)grp * G.clip_group * per_clip);
    const int per_t = grp_clips * G.tiles_per_frame;
    const int tdesc = in_grp / per_t;
    const int rem = in_grp - tdesc * per_t;
    I.t = G.n_frames - 1 - tdesc;
    I.clip = grp * G.clip_group + rem / G.tiles_per_frame;
    const int qt = rem % G.tiles_per_frame;
    I.qr0 = qt * G.QR;
    const int qr1 = min(G.H - 1, I.qr0 + G.QR - 1);
    I.qr1 = qr1;
    I.nq = (qr1 - I.qr0 + 1) * G.W;
    I.kr_lo = max(0, I.qr0 - G.radius);
    I.kr_hi = min(G.H - 1, qr1 + G.radius);
    I.nchunks = (I.kr_hi - I.kr_lo + G.RPC) / G.RPC;
    I.nctx = ctx_count(I.t, G.n_last);
    I.ntiles = I.nctx * I.nchunks;
    I.clip_row0 = (int64_t)I.clip * G.n_frames * G.N;
    I.q_row0 = (int)(I.clip_row0 + (int64_t)I.t * G.N + I.qr0 * G.W);
    return I;
}

// k-th work item of this CTA, or -1.  Items are numbered heavy-first (item_geom); row k of the schedule holds ids
// [k * grid, (k + 1) * grid) and is walked alternately left-to-right and right-to-left ("snake"), so that a CTA that got a
// heavier item in one row gets a lighter one in the next.  Pure function of (blockIdx, k): every role derives the same
// sequence without communication, and the values stay warp-uniform for ptxas.
__device__ __forceinline__ int64_t item_at(uint32_t k, int64_t total) {
    const int64_t row0 = (int64_t)k * gridDim.x;
    const int64_t id = row0 + ((k & 1u) ? (int64_t)(gridDim.x - 1u - blockIdx.x) : (int64_t)blockIdx.x);
    return (row0 < total && id < total) ? id : -1;
}

// CAP = slots per (query, group) candidate list (32, or 24 where the 16 KB that saves buy another stage of the key ring).
// Column-blocked tiles store the last, narrower column block as 4 x ncl queries (map_a3: box {64, ncl, 4}) so that the
// padding rows of the 128-row tile come last; the 64-wide K chunks of A then lie (96 + 4 ncl) * 128 bytes apart instead of
// 16 KB, and the MMA's rows past that (the next chunk's first rows) only feed TMEM lanes nobody reads.
template <bool DYN, int CAP>
__global__ void __launch_bounds__(TC_THREADS, 1)
ff_tc_persist_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a3,
                     const __grid_constant__ CUtensorMap map_b, TcGeom G,
                     uint32_t *__restrict__ cand, uint32_t *__restrict__ cand_meta, unsigned int *__restrict__ next_item) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    const uint32_t b_stage_bytes = (uint32_t)G.NT * 128u;
    const uint32_t a_in_stage = G.a_resident ? 0u : 16384u;       // D > 384: the A chunk rides in front of every B chunk
    const uint32_t stage_bytes = b_stage_bytes + a_in_stage;
    const uint32_t a_chunk = (uint32_t)G.a_chunk_bytes;
    uint8_t *sB = sA + (G.a_resident ? (size_t)G.NKC * a_chunk : 0);
    uint32_t *sList = reinterpret_cast<uint32_t *>(sB + (size_t)G.nstages * stage_bytes);
    PsCtl *ctl = reinterpret_cast<PsCtl *>(sList + TC_GROUPS * CAP * 128);
    // roles: warps 0-15 = epilogue groups (TMEM lane quadrant = warp & 3), 16 = TMA producer, 17 = MMA issuer,
    // 18 = TMEM allocator, 19 = threshold init
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int W_PROD = 16, W_MMA = 17, W_ALLOC = 18, W_INIT = 19;
    const int64_t total = G.total_tiles;          // work items = query tiles
    // Every key tile is scanned by ALL four epilogue groups (key rows dealt round-robin), not by the two groups that own
    // its TMEM buffer: a buffer is released after a quarter of the per-warp work instead of half, and 16 instead of 8
    // warps hide each other's latencies while the MMA warp fills the other buffer.  flags & 128: old mapping.
    const bool share_tiles = !(G.flags & 128);

    if (warp == W_PROD && lane == 0) {
        ptx::prefetch_tensormap(&map_a);
        ptx::prefetch_tensormap(&map_a3);
        ptx::prefetch_tensormap(&map_b);
    }
    if (warp == W_MMA && lane == 0) {
        for (int s = 0; s < G.nstages; ++s) { ptx::mbar_init(&ctl->full[s], 1); ptx::mbar_init(&ctl->empty[s], 1); }
        ptx::mbar_init(&ctl->a_full, 1);
        ptx::mbar_init(&ctl->a_free, 1);
        for (int i = 0; i < 4; ++i) ptx::mbar_init(&ctl->item_bar[i], 1);
        for (int b = 0; b < G.nbuf; ++b) { ptx::mbar_init(&ctl->tmem_full[b], 1); ptx::mbar_init(&ctl->tmem_empty[b], share_tiles ? 16 : (G.nbuf == 4 ? 4 : 8)); }
        ptx::fence_barrier_init();
    }
    if (warp == W_ALLOC) ptx::tmem_alloc<512>(&ctl->tmem_base);
    if (warp == W_INIT) for (int i = lane; i < 256; i += 32) (&ctl->thr_sh[0][0])[i] = thr_enc(-INFINITY);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    // the CTA owns all 512 TMEM columns, so the allocation starts at column 0 / lane 0; using the literal keeps the
    // accumulator addresses warp-uniform for ptxas
    if (ctl->tmem_base != 0u) __trap();
    constexpr uint32_t tmem_base = 0u;

    if (warp == W_PROD) {
        // =========================== TMA producer (32 lanes, one elected per issue) ===========================
        uint32_t stage = 0, phase = 0;
        for (uint32_t k = 0;; ++k) {
            int64_t id;
            if (DYN) {
                // dynamic work distribution: heavy-first ids from a global counter, broadcast through a 4-entry ring (the
                // producer is never more than 2 items ahead of the slowest role, see a_free / tmem_empty)
                unsigned int raw = 0;
                if (lane == 0) {
                    raw = atomicAdd(next_item, 1u);
                    ctl->item_id[k & 3u] = ((int64_t)raw < total) ? (int32_t)raw : -1;
                    ptx::mbar_arrive(&ctl->item_bar[k & 3u]);
                }
                raw = __shfl_sync(0xffffffffu, raw, 0);
                id = ((int64_t)raw < total) ? (int64_t)raw : -1;
            } else {
                id = item_at(k, total);
            }
            if (id < 0) break;
            const Item I = item_geom(G, id);
            if (G.a_resident) {
                ptx::mbar_wait(&ctl->a_free, (k & 1u) ^ 1u);          // every MMA of the previous item has read A
                if (ptx::elect_one()) {
                    ptx::mbar_expect_tx(&ctl->a_full, G.colblk ? (uint32_t)G.NKC * (3u * 4096u + (uint32_t)G.ncl * 512u) : (uint32_t)G.NKC * 16384u);
                    if (G.colblk) {
                        // column-blocked query tile: quadrant w (tile rows 32w .. 32w+31 = the TMEM lanes of epilogue warps
                        // 4g+w) holds the 4 x 8 block of queries (grid rows qr0 .. qr0+3, columns 8w .. 8w+7); one 3-D box
                        // {64 elements, 8 patch columns, 4 grid rows} = 4 KB per quadrant and 64-wide K chunk
                        const int grow0 = (I.clip * G.n_frames + I.t) * G.H + I.qr0;
                        for (int kc = 0; kc < G.NKC; ++kc) {
                            for (int w = 0; w < 3; ++w)
                                ptx::tma_load_3d(sA + kc * a_chunk + w * 4096, &map_a, kc * 64, 8 * w, grow0, &ctl->a_full);
                            ptx::tma_load_3d(sA + kc * a_chunk + 3 * 4096, &map_a3, kc * 64, 24, grow0, &ctl->a_full);   // 4 x ncl queries
                        }
                    } else {
                        for (int kc = 0; kc < G.NKC; ++kc) ptx::tma_load_2d(sA + kc * 16384, &map_a, kc * 64, I.q_row0, &ctl->a_full);
                    }
                }
            }
            for (int ci = 0; ci < I.nctx; ++ci) {
                // context frames are walked NEWEST FIRST (slot nctx-1 = frame t-1 ... slot 0 = first frame): the best matches
                // sit in the nearest frames, so the nomination threshold is close to final after the first key tiles and
                // the later ones append (and compact) almost nothing.  flags & 8 restores oldest-first for comparison.
                const int f = ctx_frame(I.t, G.n_last, (G.flags & 8) ? ci : I.nctx - 1 - ci);
                for (int ch = 0; ch < I.nchunks; ++ch) {
                    const int k_row0 = (int)(I.clip_row0 + (int64_t)f * G.N + (I.kr_lo + ch * G.RPC) * G.W);
                    for (int kc = 0; kc < G.NKC; ++kc) {
                        ptx::mbar_wait(&ctl->empty[stage], phase ^ 1u);
                        if (ptx::elect_one()) {
                            ptx::mbar_expect_tx(&ctl->full[stage], stage_bytes);
                            uint8_t *stp = sB + (size_t)stage * stage_bytes;
                            if (!G.a_resident) ptx::tma_load_2d(stp, &map_a, kc * 64, I.q_row0, &ctl->full[stage]);
                            ptx::tma_load_2d(stp + a_in_stage, &map_b, kc * 64, k_row0, &ctl->full[stage]);
                        }
                        if (++stage == (uint32_t)G.nstages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == W_MMA) {
        // =========================== MMA issuer (32 lanes, one elected per issue) ===========================
        const uint64_t da0 = ptx::umma_desc_sw128(ptx::smem_u32(sA)), db0 = ptx::umma_desc_sw128(ptx::smem_u32(sB));
        const uint32_t stage_step = stage_bytes >> 4, a_step = a_in_stage >> 4;
        uint32_t stage = 0, phase = 0, buf = 0, use = 0;
        for (uint32_t k = 0;; ++k) {
            int64_t id;
            if (DYN) {
                ptx::mbar_wait(&ctl->item_bar[k & 3u], (k >> 2) & 1u);
                id = ctl->item_id[k & 3u];
            } else {
                id = item_at(k, total);
            }
            if (id < 0) break;
            const Item I = item_geom(G, id);
            if (G.a_resident) ptx::mbar_wait(&ctl->a_full, k & 1u);
            ptx::tc_fence_after();
            int ch = 0;
            for (int tile = 0; tile < I.ntiles; ++tile) {
                const int rc = min(G.RPC, I.kr_hi + 1 - (I.kr_lo + ch * G.RPC));
                const int n_mma = min(G.NT, (rc + G.qrows - 1) / G.qrows * G.qrows * G.W);
                const uint32_t idesc = ptx::umma_idesc_f16(128, n_mma);
                ptx::mbar_wait(&ctl->tmem_empty[buf], (use & 1u) ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * (uint32_t)G.buf_cols;
                uint64_t da = da0;
                for (int kc = 0; kc < G.NKC; ++kc, da += a_chunk >> 4) {
                    ptx::mbar_wait(&ctl->full[stage], phase);
                    ptx::tc_fence_after();
                    const uint64_t ds = db0 + (uint64_t)(stage * stage_step);
                    const uint64_t db = ds + a_step;
                    const uint64_t dq = G.a_resident ? da : ds;
                    if (ptx::elect_one()) {
                        ptx::umma_f16(d_tmem, dq, db, idesc, kc != 0);
                        ptx::umma_f16(d_tmem, dq + 2, db + 2, idesc, true);
                        ptx::umma_f16(d_tmem, dq + 4, db + 4, idesc, true);
                        ptx::umma_f16(d_tmem, dq + 6, db + 6, idesc, true);
                        ptx::umma_commit(&ctl->empty[stage]);
                    }
                    if (++stage == (uint32_t)G.nstages) { stage = 0; phase ^= 1u; }
                }
                if (ptx::elect_one()) ptx::umma_commit(&ctl->tmem_full[buf]);
                if (++buf == (uint32_t)G.nbuf) { buf = 0; ++use; }
                if (++ch == I.nchunks) ch = 0;
            }
            if (G.a_resident && ptx::elect_one()) ptx::umma_commit(&ctl->a_free);   // A may be overwritten once these MMAs retired
        }
    } else if (warp < 16) {
        // =========================== epilogue groups ===========================
        const int g = warp >> 2;
        const int mybuf = (G.nbuf == 4) ? g : (g & 1);
        const int row_par = (G.nbuf == 4) ? -1 : (g >> 1);
        const int qi = ((warp & 3) << 5) + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) << 5) << 16;
        const uint32_t list0 = ptx::smem_u32(sList + qi);
        const uint32_t list = list0 + (uint32_t)g * CAP * 128u * 4u;
        constexpr int half = 16;                                       // entries a compaction may leave
        const uint32_t slot_lim = list + (uint32_t)(CAP - 8) * TC_SLOT_STRIDE;      // compaction trigger: more than CAP - 8 entries
        // column-blocked tiles: my position inside the tile's 4 x 32 grid positions (the last column block holds 4 x ncl queries)
        const int cb_w = warp & 3;
        const int cb_n = (cb_w == 3) ? G.ncl : 8;
        const int cb_lr = lane / cb_n, cb_lc = cb_w * 8 + lane % cb_n;
        const bool cb_ok = lane < 4 * cb_n;
        uint32_t gtile = 0;                                            // tiles issued before this item (all roles agree)
        for (uint32_t k = 0;; ++k) {
            int64_t id;
            if (DYN) {
                if (lane == 0) ptx::mbar_wait(&ctl->item_bar[k & 3u], (k >> 2) & 1u);
                __syncwarp();
                // every lane reads the same word; passing it through a warp collective tells ptxas so (the item geometry and all
                // loop bounds derived from it stay in uniform registers, no reconvergence code around the warp collectives)
                id = (int64_t)__reduce_max_sync(0xffffffffu, (unsigned)(ctl->item_id[k & 3u] + 1)) - 1;
            } else {
                id = item_at(k, total);
            }
            if (id < 0) break;
            const Item I = item_geom(G, id);
            uint32_t *thr_cur = ctl->thr_sh[k & 1u];
            // query of this thread (= TMEM lane qi): raster order inside the tile, or -- column-blocked tile -- position
            // (lane / 8, 8 * quadrant + lane % 8) of the tile's 4 x 32 grid positions; the last quadrant: (lane / ncl, 24 + lane % ncl)
            const int qrow = G.colblk ? I.qr0 + cb_lr : I.qr0 + qi / G.W;
            const int qcol = G.colblk ? cb_lc : qi % G.W;
            const bool valid = G.colblk ? (cb_ok && qrow <= I.qr1 && qcol < G.W) : (qi < I.nq);
            const int r_lo = qrow - G.radius, r_hi = qrow + G.radius;
            const int c_lo = qcol - G.radius;
            const int c_lo_cl = max(c_lo, 0), c_hi_cl = min(qcol + G.radius, G.W - 1);
            // Key columns / key rows any query of this warp can see (warp-uniform): only these are read from TMEM.  Raster tiles:
            // (nearly) the whole key row; column-blocked tiles: 8 + 2 * radius columns at most (20 of 28 at radius 6, 14 / 10
            // at the borders).  The column range is walked in chunks of 16 / 8 / 4 TMEM columns; its width is rounded up to a
            // multiple of 4 and, where that would run past the key row, moved left (the extra columns are outside every
            // lane's window).  M = this lane's window as a bit mask over the key columns (W <= 64).
            const int wc_lo = __reduce_min_sync(0xffffffffu, valid ? c_lo_cl : 0x7fffffff);
            const int wc_hi = __reduce_max_sync(0xffffffffu, valid ? c_hi_cl : -1);
            const int wr_lo = __reduce_min_sync(0xffffffffu, valid ? max(r_lo, I.kr_lo) : 0x7fffffff);
            const int wr_hi = __reduce_max_sync(0xffffffffu, valid ? min(r_hi, I.kr_hi) : -1);
            const int width = (wc_hi >= wc_lo) ? ((wc_hi - wc_lo + 4) & ~3) : 0;
            const int cstart = max(0, min(wc_lo, G.W - width));
            const unsigned long long M = valid ? (((2ull << c_hi_cl) - 1ull) & ~((1ull << c_lo_cl) - 1ull)) : 0ull;
            float thr = (G.flags & 2) ? INFINITY : -INFINITY;      // debug flags (TIMET_TC_PFLAGS): 1 no scan, 2 scan without appends, 4 TMEM loads only
            int cnt = 0, lost = 0;
            uint32_t slot = list;                                  // next free slot of my list (cnt is derived from it when needed)
            uint32_t slot_cmp = list;                              // ... right after the last compaction
            const uint32_t OPP_MIN = ((uint32_t)G.flags >> 20) ? ((uint32_t)G.flags >> 20) : 6u;   // experiments: TIMET_TC_PFLAGS bits 20+

            // first tile of this item that lands in my buffer: (gtile + j) % nbuf == mybuf
            const uint32_t nbuf_mask = (uint32_t)G.nbuf - 1u, nbuf_sh = (G.nbuf == 4) ? 2u : 1u;
            int j = share_tiles ? 0 : (int)(((uint32_t)mybuf - gtile) & nbuf_mask);
            const int jstep = share_tiles ? 1 : G.nbuf;
            int ci = 0, ch = j;
            while (ch >= I.nchunks) { ch -= I.nchunks; ++ci; }
            for (; j < I.ntiles; j += jstep) {
                const int kr_start = I.kr_lo + ch * G.RPC;
                const int rc = min(G.RPC, I.kr_hi + 1 - kr_start);
                const uint32_t tile_no = gtile + (uint32_t)j;                      // running tile count of this CTA
                const int buf = (int)(tile_no & nbuf_mask);
                const uint32_t use = tile_no >> nbuf_sh;
                const uint32_t t_acc = tmem_base + (uint32_t)(buf * G.buf_cols) + lane_base;
                // The key tile is usually not there yet (the MMA side is the slower one): use the wait to compact the list
                // and raise the threshold, so that the capacity check inside the scan -- which holds the TMEM buffer while it
                // compacts -- rarely fires.  Only when some lane appended OPP_MIN entries since its last compaction.
                if (!(G.flags & 8192)) {
                    const bool there = __any_sync(0xffffffffu, lane == 0 && ptx::mbar_try_wait(&ctl->tmem_full[buf], use & 1u));
                    if (!there && __any_sync(0xffffffffu, slot >= slot_cmp + OPP_MIN * TC_SLOT_STRIDE)) {
                        const float before = thr;
                        cnt = (int)((slot - list) / TC_SLOT_STRIDE);
                        tc_compact(list, cnt, thr, lost, G.topk, half);
                        slot = slot_cmp = list + (uint32_t)cnt * TC_SLOT_STRIDE;
                        if (thr > before) atomicMax(&thr_cur[qi], thr_enc(thr));
                    }
                }
                if (lane == 0) {
                    if (G.flags & 32) ptx::mbar_wait_sleep(&ctl->tmem_full[buf], use & 1u, 64);
                    else if (G.flags & 512) ptx::mbar_wait_hint(&ctl->tmem_full[buf], use & 1u, 4000);
                    else ptx::mbar_wait(&ctl->tmem_full[buf], use & 1u);
                }
                __syncwarp();
                ptx::tc_fence_after();
                thr = fmaxf(thr, thr_dec(thr_cur[qi]));
                // my key rows of this tile: rows [rlo_t, rhi_t] some lane can see, dealt round-robin to the four groups
                // (or, with the buffer-owning mapping, to the two groups of the buffer)
                // (a warp whose lanes are all padding has the empty range wr_lo > wr_hi: no rows)
                const bool scan = wr_hi >= wr_lo && !(G.flags & 1);
                const int rlo_t = scan ? max(0, wr_lo - kr_start) : 0, rhi_t = scan ? min(rc - 1, wr_hi - kr_start) : -1;
                int rr, rstep, rsh;
                if (share_tiles) { rstep = 4; rsh = 2; rr = rlo_t + ((g - rlo_t) & 3); }
                else if (row_par >= 0) { rstep = 2; rsh = 1; rr = rlo_t + ((row_par - rlo_t) & 1); }
                else { rstep = 1; rsh = 0; rr = rlo_t; }
                // Rows nearest to the query tile first: the best matches sit there, so the threshold is close to final after
                // the first row and the remaining rows of the tile append little.  A key chunk above the tile's centre is
                // walked from its last row upwards.  (flags & 32768: always top-down.)
                int n_rows = (rr <= rhi_t) ? ((rhi_t - rr) >> rsh) + 1 : 0;
                if (2 * kr_start + rc <= 2 * I.qr0 + (I.qr1 - I.qr0 + 1) && !(G.flags & 32768)) { rr += (n_rows - 1) * rstep; rstep = -rstep; }
                const int ctx_code = ((G.flags & 8) ? ci : I.nctx - 1 - ci) << 10;
                for (; n_rows > 0; --n_rows, rr += rstep) {
                    const int kr = kr_start + rr;
                    const unsigned long long Mr = (kr >= r_lo && kr <= r_hi) ? M : 0ull;
                    // pick up the other groups' progress (one LDS per key row): the four groups of a query raise one threshold
                    if (!(G.flags & 64)) thr = fmaxf(thr, thr_dec(thr_cur[qi]));
                    const int code_row = ctx_code | ((kr - r_lo) << 5);
                    const uint32_t t_row = t_acc + (uint32_t)(rr * G.W);
                    // runs of at most 8 offers, each followed by a capacity check: a list holds CAP entries and at most
                    // CAP - 8 when a run starts (a compaction leaves <= 16 <= CAP - 8), so a run can never overflow it
#define TC_CHECK()                                                                       \
    if (__any_sync(0xffffffffu, slot > slot_lim)) {                                      \
        if (G.flags & 2048) { slot = list; continue; }                                   \
        const float before = thr;                                                        \
        cnt = (int)((slot - list) / TC_SLOT_STRIDE);                                     \
        tc_compact(list, cnt, thr, lost, G.topk, half);                                  \
        slot = slot_cmp = list + (uint32_t)cnt * TC_SLOT_STRIDE;                         \
        if (thr > before) atomicMax(&thr_cur[qi], thr_enc(thr));                         \
    }
                    for (int c = cstart, rem = width; rem > 0;) {
                        const uint32_t wmask = (uint32_t)(Mr >> c);           // bit b: key column c + b is inside my window
                        const uint32_t code0 = (uint32_t)(code_row + (c - c_lo));
                        const uint32_t taddr = t_row + (uint32_t)c;
                        if (rem >= 16) {
                            uint32_t r[16];
                            ptx::tmem_ld_32x16(taddr, r);
                            c += 16; rem -= 16;
                            ptx::tmem_ld_wait();
                            if (G.flags & 4) { asm volatile("" ::"r"(r[0]), "r"(r[5]), "r"(r[10]), "r"(r[15])); continue; }
                            tc_offer4<0>(slot, r[0], r[1], r[2], r[3], thr, wmask, code0);
                            tc_offer4<4>(slot, r[4], r[5], r[6], r[7], thr, wmask, code0);
                            TC_CHECK()
                            tc_offer4<8>(slot, r[8], r[9], r[10], r[11], thr, wmask, code0);
                            tc_offer4<12>(slot, r[12], r[13], r[14], r[15], thr, wmask, code0);
                            TC_CHECK()
                        } else if (rem >= 8) {
                            uint32_t r[8];
                            ptx::tmem_ld_32x8(taddr, r);
                            c += 8; rem -= 8;
                            ptx::tmem_ld_wait();
                            if (G.flags & 4) { asm volatile("" ::"r"(r[0]), "r"(r[3]), "r"(r[5]), "r"(r[7])); continue; }
                            tc_offer4<0>(slot, r[0], r[1], r[2], r[3], thr, wmask, code0);
                            tc_offer4<4>(slot, r[4], r[5], r[6], r[7], thr, wmask, code0);
                            TC_CHECK()
                        } else {
                            uint32_t r[4];
                            ptx::tmem_ld_32x4(taddr, r);
                            c += 4; rem -= 4;
                            ptx::tmem_ld_wait();
                            if (G.flags & 4) { asm volatile("" ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])); continue; }
                            tc_offer4<0>(slot, r[0], r[1], r[2], r[3], thr, wmask, code0);
                            TC_CHECK()
                        }
                    }
#undef TC_CHECK
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&ctl->tmem_empty[buf]);
                ch += jstep;
                while (ch >= I.nchunks) { ch -= I.nchunks; ++ci; }
            }
            cnt = (int)((slot - list) / TC_SLOT_STRIDE);
            gtile += (uint32_t)I.ntiles;

            // ---- final phase of the item (the MMA warp is already working on the next one)
            if (G.flags & 1024) {                                      // attribution: no merge, nothing published
                if (g == 1) ctl->thr_sh[(k + 1u) & 1u][qi] = thr_enc(-INFINITY);
                asm volatile("bar.sync 1, 512;" ::: "memory");
                continue;
            }
            {
                const float before = thr;
                tc_compact(list, cnt, thr, lost, G.topk, half);
                if (thr > before) atomicMax(&thr_cur[qi], thr_enc(thr));
            }
            if (g == 1) ctl->thr_sh[(k + 1u) & 1u][qi] = thr_enc(-INFINITY);   // idle copy, used by the next item
            asm volatile("bar.sync 1, 512;" ::: "memory");            // (A) every group's final threshold is published
            thr = fmaxf(thr, thr_dec(thr_cur[qi]));
            tc_filter(list, cnt, thr);
            ctl->xchg[g][qi] = (uint32_t)cnt | (lost ? 0x10000u : 0u);
            asm volatile("bar.sync 1, 512;" ::: "memory");            // (B) counts known -> push offsets
            int off = 0, tot = 0, any_lost = 0;
#pragma unroll
            for (int og = 0; og < TC_GROUPS; ++og) {
                const uint32_t x = ctl->xchg[og][qi];
                if (og < g) off += (int)(x & 0xFFFFu);
                tot += (int)(x & 0xFFFFu);
                any_lost |= (int)(x >> 16);
            }
            if (g > 0) {
                for (int s = 0; s < cnt; ++s)
                    if (off + s < CAP) sts_u32(list0 + (uint32_t)(off + s) * TC_SLOT_STRIDE, lds_u32(list + s * TC_SLOT_STRIDE));
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");            // (C) list 0 holds the merged candidates
            if (g == 0) {
                lost = any_lost | (tot > CAP ? 1 : 0);
                cnt = min(tot, CAP);
                tc_compact(list0, cnt, thr, lost, G.topk, FF_CAND_STORE);
                if (valid) {
                    const int64_t q = ((int64_t)I.clip * G.nT + (I.t - G.t_begin)) * G.N + qrow * G.W + qcol;
                    uint32_t *dst = cand + q * FF_CAND_STORE;
#pragma unroll
                    for (int s4 = 0; s4 < FF_CAND_STORE; s4 += 4) {
                        if (s4 < cnt) {
                            uint4 v;
                            v.x = lds_u32(list0 + (s4 + 0) * TC_SLOT_STRIDE); v.y = lds_u32(list0 + (s4 + 1) * TC_SLOT_STRIDE);
                            v.z = lds_u32(list0 + (s4 + 2) * TC_SLOT_STRIDE); v.w = lds_u32(list0 + (s4 + 3) * TC_SLOT_STRIDE);
                            *reinterpret_cast<uint4 *>(dst + s4) = v;
                        }
                    }
                    cand_meta[q] = (uint32_t)cnt | (lost ? 0x10000u : 0u);
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == W_ALLOC) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<512>(tmem_base);
    }
}

static size_t persist_smem_bytes(const TcGeom &G, int cap) {
    const size_t a_res = G.a_resident ? (size_t)G.NKC * G.a_chunk_bytes : 0, a_stage = G.a_resident ? 0 : 16384;
    return 1024 + a_res + (size_t)G.nstages * ((size_t)G.NT * 128 + a_stage) + (size_t)TC_GROUPS * cap * 128 * 4 + sizeof(PsCtl) + 64;
}
static int persist_fit_stages(TcGeom &G, int cap, int want) {
    G.nstages = want;
    while (persist_smem_bytes(G, cap) > 227 * 1024 && G.nstages > 2) G.nstages--;
    return persist_smem_bytes(G, cap) <= 227 * 1024 ? G.nstages : 0;
}

// Configuration of the persistent kernel for a problem (host only, no CUDA call): query-tile layout, candidate-list
// capacity, key-ring depth.  TIMET_ERR_UNSUPPORTED if the shape does not qualify (the per-item kernel takes it).
static int persist_configure(const timet_ff_params &p, const FFLayout &L, TcGeom &G, int *cap_out) {
    if (!tc_geometry(p, L, &G)) return TIMET_ERR_UNSUPPORTED;
    // the epilogue keeps a lane's window as a 64-bit column mask and may read up to 3 TMEM columns past a key row whose
    // width is not a multiple of 4: wider grids / exactly filled buffers go to the per-item kernel
    if (G.W > 64 || ((G.W & 3) && G.NT + 3 > G.buf_cols)) return TIMET_ERR_UNSUPPORTED;
    const EnvCfg &E = env_cfg();
    G.flags = E.tc_pflags;
    // column-blocked query tiles (see the kernel): grids 26..32 patches wide whose tiles are 4 grid rows, resident A.
    // TIMET_TC_PFLAGS & 256 keeps the raster tile.
    G.colblk = (G.a_resident && G.W >= 26 && G.W <= 32 && G.QR == 4 && !(G.flags & 256)) ? 1 : 0;
    G.ncl = G.colblk ? G.W - 24 : 8;
    // TIMET_TC_PFLAGS & 16384: padding rows kept inside every K chunk of A (16 KB apart) and 32-slot lists
    const bool compact_a = G.colblk && !(G.flags & 16384);
    G.a_chunk_bytes = compact_a ? (96 + 4 * G.ncl + 7) / 8 * 8 * 128 : 16384;
    // Ring depth vs candidate-list capacity: the key ring is latency-bound (BASELINE configs[1]: two 28 KB stages beside
    // A and 64 KB of lists), so 24-slot lists are used where the 16 KB they free buy a deeper ring (up to 4 stages).
    const int want = (E.tc_stages >= 2 && E.tc_stages <= TC_MAX_STAGES) ? E.tc_stages : TC_MAX_STAGES;
    const int st32 = persist_fit_stages(G, 32, want), st24 = persist_fit_stages(G, 24, want);
    if (!st32 && !st24) return TIMET_ERR_UNSUPPORTED;
    const int cap = (st24 > st32 && st32 < 4 && !(G.flags & 16384)) ? 24 : 32;
    G.nstages = cap == 24 ? st24 : st32;
    if (!G.nstages) return TIMET_ERR_UNSUPPORTED;
    *cap_out = cap;
    return TIMET_OK;
}

// timet_ff_tc_plan: what ff_select will run for this problem
int ff_tc_plan(const timet_ff_params &p, int32_t *out) {
    const FFLayout L = ff_layout(p);
    for (int i = 0; i < 8; ++i) out[i] = 0;
    TcGeom G;
    if (!tc_geometry(p, L, &G)) return TIMET_OK;               // exact engine
    const EnvCfg &E = env_cfg();
    int cap = 32;
    const bool debug = E.tc_trace || E.tc_flags != 0;
    if (E.tc_persist && !debug && persist_configure(p, L, G, &cap) == TIMET_OK) {
        out[0] = 2; out[1] = G.colblk; out[2] = cap; out[3] = G.nstages; out[4] = G.a_resident ? G.NKC * G.a_chunk_bytes : 0;
        out[5] = G.NT; out[6] = G.QR; out[7] = (int32_t)persist_smem_bytes(G, cap);
        return TIMET_OK;
    }
    if (!tc_geometry(p, L, &G)) return TIMET_OK;
    out[0] = 1; out[2] = TC_CAP; out[3] = G.nstages; out[4] = G.a_resident ? G.NKC * 16384 : 0; out[5] = G.NT; out[6] = G.QR;
    out[7] = (int32_t)tc_smem_bytes(G);
    return TIMET_OK;
}

// Persistent launch; TIMET_ERR_UNSUPPORTED if the shape does not qualify (caller falls back to the per-tile kernel)
int ff_select_tc_persist_launch(const timet_ff_params &p, const FFLayout &L, char *ws, cudaStream_t st) {
    TcGeom G;
    int cap = 32;
    {
        const int rc0 = persist_configure(p, L, G, &cap);
        if (rc0 != TIMET_OK) return rc0;
    }
    const EnvCfg &E = env_cfg();
    const __half *fn16 = reinterpret_cast<const __half *>(ws + L.off_fn16);
    CUtensorMap map_a, map_a3, map_b;
    int rc;
    if (G.colblk) {
        if ((rc = tc_make_map_colblk(&map_a, fn16, (int64_t)p.n_clips * p.n_frames * G.H, L.Dp, G.W, 8)) != TIMET_OK) return rc;
        if ((rc = tc_make_map_colblk(&map_a3, fn16, (int64_t)p.n_clips * p.n_frames * G.H, L.Dp, G.W, G.ncl)) != TIMET_OK) return rc;
    } else {
        if ((rc = tc_make_map(&map_a, fn16, L.rows + 256, L.Dp, 128)) != TIMET_OK) return rc;
        map_a3 = map_a;
    }
    if ((rc = tc_make_map(&map_b, fn16, L.rows + 256, L.Dp, G.NT)) != TIMET_OK) return rc;
    const size_t smem = persist_smem_bytes(G, cap);
    auto kern = cap == 24 ? (E.tc_dyn ? ff_tc_persist_kernel<true, 24> : ff_tc_persist_kernel<false, 24>)
                          : (E.tc_dyn ? ff_tc_persist_kernel<true, 32> : ff_tc_persist_kernel<false, 32>);
    TIMET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = G.total_tiles < num_sms() ? G.total_tiles : num_sms();
    unsigned int *next_item = reinterpret_cast<unsigned int *>(ws + L.off_redo + FF_HDR_NEXT_ITEM);   // zeroed with the redo header by timet_ff_select
    kern<<<(unsigned)grid, TC_THREADS, smem, st>>>(map_a, map_a3, map_b, G, reinterpret_cast<uint32_t *>(ws + L.off_cand),
                                                                  reinterpret_cast<uint32_t *>(ws + L.off_cand_meta), next_item);
    TIMET_LAUNCHED();
    return TIMET_OK;
}

}  // namespace timet
