// Per-query selection shared by both FF engines: a sorted top list distributed over one warp
// (lane l = slot l), canonical order (affinity desc, key asc), ties of the k-th value kept
// (/root/reference/mask_propagation.py:432-436: aff[aff < kth] = 0; aff /= aff.sum()).
#pragma once
#include "common.cuh"

namespace timet {

struct TopList {
    float v;       // affinity held by this lane's slot
    int32_t key;   // frame * N + patch
    int cnt;       // warp-uniform: filled slots (<= 32)
    float kth;     // warp-uniform: value of slot k-1 once cnt >= k
    int dropped;   // warp-uniform: entries pushed past slot 31
};

__device__ __forceinline__ void list_init(TopList &L) {
    L.v = -1.f; L.key = 0x7fffffff; L.cnt = 0; L.kth = -1.f; L.dropped = 0;
}

// Every lane offers (aff, key) if `valid`.  Warp-synchronous.
__device__ __forceinline__ void list_offer(TopList &L, bool valid, float aff, int32_t key, int k, int lane) {
    unsigned mask = __ballot_sync(0xffffffffu, valid && (L.cnt < k || aff >= L.kth));
    while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        const float a = __shfl_sync(0xffffffffu, aff, b);
        const int32_t id = __shfl_sync(0xffffffffu, key, b);
        if (L.cnt >= k && a < L.kth) continue;   // the k-th value rose meanwhile
        const unsigned before = __ballot_sync(0xffffffffu, lane < L.cnt && (L.v > a || (L.v == a && L.key < id)));
        const int pos = __popc(before);
        if (pos >= 32) { L.dropped++; continue; }
        const float uv = __shfl_up_sync(0xffffffffu, L.v, 1);
        const int32_t uk = __shfl_up_sync(0xffffffffu, L.key, 1);
        if (lane > pos) { L.v = uv; L.key = uk; }
        else if (lane == pos) { L.v = a; L.key = id; }
        if (L.cnt < 32) L.cnt++; else L.dropped++;
        if (L.cnt >= k) L.kth = __shfl_sync(0xffffffffu, L.v, k - 1);
    }
}

// Number of list entries that survive the reference's selection (everything >= the k-th value).  Warp-uniform.
__device__ __forceinline__ int list_kept(const TopList &L, int k, int lane) {
    const bool keep = lane < L.cnt && (L.cnt < k || L.v >= L.kth);
    return __popc(__ballot_sync(0xffffffffu, keep));
}

// Normalise the survivors and write the sparse row.  Returns the number of kept entries
// (before truncation to kw).
__device__ __forceinline__ int list_finish(const TopList &L, int k, int kw, int lane, float *__restrict__ w_out,
                                           int32_t *__restrict__ k_out, int32_t *__restrict__ cnt_out) {
    const bool keep = lane < L.cnt && (L.cnt < k || L.v >= L.kth);
    const int m = __popc(__ballot_sync(0xffffffffu, keep));
    const bool wr = keep && lane < kw;
    const float sum = warp_sum(wr ? L.v : 0.f);
    if (lane < kw) {
        w_out[lane] = wr ? __fdiv_rn(L.v, sum) : 0.f;
        k_out[lane] = wr ? L.key : -1;
    }
    if (lane == 0) *cnt_out = m < kw ? m : kw;
    return m;
}

// CTA-level aggregation of the diagnostics counters (one atomic per counter per CTA)
struct StatAcc {
    unsigned long long selected, ties, truncated, redone;
};

}  // namespace timet
