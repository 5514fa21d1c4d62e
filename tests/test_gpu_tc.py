"""Tensor-core engine (tcgen05/TMEM/TMA): raw accumulators against a matmul of the fp16 operands, the
error bound the nomination relies on, and bit-exact agreement with the exact engine."""
import ctypes as C
import math
import os

import numpy as np
import pytest
import torch

import timetuning_b200 as tb
from timetuning_b200 import _cabi, synth
from timetuning_b200.ops import _ptr, _stream

pytestmark = pytest.mark.gpu
FF_TC_DELTA = 1.05e-3


def geometry(H, W, radius):
    q = 16 // math.gcd(W, 16)
    RPC = (256 // W) // q * q          # 2 TMEM buffers of 256 columns (default) ...
    if os.environ.get("TIMET_TC_NBUF") == "4" and (128 // W) // q >= 1:
        RPC = (128 // W) // q * q      # ... or 4 x 128
    QR = min(128 // W, H)
    return dict(QR=QR, tpf=-(-H // QR), RPC=RPC, NT=RPC * W, qrows=q)


def ctx_frames(t, n_last):
    return [0] + list(range(max(1, t - n_last), t))


@pytest.mark.parametrize("sr,D,fs,radius,tile_ids", [
    (28, 384, 3, 6, [0, 3, 6, 7, 13]),
    (14, 64, 3, 6, [0, 1, 2]),
    (60, 384, 3, 12, [0, 15, 29, 59]),
    (20, 96, 4, 4, [0, 3, 11]),
    (28, 768, 3, 6, [0, 5, 13]),          # ViT-B width: streamed query tile
])
def test_tc_raw_accumulators(sr, D, fs, radius, tile_ids):
    n_clips, C_, n_last, topk = 1, 8, 7, 5
    N = sr * sr
    feats = torch.from_numpy(synth.clip_features(n_clips, fs, sr, D, seed=3)).cuda()
    plan = tb.FFPlan(n_clips, fs, sr, sr, D, C_, n_last, radius, topk)
    if not plan.tc_supported:
        pytest.skip("shape not supported by the tensor-core engine")
    plan.prepare(feats)
    g = geometry(sr, sr, radius)
    f32 = feats[0].double()
    fn32 = (feats[0] / feats[0].norm(dim=-1, keepdim=True).clamp_min(1e-12))
    fn16 = fn32.half().double()                                     # what the GEMM consumes
    flat16 = torch.cat([fn16.reshape(fs * N, D), torch.zeros(512, D, dtype=torch.float64, device="cuda")])
    flat32 = torch.cat([fn32.double().reshape(fs * N, D), torch.zeros(512, D, dtype=torch.float64, device="cuda")])
    del f32
    for tile_id in tile_ids:
        tdesc, rem = divmod(tile_id, n_clips * g["tpf"])
        t = fs - 1 - tdesc
        qt = rem % g["tpf"]
        qr0 = qt * g["QR"]
        qr1 = min(sr - 1, qr0 + g["QR"] - 1)
        nq = (qr1 - qr0 + 1) * sr
        kr_lo, kr_hi = max(0, qr0 - radius), min(sr - 1, qr1 + radius)
        nchunks = (kr_hi - kr_lo + g["RPC"]) // g["RPC"]
        ctx = ctx_frames(t, n_last)
        ntiles = len(ctx) * nchunks
        dump = torch.full((ntiles, 128, 256), float("nan"), dtype=torch.float32, device="cuda")
        _cabi.check(_cabi.lib().timet_debug_tc_tile(C.byref(plan.params), _ptr(plan.workspace), plan.nbytes, tile_id,
                                                    _ptr(dump), _stream()), "debug_tc_tile")
        torch.cuda.synchronize()
        q0 = t * N + qr0 * sr
        for ci, f in enumerate(ctx):
            for ch in range(nchunks):
                rows_left = kr_hi + 1 - (kr_lo + ch * g["RPC"])
                rc = min(g["RPC"], rows_left)
                n_mma = min(g["NT"], -(-rc // g["qrows"]) * g["qrows"] * sr)
                k0 = f * N + (kr_lo + ch * g["RPC"]) * sr
                want = flat16[q0:q0 + nq] @ flat16[k0:k0 + n_mma].T
                got = dump[ci * nchunks + ch, :nq, :n_mma].double()
                err = (got - want).abs().max().item()
                assert err < 3e-4, f"tile {tile_id} ctx {ci} chunk {ch}: accumulator error {err:.3e}"
                n_real = min(rc * sr, n_mma)
                exact = flat32[q0:q0 + nq] @ flat32[k0:k0 + n_real].T
                bound = (got[:, :n_real] - exact).abs().max().item()
                assert bound < FF_TC_DELTA, f"|sim~ - sim| = {bound:.3e} exceeds the nomination bound"


@pytest.mark.parametrize("sr,D,fs,n_last,radius,topk,bs", [
    (28, 384, 8, 7, 6, 5, 4),        # config 2 shapes
    (14, 384, 4, 7, 6, 5, 2),        # config 1 shapes
    (60, 384, 5, 7, 12, 7, 1),       # config 4 shapes
    (20, 96, 9, 3, 4, 7, 2),         # FIFO eviction
    (16, 70, 5, 7, 3, 2, 2),         # padded dim
    (12, 64, 4, 7, 12, 5, 1),        # window = whole frame
    (28, 768, 4, 7, 6, 5, 2),        # ViT-B width (streamed query tile)
    (56, 768, 3, 7, 6, 5, 1),        # config 5 grid (ViT-B/8 448^2)
])
def test_tc_engine_is_bit_identical_to_exact(sr, D, fs, n_last, radius, topk, bs):
    N = sr * sr
    feats = torch.from_numpy(synth.clip_features(bs, fs, sr, D, seed=11)).cuda()
    plan = tb.FFPlan(bs, fs, sr, sr, D, 8, n_last, radius, topk)
    if not plan.tc_supported:
        pytest.skip("shape not supported by the tensor-core engine")
    plan.prepare(feats)
    plan.select(tb.FF_EXACT)
    ref = {(c, t): [x.clone() for x in plan.selection(c, t)] for c in range(bs) for t in range(1, fs)}
    st_exact = plan.stats()
    plan.select(tb.FF_TC)
    st = plan.stats()
    for (c, t), (w, k, n) in ref.items():
        w2, k2, n2 = plan.selection(c, t)
        assert torch.equal(n, n2), (c, t)
        assert torch.equal(k, k2), (c, t)
        assert torch.equal(w, w2), (c, t)
    assert st["selected"] == st_exact["selected"]
    assert st["tc_candidates"] >= st["selected"] - st["redone_queries"] * 32
    assert st["redone_queries"] <= 0.02 * st["queries"], st


def test_tc_random_features_worst_case():
    """No temporal coherence at all (iid gaussian features): nomination must still be exact."""
    bs, fs, sr, D = 2, 5, 28, 384
    rng = np.random.default_rng(0)
    feats = torch.from_numpy(rng.standard_normal((bs, fs, sr * sr, D)).astype(np.float32)).cuda()
    plan = tb.FFPlan(bs, fs, sr, sr, D, 8, 7, 6, 5)
    plan.prepare(feats)
    plan.select(tb.FF_EXACT)
    ref = {(c, t): [x.clone() for x in plan.selection(c, t)] for c in range(bs) for t in range(1, fs)}
    plan.select(tb.FF_TC)
    for (c, t), (w, k, n) in ref.items():
        w2, k2, n2 = plan.selection(c, t)
        assert torch.equal(n, n2) and torch.equal(k, k2) and torch.equal(w, w2), (c, t)


@pytest.mark.parametrize("env", [
    {"TIMET_TC_PERSIST": "0"},                  # one CTA per work item (ff_tc.cu)
    {"TIMET_TC_DYN": "0"},                      # persistent kernel, static snake schedule
    {"TIMET_TC_NBUF": "4"},                     # four 128-column TMEM buffers
    {"TIMET_TC_PFLAGS": "128"},                 # every TMEM buffer scanned by its own two groups
    {"TIMET_TC_PFLAGS": "72"},                  # oldest-first contexts, threshold picked up per tile
    {"TIMET_TC_PFLAGS": "256"},                 # raster query tiles instead of column-blocked ones
    {"TIMET_TC_PFLAGS": "8192"},                # no compaction while waiting for a key tile
    {"TIMET_TC_PFLAGS": "8576"},                # ... with raster tiles and buffer-owning groups
    {"TIMET_FIN_STAGED": "0"},                  # exact re-evaluation from registers instead of cp.async-staged rows
    {"TIMET_TC_PFLAGS": "16384"},               # query tile with its padding rows inside every K chunk, 32-slot lists, 2-stage key ring
], ids=lambda e: "+".join(f"{k[6:]}={v}" for k, v in e.items()))
def test_tc_kernel_variants_are_bit_identical(timet_env, env):
    """Every kernel variant / schedule behind the experiment switches nominates a superset of the exact
    top-k, so after the fp32 re-evaluation all of them reproduce the exact engine bit for bit."""
    bs, fs, sr, D = 3, 8, 28, 384
    feats = torch.from_numpy(synth.clip_features(bs, fs, sr, D, seed=5)).cuda()
    plan = tb.FFPlan(bs, fs, sr, sr, D, 8, 7, 6, 5)
    plan.prepare(feats)
    plan.select(tb.FF_EXACT)
    ref = {(c, t): [x.clone() for x in plan.selection(c, t)] for c in range(bs) for t in range(1, fs)}
    timet_env(**env)
    plan.select(tb.FF_TC)
    st = plan.stats()
    for (c, t), (w, k, n) in ref.items():
        w2, k2, n2 = plan.selection(c, t)
        assert torch.equal(n, n2) and torch.equal(k, k2) and torch.equal(w, w2), (env, c, t)
    assert st["redone_queries"] <= 0.02 * st["queries"], st


@pytest.mark.parametrize("H,W,D,fs,radius,topk,bs", [
    (28, 28, 384, 4, 1, 5, 2),       # narrow window: most TMEM columns are never read
    (28, 28, 128, 4, 13, 7, 1),      # window wider than a column block's reach
    (10, 30, 192, 4, 4, 4, 2),       # non-square, last tile has 2 of 4 grid rows
    (7, 32, 64, 5, 15, 8, 2),        # full 32 columns (no padding lanes), radius at its maximum
    (9, 26, 100, 4, 2, 3, 3),        # narrowest column-blocked grid, padded dim, last tile has 1 grid row
    (5, 28, 384, 3, 6, 5, 1),        # two tiles per frame, the second with one grid row
    (4, 30, 320, 9, 3, 5, 2),        # one tile per frame, FIFO eviction (n_last 3 below)
])
def test_tc_column_blocked_query_tiles(H, W, D, fs, radius, topk, bs):
    """Grids 26..32 patches wide: the persistent kernel arranges a query tile as four 4 x 8 column blocks (one per TMEM
    lane quadrant) and every epilogue warp reads only the key columns its block can see.  Same candidates => bit-identical
    to the exact engine, for border / partial tiles and every chunk shape (16 / 8 / 4 TMEM columns)."""
    N = H * W
    rng = np.random.default_rng(H * 100 + W)
    base = rng.standard_normal((bs, 1, N, D)).astype(np.float32)
    feats = torch.from_numpy(base + 0.7 * rng.standard_normal((bs, fs, N, D)).astype(np.float32)).cuda()
    n_last = 3 if fs > 8 else 7
    plan = tb.FFPlan(bs, fs, H, W, D, 8, n_last, radius, topk)
    assert plan.tc_supported
    plan.prepare(feats)
    plan.select(tb.FF_EXACT)
    ref = {(c, t): [x.clone() for x in plan.selection(c, t)] for c in range(bs) for t in range(1, fs)}
    plan.select(tb.FF_TC)
    st = plan.stats()
    for (c, t), (w, k, n) in ref.items():
        w2, k2, n2 = plan.selection(c, t)
        assert torch.equal(n, n2) and torch.equal(k, k2) and torch.equal(w, w2), (c, t)
    assert st["redone_queries"] <= 0.05 * st["queries"], st


def test_tc_engine_wide_features_dim_2048():
    """ResNet-width features (dim 1537..2048): the finalize kernel needs more than 48 KB of dynamic shared memory;
    FF_AUTO must pick the tensor-core engine and agree with the exact engine bit for bit."""
    bs, fs, sr, D = 1, 3, 14, 2048
    feats = torch.from_numpy(synth.clip_features(bs, fs, sr, D, seed=11)).cuda()
    plan = tb.FFPlan(bs, fs, sr, sr, D, 8, 7, 6, 5)
    assert plan.tc_supported
    plan.prepare(feats)
    plan.select(tb.FF_EXACT)
    ref = {t: [x.clone() for x in plan.selection(0, t)] for t in range(1, fs)}
    plan.select(tb.FF_AUTO)
    assert plan.stats()["tc_candidates"] > 0
    for t, (w, k, n) in ref.items():
        w2, k2, n2 = plan.selection(0, t)
        assert torch.equal(n, n2) and torch.equal(k, k2) and torch.equal(w, w2), t
