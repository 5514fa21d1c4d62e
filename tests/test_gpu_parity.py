"""Parity of the CUDA path (through the C ABI) with the oracle and the reference-made golden
fixtures.  Run on the B200 box:  python -m pytest tests -m gpu"""
import numpy as np
import pytest
import torch

import timet_oracle as O
import timetuning_b200 as tb
from conftest import assert_close
from parity import check_hard, check_soft, ff_taint
from timetuning_b200 import synth

pytestmark = pytest.mark.gpu
ENGINES = [tb.FF_EXACT, tb.FF_AUTO]


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


# ------------------------------------------------------------------ Sinkhorn
@pytest.mark.parametrize("name", ["sinkhorn_b392_k200", "sinkhorn_b1000_k37", "sinkhorn_b64_k300_eps01"])
def test_sinkhorn_golden(golden, name):
    g = golden(name)
    eps, iters = float(g["epsilon"]), int(g["iters"])
    scores = cu(g["scores"])
    q_in = torch.exp(scores / eps).t()                       # exactly what time_tuning.py:164 passes
    keep = q_in.clone()
    q = tb.sinkhorn(q_in, iters, 1)
    assert q.dtype == torch.float32 and q.shape == g["q"].shape and q.is_contiguous()
    assert torch.equal(q_in, keep), "caller's tensor must not be modified"
    assert_close(q.cpu().numpy(), g["q"], what=name)
    q2 = tb.sinkhorn_from_scores(scores, eps, iters)
    assert_close(q2.cpu().numpy(), g["q"], what=name + "/fused")
    assert (q.argmax(1).cpu().numpy() == g["q"].argmax(1)).all()


def test_sinkhorn_contiguous_kxb_and_cpu_input(golden):
    g = golden("sinkhorn_b392_k200")
    q_in = torch.exp(torch.from_numpy(g["scores"]) / float(g["epsilon"])).t().contiguous()   # CPU, physically K x B
    q = tb.sinkhorn(q_in, int(g["iters"]))
    assert q.device.type == "cpu"
    assert_close(q.numpy(), g["q"], what="cpu-input")


def test_sinkhorn_full_size_properties():
    """BASELINE configs[1] size: B = 32*784 rows, K = 200."""
    B, K = 32 * 784, 200
    scores = synth.cosine_scores(B, K, seed=71)
    q = tb.sinkhorn_from_scores(cu(scores), 0.05, 10)
    qn = q.double().cpu().numpy()
    assert np.abs(qn.sum(1) - 1).max() < 1e-5                              # my_utils.py:274
    ref = O.sinkhorn_scaling(scores, 0.05, 10, dtype=np.float64)
    assert_close(qn, ref, what="full-size vs fp64 scaling oracle")
    srt = np.sort(ref, axis=1)
    decided = (srt[:, -1] - srt[:, -2]) > 1e-5
    assert (qn.argmax(1) == ref.argmax(1))[decided].all()
    # run-to-run bit reproducibility (deterministic marginal reduction)
    q2 = tb.sinkhorn_from_scores(cu(scores), 0.05, 10)
    assert torch.equal(q, q2)


@pytest.mark.parametrize("iters", [0, 1, 2])
def test_sinkhorn_few_iterations(iters):
    scores = synth.cosine_scores(300, 64, seed=72)
    q = tb.sinkhorn_from_scores(cu(scores), 0.05, iters).cpu().numpy()
    ref = O.find_optimal_assignment(scores, 0.05, iters)
    assert_close(q, ref, what=f"iters={iters}")


def test_sinkhorn_ragged_shapes():
    for B, K in ((1, 8), (7, 5), (33, 130), (1000, 516), (50, 1000)):
        scores = synth.cosine_scores(B, K, seed=73 + B)
        q = tb.sinkhorn_from_scores(cu(scores), 0.05, 3).cpu().numpy()
        assert_close(q, O.find_optimal_assignment(scores, 0.05, 3), what=f"B={B},K={K}")


# ------------------------------------------------------------------ small routines
@pytest.mark.parametrize("name", ["restrict_8x8_s2", "restrict_14x14_s6", "restrict_5x9_s3"])
def test_restrict_neighborhood(golden, name):
    g = golden(name)
    ref = np.unpackbits(g["mask_bits"])[: int(np.prod(g["shape"]))].reshape(g["shape"]).astype(np.float32)
    m = tb.restrict_neighborhood(int(g["h"]), int(g["w"]), int(g["s"]))
    assert m.dtype == torch.float32
    assert np.array_equal(m.cpu().numpy(), ref)


def test_restrict_neighborhood_large():
    m = tb.restrict_neighborhood(60, 60, 12).cpu().numpy()
    assert np.array_equal(m, O.restrict_neighborhood(60, 60, 12))


@pytest.mark.parametrize("name", ["norm_mask", "norm_mask_f64"])
def test_norm_mask(golden, name):
    g = golden(name)
    out = tb.norm_mask(cu(g["mask"]))
    assert out.dtype == torch.from_numpy(g["out"]).dtype
    np.testing.assert_allclose(out.cpu().numpy(), g["out"], rtol=1e-6 if name == "norm_mask" else 1e-14, atol=0)


# ------------------------------------------------------------------ FF: golden fixtures (reference outputs)
class FE:
    def __init__(self, sr):
        self.spatial_resolution = sr


def test_label_propagation_golden(golden):
    g = golden("label_propagation_sr10")
    feats, segs = g["feats"], g["segs"]
    seg, feat_tar, mask = tb.label_propagation(int(g["s"]), int(g["topk"]), FE(int(g["sr"])), cu(feats[3]),
                                               [cu(feats[i]).t() for i in range(3)], [cu(s) for s in segs], None, True)
    assert seg.dtype == torch.float64 and tuple(seg.shape) == g["seg_tar"].shape
    assert np.array_equal(feat_tar.cpu().numpy(), g["feat_tar"])
    assert tuple(mask.shape) == (3, 100, 100)
    check_soft(seg.cpu().numpy(), g["seg_tar"], what="label_propagation")


@pytest.mark.parametrize("name", ["propagate_sr14_fifo", "propagate_sr12_k7", "propagate_sr9_nomask"])
def test_propagate_labels_golden(golden, name):
    g = golden(name)
    out = tb.propagate_labels(int(g["n_last"]), int(g["s"]), int(g["topk"]), FE(int(g["sr"])), cu(g["feats"]),
                              cu(g["first_seg"]), True)
    assert isinstance(out, list) and len(out) == g["segs"].shape[0]
    assert out[0].dtype == torch.float64 and out[0].is_cuda
    check_soft(torch.stack(out).cpu().numpy(), g["segs"], what=name)


def test_propagate_eval_onehot_golden(golden):
    """Eval call pattern mask_propagation.py:821: high-res one-hot first frame, CPU tensors in."""
    g = golden("propagate_eval_onehot")
    first = torch.from_numpy(O.to_one_hot(g["annotation"], int(g["n_obj"]) + 1)).unsqueeze(0)
    out = tb.propagate_labels(int(g["n_last"]), int(g["s"]), int(g["topk"]), FE(int(g["sr"])),
                              torch.from_numpy(g["feats"]), first, True)
    assert out[0].device.type == "cpu"
    out = torch.stack(out).numpy()
    check_soft(out, g["segs"], what="eval-onehot")
    check_hard(out.argmax(1), g["segs"], what="eval-onehot")


@pytest.mark.parametrize("engine", ENGINES)
def test_timet_step_cfg1_golden(golden, engine):
    """BASELINE configs[0] on reference-made fixtures: ViT-S/16 224^2, 4-frame clips, batch 2, K=200."""
    from timetuning_b200.step import ff_sinkhorn_step
    g = golden("timet_step_cfg1")
    sr = int(g["sr"])
    bq, tq, hard, labels = ff_sinkhorn_step(cu(g["head_src"]), cu(g["head_tgt"]), cu(g["backbone"]),
                                            cu(g["prototypes"]), engine=engine)
    assert_close(bq.cpu().numpy(), g["batch_q"], what="batch_q")
    assert_close(tq.cpu().numpy(), g["target_q"], what="target_q")
    soft = labels[:, -1].permute(0, 2, 1).reshape(2, -1, sr, sr).cpu().numpy()
    check_soft(soft, g["soft_last"], what="soft_last")
    assert np.array_equal(hard.cpu().numpy(), g["hard"])
    maps0 = labels[0, 1:].permute(0, 2, 1).reshape(3, -1, sr, sr).cpu().numpy()
    check_soft(maps0, g["clip0_maps"], what="clip0 all frames")


# ------------------------------------------------------------------ FF: against the oracle on seeded inputs
def _oracle_clip(feats, first_cl, n_last, radius, topk, sr):
    """feats [fs,N,D], first_cl [N,C] -> oracle segs [fs-1,C,sr,sr] (fp32 dense restatement)."""
    C = first_cl.shape[1]
    first = first_cl.T.reshape(1, C, sr, sr)
    return np.stack(O.propagate_labels(n_last, radius, topk, sr, feats, first))


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("sr,D,C,fs,n_last,radius,topk", [
    (14, 384, 200, 4, 7, 6, 5),      # config 1 shapes
    (28, 384, 200, 8, 7, 6, 5),      # config 2 shapes (2 clips)
    (20, 96, 11, 12, 3, 4, 7),       # FIFO eviction + eval-style k
    (16, 70, 6, 5, 7, 3, 2),         # dim not a multiple of 4/64, C not a multiple of 4
    (12, 64, 8, 4, 7, 12, 5),        # radius >= grid: window = whole frame
    (28, 768, 40, 3, 7, 6, 5),       # ViT-B width
])
def test_propagate_vs_oracle(engine, sr, D, C, fs, n_last, radius, topk):
    bs, N = 2, sr * sr
    feats = synth.clip_features(bs, fs, sr, D, seed=sr + D)
    first = np.stack([synth.soft_labels(N, C, seed=100 + b) for b in range(bs)])
    labels, hard = tb.propagate_labels_batched(cu(feats), cu(first), n_last, radius, topk, engine=engine)
    labels = labels.cpu().numpy()
    assert np.array_equal(labels[:, 0], first)
    for b in range(bs):
        ref = _oracle_clip(feats[b], first[b], n_last, radius, topk, sr)
        _, taint, margin = ff_taint(n_last, radius, topk, sr, feats[b], first[b].T.reshape(C, sr, sr))
        got = labels[b, 1:].transpose(0, 2, 1).reshape(fs - 1, C, sr, sr)
        check_soft(got, ref, taint, what=f"clip {b}")
        check_hard(hard[b].cpu().numpy(), ref[-1], taint[-1], what=f"clip {b} hard")


def test_selection_structure_and_engine_agreement():
    """<= k (+ties) keys per query, all inside the window and in legal context frames, weights sum
    to 1; the tensor-core engine (when supported) reproduces the exact engine bit for bit."""
    sr, D, C, fs, n_last, radius, topk, bs = 28, 384, 16, 8, 7, 6, 5, 3
    N = sr * sr
    feats = cu(synth.clip_features(bs, fs, sr, D, seed=5))
    plan = tb.FFPlan(bs, fs, sr, sr, D, C, n_last, radius, topk)
    plan.prepare(feats)
    plan.select(tb.FF_EXACT)
    st = plan.stats()
    assert st["queries"] == bs * (fs - 1) * N and st["selected"] >= st["queries"] * topk
    sel = {}
    for clip in (0, bs - 1):
        for t in (1, 4, fs - 1):
            w, k, c = (x.cpu().numpy() for x in plan.selection(clip, t))
            sel[(clip, t)] = (w, k, c)
            assert (c >= topk).all() and (c <= plan.kw).all()
            valid = np.arange(plan.kw)[None] < c[:, None]
            assert np.allclose((w * valid).sum(1), 1, atol=1e-6)
            assert (k[~valid] == -1).all() and (w[~valid] == 0).all()
            frame, patch = k // N, k % N
            ctx = O.context_frames(t, n_last)
            assert np.isin(frame[valid], ctx).all()
            qi = np.arange(N)[:, None]
            assert (np.abs(patch // sr - qi // sr)[valid] <= radius).all()
            assert (np.abs(patch % sr - qi % sr)[valid] <= radius).all()
            assert (np.diff(w, axis=1)[valid[:, 1:]] <= 0).all(), "weights sorted descending"
    if plan.tc_supported:
        plan.select(tb.FF_TC)
        for (clip, t), (w, k, c) in sel.items():
            w2, k2, c2 = (x.cpu().numpy() for x in plan.selection(clip, t))
            assert np.array_equal(c, c2) and np.array_equal(k, k2) and np.array_equal(w, w2), (clip, t)


def test_duplicate_frames_keep_ties():
    """A clip with repeated frames (the loader samples with replacement, data_loader.py:621-623) makes
    exact affinity ties across contexts; the reference keeps them all (aff < kth -> 0, :434)."""
    sr, D, C, fs, topk = 12, 64, 6, 4, 5
    N = sr * sr
    f = synth.clip_features(1, 2, sr, D, seed=9)[0]
    feats = np.stack([f[0], f[0], f[0], f[1]])[None]           # frames 0,1,2 identical
    first = synth.soft_labels(N, C, seed=10)[None]
    labels, _ = tb.propagate_labels_batched(cu(feats), cu(first), 7, 4, topk, engine=tb.FF_EXACT)
    ref = _oracle_clip(feats[0], first[0], 7, 4, topk, sr)
    got = labels[0, 1:].permute(0, 2, 1).reshape(fs - 1, C, sr, sr).cpu().numpy()
    check_soft(got, ref, what="duplicate frames")
    plan = tb.FFPlan(1, fs, sr, sr, D, C, 7, 4, topk)
    plan.prepare(cu(feats)); plan.select(tb.FF_EXACT)
    assert plan.stats()["tie_queries"] > 0


def test_full_batch_properties_config2():
    """BASELINE configs[1] at full size (32 clips): size-independent properties."""
    bs, fs, sr, D, K = 32, 8, 28, 384, 200
    N = sr * sr
    feats = cu(synth.clip_features(bs, fs, sr, D, seed=1))
    first = cu(np.stack([synth.soft_labels(N, K, seed=200 + (b % 4)) for b in range(bs)]))
    labels, hard = tb.propagate_labels_batched(feats, first, 7, 6, 5)
    s = labels.double().sum(-1)
    assert (s - 1).abs().max().item() < 1e-5, "label rows stay distributions (SURVEY.md §4.3)"
    assert labels.min().item() >= 0
    assert torch.equal(hard.view(bs, N), labels[:, -1].argmax(-1))
    # permutation equivariance over clips
    perm = torch.randperm(bs, generator=torch.Generator().manual_seed(0)).cuda()
    labels_p, hard_p = tb.propagate_labels_batched(feats[perm].contiguous(), first[perm].contiguous(), 7, 6, 5)
    assert torch.equal(labels_p, labels[perm]) and torch.equal(hard_p, hard[perm])
    # a 2-clip slice agrees with the oracle
    fn, fr = feats[:2].cpu().numpy(), first[:2].cpu().numpy()
    for b in range(2):
        ref = _oracle_clip(fn[b], fr[b], 7, 6, 5, sr)
        _, taint, _ = ff_taint(7, 6, 5, sr, fn[b], fr[b].T.reshape(K, sr, sr))
        got = labels[b, 1:].permute(0, 2, 1).reshape(fs - 1, K, sr, sr).cpu().numpy()
        check_soft(got, ref, taint, what=f"cfg2 clip {b}")


@pytest.mark.parametrize("env", [{"TIMET_GATHER_L1": "0"}, {"TIMET_FIN_STAGED": "0"}, {"TIMET_GATHER_L1": "0", "TIMET_FIN_STAGED": "0"}],
                         ids=lambda e: "+".join(f"{k[6:]}={v}" for k, v in e.items()))
def test_gather_and_finalize_variants_are_bit_identical(timet_env, env):
    """The label rows gathered through L1 / with L2-only loads and the exact re-evaluation from shared-memory staged rows /
    from registers give the same bits (same per-lane chains, same order of the weighted sums)."""
    bs, fs, sr, D, K = 6, 8, 28, 384, 200
    N = sr * sr
    feats = cu(synth.clip_features(bs, fs, sr, D, seed=31))
    first = cu(np.stack([synth.soft_labels(N, K, seed=300 + b) for b in range(bs)]))
    labels, hard = tb.propagate_labels_batched(feats, first, 7, 6, 5)
    timet_env(**env)
    labels_v, hard_v = tb.propagate_labels_batched(feats, first, 7, 6, 5)
    assert torch.equal(labels_v, labels) and torch.equal(hard_v, hard)


def test_davis_style_eval_config4_slice():
    """BASELINE configs[3] shapes (480p ViT-S/8 -> 60x60 grid, radius 12, top-k 7, n_last 7, C=11),
    first 10 frames of a synthetic video."""
    sr, D, n_obj, fs = 60, 384, 10, 10
    feats = synth.clip_features(1, fs, sr, D, seed=4)[0]
    ann = synth.blob_label_map(480, n_obj, seed=6)
    first = torch.from_numpy(O.to_one_hot(ann, n_obj + 1)).unsqueeze(0)
    out = torch.stack(tb.propagate_labels(7, 12, 7, FE(sr), cu(feats), first.cuda(), True)).cpu().numpy()
    ref = np.stack(O.propagate_labels(7, 12, 7, sr, feats, first.numpy()))
    _, taint, _ = ff_taint(7, 12, 7, sr, feats, O.nearest_resize(first.numpy().astype(np.float64), sr, sr)[0])
    check_soft(out, ref, taint, what="davis-style")
    check_hard(out.argmax(1), ref, taint, what="davis-style hard")


def test_host_pipeline_matches_device_step():
    """The chunked host-fed pipeline (bench.py e2e path) gives bit-identical results to the device-resident step."""
    from timetuning_b200.step import HostStepPipeline, ff_sinkhorn_step
    bs, fs, sr, D, K = 4, 4, 14, 384, 200
    N = sr * sr
    backbone = synth.clip_features(bs, fs, sr, D, seed=1)
    head = synth.head_features(backbone[:, [0, -1]], 256, seed=2)
    protos = cu(synth.prototypes(K, 256, seed=3))
    hs, ht = np.ascontiguousarray(head[:, 0]), np.ascontiguousarray(head[:, 1])
    q1, t1, h1, _ = ff_sinkhorn_step(cu(hs), cu(ht), cu(backbone), protos)
    pipe = HostStepPipeline(bs, fs, N, D, 256, K, chunks=2)
    pin = [torch.from_numpy(x).pin_memory() for x in (hs, ht, backbone)]
    for _ in range(2):
        q2, t2, h2 = pipe.run(pin[0], pin[1], pin[2], protos)
    torch.cuda.synchronize()
    # same kernels, same launch shapes -> same bits (the scores GEMM sees the two heads as one [2*B, dh] matrix in both)
    assert torch.equal(q1, q2) and torch.equal(t1, t2), ((q1 - q2).abs().max().item(), (t1 - t2).abs().max().item())
    assert torch.equal(h1.cpu(), h2)


@pytest.mark.parametrize("B,K,dh", [(25088, 200, 256), (392, 200, 256), (1000, 300, 256), (77, 21, 96), (300, 520, 64)])
def test_cosine_scores(B, K, dh):
    """timet_cosine_scores (fp16 hi/lo split on tensor cores) vs F.normalize(x) @ prototypes.t() in float64."""
    rng = np.random.default_rng(B + K)
    x = (rng.standard_normal((B, dh)) * rng.uniform(0.1, 5.0, size=(B, 1))).astype(np.float32)
    p = synth.prototypes(K, dh, seed=K)
    got = tb.cosine_scores(cu(x), cu(p)).cpu().numpy()
    xn = x.astype(np.float64) / np.maximum(np.linalg.norm(x.astype(np.float64), axis=1, keepdims=True), 1e-12)
    want = xn @ p.astype(np.float64).T
    err = np.abs(got - want).max()
    assert err < 2e-6, f"max abs err {err:.3e}"
    ref32 = (torch.nn.functional.normalize(cu(x), dim=-1) @ cu(p).t()).cpu().numpy()
    assert np.abs(got - ref32).max() < 2e-6


def test_eval_tail_upsample_argmax(golden):
    """mask_propagation.py:822-824 (bilinear align_corners=False + max over channels), fused, vs torch on the
    reference-made propagate_eval_onehot maps; and the one-call eval pattern."""
    g = golden("propagate_eval_onehot")
    maps = torch.from_numpy(g["segs"])                                  # [fs-1, C, sr, sr] float64 (reference output)
    R = g["annotation"].shape[-1]
    up = torch.nn.functional.interpolate(maps, size=(R, R), mode="bilinear", align_corners=False)
    want = up.max(dim=1)[1]
    got = tb.upsample_argmax(maps.cuda(), R).cpu()
    srt = up.sort(dim=1)[0]
    decided = (srt[:, -1] - srt[:, -2]) > 1e-5                           # near-ties of the interpolated maps are exempt
    assert got.shape == want.shape and got.dtype == torch.int64
    assert (got == want)[decided].all() and decided.float().mean() > 0.95
    first = torch.from_numpy(O.to_one_hot(g["annotation"], int(g["n_obj"]) + 1)).unsqueeze(0)
    pred = tb.propagate_labels_eval(int(g["n_last"]), int(g["s"]), int(g["topk"]), cu(g["feats"]), first, R).cpu()
    assert (pred == want)[decided].all()


def test_non_square_grid_through_the_c_abi():
    """The C ABI takes grid_h != grid_w (additive; the reference is square-only, mask_propagation.py:407): both
    engines against a dense fp64 restatement."""
    H, W, D, C, fs, radius, topk, n_last = 10, 24, 64, 8, 4, 3, 4, 7
    N = H * W
    rng = np.random.default_rng(3)
    feats = rng.standard_normal((1, fs, N, D)).astype(np.float32)
    first = synth.soft_labels(N, C, seed=1)
    fn = O.l2_normalize_rows(feats[0]).astype(np.float64)
    rows, cols = np.arange(N) // W, np.arange(N) % W
    win = (np.abs(rows[:, None] - rows[None]) <= radius) & (np.abs(cols[:, None] - cols[None]) <= radius)
    segs = [first.astype(np.float64)]
    for t in range(1, fs):
        ctx = O.context_frames(t, n_last)
        aff = np.concatenate([np.exp(fn[t] @ fn[c].T / 0.1) * win for c in ctx], axis=1)
        kth = np.sort(aff, axis=1)[:, -topk]
        wgt = np.where(aff >= kth[:, None], aff, 0)
        wgt /= wgt.sum(1, keepdims=True)
        segs.append(wgt @ np.concatenate([segs[c] for c in ctx], axis=0))
    for engine in ENGINES:
        plan = tb.FFPlan(1, fs, H, W, D, C, n_last, radius, topk)
        labels = torch.empty((1, fs, N, C), dtype=torch.float32, device="cuda")
        labels[0, 0] = cu(first)
        plan.propagate(cu(feats), labels, None, engine)
        got = labels[0].cpu().numpy()
        for t in range(1, fs):
            assert np.abs(got[t] - segs[t]).max() < 1e-5, (engine, t)


def test_sinkhorn_streaming_path_matches_resident(timet_env):
    """The multi-launch streaming kernels (multi-GPU NCCL fallback / shapes that do not fit shared memory) against
    the resident cooperative kernel and the oracle."""
    scores = synth.cosine_scores(4 * 784, 200, seed=91)
    q_res = tb.sinkhorn_from_scores(cu(scores), 0.05, 10)
    timet_env(TIMET_SK_STREAMING="1")
    q_str = tb.sinkhorn_from_scores(cu(scores), 0.05, 10)
    q_str2 = tb.sinkhorn_from_scores(cu(scores), 0.05, 10)
    timet_env(TIMET_SK_STREAMING=None)
    assert torch.equal(q_str, q_str2), "streaming path must be bit-reproducible"
    ref = O.sinkhorn_scaling(scores, 0.05, 10, dtype=np.float64)
    assert_close(q_str.cpu().numpy(), ref, what="streaming vs fp64 oracle")
    assert_close(q_res.cpu().numpy(), q_str.cpu().numpy(), what="resident vs streaming", atol=1e-6, rtol=1e-5)


def test_sinkhorn_rows_beyond_shared_memory():
    """configs[2] at 2 GPUs has 100 352 rows per rank: more than fits in shared memory -> streaming kernels."""
    B, K = 128 * 784, 200
    scores = synth.cosine_scores(B, K, seed=92)
    q = tb.sinkhorn_from_scores(cu(scores), 0.05, 10).double().cpu().numpy()
    assert np.abs(q.sum(1) - 1).max() < 1e-5
    col = q.sum(0) * K / B
    assert abs(col.mean() - 1) < 1e-3
    ref = O.sinkhorn_scaling(scores[:2048], 0.05, 0, dtype=np.float64)        # iters=0 rows are independent: sanity only
    assert ref.shape == (2048, K)
    full = O.sinkhorn_scaling(scores, 0.05, 10, dtype=np.float64)
    assert_close(q, full, what="B=100352 vs fp64 oracle")


@pytest.mark.parametrize("B,K,iters", [(32 * 784, 200, 10), (8 * 784, 200, 3), (4 * 196, 64, 1), (2 * 3136, 300, 10), (1000, 516, 4)])
def test_sinkhorn_pair_is_two_single_calls(timet_env, B, K, iters):
    """timet_sinkhorn_pair: the source and target assignment of a step in ONE launch.  Default = DUAL (the two problems
    side by side on half of the SMs each: same maths, another row partition -> equal to two single calls within fp32
    summation order, bit-reproducible); TIMET_SK_DUAL=0 = one after the other, bit-identical to two single calls; K = 516
    exercises the sequential fall-back."""
    s0, s1 = cu(synth.cosine_scores(B, K, seed=5)), cu(synth.cosine_scores(B, K, seed=6))
    r0, r1 = tb.sinkhorn_from_scores(s0, 0.05, iters), tb.sinkhorn_from_scores(s1, 0.05, iters)
    d0, d1 = tb.sinkhorn_pair_from_scores(s0, s1, 0.05, iters)           # default: dual
    e0, e1 = tb.sinkhorn_pair_from_scores(s0, s1, 0.05, iters)
    assert torch.equal(d0, e0) and torch.equal(d1, e1), "dual mode must be bit-reproducible"
    assert_close(d0.cpu().numpy(), r0.cpu().numpy(), atol=1e-7, rtol=2e-5, what="dual vs single (0)")
    assert_close(d1.cpu().numpy(), r1.cpu().numpy(), atol=1e-7, rtol=2e-5, what="dual vs single (1)")
    assert_close(d1.cpu().numpy(), O.sinkhorn_scaling(s1.cpu().numpy(), 0.05, iters, dtype=np.float64), what="dual vs fp64 oracle")
    timet_env(TIMET_SK_DUAL="0")
    q0, q1 = tb.sinkhorn_pair_from_scores(s0, s1, 0.05, iters)
    timet_env(TIMET_SK_DUAL=None)
    assert torch.equal(q0, r0) and torch.equal(q1, r1), ((q0 - r0).abs().max().item(), (q1 - r1).abs().max().item())


def test_sinkhorn_strided_output_into_label_frames():
    """out=labels[:, 0]: the assignment of clip b lands in frame 0 of the channel-last label tensor (no copy between
    Sinkhorn and Feature-Forwarding, time_tuning.py:144-147)."""
    bs, fs, N, K = 6, 3, 196, 200
    s0, s1 = cu(synth.cosine_scores(bs * N, K, seed=7)), cu(synth.cosine_scores(bs * N, K, seed=8))
    labels = torch.full((bs, fs, N, K), -1.0, device="cuda")
    flat = torch.empty((bs, N, K), device="cuda")
    tb.sinkhorn_pair_from_scores(s0, s1, 0.05, 10, out0=labels[:, 0], out1=flat)
    d0, d1 = tb.sinkhorn_pair_from_scores(s0, s1, 0.05, 10)
    assert torch.equal(labels[:, 0].reshape(bs * N, K), d0) and torch.equal(flat.reshape(bs * N, K), d1)
    assert_close(d0.cpu().numpy(), tb.sinkhorn_from_scores(s0, 0.05, 10).cpu().numpy(), atol=1e-7, rtol=2e-5, what="pair vs single")
    assert (labels[:, 1:] == -1).all(), "other frames untouched"
    lab2 = torch.full((bs, fs, N, K), -1.0, device="cuda")
    tb.sinkhorn_from_scores(s0, 0.05, 10, out=lab2[:, 0])
    assert torch.equal(lab2[:, 0].reshape(bs * N, K), tb.sinkhorn_from_scores(s0, 0.05, 10)) and (lab2[:, 1:] == -1).all()
    with pytest.raises(ValueError):
        tb.sinkhorn_from_scores(s0, 0.05, 10, out=labels[:, 0, :, :100])


def test_sinkhorn_share_sm_variant_matches():
    """The 512-thread resident variant (leaves half the SM to co-resident kernels) against the default and the oracle."""
    s0 = cu(synth.cosine_scores(8 * 784, 200, seed=9))
    a, b = tb.sinkhorn_from_scores(s0, 0.05, 10), tb.sinkhorn_from_scores(s0, 0.05, 10, share_sm=True)
    assert_close(b.cpu().numpy(), a.cpu().numpy(), atol=1e-7, rtol=1e-5, what="share_sm vs default")


def test_cosine_scores_multi_and_autograd():
    """One GEMM launch for several feature blocks == separate calls; the autograd wrapper's gradients == torch's."""
    from timetuning_b200 import training
    rng = np.random.default_rng(0)
    xs = [cu((rng.standard_normal((1568, 256)) * 3).astype(np.float32)) for _ in range(3)]
    p = cu(synth.prototypes(200, 256, seed=1))
    multi = tb.cosine_scores_multi(xs, p)
    for i, x in enumerate(xs):
        assert torch.equal(multi[i * 1568:(i + 1) * 1568], tb.cosine_scores(x, p))
    x = xs[0].clone().requires_grad_(True)
    pp = p.clone().requires_grad_(True)
    g = cu(rng.standard_normal((1568, 200)).astype(np.float32))
    (training.cosine_scores_autograd(x, pp) * g).sum().backward()
    x2 = xs[0].clone().requires_grad_(True)
    p2 = p.clone().requires_grad_(True)
    ((torch.nn.functional.normalize(x2, dim=-1) @ p2.t()) * g).sum().backward()
    assert_close(x.grad.cpu().numpy(), x2.grad.cpu().numpy(), atol=1e-5, rtol=1e-4, what="d/dx")
    assert_close(pp.grad.cpu().numpy(), p2.grad.cpu().numpy(), atol=1e-4, rtol=1e-4, what="d/dprototypes")


def test_non_square_grid_drop_ins():
    """Python drop-ins with a (h, w) spatial_resolution (additive; DAVIS 480x854 style) against a dense fp64 restatement."""
    H, W, D, C, fs, radius, topk = 9, 16, 48, 5, 4, 3, 4
    N = H * W
    feats = synth.clip_features(1, fs, 16, D, seed=13)[0][:, :N]        # any [fs, N, D] features
    first = synth.soft_labels(N, C, seed=2).T.reshape(1, C, H, W)

    class FEhw:
        spatial_resolution = (H, W)
    out = torch.stack(tb.propagate_labels(7, radius, topk, FEhw(), cu(feats), cu(first), True)).cpu().numpy()
    fn = O.l2_normalize_rows(feats).astype(np.float64)
    rows, cols = np.arange(N) // W, np.arange(N) % W
    win = (np.abs(rows[:, None] - rows[None]) <= radius) & (np.abs(cols[:, None] - cols[None]) <= radius)
    segs = [first[0].reshape(C, N).T.astype(np.float64)]
    for t in range(1, fs):
        ctx = O.context_frames(t, 7)
        aff = np.concatenate([np.exp(fn[t] @ fn[c].T / 0.1) * win for c in ctx], axis=1)
        kth = np.sort(aff, axis=1)[:, -topk]
        wgt = np.where(aff >= kth[:, None], aff, 0)
        wgt /= wgt.sum(1, keepdims=True)
        segs.append(wgt @ np.concatenate([segs[c] for c in ctx], axis=0))
    want = np.stack([s.T.reshape(C, H, W) for s in segs[1:]])
    assert out.shape == want.shape and np.abs(out - want).max() < 1e-5
    assert np.array_equal(tb.restrict_neighborhood(H, W, radius).cpu().numpy(), O.restrict_neighborhood(H, W, radius))
    pred = tb.propagate_labels_eval(7, radius, topk, cu(feats), cu(first), (4 * H, 4 * W), grid=(H, W))
    assert tuple(pred.shape) == (fs - 1, 4 * H, 4 * W)


def test_exact_engine_vs_oracle_config5_grid():
    """BASELINE configs[4] grid (56 x 56 patches, D = 768) through the EXACT engine against the oracle itself (the
    tensor-core engine is checked against the exact one elsewhere): one clip, three frames."""
    sr, D, C, fs = 56, 768, 12, 3
    N = sr * sr
    feats = synth.clip_features(1, fs, sr, D, seed=21)
    first = synth.soft_labels(N, C, seed=22)[None]
    for engine in (tb.FF_EXACT, tb.FF_AUTO):
        labels, hard = tb.propagate_labels_batched(cu(feats), cu(first), 7, 6, 5, engine=engine, check=True)
        ref = _oracle_clip(feats[0], first[0], 7, 6, 5, sr)
        _, taint, _ = ff_taint(7, 6, 5, sr, feats[0], first[0].T.reshape(C, sr, sr))
        got = labels[0, 1:].permute(0, 2, 1).reshape(fs - 1, C, sr, sr).cpu().numpy()
        check_soft(got, ref, taint, what=f"56x56x768 engine {engine}")
        check_hard(hard[0].cpu().numpy(), ref[-1], taint[-1], what="56x56x768 hard")


def test_sinkhorn_hybrid_kernel(timet_env):
    """Rows beyond shared memory (configs[2] at 4 GPUs: 64 clips = 50 176 rows per rank): ONE hybrid launch (resident part +
    re-read part) against the one-launch-per-pass streaming path and the fp64 oracle; bit-reproducible."""
    B, K = 64 * 784, 200
    assert tb.ops.sinkhorn_mode(B, K) == "hybrid" and tb.ops.sinkhorn_mode(32 * 784, K) == "resident"
    scores = synth.cosine_scores(B, K, seed=93)
    q = tb.sinkhorn_from_scores(cu(scores), 0.05, 10)
    assert torch.equal(q, tb.sinkhorn_from_scores(cu(scores), 0.05, 10))
    timet_env(TIMET_SK_STREAMING="1")
    q_str = tb.sinkhorn_from_scores(cu(scores), 0.05, 10)
    timet_env(TIMET_SK_STREAMING=None)
    assert_close(q.cpu().numpy(), q_str.cpu().numpy(), atol=1e-6, rtol=1e-5, what="hybrid vs streaming")
    assert_close(q.double().cpu().numpy(), O.sinkhorn_scaling(scores, 0.05, 10, dtype=np.float64), what="hybrid vs fp64 oracle")
    # strided output and K = 300 (three float4 per lane)
    s2 = synth.cosine_scores(40 * 784, 300, seed=94)
    out = torch.zeros((40, 2, 784, 300), device="cuda")
    tb.sinkhorn_from_scores(cu(s2), 0.05, 10, out=out[:, 0])
    assert_close(out[:, 0].reshape(-1, 300).double().cpu().numpy(), O.sinkhorn_scaling(s2, 0.05, 10, dtype=np.float64), what="hybrid K=300")
    assert (out[:, 1] == 0).all()
