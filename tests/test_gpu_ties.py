"""Exact affinity ties: the reference keeps EVERY key whose affinity equals the k-th largest
(mask_propagation.py:432-436: aff[aff < kth] = 0; aff /= aff.sum()), so a query can carry more than topk weights.
They arise from repeated frames (the loader samples with replacement, data_loader.py:621-623) and from repeated
patch features.  Up to kw survivors live in the regular selection slots; larger sets are "wide rows" in the pool."""
import numpy as np
import pytest
import torch

import timet_oracle as O
import timetuning_b200 as tb
from parity import check_hard, check_soft
from timetuning_b200 import synth

pytestmark = pytest.mark.gpu


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


class FE:
    def __init__(self, sr):
        self.spatial_resolution = sr


def _oracle(feats, first_cl, n_last, radius, topk, sr):
    C = first_cl.shape[1]
    return np.stack(O.propagate_labels(n_last, radius, topk, sr, feats, first_cl.T.reshape(1, C, sr, sr)))


def pattern_features(fs, sr, D, n_patterns, seed):
    """Rows with exactly four entries of +-0.5: unit norm, and every dot product is a multiple of 0.25 that float32
    adds exactly in ANY order -- ties are exact in the oracle's BLAS, in the tensor cores and in the fp32 engines."""
    rng = np.random.default_rng(seed)
    pats = np.zeros((n_patterns, D), np.float32)
    for p in range(n_patterns):
        idx = rng.choice(D, size=4, replace=False)
        pats[p, idx] = rng.choice([-0.5, 0.5], size=4)
    which = rng.integers(0, n_patterns, size=(fs, sr * sr))
    return pats[which]


@pytest.mark.parametrize("engine", [tb.FF_EXACT, tb.FF_AUTO])
@pytest.mark.parametrize("n_dup,topk", [(4, 5), (5, 5), (4, 7), (6, 7)])
def test_many_duplicate_frames(engine, n_dup, topk):
    """>= 4 identical context frames: every affinity appears n_dup times, the survivor set is n_dup * ceil(k / n_dup)."""
    sr, D, C, radius = 12, 64, 6, 4
    N = sr * sr
    f = synth.clip_features(1, 3, sr, D, seed=19)[0]
    feats = np.stack([f[0]] * n_dup + [f[1], f[2]])[None]
    fs = feats.shape[1]
    first = synth.soft_labels(N, C, seed=20)[None]
    labels, hard = tb.propagate_labels_batched(cu(feats), cu(first), 7, radius, topk, engine=engine, check=True)
    ref = _oracle(feats[0], first[0], 7, radius, topk, sr)
    got = labels[0, 1:].permute(0, 2, 1).reshape(fs - 1, C, sr, sr).cpu().numpy()
    check_soft(got, ref, what=f"{n_dup} duplicate frames, k={topk}")
    check_hard(hard[0].cpu().numpy(), ref[-1], what="hard")
    plan = tb.FFPlan(1, fs, sr, sr, D, C, 7, radius, topk)
    plan.prepare(cu(feats)); plan.select(engine)
    st = plan.stats()
    assert st["tie_queries"] > 0 and st["truncated_queries"] == 0
    # frame n_dup sees n_dup identical contexts: n_dup * ceil(k / n_dup) survivors for every query
    _, _, cnt = plan.selection(0, n_dup)
    assert (cnt.cpu().numpy() == n_dup * -(-topk // n_dup)).all()


@pytest.mark.parametrize("engine", [tb.FF_EXACT, tb.FF_AUTO])
def test_wide_rows_keep_every_tie(engine):
    """Repeated patch features: dozens to hundreds of keys tie with the k-th affinity -- far more than the kw = 16
    regular slots and than the 32-entry sorted list.  Every one of them must be kept, like the reference does."""
    sr, D, C, fs, radius, topk, n_last = 12, 64, 6, 5, 4, 5, 7
    N = sr * sr
    feats = pattern_features(fs, sr, D, n_patterns=6, seed=3)[None]
    first = synth.soft_labels(N, C, seed=4)[None]
    labels, hard = tb.propagate_labels_batched(cu(feats), cu(first), n_last, radius, topk, engine=engine, check=True)
    ref = _oracle(feats[0], first[0], n_last, radius, topk, sr)
    got = labels[0, 1:].permute(0, 2, 1).reshape(fs - 1, C, sr, sr).cpu().numpy()
    check_soft(got, ref, what="wide rows")
    check_hard(hard[0].cpu().numpy(), ref[-1], what="wide rows hard")

    plan = tb.FFPlan(1, fs, sr, sr, D, C, n_last, radius, topk)
    plan.prepare(cu(feats)); plan.select(engine)
    st = plan.stats()
    assert st["wide_rows"] > 0 and st["truncated_queries"] == 0
    # structure of the wide rows of the last frame against the fp64 explainer's survivor counts
    ex = O.ff_sparse(n_last, radius, topk, sr, feats[0], first[0].T.reshape(C, sr, sr))
    w, k, cnt = (x.cpu().numpy() for x in plan.selection(0, fs - 1))
    n_sel = np.where(cnt < 0, -cnt, cnt)
    assert np.array_equal(n_sel, ex["nnz"][-1])
    assert (n_sel > plan.kw).any() and (n_sel > 32).any()
    for i in np.nonzero(cnt < 0)[0][:8]:
        ww, kk = (x.cpu().numpy() for x in plan.wide_row(k[i, 0], -cnt[i]))
        assert abs(ww.sum() - 1) < 1e-5 and (ww > 0).all()
        frame, patch = kk // N, kk % N
        assert np.isin(frame, O.context_frames(fs - 1, n_last)).all()
        assert (np.abs(patch // sr - i // sr) <= radius).all() and (np.abs(patch % sr - i % sr) <= radius).all()
        assert len(np.unique(kk)) == len(kk)


def test_drop_in_with_wide_rows_matches_reference_semantics():
    """The same through the reference-signature shim (float64 list output)."""
    sr, D, C, fs = 10, 64, 4, 4
    feats = pattern_features(fs, sr, D, n_patterns=5, seed=8)
    first = synth.soft_labels(sr * sr, C, seed=9).T.reshape(1, C, sr, sr)
    out = torch.stack(tb.propagate_labels(7, 3, 5, FE(sr), cu(feats), cu(first), True)).cpu().numpy()
    ref = np.stack(O.propagate_labels(7, 3, 5, sr, feats, first))
    check_soft(out, ref, what="drop-in wide rows")


def test_pool_exhaustion_is_an_error_not_a_silent_truncation():
    """Constant features: every in-window key of every context ties (169 x ctx entries per query at radius 6).
    The pool cannot hold that for a whole frame batch: the shims raise instead of returning a truncated result."""
    sr, D, C, fs = 28, 64, 4, 4
    feats = np.ones((fs, sr * sr, D), np.float32)
    first = synth.soft_labels(sr * sr, C, seed=2).T.reshape(1, C, sr, sr)
    with pytest.raises(RuntimeError, match="ties"):
        tb.propagate_labels(7, 6, 5, FE(sr), cu(feats), cu(first), True)
    plan = tb.FFPlan(1, fs, sr, sr, D, C, 7, 6, 5)
    plan.prepare(cu(feats[None])); plan.select(tb.FF_AUTO)
    st = plan.stats()
    assert st["truncated_queries"] > 0
    # a small all-constant problem fits the pool and is exact: uniform average over the whole window
    sr2 = 8
    feats2 = np.ones((3, sr2 * sr2, D), np.float32)
    first2 = synth.soft_labels(sr2 * sr2, C, seed=2).T.reshape(1, C, sr2, sr2)
    out = torch.stack(tb.propagate_labels(7, 2, 5, FE(sr2), cu(feats2), cu(first2), True)).cpu().numpy()
    ref = np.stack(O.propagate_labels(7, 2, 5, sr2, feats2, first2))
    check_soft(out, ref, what="constant features")
