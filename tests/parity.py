"""Tolerance rule of the north star (BASELINE.json) with documented near-tie exemptions.

Soft labels / Q: |a-b| <= 1e-4 + 1e-5*|b|.  Hard ids: bit-exact except where the fp64 margin of the
decision is < 1e-5.  FF selection is discontinuous at the k-th/(k+1)-th neighbour
(mask_propagation.py:432-434): a query whose relative margin there is < 1e-5 in float64 may
legitimately pick a different key set under another fp32 summation order, and so
may every later query that reads its labels.  `ff_taint` marks exactly those.
"""
import numpy as np

import timet_oracle as O

ATOL, RTOL, TIE = 1e-4, 1e-5, 1e-5


def ff_taint(n_last, radius, topk, sr, feats, first_seg_crs, tie=TIE):
    """fp64 explainer -> (segs [fs-1,C,sr,sr], tainted [fs-1,N] bool, margins [fs-1,N])."""
    feats = np.asarray(feats)
    fs, N, D = feats.shape
    ex = O.ff_sparse(n_last, radius, topk, sr, feats, first_seg_crs)
    near = ex["margin"] < tie
    if not near.any():
        return ex["segs"], near, ex["margin"]
    # dependency propagation through the selected keys (recomputed densely; only when needed)
    fn = O.l2_normalize_rows(feats).astype(np.float64)
    rows, cols = np.arange(N) // sr, np.arange(N) % sr
    win = ((np.abs(rows[:, None] - rows[None]) <= radius) & (np.abs(cols[:, None] - cols[None]) <= radius)) \
        if radius > 0 else np.ones((N, N), bool)
    taint = np.zeros((fs, N), bool)
    for t in range(1, fs):
        ctx = O.context_frames(t, n_last)
        aff = np.stack([np.exp(fn[t] @ fn[c].T / 0.1) * win for c in ctx]).transpose(1, 0, 2).reshape(N, -1)
        kth = np.sort(aff, axis=1)[:, -topk]
        sel = aff >= kth[:, None] * (1 - 10 * tie)
        src_taint = np.concatenate([taint[c] for c in ctx])
        taint[t] = near[t - 1] | (sel & src_taint[None]).any(axis=1)
    return ex["segs"], taint[1:], ex["margin"]


def check_soft(actual, expected, tainted=None, what="", atol=ATOL, rtol=RTOL, max_taint_frac=0.02):
    """actual/expected [..., C, sr, sr] (or [.., N, C] if channel_last) with tainted [..., N]."""
    actual = np.asarray(actual, np.float64)
    expected = np.asarray(expected, np.float64)
    assert actual.shape == expected.shape, (what, actual.shape, expected.shape)
    err = np.abs(actual - expected)
    bad = err > atol + rtol * np.abs(expected)
    if tainted is not None:
        C = actual.shape[-3]
        t = np.asarray(tainted).reshape(*actual.shape[:-3], 1, actual.shape[-2], actual.shape[-1])
        bad &= ~np.broadcast_to(t, actual.shape)
        frac = float(np.mean(tainted))
        assert frac <= max_taint_frac, f"{what}: {frac:.3%} of queries exempted as near-ties (limit {max_taint_frac:.1%})"
        del C
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.size} outside tolerance; max abs err "
                           f"{err[bad].max():.3e} (overall max {err.max():.3e})")
    return float(err.max())


def check_hard(actual, soft_expected, tainted=None, what="", tie=TIE):
    """actual int [.., sr, sr]; soft_expected [.., C, sr, sr] float64: argmax must agree except where
    the top-2 margin is < tie or the query is tainted."""
    soft = np.asarray(soft_expected, np.float64)
    exp = soft.argmax(axis=-3)
    srt = np.sort(soft, axis=-3)
    margin = srt[..., -1, :, :] - srt[..., -2, :, :] if soft.shape[-3] > 1 else np.ones_like(exp, float)
    diff = np.asarray(actual) != exp
    ok = margin < tie
    if tainted is not None:
        ok |= np.asarray(tainted).reshape(exp.shape)
    assert not (diff & ~ok).any(), f"{what}: {int((diff & ~ok).sum())} hard-label mismatches outside near-ties"
    return int(diff.sum())
