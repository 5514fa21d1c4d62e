"""CPU oracle for TimeT's Feature-Forwarding + Sinkhorn-Knopp hot path (numpy).

TEST INFRASTRUCTURE ONLY.  This file is the *checker*, never the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  ``timetuning_b200/`` never does and
fails loudly when its CUDA library is missing.

It is an independent restatement (numpy, float32 where the reference is float32,
float64 where the reference is float64) of

* ``sinkhorn``              /root/reference/my_utils.py:246-274
* ``to_one_hot``            /root/reference/mask_propagation.py:349-361
* ``norm_mask``             /root/reference/mask_propagation.py:363-374
* ``restrict_neighborhood`` /root/reference/mask_propagation.py:377-391
* ``label_propagation``     /root/reference/mask_propagation.py:396-445
* ``propagate_labels``      /root/reference/mask_propagation.py:448-496
* ``TimeT.get_scores`` / ``find_optimal_assignment`` / ``make_seg_maps`` and the
  FF+Sinkhorn part of ``get_loss``  /root/reference/time_tuning.py:130-168,195-217,263-296

Pinning: the reference ships no tests or golden vectors for this path (SURVEY.md
§4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in
the build container by ``oracle/make_golden.py`` (which imports /root/reference
unmodified) and committed under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks every function here against those fixtures.

The arithmetic itself lives in PyTorch/ATen (pinned by the reference to
pytorch 1.13.0, environment.yml:158; fixtures were made with torch 2.11.0 CPU):
``F.normalize``, ``bmm``, ``exp``, ``topk``, ``mm`` (float64).  Their published
semantics are what is restated here.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
AFF_TEMPERATURE = F32(0.1)   # mask_propagation.py:422
NORMALIZE_EPS = F32(1e-12)   # torch.nn.functional.normalize default eps


# --------------------------------------------------------------------------- Sinkhorn
def sinkhorn(Q, nmb_iters, world_size=1, all_reduce=None):
    """Step-by-step restatement of my_utils.py:246-274.

    ``Q``: float32 ``[K, B]`` (prototypes x samples) = exp(scores/eps).T.
    ``all_reduce``: callable summing a numpy array over ranks in place (only used when
    ``world_size > 1``; my_utils.py:252,261,272).  Returns float32 ``[B, K]``.
    """
    Q = np.array(Q, dtype=F32, copy=True)                       # :249 clone
    sum_Q = Q.sum(dtype=F32)                                    # :250
    if world_size > 1:
        sum_Q = all_reduce(np.asarray(sum_Q, dtype=F32).reshape(1))[0]   # :252
    Q /= sum_Q                                                  # :253
    K, B = Q.shape
    r = np.ones(K, dtype=F32) / F32(K)                          # :256
    c = np.ones(B, dtype=F32) / F32(B * world_size)             # :257
    if world_size > 1:
        curr_sum = all_reduce(Q.sum(axis=1, dtype=F32))         # :260-261
    for _ in range(nmb_iters):
        u = curr_sum if world_size > 1 else Q.sum(axis=1, dtype=F32)     # :264-267
        Q *= (r / u)[:, None]                                   # :268
        Q *= (c / Q.sum(axis=0, dtype=F32))[None, :]            # :269
        if world_size > 1:
            curr_sum = all_reduce(Q.sum(axis=1, dtype=F32))     # :271-272
    return np.ascontiguousarray((Q / Q.sum(axis=0, keepdims=True, dtype=F32)).T.astype(F32))  # :274


def sinkhorn_scaling(scores, epsilon, nmb_iters, world_size=1, all_reduce=None, dtype=np.float64):
    """Scaling-vector form of find_optimal_assignment
    (time_tuning.py:157-168) + sinkhorn (my_utils.py:246-274).  Algebraically identical;
    this is the algorithm the CUDA kernels implement.  ``scores``: ``[B, K]`` local rows.
    In float64 it is the high-precision explainer used for near-tie margins of argmax(Q).
    """
    S = np.asarray(scores, dtype=dtype)
    B, K = S.shape
    E = np.exp(S / dtype(epsilon))
    r = dtype(1.0) / dtype(K)
    c = dtype(1.0) / dtype(B * world_size)
    b = np.ones(B, dtype=dtype)
    a = np.ones(K, dtype=dtype)
    for _ in range(nmb_iters):
        R = (E * b[:, None]).sum(axis=0)
        if world_size > 1:
            R = all_reduce(R)
        a = r / R
        b = c / (E * a[None, :]).sum(axis=1)
    Qn = E * a[None, :]
    return Qn / Qn.sum(axis=1, keepdims=True)


def find_optimal_assignment(scores, epsilon, sinkhorn_iterations, world_size=1, all_reduce=None):
    """time_tuning.py:157-168: q = exp(scores/eps).t(); sinkhorn(q, iters, world_size)."""
    q = np.exp(np.asarray(scores, dtype=F32) / F32(epsilon)).T
    return sinkhorn(q, sinkhorn_iterations, world_size, all_reduce)


# --------------------------------------------------------------------------- small helpers
def to_one_hot(y, n_dims=None):
    """mask_propagation.py:349-361: int labels [1,h,w] -> one-hot float32 [C,h,w]."""
    y = np.asarray(y)
    _, h, w = y.shape
    flat = y.astype(np.int64).reshape(-1)
    if n_dims is None:
        n_dims = int(flat.max()) + 1
    out = np.zeros((flat.size, n_dims), dtype=F32)
    out[np.arange(flat.size), flat] = 1
    return np.ascontiguousarray(out.reshape(h, w, n_dims).transpose(2, 0, 1))


def norm_mask(mask):
    """mask_propagation.py:363-374: per-channel (m - min) / max(m - min) where channel max > 0."""
    mask = np.asarray(mask)
    out = np.zeros_like(mask)
    for ch in range(mask.shape[0]):
        m = mask[ch]
        if m.max() > 0:
            m = m - m.min()
            with np.errstate(invalid="ignore", divide="ignore"):
                out[ch] = m / m.max()
    return out


def restrict_neighborhood(h, w, size_mask_neighborhood):
    """mask_propagation.py:377-391: mask[i*w+j, p*w+q] = 1 iff |i-p|<=s and |j-q|<=s."""
    s = size_mask_neighborhood
    ri = np.arange(h)[:, None]
    rj = np.arange(w)[:, None]
    row_ok = (np.abs(ri - ri.T) <= s)          # [h,h]
    col_ok = (np.abs(rj - rj.T) <= s)          # [w,w]
    mask = row_ok[:, None, :, None] & col_ok[None, :, None, :]
    return mask.reshape(h * w, h * w).astype(F32)


def nearest_resize(x, out_h, out_w):
    """F.interpolate(mode='nearest') on [..., H, W]: src = floor(dst * in / out)
    (mask_propagation.py:456)."""
    x = np.asarray(x)
    H, W = x.shape[-2:]
    ri = np.minimum((np.arange(out_h) * (H / out_h)).astype(np.int64), H - 1)
    ci = np.minimum((np.arange(out_w) * (W / out_w)).astype(np.int64), W - 1)
    if H == out_h and W == out_w:
        return x
    return x[..., ri[:, None], ci[None, :]]


def l2_normalize_rows(x):
    """F.normalize(x, dim=-1, p=2) in float32 (mask_propagation.py:418-419)."""
    x = np.asarray(x, dtype=F32)
    n = np.sqrt((x * x).sum(axis=-1, keepdims=True, dtype=F32))
    return x / np.maximum(n, NORMALIZE_EPS)


# --------------------------------------------------------------------------- FF core
def label_propagation(size_mask_neighborhood, topk, spatial_resolution, frame_tar, list_frame_feats,
                      list_segs, mask_neighborhood=None):
    """Dense restatement of mask_propagation.py:396-445 for ``features_exist=True``.

    frame_tar ``[N, D]`` float32; list_frame_feats: ctx x ``[D, N]`` (unnormalised);
    list_segs: ctx x ``[1, C, h, w]`` float64.  Returns (seg_tar ``[1,C,h,w]`` float64,
    feat_tar ``[D, N]``, mask).
    """
    h = w = spatial_resolution
    feats = np.asarray(frame_tar, dtype=F32)
    return_feat_tar = feats.T
    ncontext = len(list_frame_feats)
    feat_sources = np.stack([np.asarray(f, dtype=F32) for f in list_frame_feats])   # [ctx, D, N]
    feat_tar = l2_normalize_rows(feats)                                            # :418
    n = np.sqrt((feat_sources * feat_sources).sum(axis=1, keepdims=True, dtype=F32))
    feat_sources = feat_sources / np.maximum(n, NORMALIZE_EPS)                      # :419 (dim=1 = D)
    aff = np.exp(np.matmul(feat_tar[None], feat_sources) / AFF_TEMPERATURE)         # :422 [ctx, Nq, Nk]
    if size_mask_neighborhood > 0:
        if mask_neighborhood is None:
            mask_neighborhood = restrict_neighborhood(h, w, size_mask_neighborhood)
        m = mask_neighborhood if mask_neighborhood.ndim == 3 else mask_neighborhood[None]
        aff = aff * m                                                               # :429
    aff = np.ascontiguousarray(aff.transpose(0, 2, 1)).reshape(-1, h * w)           # :431 [ctx*Nk, Nq]
    kth = np.partition(aff, aff.shape[0] - topk, axis=0)[aff.shape[0] - topk]       # :432-433
    aff = np.where(aff < kth[None, :], F32(0), aff)                                 # :434
    aff = aff / aff.sum(axis=0, keepdims=True, dtype=F32)                           # :436
    segs = np.concatenate([np.asarray(s) for s in list_segs])                       # :440 [ctx,C,h,w]
    C = segs.shape[1]
    segs = segs.reshape(ncontext, C, -1).transpose(0, 2, 1).reshape(-1, C).T        # :442 [C, ctx*N]
    seg_tar = segs.astype(np.float64) @ aff.astype(np.float64)                      # :443
    return seg_tar.reshape(1, C, h, w), return_feat_tar, mask_neighborhood


def propagate_labels(n_last_frames, size_mask_neighborhood, topk, spatial_resolution, frame_list, first_seg):
    """mask_propagation.py:448-496 with ``features_exist=True``.

    frame_list ``[fs, N, D]`` float32 features; first_seg ``[1, C, H, W]``.
    Returns a list of fs-1 arrays ``[C, sr, sr]`` float64.
    """
    sr = spatial_resolution
    first_seg = nearest_resize(np.asarray(first_seg, dtype=np.float64), sr, sr)     # :456
    frame_list = np.asarray(frame_list, dtype=F32)
    frame1_feat = frame_list[0].T                                                   # :468
    mask = restrict_neighborhood(sr, sr, size_mask_neighborhood) if size_mask_neighborhood > 0 else None
    que = []                                                                        # FIFO, maxlen n_last_frames
    out = []
    for cnt in range(1, frame_list.shape[0]):                                       # :478
        used_feats = [frame1_feat] + [p[0] for p in que]                            # :482
        used_segs = [first_seg] + [p[1] for p in que]                               # :483
        seg, feat_tar, mask = label_propagation(size_mask_neighborhood, topk, sr, frame_list[cnt],
                                                used_feats, used_segs, mask)
        if len(que) == n_last_frames:                                               # :488-489
            que.pop(0)
        que.append((feat_tar, seg))                                                 # :493
        out.append(seg[0])                                                          # :495
    return out


def context_frames(t, n_last_frames):
    """Context set of target frame t in reference order: frame 0, then the FIFO content
    (mask_propagation.py:460,482-493).  While the FIFO is not full, frame 0 is NOT in it, so
    contexts are {0} + {1..t-1}; afterwards {0} + {t-n_last..t-1}."""
    lo = max(1, t - n_last_frames)
    return [0] + list(range(lo, t))


# --------------------------------------------------------------------------- fp64 explainer
def ff_sparse(n_last_frames, radius, topk, sr, feats, first_seg, dtype=np.float64, teacher_segs=None):
    """Sparse (gather) restatement of FF in ``dtype`` precision with tie margins.

    feats ``[fs,N,D]``; first_seg ``[C,sr,sr]`` (already at sr x sr).
    teacher_segs: optional list (len fs) of ``[C,N]`` context labels to use instead of the
    propagated ones ("teacher-forced" per-frame parity).
    Returns dict(segs=[fs-1,C,sr,sr], margin=[fs-1,N] relative gap between the k-th and
    (k+1)-th largest affinity, nnz=[fs-1,N] selected keys per query).
    """
    feats = np.asarray(feats)
    fs, N, D = feats.shape
    fn = l2_normalize_rows(feats).astype(dtype)        # fp32 normalisation as the reference, then widen
    C = first_seg.shape[0]
    rows = np.arange(N) // sr
    cols = np.arange(N) % sr
    win = (np.abs(rows[:, None] - rows[None, :]) <= radius) & (np.abs(cols[:, None] - cols[None, :]) <= radius) \
        if radius > 0 else np.ones((N, N), dtype=bool)
    segs = [np.asarray(first_seg, dtype=np.float64).reshape(C, N)]
    margins, nnzs = [], []
    for t in range(1, fs):
        ctx = context_frames(t, n_last_frames)
        sim = np.stack([fn[t] @ fn[c].T for c in ctx])                   # [ctx, Nq, Nk]
        aff = np.exp(sim / dtype(0.1)) * win[None]
        aff = aff.transpose(1, 0, 2).reshape(N, -1)                      # [Nq, ctx*Nk]
        srt = np.sort(aff, axis=1)[:, ::-1]
        kth = srt[:, topk - 1]
        nxt = srt[:, topk] if srt.shape[1] > topk else np.zeros(N, dtype=dtype)
        margins.append((kth - nxt) / kth)
        wgt = np.where(aff >= kth[:, None], aff, 0)
        nnzs.append((wgt > 0).sum(axis=1))
        wgt = wgt / wgt.sum(axis=1, keepdims=True)
        src = np.concatenate([(teacher_segs[c] if teacher_segs is not None else segs[c]) for c in ctx], axis=1)
        segs.append(src.astype(np.float64) @ wgt.astype(np.float64).T)   # [C, Nq]
    return dict(segs=np.stack(segs[1:]).reshape(fs - 1, C, sr, sr),
                margin=np.stack(margins), nnz=np.stack(nnzs))


# --------------------------------------------------------------------------- training-side step
def cosine_scores(x, prototypes):
    """time_tuning.py:130-141: F.normalize(x) @ prototypes.T (float32)."""
    return l2_normalize_rows(x) @ np.asarray(prototypes, dtype=F32).T


def get_scores(features, prototypes, epsilon, sinkhorn_iterations, world_size=1, all_reduce=None):
    """time_tuning.py:195-217 without the optional feature queue: features [bs,N,dim]."""
    bs, N, dim = features.shape
    scores = cosine_scores(np.asarray(features, dtype=F32).reshape(bs * N, dim), prototypes)
    q = find_optimal_assignment(scores, epsilon, sinkhorn_iterations, world_size, all_reduce)
    return q.reshape(bs, N, -1), scores.reshape(bs, N, -1)


def ff_sinkhorn_step(head_src, head_tgt, backbone_feats, prototypes, sr, n_last_frames=7,
                     size_mask_neighborhood=6, topk=5, epsilon=0.05, sinkhorn_iterations=10,
                     world_size=1, all_reduce=None):
    """The FF+Sinkhorn part of TimeT.get_loss (time_tuning.py:263-296), no teacher, no queue:
    head_src/head_tgt [bs,N,256] (frames 0 and -1), backbone_feats [bs,fs,N,D].
    Returns (batch_q [bs,N,K], target_q [bs,N,K], hard labels int64 [bs,sr,sr] =
    argmax over channels of the last propagated frame, last-frame soft labels [bs,K,sr,sr] f64)."""
    batch_q, _ = get_scores(head_src, prototypes, epsilon, sinkhorn_iterations, world_size, all_reduce)
    target_q, _ = get_scores(head_tgt, prototypes, epsilon, sinkhorn_iterations, world_size, all_reduce)
    bs = backbone_feats.shape[0]
    hard, soft = [], []
    for i in range(bs):                                                              # :277
        first = batch_q[i].reshape(sr, sr, -1).transpose(2, 0, 1)[None]              # :144-146
        maps = propagate_labels(n_last_frames, size_mask_neighborhood, topk, sr, backbone_feats[i], first)
        soft.append(maps[-1])                                                        # :293
        hard.append(maps[-1].argmax(axis=0))                                         # :296
    return batch_q, target_q, np.stack(hard).astype(np.int64), np.stack(soft)
