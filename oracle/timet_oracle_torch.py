"""Torch-CPU restatement of the reference's dense FF + Sinkhorn path, used as the CPU BASELINE
("port") by bench.py and cross-checked against the golden fixtures.

TEST / MEASUREMENT INFRASTRUCTURE ONLY — never imported by timetuning_b200/.

Why it exists next to the numpy oracle: the reference's arithmetic runs in ATen (bmm, exp, topk,
float64 mm — /root/reference/mask_propagation.py:418-444, my_utils.py:246-274).  Timing a numpy
restatement would under-state the reference (numpy's partition/exp are slower than ATen's), so the
baseline leg issues the same ATen operator sequence per target frame, on the same dense
[ctx*N, N] affinity, with the same per-frame re-normalisation and float64 label GEMM, using all
host threads.  It cannot import /root/reference (absent on the GPU box).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


@torch.no_grad()
def sinkhorn(Q, nmb_iters):
    """my_utils.py:246-274, world_size 1.  Q [K, B] -> [B, K]."""
    Q = Q.detach().clone()
    Q /= Q.sum()
    K, B = Q.shape
    r = torch.full((K,), 1.0 / K)
    c = torch.full((B,), 1.0 / B)
    for _ in range(nmb_iters):
        Q *= (r / Q.sum(dim=1)).unsqueeze(1)
        Q *= (c / Q.sum(dim=0)).unsqueeze(0)
    return (Q / Q.sum(dim=0, keepdim=True)).t().float()


def window_mask(sr, radius):
    """0/1 [N, N] Chebyshev window (vectorised; the reference's Python loop is excluded from timing)."""
    idx = torch.arange(sr * sr)
    r, c = idx // sr, idx % sr
    return (((r[:, None] - r[None]).abs() <= radius) & ((c[:, None] - c[None]).abs() <= radius)).float()


@torch.no_grad()
def propagate_frame(radius, topk, sr, feat_tar, ctx_feats, ctx_segs, mask):
    """One target frame, the operator sequence of mask_propagation.py:418-444.
    feat_tar [N, D]; ctx_feats list of [D, N]; ctx_segs list of [1, C, sr, sr] float64."""
    n_ctx = len(ctx_feats)
    src = F.normalize(torch.stack(ctx_feats), dim=1, p=2)
    tar = F.normalize(feat_tar, dim=1, p=2).unsqueeze(0).repeat(n_ctx, 1, 1)
    aff = torch.exp(torch.bmm(tar, src) / 0.1)
    if radius > 0:
        aff = aff * mask
    aff = aff.transpose(2, 1).reshape(-1, sr * sr)
    kth = torch.topk(aff, dim=0, k=topk)[0].min(dim=0)[0]
    aff[aff < kth] = 0
    aff = aff / aff.sum(dim=0, keepdim=True)
    segs = torch.cat(ctx_segs)
    C = segs.shape[1]
    segs = segs.reshape(n_ctx, C, -1).transpose(2, 1).reshape(-1, C).T
    return torch.mm(segs.double(), aff.double()).reshape(1, C, sr, sr)


@torch.no_grad()
def propagate_labels(n_last, radius, topk, sr, feats, first_seg, mask):
    """Eval call pattern (mask_propagation.py:821): first_seg [1, C, H, W] at any resolution (nearest-resized, :456);
    returns the list of fs-1 maps [C, sr, sr] float64."""
    first = F.interpolate(first_seg.double(), size=(sr, sr), mode="nearest")            # :456
    return propagate_clip(n_last, radius, topk, sr, feats, first, mask)


@torch.no_grad()
def propagate_clip(n_last, radius, topk, sr, feats, first_seg, mask):
    """mask_propagation.py:448-496: feats [fs, N, D], first_seg [1, C, sr, sr] float64."""
    first_feat = feats[0].T
    fifo = []
    out = []
    for t in range(1, feats.shape[0]):
        seg = propagate_frame(radius, topk, sr, feats[t], [first_feat] + [p[0] for p in fifo],
                              [first_seg] + [p[1] for p in fifo], mask)
        if len(fifo) == n_last:
            fifo.pop(0)
        fifo.append((feats[t].T, seg))
        out.append(seg[0])
    return out


@torch.no_grad()
def ff_sinkhorn_step(head_src, head_tgt, backbone, prototypes, sr, n_last=7, radius=6, topk=5, epsilon=0.05,
                     iters=10, mask=None):
    """FF + Sinkhorn part of TimeT.get_loss (time_tuning.py:263-296), no teacher / queue.
    All inputs are CPU float32 tensors.  Returns (batch_q, target_q, hard int64 [bs,sr,sr], soft_last)."""
    bs, N, dh = head_src.shape
    if mask is None:
        mask = window_mask(sr, radius)
    qs = []
    for h in (head_src, head_tgt):
        scores = torch.mm(F.normalize(h.reshape(bs * N, dh), dim=-1, p=2), prototypes.t())
        qs.append(sinkhorn(torch.exp(scores / epsilon).t(), iters).view(bs, N, -1))
    hard, soft = [], []
    for i in range(bs):
        first = qs[0][i].view(sr, sr, -1).permute(2, 0, 1).unsqueeze(0).double()
        maps = propagate_clip(n_last, radius, topk, sr, backbone[i], first, mask)
        soft.append(maps[-1])
        hard.append(maps[-1].unsqueeze(0).argmax(dim=1)[0])
    return qs[0], qs[1], torch.stack(hard), torch.stack(soft)
