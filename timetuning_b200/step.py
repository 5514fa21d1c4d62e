"""The FF + Sinkhorn step of TimeT.get_loss (time_tuning.py:263-296, no teacher, no queue) on the
CUDA path: the unit of work bench.py times and smoke() checks.

    (head feats of frame 0 and -1, backbone feats, prototypes)
        -> cosine scores (torch: library GEMM, needs autograd in training — SURVEY.md §2.2 X1)
        -> Sinkhorn x2 (fused exp, one pass per iteration)
        -> batched Feature-Forwarding of Q_source over every clip
        -> (Q_source, Q_target, last-frame hard labels)
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops


@torch.no_grad()
def ff_sinkhorn_step(head_src, head_tgt, backbone, prototypes, n_last_frames=7, size_mask_neighborhood=6, topk=5,
                     epsilon=0.05, sinkhorn_iterations=10, world_size=1, engine=ops.FF_AUTO):
    """head_src/head_tgt [bs, N, dh], backbone [bs, fs, N, D], prototypes [K, dh] — CUDA float32.
    Returns (batch_q [bs,N,K], target_q [bs,N,K], hard int64 [bs,sr,sr], labels [bs,fs,N,K])."""
    bs, N, dh = head_src.shape
    scores_src = F.normalize(head_src.reshape(bs * N, dh), dim=-1, p=2) @ prototypes.t()     # :136-140
    scores_tgt = F.normalize(head_tgt.reshape(bs * N, dh), dim=-1, p=2) @ prototypes.t()
    q_src = ops.sinkhorn_from_scores(scores_src, epsilon, sinkhorn_iterations, world_size)   # :164-165
    q_tgt = ops.sinkhorn_from_scores(scores_tgt, epsilon, sinkhorn_iterations, world_size)
    K = q_src.shape[1]
    labels, hard = ops.propagate_labels_batched(backbone, q_src.view(bs, N, K), n_last_frames,
                                                size_mask_neighborhood, topk, engine=engine)
    return q_src.view(bs, N, K), q_tgt.view(bs, N, K), hard, labels
