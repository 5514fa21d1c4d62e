"""The FF + Sinkhorn step of TimeT.get_loss (time_tuning.py:263-296, no teacher, no queue) on the
CUDA path: the unit of work bench.py times and smoke() checks.

    (head feats of frame 0 and -1, backbone feats, prototypes)
        -> cosine scores (tensor-core fp16 hi/lo split GEMM, no-grad branch; the autograd branch stays in torch)
        -> Sinkhorn x2 (fused exp, one pass per iteration)
        -> batched Feature-Forwarding of Q_source over every clip
        -> (Q_source, Q_target, last-frame hard labels)
"""
from __future__ import annotations

import torch

from . import ops


@torch.no_grad()
def ff_sinkhorn_step(head_src, head_tgt, backbone, prototypes, n_last_frames=7, size_mask_neighborhood=6, topk=5,
                     epsilon=0.05, sinkhorn_iterations=10, world_size=1, engine=ops.FF_AUTO):
    """head_src/head_tgt [bs, N, dh], backbone [bs, fs, N, D], prototypes [K, dh] — CUDA float32.
    Returns (batch_q [bs,N,K], target_q [bs,N,K], hard int64 [bs,sr,sr], labels [bs,fs,N,K])."""
    bs, N, dh = head_src.shape
    scores_src = ops.cosine_scores(head_src.reshape(bs * N, dh), prototypes)                  # :136-140 (no-grad branch)
    scores_tgt = ops.cosine_scores(head_tgt.reshape(bs * N, dh), prototypes)
    q_src = ops.sinkhorn_from_scores(scores_src, epsilon, sinkhorn_iterations, world_size)   # :164-165
    q_tgt = ops.sinkhorn_from_scores(scores_tgt, epsilon, sinkhorn_iterations, world_size)
    K = q_src.shape[1]
    labels, hard = ops.propagate_labels_batched(backbone, q_src.view(bs, N, K), n_last_frames,
                                                size_mask_neighborhood, topk, engine=engine)
    return q_src.view(bs, N, K), q_tgt.view(bs, N, K), hard, labels


class HostStepPipeline:
    """FF + Sinkhorn step fed from pinned HOST tensors (the e2e path of bench.py): the host->device copy of
    the backbone features (the bulk of the bytes) is chunked by clips on a copy stream and overlapped with
    the Sinkhorn calls and the Feature-Forwarding of the chunks that have already arrived; the hard labels
    are read back to pinned host memory at the end."""

    def __init__(self, bs, fs, N, D, dh, K, chunks=4, device=None):
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.bs, self.fs, self.N, self.D, self.dh, self.K = bs, fs, N, D, dh, K
        while bs % chunks:
            chunks -= 1
        self.chunks, self.cb = chunks, bs // chunks
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.d_head = torch.empty((2, bs, N, dh), dtype=torch.float32, device=self.device)
        self.d_backbone = torch.empty((bs, fs, N, D), dtype=torch.float32, device=self.device)
        self.labels = torch.empty((bs, fs, N, K), dtype=torch.float32, device=self.device)
        self.hard = torch.empty((bs, N), dtype=torch.int64, device=self.device)
        self.hard_host = torch.empty((bs, N), dtype=torch.int64).pin_memory()
        self.ev_head = torch.cuda.Event()
        self.ev_chunk = [torch.cuda.Event() for _ in range(chunks)]
        self.ev_done = torch.cuda.Event()

    @torch.no_grad()
    def run(self, head_src, head_tgt, backbone, prototypes, n_last_frames=7, size_mask_neighborhood=6, topk=5,
            epsilon=0.05, sinkhorn_iterations=10, world_size=1, engine=ops.FF_AUTO):
        """head_src/head_tgt [bs,N,dh], backbone [bs,fs,N,D]: pinned CPU float32; prototypes on the device.
        Returns (q_src, q_tgt on device, hard labels int64 [bs, sr, sr] in pinned host memory)."""
        bs, N, K, cb = self.bs, self.N, self.K, self.cb
        sr = int(round(N ** 0.5))
        main = torch.cuda.current_stream(self.device)
        self.copy_stream.wait_event(self.ev_done)            # previous step finished reading the device buffers
        with torch.cuda.stream(self.copy_stream):
            self.d_head[0].copy_(head_src, non_blocking=True)
            self.d_head[1].copy_(head_tgt, non_blocking=True)
            self.ev_head.record(self.copy_stream)
            for c in range(self.chunks):
                self.d_backbone[c * cb:(c + 1) * cb].copy_(backbone[c * cb:(c + 1) * cb], non_blocking=True)
                self.ev_chunk[c].record(self.copy_stream)
        main.wait_event(self.ev_head)
        scores = ops.cosine_scores(self.d_head.reshape(2 * bs * N, self.dh), prototypes)
        q_src = ops.sinkhorn_from_scores(scores[:bs * N], epsilon, sinkhorn_iterations, world_size)
        q_tgt = ops.sinkhorn_from_scores(scores[bs * N:], epsilon, sinkhorn_iterations, world_size)
        self.labels[:, 0] = q_src.view(bs, N, K)
        plan = ops._plan(cb, self.fs, sr, sr, self.D, K, n_last_frames, size_mask_neighborhood, topk, device=self.device)
        for c in range(self.chunks):
            main.wait_event(self.ev_chunk[c])
            plan.propagate(self.d_backbone[c * cb:(c + 1) * cb], self.labels[c * cb:(c + 1) * cb],
                           self.hard[c * cb:(c + 1) * cb], engine)
        self.hard_host.copy_(self.hard, non_blocking=True)
        self.ev_done.record(main)
        return q_src.view(bs, N, K), q_tgt.view(bs, N, K), self.hard_host.view(bs, sr, sr)
