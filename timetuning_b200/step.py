"""The FF + Sinkhorn step of TimeT.get_loss (time_tuning.py:263-296, no teacher, no queue) on the
CUDA path: the unit of work bench.py times and smoke() checks.

    (head feats of frame 0 and -1, backbone feats, prototypes)
        -> cosine scores (tensor-core fp16 hi/lo split GEMM, no-grad branch; the autograd branch stays in torch)
        -> Sinkhorn x2 (fused exp, one pass per iteration)
        -> batched Feature-Forwarding of Q_source over every clip
        -> (Q_source, Q_target, last-frame hard labels)

Stream choreography (StepRunner, overlap=True; OFF by default -- measured on a B200 it does not pay: the finalize and
gather kernels are bound by resident warps x loads in flight, and every warp slot / register the co-resident Sinkhorn
kernel takes slows them by as much as the overlap hides, profiles/r2_experiments.md).  The Sinkhorn chain does not depend on the affinity / top-k selection
and vice versa; only the label gather needs Q_source.  So the chain runs on a second, high-priority stream:

    main : prepare -> select (tcgen05 kernel, finalize) ------------------> [wait Q_source] -> gather
    side :            [wait tcgen05 kernel done] -> cosine scores -> Sinkhorn(source) -> Sinkhorn(target)

The resident Sinkhorn kernel is cooperative and wants every SM; the persistent tcgen05 kernel fills every SM's shared
memory.  Racing them would leave one half-resident, so the side stream is released by an event recorded right after
the tcgen05 kernel; it then co-resides (512-thread variant, half the register file) with the L2-bound finalize and gather
kernels, which leave the tensor pipe and most issue slots idle.
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
    fs, D = backbone.shape[1], backbone.shape[3]
    sr = int(round(N ** 0.5))
    runner = _runner(bs, fs, sr, D, dh, prototypes.shape[0], n_last_frames, size_mask_neighborhood, topk, epsilon,
                     sinkhorn_iterations, world_size, engine, backbone.device)
    q_src, q_tgt, hard = runner.run(head_src, head_tgt, backbone, prototypes)
    labels = runner.labels.clone()                 # the runner reuses its buffers: hand out copies
    return labels[:, 0], q_tgt.clone(), hard.clone(), labels


_runners: dict = {}


def _runner(*key):
    r = _runners.get(key)
    if r is None:
        if len(_runners) > 4:
            _runners.clear()
        *args, world_size, engine, device = key
        r = _runners[key] = StepRunner(*args, world_size=world_size, engine=engine, device=device)
    return r


class StepRunner:
    """Device-resident FF + Sinkhorn step with its buffers, plan and streams kept across calls."""

    def __init__(self, bs, fs, sr, D, dh, K, n_last_frames=7, size_mask_neighborhood=6, topk=5, epsilon=0.05,
                 sinkhorn_iterations=10, world_size=1, engine=ops.FF_AUTO, overlap=False, device=None):
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.bs, self.fs, self.sr, self.N, self.D, self.dh, self.K = bs, fs, sr, sr * sr, D, dh, K
        self.eps, self.iters, self.world_size, self.engine = float(epsilon), int(sinkhorn_iterations), int(world_size), engine
        self.plan = ops._plan(bs, fs, sr, sr, D, K, n_last_frames, size_mask_neighborhood, topk, device=self.device)
        self.labels = torch.empty((bs, fs, self.N, K), dtype=torch.float32, device=self.device)
        self.hard = torch.empty((bs, self.N), dtype=torch.int64, device=self.device)
        self.q_tgt = torch.empty((bs, self.N, K), dtype=torch.float32, device=self.device)
        self.sinkhorn_mode = ops.sinkhorn_mode(bs * self.N, K)
        self.sinkhorn_resident = self.sinkhorn_mode == "resident"
        self.sinkhorn_pair_mode = ops.sinkhorn_pair_mode(bs * self.N, K)
        # the choreography above needs the one-launch resident Sinkhorn kernel and the tensor-core engine's event hook
        self.overlap = bool(overlap) and self.sinkhorn_resident and engine != ops.FF_EXACT and self.plan.tc_supported
        with torch.cuda.device(self.device):
            self.side = torch.cuda.Stream(device=self.device, priority=-1) if self.overlap else None
            self.ev_tc_done = torch.cuda.Event()
            self.ev_tc_done.record()                   # materialise the cudaEvent_t handle (the library re-records it)
            self.ev_q_src = torch.cuda.Event()
            self.ev_side_done = torch.cuda.Event()
            self.ev_inputs = torch.cuda.Event()

    def _sinkhorn_chain(self, head_src, head_tgt, prototypes, events, share_sm):
        bs, N, K = self.bs, self.N, self.K
        if events is not None:
            events["scores0"] = ops._record()
        scores = ops.cosine_scores_multi([head_src.reshape(bs * N, self.dh), head_tgt.reshape(bs * N, self.dh)], prototypes)
        if events is not None:
            events["scores1"] = ops._record()
        # Q_source is written straight into frame 0 of the channel-last label tensor (time_tuning.py:144-147 without a copy)
        if share_sm:
            ops.sinkhorn_from_scores(scores[:bs * N], self.eps, self.iters, self.world_size, out=self.labels[:, 0], share_sm=True)
            self.ev_q_src.record()
            ops.sinkhorn_from_scores(scores[bs * N:], self.eps, self.iters, self.world_size, out=self.q_tgt, share_sm=True)
        else:       # both assignments in one resident launch: each hides the other's reduction latency
            ops.sinkhorn_pair_from_scores(scores[:bs * N], scores[bs * N:], self.eps, self.iters, self.world_size,
                                          out0=self.labels[:, 0], out1=self.q_tgt)
            self.ev_q_src.record()
        if events is not None:
            events["sinkhorn1"] = ops._record()

    @torch.no_grad()
    def run(self, head_src, head_tgt, backbone, prototypes, events=None):
        """head_src/head_tgt [bs, N, dh], backbone [bs, fs, N, D], prototypes [K, dh]: CUDA float32, contiguous.
        Returns (q_src [bs,N,K] = a view of labels[:, 0], q_tgt [bs,N,K], hard int64 [bs,sr,sr]).  The buffers are
        reused by the next call.  events: optional dict receiving CUDA timing events of the stages."""
        plan, engine = self.plan, self.engine
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream(self.device)
            if not self.overlap:
                self._sinkhorn_chain(head_src, head_tgt, prototypes, events, share_sm=False)
                plan.propagate(backbone, self.labels, self.hard, engine, events)
            else:
                self.ev_inputs.record(main)
                if events is not None:
                    events["prep0"] = ops._record()
                plan.prepare(backbone)
                ev_tc_done = self.ev_tc_done
                if events is not None:
                    events["prep1"] = ops._record()
                    events["tc0"], events["tc1"] = ops._record(), ops._record()
                    ev_tc_done = events["tc1"]
                    plan.select_timed(engine, events["tc0"], ev_tc_done)
                    events["select1"] = ops._record()
                else:
                    plan.select_timed(engine, None, ev_tc_done)     # recorded by the library right after the tcgen05 kernel
                with torch.cuda.stream(self.side):
                    self.side.wait_event(self.ev_inputs)       # inputs ready, previous step's gather finished with labels
                    self.side.wait_event(ev_tc_done)           # never race the persistent tcgen05 kernel for the SMs
                    self._sinkhorn_chain(head_src, head_tgt, prototypes, events, share_sm=True)
                    self.ev_side_done.record(self.side)
                main.wait_event(self.ev_q_src)
                if events is not None:
                    events["gather0"] = ops._record()
                plan.gather(self.labels, self.hard)
                if events is not None:
                    events["gather1"] = ops._record()
                main.wait_event(self.ev_side_done)             # q_tgt complete before the caller's stream goes on
        return self.labels[:, 0], self.q_tgt, self.hard.view(self.bs, self.sr, self.sr)


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
        self.q_tgt = torch.empty((bs, N, K), dtype=torch.float32, device=self.device)
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
        scores = ops.cosine_scores_multi([self.d_head[0].reshape(bs * N, self.dh), self.d_head[1].reshape(bs * N, self.dh)], prototypes)
        ops.sinkhorn_pair_from_scores(scores[:bs * N], scores[bs * N:], epsilon, sinkhorn_iterations, world_size,
                                      out0=self.labels[:, 0], out1=self.q_tgt)
        plan = ops._plan(cb, self.fs, sr, sr, self.D, K, n_last_frames, size_mask_neighborhood, topk, device=self.device)
        for c in range(self.chunks):
            main.wait_event(self.ev_chunk[c])
            plan.propagate(self.d_backbone[c * cb:(c + 1) * cb], self.labels[c * cb:(c + 1) * cb],
                           self.hard[c * cb:(c + 1) * cb], engine)
        self.hard_host.copy_(self.hard, non_blocking=True)
        self.ev_done.record(main)
        return self.labels[:, 0], self.q_tgt, self.hard_host.view(bs, sr, sr)
