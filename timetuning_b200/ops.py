"""Host-side mirror of the reference's hot-path callables, backed by libtimet_b200.

Same names, argument meaning, return conventions and error behaviour (Python exceptions) as

    my_utils.sinkhorn                                   /root/reference/my_utils.py:246-274
    mask_propagation.restrict_neighborhood              /root/reference/mask_propagation.py:377-391
    mask_propagation.norm_mask                          /root/reference/mask_propagation.py:363-374
    mask_propagation.label_propagation                  /root/reference/mask_propagation.py:396-445
    mask_propagation.propagate_labels                   /root/reference/mask_propagation.py:448-496

plus the additive batched entry points the training fast path uses (all clips of a batch in one
call).  PyTorch is used for device memory and streams only; every numeric step runs in the
hand-written sm_100a kernels behind the C ABI.  No CPU fallback: tensors living on the CPU are
moved to the current CUDA device, and without a CUDA device every call raises.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi
from ._cabi import FFParams, FF_AUTO, FF_EXACT, FF_TC, SK_EXP, SK_SCORES, SinkhornOpts, check

AFF_TEMPERATURE = 0.1          # mask_propagation.py:422
_comm = {"handle": None, "world_size": 1, "rank": 0, "p2p": False}


# --------------------------------------------------------------------------- plumbing
def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("timetuning_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _to_cuda(x: torch.Tensor) -> torch.Tensor:
    return x if x.is_cuda else x.to(_device(), non_blocking=True)


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _workspace(nbytes: int, device) -> torch.Tensor:
    # torch's caching allocator returns >= 512-byte aligned blocks; over-allocate to align to 1024
    buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    off = (-buf.data_ptr()) % 1024
    return buf[off:off + nbytes]


def _spatial_resolution(model):
    """Duck-typed exactly like mask_propagation.py:402-405 / :452-455.  Returns (h, w): the reference's integer
    ``spatial_resolution`` means a square grid (:407); a (h, w) pair is accepted additively for non-square grids
    (DAVIS 480x854 frames), which restrict_neighborhood (:377) already allows."""
    fe = getattr(model, "feature_extractor", None)
    sr = fe.spatial_resolution if (fe is not None and hasattr(fe, "spatial_resolution")) else model.spatial_resolution
    if isinstance(sr, (tuple, list)):
        return int(sr[0]), int(sr[1])
    return int(sr), int(sr)


# --------------------------------------------------------------------------- Sinkhorn
@torch.no_grad()
def sinkhorn(Q: torch.Tensor, nmb_iters: int, world_size: int = 1) -> torch.Tensor:
    """Drop-in for my_utils.sinkhorn: Q is [K, B] = exp(scores/eps).t() (normally a transposed
    view of a row-major [B, K] tensor, time_tuning.py:164).  Returns float32 [B, K]; the caller's
    tensor is not modified (the reference clones, my_utils.py:249)."""
    if Q.dim() != 2:
        raise ValueError(f"sinkhorn expects a 2-D [K, B] tensor, got {tuple(Q.shape)}")
    dev_in = Q.device
    E = _to_cuda(Q.detach()).t()                       # [B, K]
    if E.dtype != torch.float32:
        E = E.float()
    E = E.contiguous()                                 # no copy for the reference's transposed view
    out = _sinkhorn_launch(E, SK_EXP, 1.0, nmb_iters, world_size)
    return out if dev_in.type == "cuda" else out.to(dev_in)


@torch.no_grad()
def sinkhorn_from_scores(scores: torch.Tensor, epsilon: float, nmb_iters: int, world_size: int = 1, out=None,
                         share_sm: bool = False) -> torch.Tensor:
    """Fused form of TimeT.find_optimal_assignment (time_tuning.py:157-168): exp(scores/eps) is
    evaluated inside every pass and never stored.  scores [B, K] -> Q [B, K] float32.

    out: optional CUDA float32 destination [n_blocks, block_rows, K] whose last two dims are contiguous (blocks may be
    strided, e.g. ``labels[:, 0]`` of the channel-last label tensor [bs, fs, N, K]: the assignment of clip b is written
    straight into frame 0, no copy before Feature-Forwarding).  share_sm: see timet_sinkhorn_ex."""
    if scores.dim() != 2:
        raise ValueError(f"scores must be [B, K], got {tuple(scores.shape)}")
    dev_in = scores.device
    S = _to_cuda(scores.detach()).float().contiguous()
    res = _sinkhorn_launch(S, SK_SCORES, float(epsilon), nmb_iters, world_size, out, share_sm)
    return res if dev_in.type == "cuda" else res.to(dev_in)


@torch.no_grad()
def sinkhorn_pair_from_scores(scores0, scores1, epsilon: float, nmb_iters: int, world_size: int = 1, out0=None, out1=None):
    """Two find_optimal_assignment calls of the same shape in ONE launch (the source and target assignment of a training
    step, time_tuning.py:268,275): the two problems run side by side on half of the SMs each, so their latency-bound
    reduction chains overlap.  Equal to two sinkhorn_from_scores calls within fp32 summation order, bit-reproducible.
    scores0/1 [B, K] CUDA float32 -> (Q0, Q1); out0 / out1 as in sinkhorn_from_scores."""
    S = [_to_cuda(x.detach()).float().contiguous() for x in (scores0, scores1)]
    if S[0].dim() != 2 or S[0].shape != S[1].shape:
        raise ValueError(f"two score matrices of the same shape [B, K] expected, got {tuple(S[0].shape)} and {tuple(S[1].shape)}")
    lib = _cabi.lib()
    B, K = S[0].shape
    comm = None
    if world_size > 1:
        if _comm["handle"] is None or _comm["world_size"] != world_size:
            raise RuntimeError(f"sinkhorn(world_size={world_size}) needs timetuning_b200.dist.init_comm() first")
        comm = _comm["handle"]
    outs, opts = [], []
    with torch.cuda.device(S[0].device):
        for o in (out0, out1):
            opt = SinkhornOpts(0, 0, 0, 0)
            if o is None:
                o = torch.empty((B, K), dtype=torch.float32, device=S[0].device)
            else:
                if not (o.is_cuda and o.dtype == torch.float32 and o.dim() == 3 and o.shape[2] == K and o.stride(2) == 1
                        and o.stride(1) == K and o.shape[0] * o.shape[1] == B):
                    raise ValueError(f"out must be CUDA float32 [n_blocks, block_rows, {K}] covering {B} rows, got {tuple(o.shape)}")
                opt.out_block_rows, opt.out_block_stride = o.shape[1], o.stride(0)
            outs.append(o)
            opts.append(opt)
        nbytes = 2 * int(lib.timet_sinkhorn_workspace_bytes(B, K))
        ws = _workspace(nbytes, S[0].device)
        check(lib.timet_sinkhorn_pair(_ptr(S[0]), _ptr(S[1]), B, K, SK_SCORES, float(epsilon), int(nmb_iters), int(world_size), comm,
                                      _ptr(outs[0]), C.byref(opts[0]), _ptr(outs[1]), C.byref(opts[1]), _ptr(ws), nbytes, _stream()),
              "sinkhorn_pair")
    return outs[0], outs[1]


def sinkhorn_mode(B: int, K: int) -> str:
    """How a Sinkhorn call of this shape runs: "resident" (one launch, all rows in shared memory), "hybrid" (one launch,
    part of the rows re-read every iteration) or "streaming" (one launch per pass)."""
    return ("streaming", "resident", "hybrid")[int(_cabi.lib().timet_sinkhorn_resident(int(B), int(K)))]


def sinkhorn_pair_mode(B: int, K: int) -> str:
    """How sinkhorn_pair_from_scores runs: "dual" (one launch, the two problems side by side) or "sequential"."""
    return ("sequential", "dual")[int(_cabi.lib().timet_sinkhorn_pair_mode(int(B), int(K)))]


def sinkhorn_is_resident(B: int, K: int) -> bool:
    return sinkhorn_mode(B, K) == "resident"


def _sinkhorn_launch(x, kind, eps, iters, world_size, out=None, share_sm=False):
    lib = _cabi.lib()
    B, K = x.shape
    comm = None
    if world_size > 1:
        if _comm["handle"] is None or _comm["world_size"] != world_size:
            raise RuntimeError(f"sinkhorn(world_size={world_size}) needs timetuning_b200.dist.init_comm() first "
                               f"(communicator world size: {_comm['world_size']})")
        comm = _comm["handle"]
    opts = SinkhornOpts(0, 0, 1 if share_sm else 0, 0)
    with torch.cuda.device(x.device):
        if out is None:
            out = torch.empty((B, K), dtype=torch.float32, device=x.device)
        else:
            if not (out.is_cuda and out.dtype == torch.float32 and out.dim() == 3 and out.shape[2] == K and out.stride(2) == 1
                    and out.stride(1) == K and out.shape[0] * out.shape[1] == B):
                raise ValueError(f"out must be CUDA float32 [n_blocks, block_rows, {K}] covering {B} rows with contiguous blocks, "
                                 f"got shape {tuple(out.shape)} strides {tuple(out.stride())}")
            opts.out_block_rows, opts.out_block_stride = out.shape[1], out.stride(0)
        nbytes = lib.timet_sinkhorn_workspace_bytes(B, K)
        ws = _workspace(nbytes, x.device)
        check(lib.timet_sinkhorn_ex(_ptr(x), B, K, kind, eps, int(iters), int(world_size), comm, _ptr(out), C.byref(opts), _ptr(ws),
                                    nbytes, _stream()), "sinkhorn")
    return out


@torch.no_grad()
def cosine_scores(x: torch.Tensor, prototypes: torch.Tensor) -> torch.Tensor:
    """No-grad form of TimeT.get_feature_prototype_similarity (time_tuning.py:130-141):
    F.normalize(x, dim=-1) @ prototypes.t() for x [B, dh], prototypes [K, dh] -> float32 [B, K], on the tensor
    cores with an fp16 hi/lo split (~1e-7 from the fp32 result).  Use torch for the branch that needs autograd."""
    dev_in = x.device
    xc = _to_cuda(x.detach()).float().contiguous()
    pc = _to_cuda(prototypes.detach()).float().contiguous()
    B, dh = xc.shape
    K = pc.shape[0]
    if pc.shape[1] != dh:
        raise ValueError(f"feature dim {dh} != prototype dim {pc.shape[1]}")
    lib = _cabi.lib()
    with torch.cuda.device(xc.device):
        out = torch.empty((B, K), dtype=torch.float32, device=xc.device)
        nbytes = int(lib.timet_cosine_scores_workspace_bytes(B, K, dh))
        ws = _workspace(nbytes, xc.device)
        check(lib.timet_cosine_scores(_ptr(xc), _ptr(pc), B, K, dh, _ptr(out), _ptr(ws), nbytes, _stream()), "cosine_scores")
    return out if dev_in.type == "cuda" else out.to(dev_in)


@torch.no_grad()
def cosine_scores_multi(xs, prototypes: torch.Tensor) -> torch.Tensor:
    """cosine_scores for several equally shaped feature blocks [B, dh] against the same prototypes in ONE GEMM launch
    (get_loss scores the source and the target frame, time_tuning.py:268,275).  Returns [len(xs) * B, K]."""
    xs = [_to_cuda(x.detach()).float().contiguous() for x in xs]
    pc = _to_cuda(prototypes.detach()).float().contiguous()
    B, dh = xs[0].shape
    if any(tuple(x.shape) != (B, dh) for x in xs) or pc.shape[1] != dh or not 1 <= len(xs) <= 4:
        raise ValueError("cosine_scores_multi: 1..4 blocks of identical shape [B, dh] and prototypes [K, dh] expected")
    K = pc.shape[0]
    lib = _cabi.lib()
    ptrs = (C.c_void_p * len(xs))(*[x.data_ptr() for x in xs])
    with torch.cuda.device(pc.device):
        out = torch.empty((len(xs) * B, K), dtype=torch.float32, device=pc.device)
        nbytes = int(lib.timet_cosine_scores_workspace_bytes(len(xs) * B, K, dh))
        ws = _workspace(nbytes, pc.device)
        check(lib.timet_cosine_scores_multi(ptrs, len(xs), B, _ptr(pc), K, dh, _ptr(out), _ptr(ws), nbytes, _stream()),
              "cosine_scores_multi")
    return out


# --------------------------------------------------------------------------- small routines
def restrict_neighborhood(h: int, w: int, size_mask_neighborhood: int) -> torch.Tensor:
    """Drop-in for mask_propagation.restrict_neighborhood: float32 [h*w, h*w] 0/1 mask (built on
    the GPU in one launch instead of a 4-deep Python loop; returned on the CUDA device)."""
    dev = _device()
    out = torch.empty((h * w, h * w), dtype=torch.float32, device=dev)
    check(_cabi.lib().timet_restrict_neighborhood(int(h), int(w), int(size_mask_neighborhood), _ptr(out), _stream()),
          "restrict_neighborhood")
    return out


def norm_mask(mask: torch.Tensor) -> torch.Tensor:
    """Drop-in for mask_propagation.norm_mask: per-channel min-max normalisation of [C, h, w]."""
    c, h, w = mask.shape
    dev_in = mask.device
    m = _to_cuda(mask)
    if m.dtype not in (torch.float32, torch.float64):
        m = m.float()
    m = m.contiguous()
    out = torch.empty_like(m)
    check(_cabi.lib().timet_norm_mask(_ptr(m), _ptr(out), c, h * w, m.element_size(), _stream()), "norm_mask")
    return out if dev_in.type == "cuda" else out.to(dev_in)


@torch.no_grad()
def upsample_argmax(maps: torch.Tensor, size) -> torch.Tensor:
    """Eval tail of mask_propagation.py:822-824, fused: bilinear up-sampling (align_corners=False) of
    maps [T, C, h, w] to `size` and arg max over channels -> int64 [T, size_h, size_w]."""
    out_h, out_w = (size, size) if isinstance(size, int) else size
    T, Cc, h, w = maps.shape
    dev_in = maps.device
    cl = _to_cuda(maps.detach()).float().permute(0, 2, 3, 1).contiguous()          # channel-last frames
    return _upsample_argmax_cl(cl.view(T, h * w, Cc), h, w, out_h, out_w, dev_in)


def _upsample_argmax_cl(frames_cl, h, w, out_h, out_w, dev_out=None):
    """frames_cl: float32 [T, h*w, C] (last two dims contiguous; frames may be strided)."""
    T, N, Cc = frames_cl.shape
    assert N == h * w and frames_cl.stride(2) == 1 and frames_cl.stride(1) == Cc
    out = torch.empty((T, out_h, out_w), dtype=torch.int64, device=frames_cl.device)
    with torch.cuda.device(frames_cl.device):
        check(_cabi.lib().timet_upsample_argmax(_ptr(frames_cl), T, h, w, Cc, out_h, out_w, frames_cl.stride(0), _ptr(out),
                                                _stream()), "upsample_argmax")
    return out if dev_out is None or dev_out.type == "cuda" else out.to(dev_out)


def one_hot_first_seg(annotation: torch.Tensor, n_dims: int) -> torch.Tensor:
    """mask_propagation.to_one_hot (:349-361) + unsqueeze(0), the first_seg of the eval call (:821), on the annotation's
    device: int labels [1, H, W] -> one-hot float32 [1, C, H, W].  (stays PyTorch.)"""
    return torch.nn.functional.one_hot(annotation[0].long(), int(n_dims)).permute(2, 0, 1).unsqueeze(0).float()


@torch.no_grad()
def propagate_labels_eval(n_last_frames, size_mask_neighborhood, topk, feats, first_seg, input_resolution,
                          engine=FF_AUTO, events=None, grid=None):
    """The whole eval call pattern of mask_propagation.py:821-824 for one video on the device: feats [fs, N, D]
    backbone features, first_seg [1, C, H, W] (one-hot annotation of frame 0) -> int64 [fs-1, R_h, R_w] hard predictions
    (labels stay channel-last float32 on the GPU; no float64 maps, no [T, C, R, R] tensor).
    grid=(h, w) for non-square patch grids (DAVIS 480x854 -> 60x106 at patch 8; additive, the reference is square-only);
    input_resolution may then be a (height, width) pair."""
    feats = _to_cuda(feats.detach()).float().contiguous()
    fs, N, D = feats.shape
    gh, gw = _grid_of(N, grid)
    seg = torch.nn.functional.interpolate(_to_cuda(first_seg.detach()).to(torch.float64), size=(gh, gw), mode="nearest")
    Cc = seg.shape[1]
    first = seg[0].reshape(Cc, N).t().float()
    labels, _ = propagate_labels_batched(feats.unsqueeze(0), first.unsqueeze(0), n_last_frames, size_mask_neighborhood,
                                         topk, engine=engine, want_hard=False, events=events, grid=(gh, gw))
    Rh, Rw = (int(input_resolution),) * 2 if isinstance(input_resolution, int) else (int(input_resolution[0]), int(input_resolution[1]))
    pred = _upsample_argmax_cl(labels[0, 1:], gh, gw, Rh, Rw)
    if events is not None:
        events["tail1"] = _record()
    return pred


def _record():
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    return ev


def _grid_of(N, grid):
    """(h, w) of the patch grid: the reference's square grid (mask_propagation.py:407) unless `grid` says otherwise."""
    if grid is not None:
        gh, gw = int(grid[0]), int(grid[1])
        if gh * gw != N:
            raise ValueError(f"grid {gh}x{gw} does not match {N} patches")
        return gh, gw
    sr = int(round(N ** 0.5))
    if sr * sr != N:
        raise ValueError(f"square patch grids only (mask_propagation.py:407) unless grid=(h, w) is given, got N={N}")
    return sr, sr


# --------------------------------------------------------------------------- Feature-Forwarding
class FFPlan:
    """Shapes + workspace of one Feature-Forwarding problem batch (all clips share a shape)."""

    def __init__(self, n_clips, n_frames, grid_h, grid_w, dim, n_channels, n_last_frames, radius, topk,
                 t_begin=1, temperature=AFF_TEMPERATURE, device=None):
        self.device = device or _device()
        self.params = FFParams(int(n_clips), int(n_frames), int(grid_h), int(grid_w), int(dim), int(n_channels),
                               int(n_last_frames), int(radius), int(topk), int(t_begin), float(temperature), 0)
        lib = _cabi.lib()
        self.nbytes = int(lib.timet_ff_workspace_bytes(C.byref(self.params)))
        if self.nbytes == 0:
            raise ValueError("invalid Feature-Forwarding shape: " + lib.timet_last_error().decode())
        self.workspace = _workspace(self.nbytes, self.device)
        self.kw = int(lib.timet_ff_slots(C.byref(self.params)))
        self.N = int(grid_h) * int(grid_w)

    @property
    def tc_supported(self) -> bool:
        return bool(_cabi.lib().timet_ff_tc_supported(C.byref(self.params)))

    @property
    def tc_executed_flops(self) -> float:
        """FLOPs the tensor-core kernel issues for one select() of this plan (0 if unsupported)."""
        return float(_cabi.lib().timet_ff_tc_executed_flops(C.byref(self.params)))

    def propagate(self, feats, labels, hard=None, engine=FF_AUTO, events=None):
        """feats fp32 [n_clips, n_frames, N, D]; labels fp32 [n_clips, n_frames, N, C] with frame(s)
        < t_begin filled; writes frames >= t_begin in place and hard int64 [n_clips, N] if given.
        events: optional dict that receives CUDA timing events around the stages (prep0/prep1/tc0/tc1/select1/gather0/
        gather1) -- the stages are then issued one by one; without it ONE C call runs all three."""
        p = self.params
        assert feats.is_cuda and feats.dtype == torch.float32 and feats.is_contiguous()
        assert labels.is_cuda and labels.dtype == torch.float32 and labels.is_contiguous()
        assert tuple(feats.shape) == (p.n_clips, p.n_frames, self.N, p.dim), tuple(feats.shape)
        assert tuple(labels.shape) == (p.n_clips, p.n_frames, self.N, p.n_channels), tuple(labels.shape)
        if hard is not None:
            assert hard.is_cuda and hard.dtype == torch.int64 and hard.numel() == p.n_clips * self.N
        with torch.cuda.device(self.device):
            if events is None:
                check(_cabi.lib().timet_ff_propagate(C.byref(p), int(engine), _ptr(feats), _ptr(labels), _ptr(hard),
                                                     _ptr(self.workspace), self.nbytes, _stream()), "ff_propagate")
            else:
                events["prep0"] = _record()
                self.prepare(feats)
                events["prep1"] = _record()
                if engine != FF_EXACT and self.tc_supported:
                    events["tc0"], events["tc1"] = _record(), _record()        # handles exist; the library re-records them
                    self.select_timed(engine, events["tc0"], events["tc1"])
                else:
                    self.select(engine)
                events["select1"] = events["gather0"] = _record()
                self.gather(labels, hard)
                events["gather1"] = _record()
        return labels

    def prepare(self, feats):
        """Stage 1.  The plan keeps a reference to `feats`: stage 2 reads the fp32 rows in place."""
        assert feats.is_cuda and feats.dtype == torch.float32 and feats.is_contiguous()
        self._feats = feats
        with torch.cuda.device(self.device):
            check(_cabi.lib().timet_ff_prepare(C.byref(self.params), _ptr(feats), _ptr(self.workspace), self.nbytes,
                                               _stream()), "ff_prepare")

    def _prepared(self):
        feats = getattr(self, "_feats", None)
        if feats is None:
            raise RuntimeError("FFPlan.select() before prepare(): the selection reads the feature rows given to prepare()")
        return feats

    def select(self, engine=FF_AUTO):
        with torch.cuda.device(self.device):
            check(_cabi.lib().timet_ff_select(C.byref(self.params), int(engine), _ptr(self._prepared()), _ptr(self.workspace),
                                              self.nbytes, _stream()), "ff_select")

    def select_timed(self, engine, ev_begin, ev_end):
        """select() with two torch.cuda.Event (either may be None) recorded by the library on the stream right before /
        right after the tensor-core nomination kernel.  They must have been recorded once before so that their handles
        exist.  Used to time the dominant kernel alone and to release a second stream as soon as that kernel is done."""
        hb = C.c_void_p(ev_begin.cuda_event) if ev_begin is not None else C.c_void_p(0)
        he = C.c_void_p(ev_end.cuda_event) if ev_end is not None else C.c_void_p(0)
        with torch.cuda.device(self.device):
            check(_cabi.lib().timet_ff_select_timed(C.byref(self.params), int(engine), _ptr(self._prepared()), _ptr(self.workspace),
                                                    self.nbytes, _stream(), hb, he), "ff_select_timed")

    def gather(self, labels, hard=None):
        with torch.cuda.device(self.device):
            check(_cabi.lib().timet_ff_gather(C.byref(self.params), _ptr(labels), _ptr(hard), _ptr(self.workspace),
                                              self.nbytes, _stream()), "ff_gather")

    def stats(self) -> dict:
        out = torch.zeros(8, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            check(_cabi.lib().timet_ff_stats(C.byref(self.params), _ptr(self.workspace), self.nbytes, _ptr(out),
                                             _stream()), "ff_stats")
        v = out.tolist()
        return dict(queries=v[0], selected=v[1], tie_queries=v[2], tc_candidates=v[3], redone_queries=v[4],
                    truncated_queries=v[5], wide_rows=v[6])

    def check_complete(self):
        """Raise if the last select() had to truncate a tie set (wide-row pool exhausted): the result would differ
        from the reference, which keeps every key tied with the k-th affinity (mask_propagation.py:432-436)."""
        n = self.stats()["truncated_queries"]
        if n:
            raise RuntimeError(f"Feature-Forwarding: {n} queries have more exact affinity ties than the wide-row pool holds "
                               "(degenerate features, e.g. constant rows); the result would not match the reference")

    def wide_row(self, offset, n):
        """(weights [n], keys [n]) of a wide row (counts[i] = -n, keys[i, 0] = offset in `selection`)."""
        w = torch.empty((n,), dtype=torch.float32, device=self.device)
        k = torch.empty((n,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(_cabi.lib().timet_ff_export_wide(C.byref(self.params), _ptr(self.workspace), self.nbytes, int(offset), int(n),
                                                   _ptr(w), _ptr(k), _stream()), "ff_export_wide")
        return w, k

    def selection(self, clip, t):
        """(weights [N, kw] fp32, keys [N, kw] int32 = frame*N+patch or -1, counts [N] int32; counts[i] = -n marks a
        wide row of n > kw entries whose pool offset is keys[i, 0], see wide_row)."""
        w = torch.empty((self.N, self.kw), dtype=torch.float32, device=self.device)
        k = torch.empty((self.N, self.kw), dtype=torch.int32, device=self.device)
        c = torch.empty((self.N,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(_cabi.lib().timet_ff_export_selection(C.byref(self.params), _ptr(self.workspace), self.nbytes,
                                                        int(clip), int(t), _ptr(w), _ptr(k), _ptr(c), _stream()),
                  "ff_export_selection")
        return w, k, c


_plan_cache: dict = {}


def _plan(*key, device):
    k = (*key, device)
    pl = _plan_cache.get(k)
    if pl is None:
        if len(_plan_cache) > 8:
            _plan_cache.clear()
        pl = _plan_cache[k] = FFPlan(*key, device=device)
    return pl


@torch.no_grad()
def propagate_labels_batched(feats, first_labels, n_last_frames=7, size_mask_neighborhood=6, topk=5,
                             engine=FF_AUTO, want_hard=True, check=False, events=None, grid=None):
    """Additive fast entry: every clip of a batch in one call (replaces the per-clip Python loop of
    TimeT.get_loss, time_tuning.py:277-296).

    feats [bs, fs, N, D] float32 backbone features; first_labels [bs, N, C] (Sinkhorn Q of frame 0,
    channel-last as get_scores returns it).  Returns (labels [bs, fs, N, C] float32 with frame 0 =
    first_labels, hard int64 [bs, h, w] = argmax of the last frame, or None).
    grid=(h, w): non-square patch grids (additive; the reference supports square grids only, :407)."""
    feats = _to_cuda(feats).float().contiguous()
    first_labels = _to_cuda(first_labels).float()
    bs, fs, N, D = feats.shape
    gh, gw = _grid_of(N, grid)
    Cc = first_labels.shape[-1]
    plan = _plan(bs, fs, gh, gw, D, Cc, n_last_frames, size_mask_neighborhood, topk, device=feats.device)
    labels = torch.empty((bs, fs, N, Cc), dtype=torch.float32, device=feats.device)
    labels[:, 0] = first_labels.reshape(bs, N, Cc)
    hard = torch.empty((bs, N), dtype=torch.int64, device=feats.device) if want_hard else None
    plan.propagate(feats, labels, hard, engine, events)
    if check:               # one small device->host read: only the drop-in shims pay for it by default
        plan.check_complete()
    return labels, (hard.view(bs, gh, gw) if want_hard else None)


@torch.no_grad()
def propagate_labels(n_last_frames, size_mask_neighborhood, topk, model, frame_list, first_seg, features_exist=False):
    """Drop-in for mask_propagation.propagate_labels (:448-496).

    frame_list [fs, N, D] features (features_exist=True) or images [fs, 3, H, W]; first_seg
    [1, C, H, W].  Returns a list of fs-1 tensors [C, sr, sr], float64 like the reference (:443,:456),
    on the features' device."""
    gh, gw = _spatial_resolution(model)
    if features_exist:
        feats = frame_list
    else:   # the reference runs the backbone frame by frame (:411,:467); same call, batched by frame here
        feats = torch.stack([model(fr.unsqueeze(0), use_head=False)[0].squeeze() for fr in frame_list])
    dev_out = feats.device
    feats = _to_cuda(feats.detach()).float().contiguous()
    fs, N, D = feats.shape
    if N != gh * gw:
        raise ValueError(f"features have {N} patches but spatial_resolution is {gh}x{gw}")
    seg = _to_cuda(first_seg.detach()).to(torch.float64)
    seg = torch.nn.functional.interpolate(seg, size=(gh, gw), mode="nearest")            # :456
    Cc = seg.shape[1]
    first = seg[0].reshape(Cc, N).t().float()                                            # channel-last [N, C]
    labels, _ = propagate_labels_batched(feats.unsqueeze(0), first.unsqueeze(0), n_last_frames,
                                         size_mask_neighborhood, topk, want_hard=False, check=True, grid=(gh, gw))
    out = labels[0, 1:].permute(0, 2, 1).reshape(fs - 1, Cc, gh, gw).to(torch.float64)
    if dev_out.type != "cuda":
        out = out.to(dev_out)
    return [out[i] for i in range(fs - 1)]


@torch.no_grad()
def label_propagation(size_mask_neighborhood, topk, model, frame_tar, list_frame_feats, list_segs,
                      mask_neighborhood=None, features_exist=False):
    """Drop-in for mask_propagation.label_propagation (:396-445): one target frame against an
    explicit context list.  frame_tar [N, D] features (or an image if features_exist=False);
    list_frame_feats: ctx x [D, N] un-normalised; list_segs: ctx x [1, C, h, w].
    Returns (seg_tar [1, C, h, w] float64, feat_tar [D, N], mask_neighborhood).
    mask_neighborhood is only passed through (its content is assumed to be
    restrict_neighborhood(h, w, size_mask_neighborhood), as propagate_labels builds it)."""
    gh, gw = _spatial_resolution(model)
    if features_exist:
        features = frame_tar
    else:
        features, _ = model(frame_tar.unsqueeze(0), use_head=False)
    features = features.squeeze()
    return_feat_tar = features.T
    dev_out = features.device
    ncontext = len(list_frame_feats)
    N, D = features.shape
    ctx = torch.stack([_to_cuda(f.detach()).float().t() for f in list_frame_feats])      # [ctx, N, D]
    feats = torch.cat([ctx, _to_cuda(features.detach()).float().unsqueeze(0)]).contiguous()
    segs = torch.cat([_to_cuda(s.detach()) for s in list_segs])                          # [ctx, C, h, w]
    Cc = segs.shape[1]
    labels = torch.empty((1, ncontext + 1, N, Cc), dtype=torch.float32, device=feats.device)
    labels[0, :ncontext] = segs.reshape(ncontext, Cc, N).permute(0, 2, 1).float()
    # contexts of target `ncontext` = frame 0 + the n_last previous frames: n_last = ncontext - 1 selects all of them
    # (and keeps the reference's 8-context call inside the tensor-core engine's n_last <= 7)
    plan = FFPlan(1, ncontext + 1, gh, gw, D, Cc, max(ncontext - 1, 1), size_mask_neighborhood, topk,
                  t_begin=ncontext, device=feats.device)
    plan.propagate(feats.unsqueeze(0), labels, None, FF_AUTO)
    plan.check_complete()
    seg_tar = labels[0, ncontext].t().reshape(1, Cc, gh, gw).to(torch.float64)
    if size_mask_neighborhood > 0 and mask_neighborhood is None:                         # :424-428
        mask_neighborhood = restrict_neighborhood(gh, gw, size_mask_neighborhood).unsqueeze(0).expand(ncontext, -1, -1)
    if dev_out.type != "cuda":
        seg_tar = seg_tar.to(dev_out)
    return seg_tar, return_feat_tar, mask_neighborhood
