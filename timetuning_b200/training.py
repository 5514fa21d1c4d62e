"""Training-side callers of the hot path (§8f items 1-2), CUDA-backed.

Drop-in methods for the reference's ``TimeT`` (time_tuning.py:80-302).  ``install(..., fast_get_loss=True)`` binds
them over ``time_tuning.TimeT.get_scores`` / ``TimeT.get_loss``; nothing in the reference's source is edited.

    get_scores      TimeT.get_scores  time_tuning.py:195-217   incl. the feature-queue rows (:207-211)
    fast_get_loss   TimeT.get_loss    time_tuning.py:224-302   the FF / Sinkhorn part (:263-302) in one batched pass

What ``fast_get_loss`` does differently from the reference, with an identical loss value:
  * Only results that reach the loss are computed.  The reference runs four ``get_scores`` with the EMA teacher on
    (two without) and throws half of each away (:265-266, :272-273 keep ``[0]`` or ``[1]`` only), and neither
    ``target_batch_q`` nor the source scores are ever used by the loss (:277-301).  Here: ONE Sinkhorn (the assignment of
    the source frame: teacher features x teacher prototypes when the teacher is set) and ONE cosine-score matrix with
    autograd (student target frame).
  * The per-clip Python loop over ``make_seg_maps`` (:277-296, one device->host bounce per clip, :456) is one batched
    Feature-Forwarding call; the last-frame ``argmax`` is fused into it (hard labels int64 [bs, sr, sr]).
  * The cross entropy is evaluated for all clips at once (per-clip means averaged over the batch = the reference's
    ``batch_loss / bs``).
  * ONE student backbone forward instead of two (``_extract``): the tokens feed the head with autograd
    and, detached, the Feature-Forwarding, when the extractor has the reference's ``get_features`` / ``head`` structure and
    no active dropout.
PyTorch keeps what needs autograd (feature extractor, student scores, cross entropy).
"""
from __future__ import annotations

import sys

import torch
import torch.nn.functional as F

from . import ops


# --------------------------------------------------------------------------- cosine scores with autograd
class _CosineScores(torch.autograd.Function):
    """scores = F.normalize(x, dim=-1) @ prototypes.t()  (time_tuning.py:136-140).  Forward on the tensor cores
    (timet_cosine_scores, fp16 hi/lo split, ~1e-7 from fp32); backward = two torch matmuls + the normalisation Jacobian."""

    @staticmethod
    def forward(ctx, x, prototypes):
        scores = ops.cosine_scores(x, prototypes)
        ctx.save_for_backward(x, prototypes)
        return scores

    @staticmethod
    def backward(ctx, grad):
        x, prototypes = ctx.saved_tensors
        grad = grad.contiguous()
        gx = gp = None
        norm = x.norm(dim=-1, keepdim=True).clamp_min(1e-12)             # F.normalize eps
        xn = x / norm
        if ctx.needs_input_grad[1]:
            gp = grad.t() @ xn                                            # d scores / d prototypes
        if ctx.needs_input_grad[0]:
            gn = grad @ prototypes                                        # d scores / d x_hat
            gx = (gn - xn * (gn * xn).sum(dim=-1, keepdim=True)) / norm   # through x_hat = x / ||x||
        return gx, gp


def cosine_scores_autograd(x: torch.Tensor, prototypes: torch.Tensor) -> torch.Tensor:
    """Differentiable form of TimeT.get_feature_prototype_similarity (time_tuning.py:130-141) for CUDA float32 inputs."""
    return _CosineScores.apply(x.contiguous(), prototypes.contiguous())


# --------------------------------------------------------------------------- reference module globals
def _ref_module(model):
    """The reference module the model's class lives in (time_tuning): `world_size` is a module global there (:75,
    :511-512) and `apply_attention_mask` is imported into it (:50)."""
    return sys.modules.get(type(model).__module__)


def _world_size(model) -> int:
    mod = _ref_module(model)
    return int(getattr(mod, "world_size", 1)) if mod is not None else 1


def _similarity(model, x, use_teacher, tensor_cores):
    """get_feature_prototype_similarity (:130-141); `tensor_cores` only for the branches that need no gradient."""
    protos = model.teacher_prototypes if use_teacher else model.prototypes
    if tensor_cores and x.is_cuda and not (torch.is_grad_enabled() and (x.requires_grad or protos.requires_grad)):
        return ops.cosine_scores(x, protos)
    return torch.mm(F.normalize(x, dim=-1, p=2), protos.t())


def _queue_rows(model, dim):
    """The feature-queue rows the reference appends to the score matrix (:206-211), or None."""
    q = getattr(model, "queue", None)
    if q is None or int(q[-1].count_nonzero()) == 0:      # the reference's own (synchronising) fullness test, :207
        return None
    return q.view(-1, dim)


# --------------------------------------------------------------------------- TimeT.get_scores
def get_scores(self, features, epsilon, sinkhorn_iterations, use_teacher=False):
    """Drop-in for TimeT.get_scores (time_tuning.py:195-217): features [bs, N, dim] -> (batch_q [bs, N, K],
    batch_scores [bs, N, K]).  Scores stay in torch (they carry the gradient of the loss); exp + Sinkhorn-Knopp run
    fused in the CUDA kernel over the batch rows AND the feature-queue rows when the queue is filled (:207-211); the
    queue rows only shape the marginals and are sliced off again (:213)."""
    bs, num_patches, dim = features.shape
    flat = features.contiguous().view(bs * num_patches, dim)
    batch_scores = self.get_feature_prototype_similarity(flat, use_teacher)
    scores = batch_scores
    queue = _queue_rows(self, dim)
    if queue is not None:
        scores = torch.cat([batch_scores, self.get_feature_prototype_similarity(queue, use_teacher)], dim=0)
    q = ops.sinkhorn_from_scores(scores.detach(), epsilon, sinkhorn_iterations, _world_size(self))
    batch_q = q[:bs * num_patches].view(bs, num_patches, -1)
    return batch_q, batch_scores.view(bs, num_patches, -1)


def assignment(self, features, epsilon, sinkhorn_iterations, use_teacher=False):
    """``get_scores(...)[0]`` without the autograd graph: tensor-core cosine scores -> fused exp + Sinkhorn."""
    bs, num_patches, dim = features.shape
    with torch.no_grad():
        flat = features.detach().contiguous().view(bs * num_patches, dim)
        scores = _similarity(self, flat, use_teacher, tensor_cores=True)
        queue = _queue_rows(self, dim)
        if queue is not None:
            scores = torch.cat([scores, _similarity(self, queue, use_teacher, tensor_cores=True)], dim=0)
        q = ops.sinkhorn_from_scores(scores, epsilon, sinkhorn_iterations, _world_size(self))
    return q[:bs * num_patches].view(bs, num_patches, -1)


# --------------------------------------------------------------------------- one backbone forward instead of two
def _stochastic(module) -> bool:
    """True if a forward pass of `module` draws random numbers (active Dropout / stochastic depth): two forward passes of
    the reference then differ, and sharing one would change its statistics."""
    for m in module.modules():
        if m.training and ((isinstance(m, torch.nn.modules.dropout._DropoutNd) and m.p > 0) or float(getattr(m, "drop_prob", 0) or 0) > 0):
            return True
    return False


def _extract(fe, frames, dedup):
    """(head features, attentions, backbone features) of `frames`.

    The reference calls the extractor twice per step (time_tuning.py:237-239): `fe(x)` and, under no_grad, `fe(x,
    use_head=False)` -- two full backbone forwards whose token outputs are identical (models.py:1070-1078: forward =
    get_features, then the head).  SURVEY.md §8f item 4: when the extractor exposes that structure (`get_features`, `head`)
    and is deterministic, ONE backbone forward serves both: the tokens feed the head (with autograd) and, detached, the
    Feature-Forwarding.  Anything else falls back to the reference's two calls."""
    if dedup and callable(getattr(fe, "get_features", None)) and hasattr(fe, "head") and not _stochastic(fe):
        tokens, attentions = fe.get_features(frames)
        head_out = tokens
        if fe.head is not None:
            n, p, d = tokens.shape
            head_out = fe.head(tokens.reshape(n * p, d)).view(n, p, -1)               # models.py:1072-1077
        return head_out, attentions, tokens.detach()
    head_out, attentions = fe(frames)
    with torch.no_grad():
        backbone_out, _ = fe(frames, use_head=False)
    return head_out, attentions, backbone_out


# --------------------------------------------------------------------------- TimeT.get_loss
def fast_get_loss(self, x, annotations=None, n_last_frames=7, size_mask_neighborhood=6, topk=5, epsilon=0.05,
                  sinkhorn_iterations=10, mask_features=False, return_aux=False, dedup_backbone=True):
    """Drop-in for TimeT.get_loss (time_tuning.py:224-302), same arguments and defaults, same loss value.

    Feature extraction, attention masking and the queue update make the same calls in the same order as :231-261 (same RNG
    consumption, same queue content); everything from :263 on is the batched path described in the module docstring."""
    bs, fs = x.shape[:2]
    frames = x.flatten(0, 1)                                                          # [bs * fs, c, h, w]
    fe = self.feature_extractor
    sr = fe.spatial_resolution
    mod = _ref_module(self)
    has_teacher = self.teacher is not None

    def per_clip(t):                                                                  # [bs * fs, N, d] -> [bs, fs, N, d]
        return t.view(bs, fs, t.shape[-2], t.shape[-1])

    # ---- feature extraction in the order of :231-246 (teacher, then student head + backbone features)
    teacher_feats = None
    if has_teacher:
        t_out, t_attn = self.teacher(frames)
        teacher_feats = per_clip(t_out)
        if mask_features:
            teacher_feats, t_attn = mod.apply_attention_mask(teacher_feats, t_attn, sr)
    head_out, attentions, backbone_out = _extract(fe, frames, dedup_backbone)          # one backbone forward where possible
    features, backbone_features = per_clip(head_out), per_clip(backbone_out)
    num_patches, dim = features.shape[-2:]
    if mask_features:
        features, attentions = mod.apply_attention_mask(features, attentions, sr)
        attentions = attentions.view(bs, fs, sr, sr)

    # ---- feature queue (:250-261): push a random subset of the source-frame vectors at the front, oldest fall off the end.
    # One randperm of the same length as the reference draws, so the RNG stream and the queue content are identical.
    if self.queue is not None:
        src_vectors = (teacher_feats if has_teacher else features)[:, 0].reshape(-1, dim).detach()
        n_new = min(10 * bs, self.queue.size(0))
        pick = torch.randperm(src_vectors.size(0))[:n_new]
        self.queue.copy_(torch.cat([src_vectors[pick], self.queue[:self.queue.size(0) - n_new]], dim=0))

    # ---- assignment of the source frame (:263-268): the only Sinkhorn result the loss consumes
    if has_teacher:
        batch_q = assignment(self, teacher_feats[:, 0], epsilon, sinkhorn_iterations, use_teacher=True)
    else:
        batch_q = assignment(self, features[:, 0], epsilon, sinkhorn_iterations)
    # ---- student scores of the target frame (:269-275): the only scores the loss consumes; autograd stays in torch
    target_scores = self.get_feature_prototype_similarity(features[:, -1].contiguous().view(bs * num_patches, dim))
    # ---- Feature-Forwarding of every clip at once + fused last-frame argmax (:277-296)
    _, hard = ops.propagate_labels_batched(backbone_features, batch_q, n_last_frames, size_mask_neighborhood, topk)
    # ---- cross entropy (:294-301): per-clip mean over the sr x sr positions, averaged over the batch
    logits = target_scores.view(bs, sr, sr, -1).permute(0, 3, 1, 2) / 0.1
    per_pixel = F.cross_entropy(logits, hard, reduction="none")                       # [bs, sr, sr]
    if mask_features:
        per_pixel = per_pixel * attentions[:, -1]
    loss = per_pixel.reshape(bs, -1).mean(dim=1).sum() / bs
    if return_aux:
        return loss, dict(batch_q=batch_q, hard=hard)
    return loss
