"""Run the UNMODIFIED reference's FF + Sinkhorn path on given features.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/ref_loader.py): used by oracle/make_golden.py, by
tests/test_gpu_reference_dropin.py and by bench.py's reference arm (``--impl reference``, ``cpu_baseline``).

The reference's ``TimeT`` (time_tuning.py:80-302) wraps a ``models.FeatureExtractor``; here it wraps ``StoredFE``,
which returns stored tensors through the same ``forward(x, use_head)`` signature (models.py:1070-1078), so that
``TimeT.get_loss`` / ``get_scores`` / ``make_seg_maps`` run their own code on synthetic features without a ViT.
"""
from __future__ import annotations

import torch

import ref_loader


class StoredFE(torch.nn.Module):
    """Feature extractor returning stored tensors (head [bs,fs,N,dh], backbone [bs,fs,N,D])."""

    def __init__(self, head, backbone, sr):
        super().__init__()
        self.head_feats, self.backbone_feats = head, backbone
        self.spatial_resolution, self.feature_dim = sr, head.shape[-1]
        self.dummy = torch.nn.Parameter(torch.zeros(1, device=head.device))

    def forward(self, x, use_head=True):
        f = self.head_feats if use_head else self.backbone_feats
        return f.reshape(-1, f.shape[-2], f.shape[-1]) + 0 * self.dummy, None


def build_timet(head, backbone, prototypes, sr):
    """Reference TimeT over stored features.  head [bs,fs,N,dh], backbone [bs,fs,N,D], prototypes [K,dh] (torch)."""
    mu, mp, tt, _ = ref_loader.load()
    fe = StoredFE(head, backbone, sr)
    model = tt.TimeT(fe, prototypes.shape[0], prototype_init=prototypes.clone())
    tt.world_size = 1
    mp.mask_neighborhood = None                      # module-global cache (mask_propagation.py:85,473-476)
    return model.to(head.device)


def ff_sinkhorn_step(model, head_src, head_tgt, backbone, n_last=7, radius=6, topk=5, epsilon=0.05, iters=10):
    """The FF + Sinkhorn part of TimeT.get_loss (time_tuning.py:263-296, no teacher / queue) through the reference's own
    methods: get_scores x2, then the per-clip make_seg_maps loop and the last-frame argmax.
    Returns (batch_q, target_q, hard [bs,sr,sr] int64, soft_last [bs,K,sr,sr] float64)."""
    with torch.no_grad():
        batch_q, _ = model.get_scores(head_src, epsilon, iters)                       # :268
        target_q, _ = model.get_scores(head_tgt, epsilon, iters)                      # :275
        hard, soft = [], []
        for i in range(backbone.shape[0]):                                            # :277
            maps = model.make_seg_maps(batch_q[i], backbone[i], n_last, radius, topk, features_exist=True)   # :285
            soft.append(maps[-1])
            hard.append(maps[-1].unsqueeze(0).argmax(dim=1).long()[0])                # :296
    return batch_q, target_q, torch.stack(hard), torch.stack(soft)
