"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference).

TEST INFRASTRUCTURE ONLY; runs in the build container only (the reference cannot travel to
the GPU box).  Each fixture stores seeded inputs and the reference's own outputs, so the
numpy oracle (and through it the CUDA path) is pinned to the reference, not to itself.

    python oracle/make_golden.py            # rewrites every fixture (deterministic)

Reference entry points exercised (file:line in /root/reference):
  my_utils.sinkhorn                         my_utils.py:246-274   (world_size 1, and 2 under gloo)
  mask_propagation.restrict_neighborhood    mask_propagation.py:377-391
  mask_propagation.norm_mask                mask_propagation.py:363-374
  mask_propagation.to_one_hot               mask_propagation.py:349-361
  mask_propagation.label_propagation        mask_propagation.py:396-445
  mask_propagation.propagate_labels         mask_propagation.py:448-496
  time_tuning.TimeT.get_scores / make_seg_maps / get_loss   time_tuning.py:143-302
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402
from timetuning_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def save(name, **arrays):
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def gold_sinkhorn(mu):
    for name, B, K, iters, eps, seed in (("sinkhorn_b392_k200", 392, 200, 10, 0.05, 11),
                                         ("sinkhorn_b1000_k37", 1000, 37, 3, 0.05, 12),
                                         ("sinkhorn_b64_k300_eps01", 64, 300, 10, 0.1, 13)):
        scores = synth.cosine_scores(B, K, seed=seed)
        q_in = torch.exp(T(scores) / eps).t()                      # time_tuning.py:164 (transposed view)
        q = mu.sinkhorn(q_in, iters, 1)
        save(name, scores=scores, epsilon=np.float32(eps), iters=np.int64(iters), q=q.numpy())


def _dist_worker(rank, ws, port, scores, eps, iters, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, HERE)
    import ref_loader as rl
    mu = rl.load()[0]
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=ws)
    B = scores.shape[0] // ws
    local = torch.from_numpy(scores[rank * B:(rank + 1) * B])
    q = mu.sinkhorn(torch.exp(local / eps).t(), iters, ws)
    np.save(os.path.join(out_dir, f"q{rank}.npy"), q.numpy())
    dist.destroy_process_group()


def gold_sinkhorn_distributed():
    import tempfile
    import torch.multiprocessing as mp
    ws, B, K, iters, eps = 2, 512, 200, 10, 0.05
    scores = synth.cosine_scores(B, K, seed=21)
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_dist_worker, args=(ws, 29431, scores, eps, iters, d), nprocs=ws, join=True)
        q = np.concatenate([np.load(os.path.join(d, f"q{r}.npy")) for r in range(ws)])
    save("sinkhorn_ws2_b512_k200", scores=scores, epsilon=np.float32(eps), iters=np.int64(iters),
         world_size=np.int64(ws), q=q)


def gold_small(mp):
    for h, w, s in ((8, 8, 2), (14, 14, 6), (5, 9, 3)):
        m = mp.restrict_neighborhood(h, w, s).numpy()
        save(f"restrict_{h}x{w}_s{s}", h=np.int64(h), w=np.int64(w), s=np.int64(s),
             mask_bits=np.packbits(m.astype(np.uint8)), shape=np.array(m.shape))
    rng = np.random.default_rng(31)
    m = rng.standard_normal((6, 7, 7)).astype(np.float32)
    m[1] = -np.abs(m[1])          # channel max <= 0 -> zeros
    m[2] = 0
    m[4] = np.abs(m[4]) + 0.5
    save("norm_mask", mask=m, out=mp.norm_mask(T(m)).numpy())
    md = rng.random((3, 5, 5))
    save("norm_mask_f64", mask=md, out=mp.norm_mask(T(md)).numpy())
    y = rng.integers(0, 5, size=(1, 6, 6))
    save("to_one_hot", y=y, n_dims=np.int64(5), out=mp.to_one_hot(T(y), 5).numpy())


def gold_label_propagation(mp):
    """One target frame, 3 contexts, mask passed in (the propagate_labels call pattern :485)."""
    sr, D, C, s, k = 10, 48, 6, 3, 5
    feats = synth.clip_features(1, 4, sr, D, seed=41)[0]
    segs = [synth.soft_labels(sr * sr, C, seed=42 + i).T.reshape(1, C, sr, sr).astype(np.float64) for i in range(3)]
    model = ref_loader.fake_feature_extractor(sr)
    mask = mp.restrict_neighborhood(sr, sr, s)
    seg, feat_tar, _ = mp.label_propagation(s, k, model, T(feats[3]), [T(feats[i]).T for i in range(3)],
                                            [T(x) for x in segs], mask, True)
    save("label_propagation_sr10", feats=feats, segs=np.stack(segs), sr=np.int64(sr), s=np.int64(s),
         topk=np.int64(k), seg_tar=seg.numpy(), feat_tar=feat_tar.numpy())


def gold_propagate(mp):
    cases = (("propagate_sr14_fifo", 14, 64, 10, 6, 2, 3, 5, 51),     # FIFO eviction (n_last=2 < fs-1)
             ("propagate_sr12_k7", 12, 32, 5, 5, 7, 4, 7, 52),       # eval-style topk 7
             ("propagate_sr9_nomask", 9, 24, 4, 4, 7, 0, 3, 53))     # size_mask_neighborhood = 0
    for name, sr, D, C, fs, n_last, s, k, seed in cases:
        feats = synth.clip_features(1, fs, sr, D, seed=seed)[0]
        first = synth.soft_labels(sr * sr, C, seed=seed + 100).T.reshape(1, C, sr, sr)
        mp.mask_neighborhood = None                                  # module-global cache :85
        out = mp.propagate_labels(n_last, s, k, ref_loader.fake_feature_extractor(sr), T(feats), T(first), True)
        save(name, feats=feats, first_seg=first, sr=np.int64(sr), n_last=np.int64(n_last), s=np.int64(s),
             topk=np.int64(k), segs=torch.stack(out).numpy())
    # eval call pattern (mask_propagation.py:821): one-hot first frame at higher resolution -> nearest resize
    sr, D, fs, n_obj = 12, 32, 4, 3
    feats = synth.clip_features(1, fs, sr, D, seed=54)[0]
    ann = synth.blob_label_map(4 * sr, n_obj, seed=55)
    first = mp.to_one_hot(T(ann), n_obj + 1).unsqueeze(0)
    mp.mask_neighborhood = None
    out = mp.propagate_labels(7, 4, 5, ref_loader.fake_feature_extractor(sr), T(feats), first, True)
    save("propagate_eval_onehot", feats=feats, annotation=ann, n_obj=np.int64(n_obj), sr=np.int64(sr),
         n_last=np.int64(7), s=np.int64(4), topk=np.int64(5), segs=torch.stack(out).numpy())


class _StoredFE(torch.nn.Module):
    """Feature extractor returning stored tensors: lets the reference's TimeT.get_loss run its
    FF + Sinkhorn path on given features (models.FeatureExtractor.forward signature, models.py:1070)."""

    def __init__(self, head, backbone, sr):
        super().__init__()
        self.head_feats, self.backbone_feats = head, backbone
        self.spatial_resolution, self.feature_dim = sr, head.shape[-1]
        self.dummy = torch.nn.Parameter(torch.zeros(1))

    def forward(self, x, use_head=True):
        f = self.head_feats if use_head else self.backbone_feats
        return f.reshape(-1, f.shape[-2], f.shape[-1]) + 0 * self.dummy, None


def gold_timet_step(mp, tt, dv):
    """Config-1 shapes (BASELINE.json configs[0]): ViT-S/16 224^2, 4-frame clips, batch 2, K=200,
    random-init backbone from the reference's dino_vision_transformer.vit_small."""
    torch.manual_seed(1)                                            # time_tuning.py:68
    bs, fs, sr, K = 2, 4, 14, 200
    vit = dv.vit_small(patch_size=16).eval()
    clips = np.stack([synth.video_clip(fs, 224, seed=61 + b) for b in range(bs)])
    with torch.no_grad():
        bb = vit.get_intermediate_layers(T(clips).reshape(bs * fs, 3, 224, 224), n=1)[0][:, 1:]   # models.py:965-967
    backbone = bb.reshape(bs, fs, sr * sr, 384).numpy().copy()
    head = synth.head_features(backbone, 256, seed=62)
    protos = synth.prototypes(K, 256, seed=63)
    fe = _StoredFE(T(head), T(backbone), sr)
    model = tt.TimeT(fe, K, prototype_init=T(protos))
    tt.world_size = 1
    mp.mask_neighborhood = None
    x = torch.zeros(bs, fs, 3, 8, 8)
    loss = model.get_loss(x)                                        # time_tuning.py:224-302 with its defaults
    with torch.no_grad():
        batch_q, batch_scores = model.get_scores(T(head)[:, 0], 0.05, 10)
        target_q, target_scores = model.get_scores(T(head)[:, -1], 0.05, 10)
        soft_last, hard = [], []
        for i in range(bs):
            maps = model.make_seg_maps(batch_q[i], T(backbone)[i], 7, 6, 5, True)
            soft_last.append(maps[-1].numpy())
            hard.append(maps[-1].unsqueeze(0).argmax(dim=1).long().numpy()[0])
        all_maps = model.make_seg_maps(batch_q[0], T(backbone)[0], 7, 6, 5, True).numpy()
    save("timet_step_cfg1", backbone=backbone.astype(np.float32), head_src=head[:, 0].copy(), head_tgt=head[:, -1].copy(),
         prototypes=protos, sr=np.int64(sr), batch_q=batch_q.numpy(), target_q=target_q.numpy(),
         target_scores=target_scores.numpy(), soft_last=np.stack(soft_last), hard=np.stack(hard),
         clip0_maps=all_maps, loss=np.float64(loss.item()))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    mu, mp, tt, dv = ref_loader.load()
    gold_sinkhorn(mu)
    gold_small(mp)
    gold_label_propagation(mp)
    gold_propagate(mp)
    gold_timet_step(mp, tt, dv)
    gold_sinkhorn_distributed()


if __name__ == "__main__":
    main()
