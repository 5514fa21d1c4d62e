"""Device-resident FF (+ Sinkhorn where the config has prototypes) timings for every BASELINE.json config shape,
with the selection diagnostics.  Not the bench line (bench.py = configs[1]); evidence that the other configs run
on the tensor-core engine and how many queries needed the exact re-do."""
import json
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import timetuning_b200 as tb
from timetuning_b200 import synth

CONFIGS = [
    # name, clips, frames, grid, D, C, n_last, radius, topk, sinkhorn K (0 = none)
    ("configs[0] ViT-S/16 224^2, 4 frames, batch 2", 2, 4, 14, 384, 200, 7, 6, 5, 200),
    ("configs[1] ViT-S/16 448^2, 8 frames, batch 32", 32, 8, 28, 384, 200, 7, 6, 5, 200),
    ("configs[3] DAVIS-style 480p ViT-S/8, 80 frames, 1 video", 1, 80, 60, 384, 11, 7, 12, 7, 0),
    ("configs[4] ViT-B/8 448^2, 16 frames, 8 clips per GPU, K=300", 8, 16, 56, 768, 300, 7, 6, 5, 300),
]


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, bs, fs, sr, D, C, n_last, radius, topk, K in CONFIGS:
    N = sr * sr
    feats = torch.from_numpy(synth.clip_features(bs, fs, sr, D, seed=1)).cuda()
    first = torch.from_numpy(np.stack([synth.soft_labels(N, C, seed=9 + b) for b in range(bs)])).cuda()
    plan = tb.FFPlan(bs, fs, sr, sr, D, C, n_last, radius, topk)
    labels = torch.empty((bs, fs, N, C), dtype=torch.float32, device="cuda")
    labels[:, 0] = first
    hard = torch.empty((bs, N), dtype=torch.int64, device="cuda")
    out = {"config": name, "tc_engine": plan.tc_supported}
    out["prepare_ms"] = timed(lambda: plan.prepare(feats))
    out["select_ms"] = timed(lambda: plan.select(tb.FF_AUTO))
    out["gather_ms"] = timed(lambda: plan.gather(labels, hard))
    out["stats"] = plan.stats()
    if K:
        scores = torch.from_numpy(synth.cosine_scores(bs * N, K, seed=4)).cuda()
        out["sinkhorn_ms"] = timed(lambda: tb.sinkhorn_from_scores(scores, 0.05, 10))
    sigma_ctx = sum(1 + (t - max(1, t - n_last)) for t in range(1, fs))
    out["dense_TFLOPs_select"] = 2.0 * N * N * D * sigma_ctx * bs / (out["select_ms"] * 1e-3) / 1e12
    out["ff_clips_per_s"] = bs / ((out["prepare_ms"] + out["select_ms"] + out["gather_ms"]) * 1e-3)
    print(json.dumps(out))
    del plan, feats, labels
    torch.cuda.empty_cache()
