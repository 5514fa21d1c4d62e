"""Seeded synthetic inputs for the FF + Sinkhorn hot path (tests, smoke, bench).

There are no datasets or checkpoints in this environment, so every
measurement uses inputs of the reference's shapes generated here.  All generators are
deterministic in ``seed`` and run on CPU (numpy) so the oracle and the CUDA path see
bit-identical inputs.

``clip_features`` does not need a ViT: it samples a smooth random feature field under a
per-frame translation plus noise, which reproduces the two properties of real backbone
features that matter for Feature-Forwarding — temporal coherence (the best matches of a
patch lie near its own location in earlier frames) and a dense spectrum of near-equal
similarities around the k-th neighbour.
"""
from __future__ import annotations

import numpy as np


def _smooth_field(rng, channels, size, cells):
    """Bilinear up-sampling of a coarse ``[channels, cells, cells]`` gaussian grid to ``size``."""
    coarse = rng.standard_normal((channels, cells + 1, cells + 1)).astype(np.float32)
    pos = np.linspace(0, cells, size, endpoint=False, dtype=np.float32)
    i0 = np.floor(pos).astype(np.int64)
    f = (pos - i0).astype(np.float32)
    rows = coarse[:, i0, :] * (1 - f)[None, :, None] + coarse[:, i0 + 1, :] * f[None, :, None]
    return rows[:, :, i0] * (1 - f)[None, None, :] + rows[:, :, i0 + 1] * f[None, None, :]


def clip_features(bs, fs, sr, dim, seed=1, noise=0.35, max_shift=1.25, cells=None):
    """Backbone-like features ``[bs, fs, sr*sr, dim]`` float32 (un-normalised, as
    time_tuning.py:238-239 hands them to make_seg_maps)."""
    rng = np.random.default_rng(seed)
    cells = cells or max(4, sr // 3)
    pad = int(np.ceil(max_shift * fs)) + 2
    out = np.empty((bs, fs, sr * sr, dim), dtype=np.float32)
    for b in range(bs):
        field = _smooth_field(rng, dim, (sr + 2 * pad) * 2, cells * 2)      # 2x oversampled canvas
        vel = rng.uniform(-max_shift, max_shift, size=2)
        for t in range(fs):
            oy = int(round((pad + vel[0] * t) * 2))
            ox = int(round((pad + vel[1] * t) * 2))
            crop = field[:, oy:oy + 2 * sr:2, ox:ox + 2 * sr:2]
            frame = crop.reshape(dim, sr * sr).T
            frame = frame + noise * rng.standard_normal(frame.shape).astype(np.float32)
            out[b, t] = frame * np.float32(1.7) + np.float32(0.05)          # arbitrary scale: FF must normalise
    return out


def head_features(backbone, out_dim=256, seed=2):
    """A fixed random 2-layer MLP standing in for the projection head (time_tuning.py:574):
    ``[..., D] -> [..., out_dim]`` float32."""
    rng = np.random.default_rng(seed)
    d = backbone.shape[-1]
    w1 = (rng.standard_normal((d, 512)) / np.sqrt(d)).astype(np.float32)
    w2 = (rng.standard_normal((512, out_dim)) / np.sqrt(512)).astype(np.float32)
    return (np.tanh(backbone @ w1) @ w2).astype(np.float32)


def prototypes(k, dim=256, seed=3):
    """L2-normalised gaussian prototypes (time_tuning.py:91-92)."""
    rng = np.random.default_rng(seed)
    p = rng.standard_normal((k, dim)).astype(np.float32)
    return p / np.linalg.norm(p, axis=1, keepdims=True).astype(np.float32)


def cosine_scores(n_rows, k, seed=4, dim=256):
    """Cosine scores ``[n_rows, k]`` in [-1, 1] (time_tuning.py:130-141) from random unit vectors."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n_rows, dim)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True).astype(np.float32)
    # sharpen so that exp(s/eps) spans many decades like trained prototypes do
    return np.clip(3.0 * (x @ prototypes(k, dim, seed + 1).T), -1.0, 1.0).astype(np.float32)


def soft_labels(n, c, seed=5):
    """First-frame soft labels ``[n, c]``: rows are distributions (a Sinkhorn Q row)."""
    rng = np.random.default_rng(seed)
    z = 3.0 * rng.standard_normal((n, c)).astype(np.float32)
    z = np.exp(z - z.max(axis=1, keepdims=True))
    return (z / z.sum(axis=1, keepdims=True)).astype(np.float32)


def blob_label_map(sr, n_objects, seed=6):
    """DAVIS-style first-frame annotation ``[1, sr, sr]`` int64 with ``n_objects`` blobs + background."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:sr, 0:sr]
    lab = np.zeros((sr, sr), dtype=np.int64)
    for o in range(1, n_objects + 1):
        cy, cx = rng.uniform(0, sr, size=2)
        r = rng.uniform(sr / 12, sr / 5)
        lab[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = o
    return lab[None]


def video_clip(fs, size, seed=7):
    """Image clip ``[fs, 3, size, size]`` float32: smooth field translated 2-3 px/frame + noise,
    ImageNet-normalised.  Only used to feed a ViT in the build container."""
    rng = np.random.default_rng(seed)
    pad = 3 * fs + 4
    canvas = _smooth_field(rng, 3, size + 2 * pad, 32)
    vel = rng.uniform(2, 3, size=2) * rng.choice([-1, 1], size=2)
    mean = np.array([0.485, 0.456, 0.406], dtype=np.float32)[:, None, None]
    std = np.array([0.229, 0.224, 0.225], dtype=np.float32)[:, None, None]
    out = np.empty((fs, 3, size, size), dtype=np.float32)
    for t in range(fs):
        oy = int(round(pad + vel[0] * t))
        ox = int(round(pad + vel[1] * t))
        img = canvas[:, oy:oy + size, ox:ox + size] * 0.25 + 0.5
        img = img + 0.1 * rng.standard_normal(img.shape).astype(np.float32)
        out[t] = (img - mean) / std
    return out
