"""Multi-GPU check, launched with torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py

Sinkhorn with clips (rows) sharded by rank and the K-vector of prototype marginals all-reduced by the
library's own communicator must equal the single-process result on the concatenated batch — the
reference's distributed semantics (my_utils.py:250-272; fixture sinkhorn_ws2_b512_k200 was produced by
the reference itself under gloo).  Rank 0 prints PASS/FAIL and exits non-zero on failure."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import timet_oracle as O  # noqa: E402
import timetuning_b200 as tb  # noqa: E402
from timetuning_b200 import dist as tdist, synth  # noqa: E402


def main():
    rank, ws = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    tdist.init_comm()
    ok = True
    # (1) golden fixture made by the reference with 2 gloo ranks (only meaningful for ws == 2)
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "sinkhorn_ws2_b512_k200.npz")))
    cases = [("golden-ws2", g["scores"], float(g["epsilon"]), int(g["iters"]), g["q"])] if ws == 2 else []
    # (2) BASELINE configs[2]-style: 784 rows per clip, 4 clips per rank, K = 200
    B = 4 * 784 * ws
    scores = synth.cosine_scores(B, 200, seed=123)
    cases.append(("cfg3-slice", scores, 0.05, 10, O.sinkhorn_scaling(scores, 0.05, 10, dtype=np.float64)))
    # rows beyond shared memory on every rank (configs[2] at 4 GPUs: 64 clips per rank): the hybrid one-launch kernel
    big = synth.cosine_scores(64 * 784 * ws, 200, seed=124)
    cases.append(("hybrid-64-clips-per-rank", big, 0.05, 10, O.sinkhorn_scaling(big, 0.05, 10, dtype=np.float64)))
    if rank == 0:
        print("sinkhorn path:", "p2p" if tb.ops._comm.get("p2p") else "nccl")
    for name, sc, eps, iters, want in cases:
        rows = tdist.shard_range(sc.shape[0], rank, ws)
        local = torch.from_numpy(sc[rows.start:rows.stop]).cuda()
        q = tb.sinkhorn_from_scores(local, eps, iters, world_size=ws)
        q2 = tb.sinkhorn(torch.exp(local / eps).t(), iters, ws)
        gathered = [torch.empty_like(q) for _ in range(ws)]
        dist.all_gather(gathered, q)
        full = torch.cat(gathered).cpu().numpy()
        err = np.abs(full - want)
        bad = err > 1e-4 + 1e-5 * np.abs(want)
        err2 = (q - q2).abs().max().item()
        if rank == 0:
            print(f"{name}: max abs err {err.max():.3e}, outside tol {int(bad.sum())}, fused-vs-exp {err2:.2e}, "
                  f"col sums*K/B in [{(full.sum(0) * full.shape[1] / full.shape[0]).min():.6f}, "
                  f"{(full.sum(0) * full.shape[1] / full.shape[0]).max():.6f}]")
        ok &= not bad.any() and err2 < 1e-6
    # (3) clips sharded by rank give the same FF result as one process doing all clips
    bs_local, fs, sr, D, C = 2, 4, 14, 64, 8
    feats = synth.clip_features(bs_local * ws, fs, sr, D, seed=7)
    first = np.stack([synth.soft_labels(sr * sr, C, seed=30 + b) for b in range(bs_local * ws)])
    clips = tdist.shard_range(bs_local * ws, rank, ws)
    lab, hard = tb.propagate_labels_batched(torch.from_numpy(feats[clips.start:clips.stop]).cuda(),
                                            torch.from_numpy(first[clips.start:clips.stop]).cuda(), 7, 6, 5)
    gl = [torch.empty_like(lab) for _ in range(ws)]
    dist.all_gather(gl, lab)
    if rank == 0:
        ref, _ = tb.propagate_labels_batched(torch.from_numpy(feats).cuda(), torch.from_numpy(first).cuda(), 7, 6, 5)
        same = torch.equal(torch.cat(gl), ref)
        print("FF sharded == single process:", same)
        ok &= same
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("PASS" if flag.item() else "FAIL")
    tdist.destroy_comm()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
