"""Host-side multi-process logic on CPU (gloo, world_size 2): clip sharding, the byte broadcast that
carries the NCCL unique id, and the distributed-Sinkhorn semantics (rows sharded by rank + summed
K-vector == global batch) with the oracle's world_size > 1 path over a real gloo all-reduce."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, ws, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import timet_oracle as O
    from timetuning_b200 import dist as tdist, synth
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=ws)
    # byte broadcast (carrier of the 128-byte unique id)
    payload = bytes(range(128)) if rank == 0 else None
    got = tdist.broadcast_bytes(payload, 128, src=0)
    assert got == bytes(range(128))
    # sharding
    r = tdist.shard_range(8, rank, ws)
    assert (r.start, r.stop) == (rank * 4, rank * 4 + 4)
    # distributed Sinkhorn semantics through a real all-reduce
    scores = synth.cosine_scores(256, 48, seed=5)
    rows = tdist.shard_range(256, rank, ws)

    def all_reduce(x):
        t = torch.from_numpy(np.array(x, dtype=np.float32, copy=True))
        dist.all_reduce(t)
        return t.numpy()
    q = O.find_optimal_assignment(scores[rows.start:rows.stop], 0.05, 5, ws, all_reduce)
    np.save(os.path.join(out_dir, f"q{rank}.npy"), q)
    dist.destroy_process_group()


def test_gloo_world_size_2(tmp_path):
    ws = 2
    mp.spawn(_worker, args=(ws, 29533, str(tmp_path)), nprocs=ws, join=True)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import timet_oracle as O
    from timetuning_b200 import synth
    q = np.concatenate([np.load(tmp_path / f"q{r}.npy") for r in range(ws)])
    ref = O.find_optimal_assignment(synth.cosine_scores(256, 48, seed=5), 0.05, 5)
    np.testing.assert_allclose(q, ref, rtol=2e-5, atol=2e-7)


def test_shard_range_rejects_ragged():
    from timetuning_b200 import dist as tdist
    with pytest.raises(ValueError):
        tdist.shard_range(7, 0, 2)
