"""The numpy oracle against outputs of the reference itself (tests/golden, made by
oracle/make_golden.py from /root/reference).  CPU only."""
import numpy as np
import pytest

import timet_oracle as O
from conftest import assert_close

TIGHT = dict(atol=2e-6, rtol=2e-5)     # oracle vs reference: same maths, different BLAS/summation order


@pytest.mark.parametrize("name", ["sinkhorn_b392_k200", "sinkhorn_b1000_k37", "sinkhorn_b64_k300_eps01"])
def test_sinkhorn_stepwise(golden, name):
    g = golden(name)
    q = O.find_optimal_assignment(g["scores"], float(g["epsilon"]), int(g["iters"]))
    assert q.dtype == np.float32 and q.shape == g["q"].shape
    assert_close(q, g["q"], what=name, **TIGHT)
    assert (q.argmax(1) == g["q"].argmax(1)).all()


@pytest.mark.parametrize("name", ["sinkhorn_b392_k200", "sinkhorn_b1000_k37", "sinkhorn_b64_k300_eps01"])
def test_sinkhorn_scaling_form(golden, name):
    g = golden(name)
    for dt in (np.float32, np.float64):
        q = O.sinkhorn_scaling(g["scores"], float(g["epsilon"]), int(g["iters"]), dtype=dt)
        assert_close(q, g["q"], what=f"{name}/{dt.__name__}", **TIGHT)


def test_sinkhorn_distributed_equals_global(golden):
    """Reference run with 2 gloo ranks (my_utils.py:250-272) == single-process on the concatenated batch."""
    g = golden("sinkhorn_ws2_b512_k200")
    q = O.find_optimal_assignment(g["scores"], float(g["epsilon"]), int(g["iters"]))
    assert_close(q, g["q"], what="ws2-vs-global", **TIGHT)
    # and the oracle's own world_size>1 code path with an in-process all-reduce
    ws = int(g["world_size"])
    B = g["scores"].shape[0] // ws
    shards = [np.exp(g["scores"][r * B:(r + 1) * B] / g["epsilon"]).T for r in range(ws)]
    outs = _lockstep_sinkhorn(shards, int(g["iters"]), ws)
    assert_close(np.concatenate(outs), g["q"], what="ws2-oracle", **TIGHT)


def _lockstep_sinkhorn(shards, iters, ws):
    """Run O.sinkhorn for every rank in lock-step threads with a barrier-based sum all-reduce."""
    import threading
    barrier = threading.Barrier(ws)
    slots = [None] * ws
    outs = [None] * ws

    def make_allreduce(rank):
        def all_reduce(x):
            slots[rank] = np.array(x, copy=True)
            barrier.wait()
            tot = sum(slots[1:], slots[0].copy())
            barrier.wait()
            return tot.astype(np.float32)
        return all_reduce

    def run(rank):
        outs[rank] = O.sinkhorn(shards[rank], iters, ws, make_allreduce(rank))

    th = [threading.Thread(target=run, args=(r,)) for r in range(ws)]
    [t.start() for t in th]
    [t.join() for t in th]
    return outs


@pytest.mark.parametrize("name", ["restrict_8x8_s2", "restrict_14x14_s6", "restrict_5x9_s3"])
def test_restrict_neighborhood(golden, name):
    g = golden(name)
    ref = np.unpackbits(g["mask_bits"])[: int(np.prod(g["shape"]))].reshape(g["shape"]).astype(np.float32)
    m = O.restrict_neighborhood(int(g["h"]), int(g["w"]), int(g["s"]))
    assert m.dtype == np.float32
    assert np.array_equal(m, ref)


@pytest.mark.parametrize("name", ["norm_mask", "norm_mask_f64"])
def test_norm_mask(golden, name):
    g = golden(name)
    out = O.norm_mask(g["mask"])
    assert out.dtype == g["out"].dtype
    assert np.array_equal(out, g["out"])


def test_to_one_hot(golden):
    g = golden("to_one_hot")
    assert np.array_equal(O.to_one_hot(g["y"], int(g["n_dims"])), g["out"])


def test_label_propagation(golden):
    g = golden("label_propagation_sr10")
    feats, segs = g["feats"], g["segs"]
    seg, feat_tar, _ = O.label_propagation(int(g["s"]), int(g["topk"]), int(g["sr"]), feats[3],
                                           [feats[i].T for i in range(3)], list(segs),
                                           O.restrict_neighborhood(10, 10, int(g["s"])))
    assert seg.dtype == np.float64 and seg.shape == g["seg_tar"].shape
    assert np.array_equal(feat_tar, g["feat_tar"])
    assert_close(seg, g["seg_tar"], what="label_propagation", **TIGHT)


@pytest.mark.parametrize("name", ["propagate_sr14_fifo", "propagate_sr12_k7", "propagate_sr9_nomask"])
def test_propagate_labels(golden, name):
    g = golden(name)
    out = O.propagate_labels(int(g["n_last"]), int(g["s"]), int(g["topk"]), int(g["sr"]), g["feats"], g["first_seg"])
    out = np.stack(out)
    assert out.dtype == np.float64
    assert_close(out, g["segs"], what=name, **TIGHT)
    # the sparse fp64 explainer agrees too and reports sane margins
    C = g["first_seg"].shape[1]
    ex = O.ff_sparse(int(g["n_last"]), int(g["s"]), int(g["topk"]), int(g["sr"]), g["feats"], g["first_seg"][0])
    assert ex["margin"].min() > 1e-5, "fixture contains a near-tie; regenerate with another seed"
    assert_close(ex["segs"], g["segs"], what=name + "/sparse", **TIGHT)
    assert (ex["nnz"] == int(g["topk"])).all()
    assert C == out.shape[1]


def test_propagate_eval_onehot(golden):
    g = golden("propagate_eval_onehot")
    first = O.to_one_hot(g["annotation"], int(g["n_obj"]) + 1)[None]
    out = np.stack(O.propagate_labels(int(g["n_last"]), int(g["s"]), int(g["topk"]), int(g["sr"]), g["feats"], first))
    assert_close(out, g["segs"], what="eval-onehot", **TIGHT)
    assert (out.argmax(1) == g["segs"].argmax(1)).mean() > 0.999


def test_context_frames():
    assert O.context_frames(1, 7) == [0]
    assert O.context_frames(3, 7) == [0, 1, 2]
    assert O.context_frames(8, 7) == [0, 1, 2, 3, 4, 5, 6, 7]
    assert O.context_frames(9, 7) == [0, 2, 3, 4, 5, 6, 7, 8]
    assert O.context_frames(5, 2) == [0, 3, 4]


def test_timet_step_cfg1(golden):
    """BASELINE.json configs[0]: the FF+Sinkhorn part of TimeT.get_loss on ViT-S/16 224^2 features."""
    g = golden("timet_step_cfg1")
    bq, tq, hard, soft = O.ff_sinkhorn_step(g["head_src"], g["head_tgt"], g["backbone"], g["prototypes"], int(g["sr"]))
    assert_close(bq, g["batch_q"], what="batch_q", **TIGHT)
    assert_close(tq, g["target_q"], what="target_q", **TIGHT)
    assert_close(soft, g["soft_last"], what="soft_last", atol=1e-5, rtol=1e-5)
    assert np.array_equal(hard, g["hard"])


def test_torch_baseline_port_matches_reference(golden):
    """The CPU-baseline port (same ATen op sequence) reproduces the reference's outputs on config 1."""
    import torch
    import timet_oracle_torch as OT
    g = golden("timet_step_cfg1")
    T = torch.from_numpy
    bq, tq, hard, soft = OT.ff_sinkhorn_step(T(g["head_src"]), T(g["head_tgt"]), T(g["backbone"]), T(g["prototypes"]),
                                             int(g["sr"]))
    assert_close(bq.numpy(), g["batch_q"], what="batch_q", atol=1e-7, rtol=1e-6)
    assert_close(tq.numpy(), g["target_q"], what="target_q", atol=1e-7, rtol=1e-6)
    assert_close(soft.numpy(), g["soft_last"], what="soft", atol=1e-7, rtol=1e-6)
    assert np.array_equal(hard.numpy(), g["hard"])
    gp = golden("propagate_sr14_fifo")
    sr = int(gp["sr"])
    out = OT.propagate_clip(int(gp["n_last"]), int(gp["s"]), int(gp["topk"]), sr, T(gp["feats"]),
                            T(gp["first_seg"]).double(), OT.window_mask(sr, int(gp["s"])))
    assert_close(torch.stack(out).numpy(), gp["segs"], what="fifo", atol=1e-7, rtol=1e-6)
