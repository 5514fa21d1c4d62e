"""The drop-in boundary against the REAL reference modules: import time_tuning / mask_propagation /
my_utils (in place here, from the shipped git-ignored baseline/_ref copy on the GPU box), run the reference's own
callers unpatched, call timetuning_b200.install(...), run the same callers again through the patched names, compare.

Callers exercised (file:line in the reference):
  TimeT.get_loss            time_tuning.py:224-302   (-> get_scores :195, find_optimal_assignment :157 -> sinkhorn;
                                                       make_seg_maps :143 -> propagate_labels -> label_propagation)
  mask_propagation eval     mask_propagation.py:817-824  (propagate_labels + bilinear up-sampling + argmax)
Skipped cleanly when no copy of the reference is available."""
import numpy as np
import pytest
import torch

import ref_loader
import timetuning_b200 as tb
from conftest import assert_close
from parity import check_hard, check_soft
from timetuning_b200 import synth, training

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="reference copy not available")]


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


def _model(ref, g, teacher=False, queue=0, device="cuda"):
    import ref_step
    # the fixture stores the head features of frames 0 and -1; the frames in between never reach the loss (:248, :269)
    bs, fs = g["backbone"].shape[:2]
    full = torch.zeros((bs, fs) + g["head_src"].shape[1:])
    full[:, 0], full[:, -1] = torch.from_numpy(g["head_src"]), torch.from_numpy(g["head_tgt"])
    model = ref_step.build_timet(full.to(device), torch.from_numpy(g["backbone"]).to(device),
                                 torch.from_numpy(g["prototypes"]).to(device), int(g["sr"]))
    if teacher:
        model.init_momentum_teacher()
        with torch.no_grad():                       # a teacher that differs from the student
            model.teacher_prototypes.copy_(torch.nn.functional.normalize(
                model.teacher_prototypes + 0.05 * torch.randn_like(model.teacher_prototypes), dim=1))
            model.teacher.head_feats = model.teacher.head_feats * 1.0 + 0.01
    if queue:
        model.init_queue(queue)
        with torch.no_grad():
            model.queue.copy_(torch.randn_like(model.queue))
    return model


@pytest.mark.parametrize("teacher,queue", [(False, 0), (True, 0), (False, 1024), (True, 640)])
def test_get_loss_through_the_patched_names(ref, golden, teacher, queue):
    """TimeT.get_loss unpatched (the reference's dense path, here on the GPU) vs patched by install(): same loss; and
    vs install(fast_get_loss=True): same loss, same gradient.  With the teacher and the feature queue too."""
    mu, mp, tt, _ = ref
    g = golden("timet_step_cfg1")
    x = torch.zeros(2, 4, 3, 8, 8, device="cuda")

    def run(patch):
        torch.manual_seed(7)
        model = _model(ref, g, teacher, queue)
        mp.mask_neighborhood = None
        torch.manual_seed(11)                        # the queue update draws a randperm (:257)
        if patch is None:
            loss = model.get_loss(x)
        else:
            with tb.install(tt, mp, mu, fast_get_loss=(patch == "fast")):
                loss = model.get_loss(x)
        loss.backward()
        return loss.item(), model.prototypes.grad.clone(), (model.queue.clone() if queue else None)

    l_ref, g_ref, q_ref = run(None)
    if not teacher and not queue:
        assert abs(l_ref - float(g["loss"])) < 1e-5, "unpatched reference on the GPU reproduces its own CPU fixture"
    for patch in ("shim", "fast"):
        l, gr, q = run(patch)
        assert abs(l - l_ref) < 2e-5 * max(1.0, abs(l_ref)), (patch, l, l_ref)
        assert_close(gr.cpu().numpy(), g_ref.cpu().numpy(), atol=1e-6, rtol=1e-4, what=f"{patch}: d loss / d prototypes")
        if queue:
            assert torch.equal(q, q_ref), "queue update (:250-261) must be untouched"


def test_fast_get_loss_aux_matches_reference_step(ref, golden):
    """The batched fast path's intermediate results against the reference-made fixture of the same step."""
    mu, mp, tt, _ = ref
    g = golden("timet_step_cfg1")
    model = _model(ref, g)
    x = torch.zeros(2, 4, 3, 8, 8, device="cuda")
    loss, aux = training.fast_get_loss(model, x, return_aux=True)
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    assert_close(aux["batch_q"].cpu().numpy(), g["batch_q"], what="batch_q")
    assert np.array_equal(aux["hard"].cpu().numpy(), g["hard"])


def test_get_scores_with_feature_queue(ref, golden):
    """TimeT.get_scores with a filled feature queue (:207-211): the queue rows take part in the Sinkhorn marginals."""
    mu, mp, tt, _ = ref
    g = golden("timet_step_cfg1")
    model = _model(ref, g, queue=1536)
    feats = cu(g["head_src"])
    with torch.no_grad():
        q_ref, s_ref = model.get_scores(feats, 0.05, 10)
        q_noq = tt.TimeT.get_scores(_model(ref, g), feats, 0.05, 10)[0]
        with tb.install(tt, mp, mu, fast_get_loss=True):
            q_new, s_new = model.get_scores(feats, 0.05, 10)
        q_asg = training.assignment(model, feats, 0.05, 10)
    assert (q_ref - q_noq).abs().max().item() > 1e-4, "the queue must matter in this test"
    assert_close(q_new.cpu().numpy(), q_ref.cpu().numpy(), what="get_scores with queue")
    assert_close(q_asg.cpu().numpy(), q_ref.cpu().numpy(), what="assignment with queue")
    assert torch.equal(s_new, s_ref)


def test_propagation_eval_call_pattern(ref, golden):
    """mask_propagation.py:817-824 on one synthetic video: propagate_labels (via the module attribute the eval driver
    resolves, :821) + stack + bilinear up-sampling + argmax, unpatched vs patched; and label_propagation (:485)."""
    mu, mp, tt, _ = ref
    sr, D, n_obj, fs, R = 20, 96, 4, 6, 80
    feats = cu(synth.clip_features(1, fs, sr, D, seed=44)[0])
    ann = torch.from_numpy(synth.blob_label_map(R, n_obj, seed=45))
    first = mp.to_one_hot(ann, n_obj + 1).unsqueeze(0)                       # :821 (CPU one-hot, like the eval driver)

    class FE:
        spatial_resolution = sr

    def eval_tail():
        mp.mask_neighborhood = None
        maps = mp.propagate_labels(7, 4, 5, FE(), feats, first, True)         # :821
        maps = torch.stack(maps, dim=0)                                       # :822
        up = torch.nn.functional.interpolate(maps, size=(R, R), mode="bilinear", align_corners=False)   # :823
        return maps, up.max(dim=1)[1]                                         # :824

    maps_ref, pred_ref = eval_tail()
    with tb.install(tt, mp, mu):
        maps_new, pred_new = eval_tail()
        assert maps_new.dtype == maps_ref.dtype and maps_new.shape == maps_ref.shape
    check_soft(maps_new.cpu().numpy(), maps_ref.cpu().numpy(), what="eval maps")
    up = torch.nn.functional.interpolate(maps_ref, size=(R, R), mode="bilinear", align_corners=False)
    srt = up.sort(dim=1)[0]
    decided = (srt[:, -1] - srt[:, -2]) > 1e-5
    assert (pred_new == pred_ref)[decided].all() and decided.float().mean() > 0.95

    # one explicit label_propagation call (the inner call of propagate_labels, :485) with the reference's own mask
    mask = mp.restrict_neighborhood(sr, sr, 4)
    segs = [first_to_sr(first, sr).cuda()] + [maps_ref[i:i + 1] for i in range(2)]
    ctx_feats = [feats[i].t() for i in range(3)]
    seg_ref, ft_ref, _ = mp.label_propagation(4, 5, FE(), feats[3], ctx_feats, segs, mask.cuda(), True)
    with tb.install(tt, mp, mu):
        seg_new, ft_new, m_new = mp.label_propagation(4, 5, FE(), feats[3], ctx_feats, segs, mask.cuda(), True)
        assert np.array_equal(mp.restrict_neighborhood(sr, sr, 4).cpu().numpy(), mask.numpy())
    assert torch.equal(ft_new, ft_ref)
    check_soft(seg_new.cpu().numpy(), seg_ref.cpu().numpy(), what="label_propagation")
    check_hard(seg_new.argmax(1).cpu().numpy(), seg_ref.cpu().numpy(), what="label_propagation hard")


def first_to_sr(first, sr):
    return torch.nn.functional.interpolate(first.type(torch.DoubleTensor), size=(sr, sr), mode="nearest")


def test_sinkhorn_through_my_utils(ref, golden):
    """my_utils.sinkhorn patched: called the way TimeT.find_optimal_assignment calls it (:164-165)."""
    mu, mp, tt, _ = ref
    g = golden("sinkhorn_b392_k200")
    q_in = torch.exp(cu(g["scores"]) / float(g["epsilon"])).t()
    q_ref = mu.sinkhorn(q_in, int(g["iters"]), 1)
    with tb.install(tt, mp, mu):
        q_new = mu.sinkhorn(q_in, int(g["iters"]), 1)
        assert mu.sinkhorn is tb.sinkhorn
    assert_close(q_new.cpu().numpy(), q_ref.cpu().numpy(), what="sinkhorn")
    assert_close(q_ref.cpu().numpy(), g["q"], what="reference on GPU vs its CPU fixture")


class _TinyFE(torch.nn.Module):
    """A small extractor with the structure of the reference's models.FeatureExtractor (models.py:903-1078): a backbone
    reached through get_features(), an MLP head, forward(x, use_head) = get_features then head."""

    def __init__(self, sr, dim=48, out=32):
        super().__init__()
        self.spatial_resolution, self.feature_dim = sr, out
        self.patch = torch.nn.Conv2d(3, dim, kernel_size=4, stride=4)
        self.mix = torch.nn.Linear(dim, dim)
        self.head = torch.nn.Sequential(torch.nn.Linear(dim, 64), torch.nn.GELU(), torch.nn.Linear(64, out))
        self.calls = 0

    def get_features(self, x):
        self.calls += 1
        t = self.patch(x).flatten(2).permute(0, 2, 1)
        return t + torch.tanh(self.mix(t)), None

    def forward(self, x, use_head=True):
        t, a = self.get_features(x)
        if use_head:
            n, p, d = t.shape
            t = self.head(t.reshape(n * p, d)).view(n, p, -1)
        return t, a


def test_fast_get_loss_shares_one_backbone_forward(ref):
    """the reference runs the student backbone twice per step (time_tuning.py:237-239); the fast path
    runs it once when the extractor has the reference's structure -- same loss, same gradients (head AND backbone)."""
    mu, mp, tt, _ = ref
    sr, bs, fs, K = 8, 3, 4, 24

    def run(fast):
        torch.manual_seed(3)
        fe = _TinyFE(sr).cuda()
        model = tt.TimeT(fe, K).cuda()
        tt.world_size = 1
        mp.mask_neighborhood = None
        x = torch.randn(bs, fs, 3, 4 * sr, 4 * sr, device="cuda")
        if fast:
            with tb.install(tt, mp, mu, fast_get_loss=True):
                loss = model.get_loss(x, size_mask_neighborhood=3, topk=3)
        else:
            loss = model.get_loss(x, size_mask_neighborhood=3, topk=3)
        loss.backward()
        return loss.item(), fe.calls, fe.patch.weight.grad.clone(), fe.head[0].weight.grad.clone(), model.prototypes.grad.clone()

    l0, c0, gp0, gh0, gq0 = run(False)
    l1, c1, gp1, gh1, gq1 = run(True)
    assert c0 == 2 and c1 == 1, (c0, c1)
    assert abs(l0 - l1) < 2e-5 * max(1.0, abs(l0)), (l0, l1)
    for a, b, what in ((gp1, gp0, "backbone grad"), (gh1, gh0, "head grad"), (gq1, gq0, "prototype grad")):
        assert_close(a.cpu().numpy(), b.cpu().numpy(), atol=1e-6, rtol=1e-4, what=what)
