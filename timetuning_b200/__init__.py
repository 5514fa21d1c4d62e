"""timetuning_b200 — B200-native Feature-Forwarding + Sinkhorn-Knopp for TimeT (SMSD75/Timetuning).

Only the one data-parallel hot path named by BASELINE.json's north_star lives here: hand-written
sm_100a CUDA (csrc/) behind a C ABI (include/timet_b200.h) and the Python mirror of the reference's
callables (ops.py).  `install()` binds them over the reference's module attributes.
"""
from .ops import (FFPlan, FF_AUTO, FF_EXACT, FF_TC, cosine_scores, label_propagation, norm_mask, propagate_labels,  # noqa: F401
                  propagate_labels_batched, propagate_labels_eval, upsample_argmax, restrict_neighborhood, sinkhorn, sinkhorn_from_scores,
                  sinkhorn_pair_from_scores, cosine_scores_multi)

__all__ = ["sinkhorn", "sinkhorn_from_scores", "sinkhorn_pair_from_scores", "cosine_scores_multi", "cosine_scores", "restrict_neighborhood", "norm_mask", "label_propagation",
           "propagate_labels", "propagate_labels_batched", "propagate_labels_eval", "upsample_argmax", "FFPlan", "FF_AUTO", "FF_EXACT", "FF_TC", "install", "Installation"]


_BINDINGS = (("my_utils", "sinkhorn", "sinkhorn"),                          # my_utils.py:246
             ("mask_propagation", "label_propagation", "label_propagation"),   # mask_propagation.py:396
             ("mask_propagation", "propagate_labels", "propagate_labels"),     # :448
             ("mask_propagation", "restrict_neighborhood", "restrict_neighborhood"),   # :377
             ("mask_propagation", "norm_mask", "norm_mask"),                   # :363
             ("time_tuning", "sinkhorn", "sinkhorn"),                          # time_tuning.py:49 binds by name at import
             ("time_tuning", "propagate_labels", "propagate_labels"))          # time_tuning.py:51


class Installation:
    """What install() replaced; uninstall() puts the reference's own callables back."""

    def __init__(self):
        self._saved = []

    def _bind(self, obj, name, value):
        self._saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)

    def uninstall(self):
        for obj, name, old in reversed(self._saved):
            setattr(obj, name, old)
        self._saved.clear()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.uninstall()


def install(time_tuning=None, mask_propagation=None, my_utils=None, fast_get_loss=False) -> Installation:
    """Bind the CUDA-backed callables over the reference's module attributes (SURVEY.md §8b "install points"); no
    reference source is edited.  Pass the already-imported reference modules.

    fast_get_loss=True additionally rebinds ``time_tuning.TimeT.get_scores`` and ``TimeT.get_loss`` to the batched
    training fast path (timetuning_b200/training.py: one Sinkhorn instead of two/four, all clips in one Feature-Forwarding
    call, fused argmax; same loss value).  Returns a handle whose ``uninstall()`` restores the reference."""
    from . import ops, training
    mods = {"time_tuning": time_tuning, "mask_propagation": mask_propagation, "my_utils": my_utils}
    inst = Installation()
    for mod_name, attr, ours in _BINDINGS:
        if mods[mod_name] is not None:
            inst._bind(mods[mod_name], attr, getattr(ops, ours))
    if fast_get_loss:
        if time_tuning is None:
            raise ValueError("fast_get_loss=True needs the time_tuning module")
        inst._bind(time_tuning.TimeT, "get_scores", training.get_scores)
        inst._bind(time_tuning.TimeT, "get_loss", training.fast_get_loss)
    return inst
