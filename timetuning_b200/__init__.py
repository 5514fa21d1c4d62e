"""timetuning_b200 — B200-native Feature-Forwarding + Sinkhorn-Knopp for TimeT (SMSD75/Timetuning).

Only the one data-parallel hot path named by BASELINE.json's north_star lives here: hand-written
sm_100a CUDA (csrc/) behind a C ABI (include/timet_b200.h) and the Python mirror of the reference's
callables (ops.py).  `install()` binds them over the reference's module attributes.
"""
from .ops import (FFPlan, FF_AUTO, FF_EXACT, FF_TC, cosine_scores, label_propagation, norm_mask, propagate_labels,  # noqa: F401
                  propagate_labels_batched, propagate_labels_eval, upsample_argmax, restrict_neighborhood, sinkhorn, sinkhorn_from_scores)

__all__ = ["sinkhorn", "sinkhorn_from_scores", "cosine_scores", "restrict_neighborhood", "norm_mask", "label_propagation",
           "propagate_labels", "propagate_labels_batched", "propagate_labels_eval", "upsample_argmax", "FFPlan", "FF_AUTO", "FF_EXACT", "FF_TC", "install"]


def install(time_tuning=None, mask_propagation=None, my_utils=None):
    """Bind the CUDA-backed callables over the reference's module attributes (SURVEY.md §8b
    "install points"); no reference source is edited.  Pass the already-imported reference modules."""
    from . import ops
    if my_utils is not None:
        my_utils.sinkhorn = ops.sinkhorn
    if mask_propagation is not None:
        mask_propagation.label_propagation = ops.label_propagation
        mask_propagation.propagate_labels = ops.propagate_labels
        mask_propagation.restrict_neighborhood = ops.restrict_neighborhood
        mask_propagation.norm_mask = ops.norm_mask
    if time_tuning is not None:
        time_tuning.sinkhorn = ops.sinkhorn                    # time_tuning.py:49 binds by name at import
        time_tuning.propagate_labels = ops.propagate_labels    # time_tuning.py:51
