"""mixstage_b200 -- B200-native (sm_100a) drop-in for Mix-StAGE's generator hot path.

Public surface mirrors the reference's names (src/model/__init__.py:5-13):
``JointLateClusterSoftStyle4_G``, ``JointLateClusterSoftStyle4_D`` (= ``Speech2Gesture_D``)
and ``GAN``.  ``install()`` patches them into the reference's ``model.trainer`` namespace,
which is where ``eval(args.model)`` resolves them (trainer.py:1049,1076)."""
from ._lib import MixStageError, load as load_library            # noqa: F401
from .gan import GAN                                              # noqa: F401
from .joint_late_cluster_soft_style import (JointLateClusterSoftStyle4_D,      # noqa: F401
                                            JointLateClusterSoftStyle4_G)
from .speech2gesture import Speech2Gesture_D, Speech2Gesture_G    # noqa: F401
from .ops import get_precision, precision_scope, set_precision    # noqa: F401
from .train_step import FlatState, TrainStep                      # noqa: F401
from .preprocess import PoseMetrics, PosePreprocessor             # noqa: F401

__all__ = ["JointLateClusterSoftStyle4_G", "JointLateClusterSoftStyle4_D", "Speech2Gesture_D", "Speech2Gesture_G", "GAN",
           "install", "MixStageError", "TrainStep", "FlatState", "PosePreprocessor", "PoseMetrics", "set_precision", "get_precision", "precision_scope"]


def install(namespace=None):
    """Make the reference's trainer build the B200 classes.

    ``namespace`` is a module or dict (default: the already-imported ``model.trainer`` and
    ``model`` modules of the reference).  No reference file is edited."""
    import sys
    targets = []
    if namespace is not None:
        targets.append(namespace)
    else:
        for name in ("model.trainer", "model"):
            if name in sys.modules:
                targets.append(sys.modules[name])
        if not targets:
            raise MixStageError("install(): import the reference's `model.trainer` first or pass a namespace")
    for t in targets:
        d = t if isinstance(t, dict) else t.__dict__
        d["JointLateClusterSoftStyle4_G"] = JointLateClusterSoftStyle4_G
        d["JointLateClusterSoftStyle4_D"] = JointLateClusterSoftStyle4_D
        d["Speech2Gesture_D"] = Speech2Gesture_D
        d["Speech2Gesture_G"] = Speech2Gesture_G
        d["GAN"] = GAN
    return targets
