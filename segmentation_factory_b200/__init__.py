"""segmentation_factory_b200 -- B200-native per-pixel head-to-metric path of Segmentation_Factory.

Mirrors the reference's util/ API for this path only:
    CrossEntropy, OhemCrossEntropy, get_loss        (util/losses.py)
    Metrics                                         (util/metrics.py)
    ConfusionMatrix                                 (util/utils.py)
    criterion (CE + dice)                           (engine.py:10-15)
    LabelLUT (id_to_train_id / label_map on device) (datasets/cityscapes.py:92,130-131; coco_stuff.py:66-69,98-100)
    HeadClassifier (the heads' final 1x1 conv)      (models/heads/segformer.py:39,57; upernet.py:32,50; fpn.py:23,37)
backed by hand-written sm_100a kernels in libsegfactory.so (C ABI: include/segfactory.h).
"""
from .head import HeadClassifier, head_classifier
from .labels import LabelLUT
from .losses import CrossEntropy, FocalLoss, OhemCrossEntropy, get_loss
from .metrics import ConfusionMatrix, Metrics, fused_eval_update, reduce_from_all_processes
from .engine import criterion
from .functional import eval_hist, fused_criterion, fused_cross_entropy, fused_focal_loss, hist_pairs

__all__ = ["CrossEntropy", "OhemCrossEntropy", "FocalLoss", "fused_focal_loss", "get_loss", "Metrics", "ConfusionMatrix",
           "fused_eval_update", "reduce_from_all_processes", "fused_cross_entropy", "fused_criterion", "criterion", "eval_hist", "hist_pairs", "LabelLUT", "head_classifier", "HeadClassifier"]
__version__ = "0.2.0"
