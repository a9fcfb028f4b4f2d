"""Drop-in mirror of util/metrics.py:Metrics (semantic-segmentation part) and util/utils.py:ConfusionMatrix.

`update` runs the fused upsample+argmax+histogram kernel; everything O(C^2) (compute_*, reduce) stays
as the reference's own tensor expressions on the accumulated matrix, so the derived numbers are
bit-identical once the matrix is.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import Tensor

from .functional import accumulate_counts, eval_hist, eval_update, hist_pairs

__all__ = ["Metrics", "ConfusionMatrix", "fused_eval_update"]


class Metrics:
    """util/metrics.py:9-49,108-114.

    `hist` is float32 like the reference's and receives the same per-call `hist += int64_counts`
    (inexact past 2**24 per cell, reproduced as is).  `hist_i64` is an exact shadow;
    it is what gets all-reduced (integer SUM is order-independent), after which `hist` is refreshed from it.
    """

    def __init__(self, num_classes: int, ignore_label: int, device) -> None:
        self.ignore_label = ignore_label
        self.num_classes = num_classes
        self.hist = torch.zeros(num_classes, num_classes).to(device)
        self.hist_i64 = torch.zeros(num_classes, num_classes, dtype=torch.int64).to(device)
        self.device = device
        self.pq_hist = {"TP": 0, "FP": 0, "FN": 0, "iou_sum": 0.0}   # util/metrics.py:17-22: kept for attribute parity only
        self._counts = None               # per-call matrix (device scratch, left zeroed by sf_accumulate_counts)

    def _accumulate(self, counts: Tensor) -> None:
        self.hist += counts                   # float32 += int64, exactly the reference's update expression
        self.hist_i64 += counts

    def update(self, pred: Tensor, target: Tensor, size=None) -> None:
        """pred: [B, C, h, w] logits at any resolution; target: [B, H, W] labels, or the flattened vector
        engine.py:91 passes (then H, W = `size`, defaulting to pred's own h, w)."""
        eval_update(pred, target, self.ignore_label, self._scratch(pred.device), self.hist, self.hist_i64, None, size)

    def update_pq(self, gt_masks, pred_masks, iou_threshold=0.5):
        """util/metrics.py:50-91 (instance-mask matching for panoptic quality).  Nothing in the reference calls it and it is not
        part of the per-pixel head-to-metric path this package implements (out of scope)."""
        raise NotImplementedError("Metrics.update_pq is outside the per-pixel head-to-metric path of segmentation_factory_b200; "
                                  "use the reference's util.metrics.Metrics for instance / panoptic quality")

    def compute_pq(self):
        """util/metrics.py:94-106 on whatever `pq_hist` holds (pure host arithmetic)."""
        tp, fp, fn, iou_sum = (self.pq_hist[k] for k in ("TP", "FP", "FN", "iou_sum"))
        sq = iou_sum / tp if tp > 0 else 0.0
        den = tp + 0.5 * fp + 0.5 * fn
        rq = tp / den if den > 0 else 0.0
        return round(sq * rq * 100, 2), round(sq * 100, 2), round(rq * 100, 2)

    def _scratch(self, device):
        if self._counts is None or self._counts.device != device:
            self._counts = torch.zeros(self.num_classes, self.num_classes, dtype=torch.int64, device=device)
        return self._counts

    def compute_iou(self):
        ious = self.hist.diag() / (self.hist.sum(0) + self.hist.sum(1) - self.hist.diag())
        miou = ious[~ious.isnan()].mean().item()
        ious *= 100
        miou *= 100
        return ious.cpu().numpy().round(2).tolist(), round(miou, 2)

    def compute_f1(self):
        f1 = 2 * self.hist.diag() / (self.hist.sum(0) + self.hist.sum(1))
        mf1 = f1[~f1.isnan()].mean().item()
        f1 *= 100
        mf1 *= 100
        return f1.cpu().numpy().round(2).tolist(), round(mf1, 2)

    def compute_pixel_acc(self):
        acc = self.hist.diag() / self.hist.sum(1)
        macc = acc[~acc.isnan()].mean().item()
        acc *= 100
        macc *= 100
        return acc.cpu().numpy().round(2).tolist(), round(macc, 2)

    def reduce_from_all_processes(self):
        if not dist.is_available() or not dist.is_initialized():
            return
        dist.all_reduce(self.hist_i64)        # one int64 [C, C] SUM over NVLink; no barrier needed
        self.hist = self.hist_i64.to(torch.float32)


class ConfusionMatrix(object):
    """util/utils.py:94-143 (int64 matrix; the source of the reported mIoU and of best-checkpoint selection)."""

    def __init__(self, num_classes):
        self.num_classes = num_classes
        self.mat = None

    def _ensure(self, device):
        if self.mat is None:
            n = self.num_classes
            self.mat = torch.zeros((n, n), dtype=torch.int64, device=device)

    def update(self, a, b):
        """a: flattened labels, b: flattened predicted class indices (engine.py:90)."""
        self._ensure(a.device)
        with torch.no_grad():
            hist_pairs(a, b, self.num_classes, self.mat)

    def update_from_logits(self, labels: Tensor, logits: Tensor, ignore_label: int = 255, size=None):
        """Fused form: argmax of the (upsampled) logits is taken in-kernel, nothing is materialised."""
        self._ensure(logits.device)
        with torch.no_grad():
            eval_hist(logits, labels, ignore_label, self.mat, size)

    def reset(self):
        if self.mat is not None:
            self.mat.zero_()

    def compute(self):
        h = self.mat.float()
        acc_global = torch.diag(h).sum() / h.sum()
        acc = torch.diag(h) / h.sum(1)
        iu = torch.diag(h) / (h.sum(1) + h.sum(0) - torch.diag(h))
        return acc_global, acc, iu

    def reduce_from_all_processes(self):
        if not dist.is_available() or not dist.is_initialized():
            return
        dist.all_reduce(self.mat)

    def __str__(self):
        acc_global, acc, iu = self.compute()
        return (
            'global correct: {:.1f}\n'
            'average row correct: {}\n'
            'IoU: {}\n'
            'mean IoU: {:.1f}').format(
                acc_global.item() * 100,
                ['{:.1f}'.format(i) for i in (acc * 100).tolist()],
                ['{:.1f}'.format(i) for i in (iu * 100).tolist()],
                iu.mean().item() * 100)


def reduce_from_all_processes(metric: Metrics, confmat: ConfusionMatrix) -> None:
    """engine.py:100-101 (confmat.reduce_from_all_processes(); metric.reduce_from_all_processes()) as ONE collective:
    both int64 matrices travel in a single packed all_reduce(SUM) -- the message is a few KB, so the cost is latency and
    two collectives cost twice that.  Integer SUM is order-independent, so every rank ends with identical matrices."""
    if not dist.is_available() or not dist.is_initialized():
        return
    confmat._ensure(metric.hist_i64.device)
    n = metric.hist_i64.numel()
    packed = torch.cat([metric.hist_i64.reshape(-1), confmat.mat.reshape(-1).to(metric.hist_i64.device)])
    dist.all_reduce(packed)
    metric.hist_i64.copy_(packed[:n].view_as(metric.hist_i64))
    metric.hist = metric.hist_i64.to(torch.float32)
    confmat.mat.copy_(packed[n:].view_as(confmat.mat))


def fused_eval_update(logits: Tensor, labels: Tensor, metric: Metrics, confmat: ConfusionMatrix, size=None) -> None:
    """One kernel launch feeding both accumulators of engine.evaluate (engine.py:90-91): the reference takes
    two argmax passes over materialised full-resolution logits and builds two histograms.

    The two reference accumulators mask differently: Metrics drops `ignore_label`, ConfusionMatrix keeps every label in
    [0, C).  They agree exactly when the ignore label lies outside [0, C) (255 everywhere in the reference); an in-range
    ignore label would silently drop that row from the confusion matrix, so it is refused here -- use
    confmat.update(a, b) + metric.update(...) for that case."""
    if 0 <= int(metric.ignore_label) < int(confmat.num_classes):
        raise ValueError(f"fused_eval_update needs an ignore label outside [0, {confmat.num_classes}); "
                         f"got {metric.ignore_label}: update the two accumulators separately")
    confmat._ensure(logits.device)
    eval_update(logits, labels, metric.ignore_label, metric._scratch(logits.device), metric.hist, metric.hist_i64, confmat.mat,
                size)                                                       # one C-ABI call, 2 launches per batch
