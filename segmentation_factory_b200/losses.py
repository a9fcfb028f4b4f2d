"""Drop-in mirror of the reference's util/losses.py loss classes for this path.

Same constructor signatures, same forward(preds, labels) contract (tuple preds -> aux-weighted sum),
same return value (0-dim fp32 tensor).  Extension: `preds` may have a smaller spatial size than
`labels`; the bilinear upsample of models/build_models.py:65 then happens inside the kernel and the
full-resolution logits are never materialised.  With equal sizes the result equals the reference's.
"""
from __future__ import annotations

import torch
from torch import Tensor, nn

from .functional import fused_cross_entropy, fused_focal_loss

__all__ = ["CrossEntropy", "OhemCrossEntropy", "FocalLoss", "get_loss"]


class CrossEntropy(nn.Module):
    """util/losses.py:28-41 -- nn.CrossEntropyLoss(weight, ignore_index) mean; aux_weights [1, 0.4, 0.4]."""

    def __init__(self, ignore_label: int = 255, weight: Tensor = None, aux_weights: list = [1, 0.4, 0.4]) -> None:
        super().__init__()
        self.ignore_label = ignore_label
        self.aux_weights = aux_weights
        self.weight = weight

    def _forward(self, preds: Tensor, labels: Tensor) -> Tensor:
        return fused_cross_entropy(preds, labels, self.weight, self.ignore_label, None)

    def forward(self, preds, labels: Tensor) -> Tensor:
        if isinstance(preds, tuple):
            return sum([w * self._forward(pred, labels) for (pred, w) in zip(preds, self.aux_weights)])
        return self._forward(preds, labels)


class OhemCrossEntropy(nn.Module):
    """util/losses.py:44-66 -- keep pixels with loss > -log(thresh); if fewer than n_valid//16, the top n_min."""

    def __init__(self, ignore_label: int = 255, weight: Tensor = None, thresh: float = 0.7,
                 aux_weights: list = [1, 1]) -> None:
        super().__init__()
        self.ignore_label = ignore_label
        self.aux_weights = aux_weights
        self.weight = weight
        # same arithmetic as the reference: -log of the fp32 value, kept as a 0-dim CPU fp32 tensor
        self.thresh = -torch.log(torch.tensor(thresh, dtype=torch.float))

    def _forward(self, preds: Tensor, labels: Tensor) -> Tensor:
        return fused_cross_entropy(preds, labels, self.weight, self.ignore_label, float(self.thresh))

    def forward(self, preds, labels: Tensor) -> Tensor:
        if isinstance(preds, tuple):
            return sum([w * self._forward(pred, labels) for (pred, w) in zip(preds, self.aux_weights)])
        return self._forward(preds, labels)


class FocalLoss(nn.Module):
    """util/losses.py:9-25: an epilogue on the per-pixel cross-entropy.
    Same constructor; the mean runs over all pixels, ignored ones included as zeros, like the reference."""

    def __init__(self, alpha=1, gamma=0, size_average=True, ignore_index=255):
        super().__init__()
        self.alpha = alpha
        self.gamma = gamma
        self.ignore_index = ignore_index
        self.size_average = size_average

    def forward(self, inputs, targets):
        return fused_focal_loss(inputs, targets, self.alpha, self.gamma, self.size_average, self.ignore_index)


def get_loss(loss_fn_name: str = "CrossEntropy", ignore_label: int = 255, cls_weights: Tensor = None):
    """util/losses.py:118-122.  Dice is outside this path; the reference's own get_loss would
    call FocalLoss(ignore_label, cls_weights) positionally, i.e. as (alpha, gamma) -- kept as is."""
    table = {"CrossEntropy": CrossEntropy, "OhemCrossEntropy": OhemCrossEntropy, "FocalLoss": FocalLoss}
    assert loss_fn_name in table, f"Unavailable loss function name >> {loss_fn_name}.\nAvailable loss functions: {list(table)}"
    return table[loss_fn_name](ignore_label, cls_weights)
