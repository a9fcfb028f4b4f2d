"""The reference path restated with the same torch ops it calls -- TEST INFRASTRUCTURE ONLY.

/root/reference does not exist on the GPU box, so the live differential tests there cannot import the
reference's util/ package.  These few functions issue the same ATen calls in the same order
(models/build_models.py:65, util/losses.py:32-36,49-61, util/metrics.py:25-27, util/utils.py:106-109);
tests/test_oracle_golden.py::test_torch_restatement_matches_golden checks them against the vectors the real
reference produced.  Runs on CPU (bit-exact comparator for the confusion matrix) or on CUDA
(the "torch eager on the same B200" comparator of BASELINE.md section 5.2).
"""
import torch
import torch.nn.functional as F


def upsample(low, size):
    return F.interpolate(low, size=size, mode="bilinear", align_corners=False)


def cross_entropy(up, labels, weight=None, ignore=255):
    return F.cross_entropy(up, labels, weight=weight, ignore_index=ignore)


def ohem_cross_entropy(up, labels, weight=None, ignore=255, thresh=0.7):
    t = -torch.log(torch.tensor(thresh, dtype=torch.float))
    n_min = labels[labels != ignore].numel() // 16
    loss = F.cross_entropy(up, labels, weight=weight, ignore_index=ignore, reduction="none").view(-1)
    hard = loss[loss > t]
    if hard.numel() < n_min:
        hard, _ = loss.topk(n_min)
    return torch.mean(hard)


def confusion_counts(up, labels, C, ignore=255):
    pred = up.argmax(1).flatten()
    tgt = labels.flatten()
    keep = tgt != ignore
    return torch.bincount(tgt[keep] * C + pred[keep], minlength=C * C).view(C, C)


def loss_and_grad(low, labels, kind, weight=None, ignore=255, grad_scale=1.0):
    """(loss, d loss / d low) through interpolate, the way engine.py + autograd would compute it."""
    low = low.detach().clone().requires_grad_(True)
    up = upsample(low, labels.shape[1:]).float()
    lab = labels.long()
    loss = ohem_cross_entropy(up, lab, weight, ignore) if kind == "ohem" else cross_entropy(up, lab, weight, ignore)
    (loss * grad_scale).backward()
    return loss.detach(), low.grad.detach()
