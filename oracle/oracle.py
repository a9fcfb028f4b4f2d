"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (segmentation_factory_b200/) never does.

All functions take/return numpy arrays (torch CPU tensors are converted by the callers).
bf16 / fp16 logits are passed as uint16 bit patterns together with the dtype code.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

F32, BF16, F16 = 0, 1, 2
I64, U8 = 0, 1

_lib = None


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc only; no reference sources involved)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "segpath_oracle.c")):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        vp, i, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        L.orc_upsample.argtypes = [vp, i, i, i, i, i, i, i, vp]
        L.orc_path_forward.argtypes = [vp, i, vp, i, i, i, i, i, i, i, i, vp, f, vp, vp, vp, vp, vp]
        L.orc_ce_reduce.argtypes = [vp]
        L.orc_ce_reduce.restype = f
        L.orc_ohem_reduce.argtypes = [vp, ctypes.c_int64, ctypes.c_int64, f, vp, vp, vp, vp]
        L.orc_path_backward.argtypes = [vp, i, vp, i, i, i, i, i, i, i, i, vp, vp, vp, vp]
        L.orc_hist_pairs.argtypes = [vp, vp, ctypes.c_int64, i, vp]
        L.orc_dice_stats.argtypes = [vp, i, vp, i, i, i, i, i, i, i, i, vp, vp]
        L.orc_dice_backward.argtypes = [vp, i, vp, i, i, i, i, i, i, i, i, vp, vp, ctypes.c_double, vp, vp, vp]
        L.orc_num_threads.restype = i
        L.orc_set_num_threads.argtypes = [i]
        L.orc_set_num_threads.restype = None
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def logits_code(logits: np.ndarray, ldtype: int | None) -> int:
    if ldtype is not None:
        return ldtype
    if logits.dtype == np.float32:
        return F32
    if logits.dtype == np.float16:
        return F16
    raise TypeError("bf16 logits must be passed as uint16 bit patterns with ldtype=BF16")


def labels_code(labels: np.ndarray) -> int:
    if labels.dtype == np.int64:
        return I64
    if labels.dtype == np.uint8:
        return U8
    raise TypeError(f"labels must be int64 or uint8, got {labels.dtype}")


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int) -> None:
    """Override OMP_NUM_THREADS (torch.distributed.run sets it to 1 for its workers)."""
    lib().orc_set_num_threads(int(n))


def upsample(logits: np.ndarray, size, ldtype: int | None = None) -> np.ndarray:
    """F.interpolate(bilinear, align_corners=False) restated; returns fp32 holding dtype-rounded values."""
    logits = np.ascontiguousarray(logits)
    B, C, h, w = logits.shape
    H, W = size
    out = np.empty((B, C, H, W), np.float32)
    rc = lib().orc_upsample(_p(logits), logits_code(logits, ldtype), B, C, h, w, H, W, _p(out))
    assert rc == 0
    return out


def path_forward(logits, labels, ignore=255, class_w=None, thresh=0.0, ldtype=None,
                 want_loss=True, want_lse=True, want_arg=True, want_hist=True):
    """One pass over all label pixels. Returns dict(loss_px, lse_px, arg_px, hist, sums)."""
    logits = np.ascontiguousarray(logits)
    labels = np.ascontiguousarray(labels)
    B, C, h, w = logits.shape
    _, H, W = labels.shape
    n = B * H * W
    loss_px = np.empty(n, np.float32) if want_loss else None
    lse_px = np.empty(n, np.float32) if want_lse else None
    arg_px = np.empty(n, np.int32) if want_arg else None
    hist = np.zeros((C, C), np.int64) if want_hist else None
    sums = np.zeros(4, np.float64)
    cw = None if class_w is None else np.ascontiguousarray(class_w, np.float32)
    rc = lib().orc_path_forward(_p(logits), logits_code(logits, ldtype), _p(labels), labels_code(labels),
                                B, C, h, w, H, W, int(ignore), _p(cw), float(thresh),
                                _p(loss_px), _p(lse_px), _p(arg_px), _p(hist), _p(sums))
    assert rc == 0
    return dict(loss_px=loss_px, lse_px=lse_px, arg_px=arg_px, hist=hist, sums=sums)


def ce_reduce(sums: np.ndarray) -> float:
    return float(lib().orc_ce_reduce(_p(np.ascontiguousarray(sums, np.float64))))


def ohem_reduce(loss_px: np.ndarray, n_valid: int, thresh: float):
    """Returns (loss, kth, kept, branch) following util/losses.py:52-61."""
    loss_px = np.ascontiguousarray(loss_px, np.float32)
    loss = ctypes.c_float()
    kth = ctypes.c_float()
    kept = ctypes.c_int64()
    branch = ctypes.c_int()
    rc = lib().orc_ohem_reduce(_p(loss_px), loss_px.size, int(n_valid), float(thresh),
                               ctypes.byref(loss), ctypes.byref(kth), ctypes.byref(kept), ctypes.byref(branch))
    assert rc == 0
    return loss.value, kth.value, kept.value, branch.value


def path_backward(logits, labels, lse_px, coef_px, ignore=255, class_w=None, ldtype=None) -> np.ndarray:
    logits = np.ascontiguousarray(logits)
    labels = np.ascontiguousarray(labels)
    B, C, h, w = logits.shape
    _, H, W = labels.shape
    grad = np.empty((B, C, h, w), np.float32)
    cw = None if class_w is None else np.ascontiguousarray(class_w, np.float32)
    rc = lib().orc_path_backward(_p(logits), logits_code(logits, ldtype), _p(labels), labels_code(labels),
                                 B, C, h, w, H, W, int(ignore), _p(cw),
                                 _p(np.ascontiguousarray(lse_px, np.float32)),
                                 _p(np.ascontiguousarray(coef_px, np.float32)), _p(grad))
    assert rc == 0
    return grad


def hist_pairs(a: np.ndarray, b: np.ndarray, C: int) -> np.ndarray:
    a = np.ascontiguousarray(a, np.int64)
    b = np.ascontiguousarray(b, np.int64)
    mat = np.zeros((C, C), np.int64)
    rc = lib().orc_hist_pairs(_p(a), _p(b), a.size, C, _p(mat))
    assert rc == 0, "prediction index outside [0, C)"
    return mat


# ---------------------------------------------------------------------------------------------
# Whole-path helpers (what the tests compare the CUDA path with)
# ---------------------------------------------------------------------------------------------
OHEM_THRESH = float(-np.log(np.float32(0.7)))   # util/losses.py:49  (-log(float32(0.7)), fp32)


def ce_loss_and_grad(logits, labels, ignore=255, class_w=None, ldtype=None, grad_out=1.0, want_grad=True):
    """CrossEntropy (util/losses.py:28-41) through the low-res logits."""
    f = path_forward(logits, labels, ignore, class_w, 0.0, ldtype, want_arg=False, want_hist=False)
    loss = ce_reduce(f["sums"])
    grad = None
    if want_grad:
        coef = np.full(f["loss_px"].shape, grad_out / f["sums"][1], np.float32)   # d mean / d loss_px = 1 / sum(w_y)
        grad = path_backward(logits, labels, f["lse_px"], coef, ignore, class_w, ldtype)
    return loss, grad


def ohem_loss_and_grad(logits, labels, ignore=255, class_w=None, thresh=OHEM_THRESH, ldtype=None,
                       grad_out=1.0, want_grad=True):
    """OhemCrossEntropy (util/losses.py:44-66). Ties at the k-th value share the remaining slots equally."""
    f = path_forward(logits, labels, ignore, class_w, thresh, ldtype, want_arg=False, want_hist=False)
    loss, kth, kept, branch = ohem_reduce(f["loss_px"], int(f["sums"][2]), thresh)
    grad = None
    if want_grad:
        lp = f["loss_px"]
        if branch == 0:
            coef = np.where(lp > thresh, np.float32(grad_out / kept), 0).astype(np.float32)
        else:
            gt = lp > kth
            tie = lp == kth
            share = (kept - int(gt.sum())) / max(int(tie.sum()), 1)
            coef = (gt.astype(np.float64) + tie.astype(np.float64) * share) * (grad_out / kept)
            coef = coef.astype(np.float32)
        grad = path_backward(logits, labels, f["lse_px"], coef, ignore, class_w, ldtype)
    return loss, grad, dict(kth=kth, kept=kept, branch=branch, n_valid=int(f["sums"][2]), n_hard=int(f["sums"][3]))


def focal_loss_and_grad(logits, labels, alpha=1.0, gamma=0.0, size_average=True, ignore=255, ldtype=None, want_grad=True):
    """FocalLoss (util/losses.py:9-25): alpha * (1 - exp(-ce))**gamma * ce, mean over ALL pixels (ignored ones are
    zeros) or sum.  ce is the unweighted per-pixel cross-entropy; the gradient flows through ce only."""
    f = path_forward(logits, labels, ignore, None, 0.0, ldtype, want_arg=False, want_hist=False)
    ce = f["loss_px"].astype(np.float64)
    pt = np.exp(-ce)
    om = 1.0 - pt
    with np.errstate(divide="ignore", invalid="ignore"):
        w = np.where(om > 0, om ** gamma, 1.0 if gamma == 0 else 0.0)
        dw = np.where(om > 0, gamma * om ** (gamma - 1.0) * pt, 0.0)         # d (1-pt)^g / d ce
    val = alpha * w * ce
    denom = ce.size if size_average else 1.0
    loss = float(val.sum() / denom)
    grad = None
    if want_grad:
        coef = (alpha * (w + ce * dw) / denom).astype(np.float32)
        grad = path_backward(logits, labels, f["lse_px"], coef, ignore, None, ldtype)
    return loss, grad


def dice_stats(logits, labels, lse_px, ignore=255, ldtype=None) -> np.ndarray:
    """[B, C, 3] double: sum p_c, sum p_c [y == c], count [y == c] over the non-ignored pixels of each image."""
    logits = np.ascontiguousarray(logits)
    labels = np.ascontiguousarray(labels)
    B, C, h, w = logits.shape
    _, H, W = labels.shape
    stats = np.zeros((B, C, 3), np.float64)
    rc = lib().orc_dice_stats(_p(logits), logits_code(logits, ldtype), _p(labels), labels_code(labels),
                              B, C, h, w, H, W, int(ignore), _p(np.ascontiguousarray(lse_px, np.float32)), _p(stats))
    assert rc == 0
    return stats


def dice_from_stats(stats: np.ndarray, eps: float = 1e-6):
    """util/losses.py:147-177 on the per-image sums: returns (dice_loss, ta, tb) with the backward tables of
    orc_dice_backward (zero where sets_sum == 0 makes the coefficient a constant)."""
    B, C, _ = stats.shape
    sp, si, sn = stats[..., 0], stats[..., 1], stats[..., 2]
    sets = sp + sn
    dead = sets == 0
    den = np.where(dead, 2 * si, sets) + eps
    D = (2 * si + eps) / den
    loss = 1.0 - D.sum() / (B * C)
    ta = np.where(dead, 0.0, 2.0 / den) / (B * C)
    tb = np.where(dead, 0.0, D / den) / (B * C)
    return float(loss), np.ascontiguousarray(ta), np.ascontiguousarray(tb)


def criterion_loss_and_grad(logits, labels, class_w=None, dice=True, ignore=255, ldtype=None, want_grad=True, eps=1e-6):
    """engine.criterion (engine.py:10-15): mean CE (+ multiclass dice loss) through the low-res logits."""
    f = path_forward(logits, labels, ignore, class_w, 0.0, ldtype, want_arg=False, want_hist=False)
    ce = ce_reduce(f["sums"])
    if not dice:
        loss, grad = ce_loss_and_grad(logits, labels, ignore, class_w, ldtype, want_grad=want_grad)
        return loss, grad, None
    stats = dice_stats(logits, labels, f["lse_px"], ignore, ldtype)
    dl, ta, tb = dice_from_stats(stats, eps)
    grad = None
    if want_grad:
        logits_c = np.ascontiguousarray(logits)
        labels_c = np.ascontiguousarray(labels)
        B, C, h, w = logits_c.shape
        _, H, W = labels_c.shape
        grad = np.empty((B, C, h, w), np.float32)
        cw = None if class_w is None else np.ascontiguousarray(class_w, np.float32)
        rc = lib().orc_dice_backward(_p(logits_c), logits_code(logits_c, ldtype), _p(labels_c), labels_code(labels_c),
                                     B, C, h, w, H, W, int(ignore), _p(cw), _p(f["lse_px"]),
                                     1.0 / f["sums"][1], _p(ta), _p(tb), _p(grad))
        assert rc == 0
    return ce + dl, grad, dict(stats=stats, ce=ce, dice=dl)
