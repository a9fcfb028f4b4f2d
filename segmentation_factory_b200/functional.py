"""Tensor-level entry points over the C ABI: pointers, shapes and the current CUDA stream.

PyTorch is used for device memory, streams and autograd bookkeeping only; all per-pixel work
happens in libsegfactory.so.  Nothing here synchronises with the host.
"""
from __future__ import annotations

import math
import os

import torch
from torch.autograd.function import once_differentiable

from . import _lib

_LOGIT_CODE = {torch.float32: _lib.SF_F32, torch.bfloat16: _lib.SF_BF16, torch.float16: _lib.SF_F16}
_LABEL_CODE = {torch.int64: _lib.SF_I64, torch.uint8: _lib.SF_U8}


def _stream(device) -> int:
    """Raw cudaStream_t of torch's current stream on `device` (torch.cuda.current_stream() builds a Stream object: 3.5 us)."""
    idx = device.index
    return torch._C._cuda_getCurrentRawStream(idx if idx is not None else torch.cuda.current_device())


class _on_device:
    """Make `device` current around a C-ABI call, skipping the guard (2 us) when it already is -- the per-call host cost
    matters at validation batch size 1 (engine.py:83-91 runs one image per step)."""
    __slots__ = ("idx", "prev")

    def __init__(self, device):
        self.idx = device.index
        self.prev = -1

    def __enter__(self):
        if self.idx is not None:
            cur = torch.cuda.current_device()
            if cur != self.idx:
                self.prev = cur
                torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.prev >= 0:
            torch.cuda.set_device(self.prev)
        return False


def _ptr(t):
    return None if t is None else t.data_ptr()


def prep_logits(logits: torch.Tensor) -> torch.Tensor:
    if not isinstance(logits, torch.Tensor) or logits.dim() != 4:
        raise ValueError("logits must be a [B, C, h, w] tensor")
    if not logits.is_cuda:
        raise _lib.SegFactoryError("segmentation_factory_b200 has no CPU path: logits must be a CUDA tensor")
    if logits.dtype not in _LOGIT_CODE:
        raise TypeError(f"logits dtype {logits.dtype} not supported (float32, bfloat16, float16)")
    return logits.contiguous()


def prep_labels(labels: torch.Tensor, B: int, size=None) -> torch.Tensor:
    """Labels as [B, H, W] int64/uint8 on the device.  A flattened vector (engine.py:91 passes
    labels.flatten()) is reshaped with `size`, which defaults to the logits' own spatial size."""
    if not labels.is_cuda:
        raise _lib.SegFactoryError("segmentation_factory_b200 has no CPU path: labels must be a CUDA tensor")
    if labels.dtype not in _LABEL_CODE:
        labels = labels.long()
    if labels.dim() != 3:
        if size is None:
            raise ValueError("flattened labels need size=(H, W)")
        H, W = size
        if labels.numel() != B * H * W:
            raise ValueError(f"labels have {labels.numel()} elements, expected {B}*{H}*{W}")
        labels = labels.reshape(B, H, W)
    if labels.shape[0] != B:
        raise ValueError(f"batch mismatch: logits {B}, labels {labels.shape[0]}")
    return labels.contiguous()


def _class_w(weight, C, device):
    if weight is None:
        return None
    w = weight.detach().to(device=device, dtype=torch.float32).contiguous()
    if w.numel() != C:
        raise ValueError(f"class weight has {w.numel()} entries, expected {C}")
    return w


def eval_hist(logits: torch.Tensor, labels: torch.Tensor, ignore: int, hist: torch.Tensor, size=None) -> None:
    """hist[C, C] (int64, device) += confusion counts of argmax(upsample(logits)) vs labels."""
    logits = prep_logits(logits)
    B, C, h, w = logits.shape
    labels = prep_labels(labels, B, size if size is not None else (h, w))
    if hist.dtype != torch.int64 or not hist.is_contiguous() or hist.numel() != C * C or hist.device != logits.device:
        raise ValueError("hist must be a contiguous int64 [C, C] tensor on the logits' device")
    _, H, W = labels.shape
    with _on_device(logits.device):
        rc = _lib.lib().sf_eval_hist(_ptr(logits), _LOGIT_CODE[logits.dtype], _ptr(labels), _LABEL_CODE[labels.dtype],
                                     B, C, h, w, H, W, int(ignore), _ptr(hist), _stream(logits.device))
    _lib.check(rc, "sf_eval_hist")


def _check_cells(name, t, dtype, like) -> None:
    """A raw pointer is about to be handed to a kernel: a host tensor, a tensor on another GPU, a strided view or a
    wrong dtype would be an illegal address (sticky context error) instead of a Python exception."""
    if t is None:
        return
    if not t.is_cuda or t.device != like.device or t.dtype != dtype or not t.is_contiguous() or t.numel() != like.numel():
        raise ValueError(f"{name} must be a contiguous {dtype} tensor with {like.numel()} cells on {like.device} "
                         f"(got {t.dtype}, {tuple(t.shape)}, {t.device})")


def eval_update(logits, labels, ignore, counts, hist_f32=None, hist_i64=None, mat=None, size=None) -> None:
    """eval_hist + accumulate_counts through ONE C-ABI call (sf_eval_update): `counts` is the zeroed int64 [C, C] scratch."""
    logits = prep_logits(logits)
    B, C, h, w = logits.shape
    labels = prep_labels(labels, B, size if size is not None else (h, w))
    if counts.dtype != torch.int64 or not counts.is_contiguous() or counts.numel() != C * C or counts.device != logits.device:
        raise ValueError("counts must be a contiguous int64 [C, C] tensor on the logits' device")
    for name, t, dt in (("hist_f32", hist_f32, torch.float32), ("hist_i64", hist_i64, torch.int64), ("mat", mat, torch.int64)):
        _check_cells(name, t, dt, counts)
    _, H, W = labels.shape
    with _on_device(logits.device):
        rc = _lib.lib().sf_eval_update(_ptr(logits), _LOGIT_CODE[logits.dtype], _ptr(labels), _LABEL_CODE[labels.dtype],
                                       B, C, h, w, H, W, int(ignore), _ptr(counts), _ptr(hist_f32), _ptr(hist_i64), _ptr(mat),
                                       _stream(logits.device))
    _lib.check(rc, "sf_eval_update")


def hist_pairs(a: torch.Tensor, b: torch.Tensor, C: int, mat: torch.Tensor) -> None:
    """mat[C, C] (int64) += bincount(C * a[k] + b[k]) with k = (a >= 0) & (a < C)  (util/utils.py:106-109)."""
    if not (a.is_cuda and b.is_cuda):
        raise _lib.SegFactoryError("segmentation_factory_b200 has no CPU path: index vectors must be CUDA tensors")
    a = a.reshape(-1)
    b = b.reshape(-1)
    if a.dtype not in _LABEL_CODE:
        a = a.long()
    a = a.contiguous()
    b = b.long().contiguous()
    if a.numel() != b.numel():
        raise ValueError("a and b must have the same number of elements")
    if (not mat.is_cuda or mat.device != a.device or b.device != a.device or mat.dtype != torch.int64
            or not mat.is_contiguous() or mat.numel() != int(C) * int(C)):
        raise ValueError("mat must be a contiguous int64 [C, C] tensor on the device of a and b")
    with _on_device(a.device):
        rc = _lib.lib().sf_hist_pairs(_ptr(a), _LABEL_CODE[a.dtype], _ptr(b), a.numel(), int(C), _ptr(mat), _stream(a.device))
    _lib.check(rc, "sf_hist_pairs")


def accumulate_counts(counts, hist_f32=None, hist_i64=None, mat=None) -> None:
    """One launch for the per-call bookkeeping of Metrics / ConfusionMatrix: hist_f32 += counts (float32 += int64, the
    reference's expression), hist_i64 += counts, mat += counts, counts = 0."""
    if not counts.is_cuda or counts.dtype != torch.int64 or not counts.is_contiguous():
        raise ValueError("counts must be a contiguous int64 CUDA tensor")
    for name, t, dt in (("hist_f32", hist_f32, torch.float32), ("hist_i64", hist_i64, torch.int64), ("mat", mat, torch.int64)):
        _check_cells(name, t, dt, counts)
    with _on_device(counts.device):
        rc = _lib.lib().sf_accumulate_counts(_ptr(counts), _ptr(hist_f32), _ptr(hist_i64), _ptr(mat), counts.numel(),
                                             _stream(counts.device))
    _lib.check(rc, "sf_accumulate_counts")


def _onepass_enabled() -> bool:
    """SF_NO_ONEPASS=1 forces the forward / finalize / backward kernel route (A/B measurement, tests)."""
    return os.environ.get("SF_NO_ONEPASS", "0") != "1"


class _BranchHint:
    """Last OHEM decision seen on a device (sf_loss_state.branch), learnt WITHOUT a host sync: every backward enqueues a
    4-byte copy of the state into pinned memory behind an event; the next forward uses the value only if that event has
    already completed.  The one-pass walk speculates on the threshold branch (util/losses.py:56); when training has
    settled on the top-k branch (:58-59) its gradient sums would be thrown away every step, so those steps take the
    forward / finalize / backward kernels instead.  A stale or missing hint only costs time, never correctness: both
    routes handle both branches (their results agree to float rounding, not bit for bit: SF_NO_BRANCH_HINT=1 pins the route
    when bitwise run-to-run reproducibility across a branch change matters)."""
    _by_device = {}

    def __init__(self, dev):
        self.host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.event = None

    @classmethod
    def last_branch(cls, dev):
        if os.environ.get("SF_NO_BRANCH_HINT", "0") == "1":          # fixed route: bitwise-identical steps regardless of history
            return None
        h = cls._by_device.get(dev)
        if h is None or h.event is None or torch.cuda.is_current_stream_capturing() or not h.event.query():
            return None
        return int(h.host[0])

    @classmethod
    def record(cls, dev, ws):
        if torch.cuda.is_current_stream_capturing():
            return
        h = cls._by_device.get(dev)
        if h is None:
            h = cls._by_device[dev] = cls(dev)
        h.host.copy_(ws[16:20].view(torch.int32), non_blocking=True)     # sf_loss_state.branch
        h.event = torch.cuda.Event()
        h.event.record(torch.cuda.current_stream(dev))


class _FusedCE(torch.autograd.Function):
    """upsample + (OHEM) cross-entropy.

    Training (a gradient is wanted, shape covered by sf_onepass.cu): forward = sf_ce_onepass + sf_loss_finalize -- one walk
    computes loss, confusion counts and the unscaled gradient -- and backward = sf_ce_onepass_backward (the global 1/K
    scale, or the top-k re-walk when the device-side OHEM decision was branch 2).  Otherwise forward = sf_ce_forward +
    sf_loss_finalize and backward = sf_ce_backward.  Second-order differentiation is not supported (once_differentiable):
    the reference's `create_graph=is_second_order` path (engine.py:50-53, timm adahessian) raises here instead of silently
    dropping the loss-curvature term.
    """

    @staticmethod
    def forward(ctx, logits, labels, class_w, ignore, thresh, hist, focal=None):
        L = _lib.lib()
        B, C, h, w = logits.shape
        _, H, W = labels.shape
        dev = logits.device
        n_px = B * H * W
        ohem = thresh >= 0.0
        need_grad = ctx.needs_input_grad[0]
        keep_loss = ohem or focal is not None            # focal: the per-pixel CE is what the epilogue transforms
        ws_bytes = L.sf_workspace_bytes()
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        loss_px = torch.empty(n_px, dtype=torch.float32, device=dev) if keep_loss else None
        out = torch.empty((), dtype=torch.float32, device=dev)
        lc, yc = _LOGIT_CODE[logits.dtype], _LABEL_CODE[labels.dtype]
        onepass = (need_grad and focal is None and _onepass_enabled()
                   and not (ohem and _BranchHint.last_branch(dev) == 2)
                   and L.sf_ce_onepass_supported(lc, _ptr(logits), C, h, w, H, W) == 1)
        lse_px = grad_acc = None
        with _on_device(dev):
            st = _stream(dev)
            if onepass:
                grad_acc = torch.empty((B, C, h, w), dtype=torch.float32, device=dev)
                rc = L.sf_ce_onepass(_ptr(logits), lc, _ptr(labels), yc, B, C, h, w, H, W, int(ignore), _ptr(class_w),
                                     float(thresh), _ptr(loss_px), _ptr(hist), _ptr(grad_acc), _ptr(ws), ws_bytes, st)
                _lib.check(rc, "sf_ce_onepass")
            else:
                lse_px = torch.empty(n_px, dtype=torch.float32, device=dev) if need_grad else None
                rc = L.sf_ce_forward(_ptr(logits), lc, _ptr(labels), yc, B, C, h, w, H, W, int(ignore), _ptr(class_w),
                                     float(thresh), _ptr(loss_px), _ptr(lse_px), _ptr(hist), _ptr(ws), ws_bytes, st)
                _lib.check(rc, "sf_ce_forward")
            if focal is None:
                rc = L.sf_loss_finalize(_ptr(loss_px), n_px, float(thresh), _ptr(ws), ws_bytes, _ptr(out), st)
                _lib.check(rc, "sf_loss_finalize")
            else:
                alpha, gamma, size_average = focal
                rc = L.sf_focal_finalize(_ptr(loss_px), n_px, float(alpha), float(gamma), int(bool(size_average)),
                                         _ptr(ws), ws_bytes, _ptr(out), st)
                _lib.check(rc, "sf_focal_finalize")
        ctx.ignore = int(ignore)
        ctx.ws = ws                      # plain attribute: the workspace is an opaque device scratch
        ctx.onepass = onepass
        ctx.ohem = ohem
        ctx.save_for_backward(logits, labels, class_w, loss_px, grad_acc if onepass else lse_px)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        logits, labels, class_w, loss_px, aux = ctx.saved_tensors
        L = _lib.lib()
        B, C, h, w = logits.shape
        _, H, W = labels.shape
        dev = logits.device
        g = grad_out.detach().to(device=dev, dtype=torch.float32).contiguous()
        grad = torch.empty_like(logits)
        lc, yc = _LOGIT_CODE[logits.dtype], _LABEL_CODE[labels.dtype]
        with _on_device(dev):
            if ctx.onepass:
                rc = L.sf_ce_onepass_backward(_ptr(logits), lc, _ptr(labels), yc, B, C, h, w, H, W, ctx.ignore,
                                              _ptr(class_w), _ptr(loss_px), _ptr(aux), _ptr(g), _ptr(ctx.ws), _ptr(grad),
                                              _stream(dev))
                _lib.check(rc, "sf_ce_onepass_backward")
            else:
                rc = L.sf_ce_backward(_ptr(logits), lc, _ptr(labels), yc, B, C, h, w, H, W, ctx.ignore, _ptr(class_w),
                                      _ptr(loss_px), _ptr(aux), _ptr(g), _ptr(ctx.ws), _ptr(grad), _stream(dev))
                _lib.check(rc, "sf_ce_backward")
            if loss_px is not None and ctx.ohem:
                _BranchHint.record(dev, ctx.ws)
        return grad, None, None, None, None, None, None


def fused_cross_entropy(logits, labels, weight=None, ignore_index=255, ohem_thresh=None, hist=None):
    """Scalar loss of CrossEntropy (ohem_thresh=None) or OhemCrossEntropy (ohem_thresh = -log(0.7))
    on logits that may be lower-resolution than the labels.  `hist`, if given, is an int64 [C, C]
    tensor that additionally receives the confusion counts of the same pass."""
    logits = prep_logits(logits)
    B, C, h, w = logits.shape
    labels = prep_labels(labels, B, (h, w))
    cw = _class_w(weight, C, logits.device)
    thresh = -1.0 if ohem_thresh is None else float(ohem_thresh)
    if ohem_thresh is not None and not (thresh >= 0.0 and math.isfinite(thresh)):
        raise ValueError("ohem_thresh must be a finite non-negative loss value")
    if hist is not None and (hist.dtype != torch.int64 or hist.numel() != C * C or not hist.is_contiguous()
                             or hist.device != logits.device):
        raise ValueError("hist must be a contiguous int64 [C, C] tensor on the logits' device")
    return _FusedCE.apply(logits, labels, cw, int(ignore_index), thresh, hist, None)


def fused_focal_loss(logits, labels, alpha=1.0, gamma=0.0, size_average=True, ignore_index=255):
    """FocalLoss of util/losses.py:9-25 on logits that may be lower-resolution than the labels:
    alpha * (1 - exp(-ce))**gamma * ce per pixel (ce = 0 where ignored), mean over ALL pixels or sum."""
    logits = prep_logits(logits)
    B, C, h, w = logits.shape
    labels = prep_labels(labels, B, (h, w))
    return _FusedCE.apply(logits, labels, None, int(ignore_index), -1.0, None, (float(alpha), float(gamma), bool(size_average)))


class _FusedCriterion(torch.autograd.Function):
    """engine.criterion with dice=True: mean CE + multiclass dice loss.  Forward = sf_ce_dice_forward + sf_loss_finalize +
    sf_dice_finalize (+ sf_dice_coef when a gradient is wanted), backward = sf_dice_backward."""

    @staticmethod
    def forward(ctx, logits, labels, class_w, ignore, eps):
        L = _lib.lib()
        B, C, h, w = logits.shape
        _, H, W = labels.shape
        dev = logits.device
        n_px = B * H * W
        need_grad = ctx.needs_input_grad[0]
        ws_bytes = L.sf_workspace_bytes()
        dws_bytes = L.sf_dice_workspace_bytes(B, C)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        dws = torch.empty(dws_bytes, dtype=torch.uint8, device=dev)
        lse_px = torch.empty(n_px, dtype=torch.float32, device=dev)
        out = torch.empty((), dtype=torch.float32, device=dev)
        geo = (B, C, h, w, H, W, int(ignore))
        coef = None
        with _on_device(dev):
            st = _stream(dev)
            lc, yc = _LOGIT_CODE[logits.dtype], _LABEL_CODE[labels.dtype]
            rc = L.sf_ce_dice_forward(_ptr(logits), lc, _ptr(labels), yc, *geo, _ptr(class_w), _ptr(lse_px),
                                      _ptr(ws), ws_bytes, _ptr(dws), dws_bytes, st)
            _lib.check(rc, "sf_ce_dice_forward")
            rc = L.sf_loss_finalize(None, n_px, -1.0, _ptr(ws), ws_bytes, None, st)
            _lib.check(rc, "sf_loss_finalize")
            rc = L.sf_dice_finalize(B, C, float(eps), _ptr(ws), _ptr(dws), dws_bytes, _ptr(out), st)
            _lib.check(rc, "sf_dice_finalize")
            if need_grad:
                coef = torch.empty((2, n_px), dtype=torch.float32, device=dev)
                rc = L.sf_dice_coef(_ptr(logits), lc, _ptr(labels), yc, *geo, _ptr(class_w), _ptr(lse_px), _ptr(ws),
                                    _ptr(dws), dws_bytes, _ptr(coef[0]), _ptr(coef[1]), st)
                _lib.check(rc, "sf_dice_coef")
        ctx.geo = geo
        ctx.dws = dws
        ctx.save_for_backward(logits, labels, lse_px, coef)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        logits, labels, lse_px, coef = ctx.saved_tensors
        L = _lib.lib()
        dev = logits.device
        g = grad_out.detach().to(device=dev, dtype=torch.float32).contiguous()
        grad = torch.empty_like(logits)
        with _on_device(dev):
            rc = L.sf_dice_backward(_ptr(logits), _LOGIT_CODE[logits.dtype], _ptr(labels), _LABEL_CODE[labels.dtype],
                                    *ctx.geo, _ptr(lse_px), _ptr(coef[0]), _ptr(coef[1]), _ptr(ctx.dws), ctx.dws.numel(),
                                    _ptr(g), _ptr(grad), _stream(dev))
        _lib.check(rc, "sf_dice_backward")
        return grad, None, None, None, None


def fused_criterion(logits, labels, weight=None, dice=True, ignore_index=255, epsilon=1e-6):
    """engine.criterion (engine.py:10-15): F.cross_entropy(ignore_index, weight) + dice_loss(multiclass) on logits that
    may be lower-resolution than the labels."""
    if not dice:
        return fused_cross_entropy(logits, labels, weight, ignore_index)
    if 0 <= int(ignore_index) <= 1:
        raise ValueError("ignore_index 0 or 1 collides with the one-hot values build_target compares against")
    logits = prep_logits(logits)
    B, C, h, w = logits.shape
    labels = prep_labels(labels, B, (h, w))
    return _FusedCriterion.apply(logits, labels, _class_w(weight, C, logits.device), int(ignore_index), float(epsilon))
