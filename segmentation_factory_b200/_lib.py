"""ctypes loader for libsegfactory.so (the C ABI in include/segfactory.h).

There is no fallback: if the library has not been built, or a tensor is not on a CUDA device,
the callers raise.  Build with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C segmentation_factory_b200/csrc`.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SEGFACTORY_LIB") or os.path.join(_HERE, "libsegfactory.so")   # override: experiment builds

SF_F32, SF_BF16, SF_F16 = 0, 1, 2
SF_I64, SF_U8 = 0, 1

# every symbol include/segfactory.h declares (tests/test_host_cpu.py checks the .so exports them)
SYMBOLS = (
    "sf_version", "sf_strerror", "sf_workspace_bytes", "sf_eval_hist", "sf_ce_forward",
    "sf_loss_finalize", "sf_focal_finalize", "sf_ce_backward", "sf_hist_pairs", "sf_accumulate_counts",
    "sf_label_lut", "sf_eval_update", "sf_head_classifier", "sf_ce_onepass_supported", "sf_ce_onepass", "sf_ce_onepass_backward", "sf_dice_workspace_bytes", "sf_ce_dice_forward", "sf_dice_stats", "sf_dice_finalize", "sf_dice_coef", "sf_dice_backward",
)

_lib = None


class SegFactoryError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SegFactoryError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built and there is no CPU fallback. "
            "Run `make -C segmentation_factory_b200/csrc` (needs nvcc, targets sm_100a).")
    L = ctypes.CDLL(LIB_PATH)
    vp, i, f, ll, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong, ctypes.c_size_t
    L.sf_version.restype = i
    L.sf_strerror.restype = ctypes.c_char_p
    L.sf_strerror.argtypes = [i]
    L.sf_workspace_bytes.restype = sz
    L.sf_eval_hist.argtypes = [vp, i, vp, i, i, i, i, i, i, i, i, vp, vp]
    L.sf_ce_forward.argtypes = [vp, i, vp, i, i, i, i, i, i, i, i, vp, f, vp, vp, vp, vp, sz, vp]
    L.sf_loss_finalize.argtypes = [vp, ll, f, vp, sz, vp, vp]
    L.sf_focal_finalize.argtypes = [vp, ll, f, f, i, vp, sz, vp, vp]
    L.sf_ce_backward.argtypes = [vp, i, vp, i, i, i, i, i, i, i, i, vp, vp, vp, vp, vp, vp, vp]
    L.sf_ce_onepass_supported.argtypes = [i, vp, i, i, i, i, i]
    L.sf_ce_onepass.argtypes = [vp, i, vp, i, i, i, i, i, i, i, i, vp, f, vp, vp, vp, vp, sz, vp]
    L.sf_ce_onepass_backward.argtypes = [vp, i, vp, i, i, i, i, i, i, i, i, vp, vp, vp, vp, vp, vp, vp]
    for name in ("sf_ce_onepass_supported", "sf_ce_onepass", "sf_ce_onepass_backward"):
        getattr(L, name).restype = i
    L.sf_eval_update.argtypes = [vp, i, vp, i, i, i, i, i, i, i, i, vp, vp, vp, vp, vp]
    L.sf_eval_update.restype = i
    L.sf_head_classifier.argtypes = [vp, vp, vp, vp, i, i, i, i, i, i, i, vp]
    L.sf_head_classifier.restype = i
    L.sf_hist_pairs.argtypes = [vp, i, vp, ll, i, vp, vp]
    L.sf_accumulate_counts.argtypes = [vp, vp, vp, vp, i, vp]
    L.sf_label_lut.argtypes = [vp, ll, vp, vp, i, vp]
    L.sf_label_lut.restype = i
    L.sf_dice_workspace_bytes.restype = sz
    L.sf_dice_workspace_bytes.argtypes = [i, i]
    L.sf_ce_dice_forward.argtypes = [vp, i, vp, i, i, i, i, i, i, i, i, vp, vp, vp, sz, vp, sz, vp]
    L.sf_dice_stats.argtypes = [vp, i, vp, i, i, i, i, i, i, i, i, vp, vp, sz, vp]
    L.sf_dice_finalize.argtypes = [i, i, f, vp, vp, sz, vp, vp]
    L.sf_dice_coef.argtypes = [vp, i, vp, i, i, i, i, i, i, i, i, vp, vp, vp, vp, sz, vp, vp, vp]
    L.sf_dice_backward.argtypes = [vp, i, vp, i, i, i, i, i, i, i, i, vp, vp, vp, vp, sz, vp, vp, vp]
    for name in ("sf_ce_dice_forward", "sf_dice_stats", "sf_dice_finalize", "sf_dice_coef", "sf_dice_backward"):
        getattr(L, name).restype = i
    for name in ("sf_eval_hist", "sf_ce_forward", "sf_loss_finalize", "sf_focal_finalize", "sf_ce_backward", "sf_hist_pairs", "sf_accumulate_counts"):
        getattr(L, name).restype = i
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise SegFactoryError(f"{what} failed: {lib().sf_strerror(rc).decode()} (code {rc})")
