"""The decode head's 1x1 classifier on tcgen05: drop-in for the `nn.Conv2d(embed_dim,
num_classes, 1)` that ends SegFormerHead / UPerHead / FPNHead (models/heads/segformer.py:39,57; upernet.py:32,50;
fpn.py:23,37), forward only.  It writes NCHW logits in the dtype the loss / metric kernels read (bf16 / fp16 / fp32), so
`logits = classifier(features); loss = OhemCrossEntropy()(logits, labels)` involves no conversion pass.
"""
from __future__ import annotations

import torch
from torch import Tensor, nn

from . import _lib
from .functional import _LOGIT_CODE, _on_device, _ptr, _stream

__all__ = ["head_classifier", "HeadClassifier"]


def head_classifier(x: Tensor, weight: Tensor, bias: Tensor | None = None, out_dtype: torch.dtype | None = None) -> Tensor:
    """logits[b, c, i, j] = bias[c] + sum_k weight[c, k] * x[b, k, i, j]   (F.conv2d with a 1x1 kernel).

    x [B, K, h, w] and weight [C, K] or [C, K, 1, 1] in bf16 / fp16 on a CUDA device; bias [C] (any float dtype) or None;
    the result is [B, C, h, w] in `out_dtype` (default: x.dtype).  No autograd: use it under torch.no_grad() /
    inference_mode (validation, export) or for the forward of a frozen head."""
    if not (isinstance(x, Tensor) and x.dim() == 4):
        raise ValueError("x must be a [B, K, h, w] tensor")
    if not x.is_cuda:
        raise _lib.SegFactoryError("segmentation_factory_b200 has no CPU path: x must be a CUDA tensor")
    if x.dtype not in (torch.bfloat16, torch.float16):
        raise TypeError(f"head_classifier takes bf16 / fp16 activations (got {x.dtype}); keep torch's conv for fp32")
    if torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad or (bias is not None and bias.requires_grad)):
        raise RuntimeError("head_classifier is forward-only: call it under torch.no_grad() or detach its inputs")
    B, K, h, w = x.shape
    wt = weight.detach()
    if wt.dim() == 4:
        if wt.shape[2:] != (1, 1):
            raise ValueError("weight must be a 1x1 convolution kernel")
        wt = wt.reshape(wt.shape[0], wt.shape[1])
    if wt.dim() != 2 or wt.shape[1] != K:
        raise ValueError(f"weight must be [C, {K}] (or [C, {K}, 1, 1])")
    C = wt.shape[0]
    wt = wt.to(device=x.device, dtype=x.dtype).contiguous()
    b32 = None if bias is None else bias.detach().to(device=x.device, dtype=torch.float32).contiguous()
    if b32 is not None and b32.numel() != C:
        raise ValueError(f"bias has {b32.numel()} entries, expected {C}")
    x = x.detach().contiguous()
    out_dtype = x.dtype if out_dtype is None else out_dtype
    if out_dtype not in _LOGIT_CODE:
        raise TypeError(f"out_dtype {out_dtype} not supported (float32, bfloat16, float16)")
    out = torch.empty((B, C, h, w), dtype=out_dtype, device=x.device)
    with _on_device(x.device):
        rc = _lib.lib().sf_head_classifier(_ptr(x), _ptr(wt), _ptr(b32), _ptr(out), _LOGIT_CODE[x.dtype], _LOGIT_CODE[out_dtype],
                                           B, C, K, h, w, _stream(x.device))
    _lib.check(rc, "sf_head_classifier")
    return out


class HeadClassifier(nn.Module):
    """nn.Conv2d(in_channels, num_classes, 1) with the same parameter names and shapes (`weight` [C, K, 1, 1], `bias` [C]),
    so a head's state_dict entry (`linear_pred.*`, `conv_seg.*`) loads unchanged; forward runs sf_head_classifier."""

    def __init__(self, in_channels: int, num_classes: int, bias: bool = True, out_dtype: torch.dtype | None = None):
        super().__init__()
        ref = nn.Conv2d(in_channels, num_classes, 1, bias=bias)      # same initialisation as the layer it replaces
        self.weight = ref.weight
        self.bias = ref.bias
        self.out_dtype = out_dtype

    @torch.no_grad()
    def forward(self, x: Tensor) -> Tensor:
        return head_classifier(x, self.weight, self.bias, self.out_dtype)
