"""Label ingest on the device: raw uint8 annotation ids -> training ids through a 256-entry table.

Mirrors what the reference's datasets do per sample on the CPU with numpy --
    Cityscapes.encode_target(target) = id_to_train_id[np.array(target)]        (datasets/cityscapes.py:92,130-131)
    COCOStuff.encode(label)          = label_map[label]                        (datasets/coco_stuff.py:66-69,98-100)
-- so that the annotation bytes can travel to the GPU as uint8 (1 B/px instead of the int64 collate's 8 B/px) and stay
uint8 for the loss / metric kernels (which read uint8 labels directly).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


class LabelLUT:
    """table[k] is the training id of raw id k (the reference's `id_to_train_id` / `label_map` array, any integer
    dtype, up to 256 entries with values in [0, 255]).  Raw ids past the table map to `fill` (numpy would raise an
    IndexError there; a kernel cannot without a host sync)."""

    def __init__(self, table, device="cuda", fill: int = 255):
        t = np.asarray(table).astype(np.int64).reshape(-1)
        if t.size == 0 or t.size > 256:
            raise ValueError("label table must have 1..256 entries")
        if t.min() < 0 or t.max() > 255:
            raise ValueError("label table values must fit uint8 (the kernels' label type)")
        full = np.full(256, int(fill), np.uint8)
        full[:t.size] = t.astype(np.uint8)
        self.table = full
        self.device = torch.device(device)
        self._lut = None

    def _device_lut(self) -> torch.Tensor:
        if self._lut is None:
            self._lut = torch.from_numpy(self.table).to(self.device)
        return self._lut

    def __call__(self, raw: torch.Tensor, dtype: torch.dtype = torch.uint8) -> torch.Tensor:
        if not isinstance(raw, torch.Tensor) or raw.dtype != torch.uint8:
            raise TypeError("raw ids must be a uint8 tensor (the annotation PNG's bytes)")
        if not raw.is_cuda:
            raise _lib.SegFactoryError("segmentation_factory_b200 has no CPU path: copy the raw uint8 ids to the GPU first")
        if dtype not in (torch.uint8, torch.int64):
            raise TypeError("output dtype must be uint8 or int64")
        raw = raw.contiguous()
        out = torch.empty(raw.shape, dtype=dtype, device=raw.device)
        if raw.numel() == 0:
            return out
        lut = self._device_lut()
        if lut.device != raw.device:
            lut = lut.to(raw.device)
        with torch.cuda.device(raw.device):
            rc = _lib.lib().sf_label_lut(raw.data_ptr(), raw.numel(), lut.data_ptr(), out.data_ptr(),
                                         _lib.SF_U8 if dtype == torch.uint8 else _lib.SF_I64,
                                         torch.cuda.current_stream(raw.device).cuda_stream)
        _lib.check(rc, "sf_label_lut")
        return out
