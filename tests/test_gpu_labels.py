"""Label ingest on the GPU: sf_label_lut against the reference datasets' own encoders (golden) and
numpy at the full Cityscapes batch size.  Run on the B200 box: python -m pytest tests -m gpu"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sf():
    import segmentation_factory_b200 as m
    from segmentation_factory_b200 import _lib
    _lib.lib()
    assert torch.cuda.is_available()
    return m


@pytest.mark.parametrize("key", ["cs", "coco"])
@pytest.mark.parametrize("dtype", [torch.uint8, torch.int64])
def test_label_lut_golden(sf, key, dtype):
    z = np.load(os.path.join(GOLDEN_DIR, "label_lut.npz"))
    lut = sf.LabelLUT(z[key + "_table"])
    out = lut(torch.from_numpy(z[key + "_raw"]).cuda(), dtype)
    assert out.dtype == dtype and out.shape == z[key + "_raw"].shape
    assert np.array_equal(out.cpu().numpy().astype(np.int64), z[key + "_enc"].astype(np.int64))   # bit-exact


def test_label_lut_full_size_and_ragged(sf):
    z = np.load(os.path.join(GOLDEN_DIR, "label_lut.npz"))
    lut = sf.LabelLUT(z["cs_table"])
    g = torch.Generator().manual_seed(3)
    raw = torch.randint(0, 256, (8, 1024, 2048), generator=g, dtype=torch.uint8)        # incl. ids past the table -> 255
    ref = lut.table[raw.numpy()]
    dev = raw.cuda()
    assert np.array_equal(lut(dev).cpu().numpy(), ref)
    assert np.array_equal(lut(dev, torch.int64).cpu().numpy(), ref.astype(np.int64))
    # idempotence of the Cityscapes table on its own output range is NOT expected; a second property instead:
    # mapping commutes with slicing (ragged lengths and unaligned starts take the scalar kernel / tail path)
    flat = dev.flatten()
    for off, n in [(1, 1000), (3, 17), (16, 15), (0, 1), (5, 4099), (0, 0)]:
        part = flat[off:off + n]
        assert np.array_equal(lut(part).cpu().numpy(), ref.reshape(-1)[off:off + n])
        assert np.array_equal(lut(part, torch.int64).cpu().numpy(), ref.reshape(-1)[off:off + n].astype(np.int64))


def test_label_lut_feeds_the_loss_kernels(sf):
    """uint8 end to end: raw ids -> LUT (uint8) -> OhemCrossEntropy / Metrics give the same numbers as the int64 route."""
    z = np.load(os.path.join(GOLDEN_DIR, "label_lut.npz"))
    lut = sf.LabelLUT(z["cs_table"])
    g = torch.Generator().manual_seed(4)
    coarse = torch.randint(0, 34, (2, 16, 32), generator=g, dtype=torch.uint8)
    raw = coarse.repeat_interleave(16, 1).repeat_interleave(16, 2).contiguous().cuda()
    low = (3.0 * torch.randn((2, 19, 64, 128), generator=g)).cuda()
    y8, y64 = lut(raw), lut(raw, torch.int64)
    crit = sf.OhemCrossEntropy(255)
    assert crit(low, y8).item() == crit(low, y64).item()
    m8, m64 = sf.Metrics(19, 255, "cuda"), sf.Metrics(19, 255, "cuda")
    m8.update(low, y8)
    m64.update(low, y64)
    assert torch.equal(m8.hist, m64.hist) and float(m8.hist.sum()) == float((y64 != 255).sum())
