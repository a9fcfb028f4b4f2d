"""CPU-side checks: the C-ABI library loads and exports what include/segfactory.h declares, the host
mirror validates its inputs and refuses to run without CUDA, Metrics' float32 accumulation and the
compute_* formulas match the reference's golden numbers, and the torch restatement used on the GPU box
is itself pinned to the golden vectors."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, ROOT, golden_names, load_golden


def test_library_exports_every_declared_symbol():
    from segmentation_factory_b200 import _lib
    header = open(os.path.join(ROOT, "include", "segfactory.h")).read()
    declared = set(re.findall(r"^(?:int|size_t|const char\*)\s+(sf_[a-z_]+)\s*\(", header, flags=re.M))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.sf_version() == 200
    assert L.sf_dice_workspace_bytes(8, 19) >= 64 + 8 * 19 * 32 and L.sf_dice_workspace_bytes(0, 19) == 0
    assert L.sf_workspace_bytes() > 64
    assert L.sf_strerror(0) == b"ok" and b"argument" in L.sf_strerror(-1)


def test_no_cpu_fallback_and_input_validation():
    import segmentation_factory_b200 as sf
    from segmentation_factory_b200._lib import SegFactoryError
    x, y = torch.randn(1, 19, 8, 8), torch.zeros(1, 32, 32, dtype=torch.long)
    with pytest.raises(SegFactoryError):
        sf.OhemCrossEntropy()(x, y)
    with pytest.raises(SegFactoryError):
        sf.CrossEntropy()(x, y)
    with pytest.raises(SegFactoryError):
        sf.criterion(x, y, None, num_classes=19, dice=True, ignore_index=255)
    with pytest.raises(ValueError):
        sf.criterion(x, y, None, num_classes=2, dice=True, ignore_index=255)
    with pytest.raises(SegFactoryError):
        sf.Metrics(19, 255, "cpu").update(x, y)
    with pytest.raises(SegFactoryError):
        sf.ConfusionMatrix(19).update(y.flatten(), y.flatten())
    with pytest.raises(AssertionError):
        sf.get_loss("Dice")
    with pytest.raises(SegFactoryError):
        sf.FocalLoss(0.25, 2.0)(x, y)
    assert isinstance(sf.get_loss("CrossEntropy", 255, None), sf.CrossEntropy)
    o = sf.OhemCrossEntropy(thresh=0.7)
    assert o.thresh.dtype == torch.float32 and abs(float(o.thresh) - 0.35667494) < 1e-7   # util/losses.py:49
    assert o.aux_weights == [1, 1] and sf.CrossEntropy().aux_weights == [1, 0.4, 0.4]


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "segmentation_factory_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("oracle/segpath_oracle.c", ""), f


def test_metrics_float32_accumulation_and_formulas():
    """util/metrics.py:13,27,30-49 + util/utils.py:115-143 on the golden count sequence (cells pass 2**24)."""
    import segmentation_factory_b200 as sf
    z = np.load(os.path.join(GOLDEN_DIR, "hist_f32_accum.npz"))
    m = sf.Metrics(19, 255, "cpu")
    for c in z["counts"]:
        m._accumulate(torch.from_numpy(c))
    assert np.array_equal(m.hist.numpy(), z["hist_f32"])
    assert (m.hist_i64.numpy() == z["counts"].sum(0)).all()
    assert not np.array_equal(m.hist.numpy().astype(np.int64), m.hist_i64.numpy())   # float32 really is inexact here
    iou, miou = m.compute_iou()
    f1, mf1 = m.compute_f1()
    acc, macc = m.compute_pixel_acc()
    assert np.array_equal(np.array(iou), z["iou"], equal_nan=True) and miou == float(z["miou"])
    assert np.array_equal(np.array(f1), z["f1"], equal_nan=True) and mf1 == float(z["mf1"])
    assert np.array_equal(np.array(acc), z["acc"], equal_nan=True) and macc == float(z["macc"])
    cm = sf.ConfusionMatrix(19)
    cm.mat = torch.from_numpy(z["counts"].sum(0))
    ag, ac, iu = cm.compute()
    assert np.array_equal(iu.numpy(), z["cm_iu"], equal_nan=True)
    assert np.array_equal(ac.numpy(), z["cm_acc"], equal_nan=True)
    assert float(ag) == float(z["cm_acc_global"])
    assert str(cm) == str(z["cm_str"])
    cm.reset()
    assert int(cm.mat.sum()) == 0


@pytest.mark.parametrize("name", [n for n in golden_names() if "all_ignored" not in n])
def test_torch_restatement_matches_golden(name):
    from oracle import torch_restatement as tr
    g = load_golden(name)
    if g["logits_dtype"] == "float32":
        low = torch.from_numpy(g["logits"])
    else:
        low = torch.from_numpy(g["logits"].view(np.int16)).view({"bfloat16": torch.bfloat16, "float16": torch.float16}[g["logits_dtype"]])
    lab = torch.from_numpy(g["labels"])
    cw = None if g["class_w"] is None else torch.from_numpy(g["class_w"])
    C = low.shape[1]
    up = tr.upsample(low, tuple(g["size"])).float()
    assert np.array_equal(tr.confusion_counts(up, lab.long(), C).numpy(), g["mat_i64"])
    loss, grad = tr.loss_and_grad(low, lab, g["loss_kind"], cw)
    assert loss.item() == float(g["loss"])
    assert np.array_equal(grad.float().numpy(), g["grad"])


def test_label_lut_table_matches_reference_encoders():
    """Host side of the label ingest: the 256-entry table reproduces Cityscapes.encode_target and
    COCOStuff.encode on the golden raw ids; the mapping itself has no CPU path."""
    import segmentation_factory_b200 as sf
    from segmentation_factory_b200._lib import SegFactoryError
    z = np.load(os.path.join(GOLDEN_DIR, "label_lut.npz"))
    for key in ("cs", "coco"):
        lut = sf.LabelLUT(z[key + "_table"], device="cpu")
        assert lut.table.dtype == np.uint8 and lut.table.shape == (256,)
        assert np.array_equal(lut.table[z[key + "_raw"]].astype(np.int64), z[key + "_enc"].astype(np.int64))
    assert (sf.LabelLUT(z["cs_table"], device="cpu").table[35:] == 255).all()          # ids past the table -> ignore
    with pytest.raises(SegFactoryError):
        sf.LabelLUT(z["cs_table"], device="cpu")(torch.zeros(4, dtype=torch.uint8))
    with pytest.raises(TypeError):
        sf.LabelLUT(z["cs_table"], device="cpu")(torch.zeros(4, dtype=torch.int64))
    with pytest.raises(ValueError):
        sf.LabelLUT(np.arange(300))


def test_metrics_pq_surface():
    """Attribute parity with util/metrics.py: pq_hist exists, compute_pq is the reference's formula, update_pq says why it is absent."""
    import segmentation_factory_b200 as sf
    m = sf.Metrics(3, 255, "cpu")
    assert m.compute_pq() == (0.0, 0.0, 0.0)
    m.pq_hist.update(TP=3, FP=1, FN=2, iou_sum=2.4)
    assert m.compute_pq() == (round(0.8 * (3 / 4.5) * 100, 2), 80.0, round(3 / 4.5 * 100, 2))
    with pytest.raises(NotImplementedError):
        m.update_pq([], [])
