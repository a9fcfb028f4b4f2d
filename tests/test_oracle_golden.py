"""Pins oracle/segpath_oracle.c against vectors produced by the reference's own code
(tests/golden/make_golden.py).  CPU only."""
import hashlib

import numpy as np
import pytest

from conftest import criterion_names, focal_names, golden_names, load_criterion, load_focal, load_golden

DT = {"float32": 0, "bfloat16": 1, "float16": 2}


def _up_bits(up_f32, dtype):
    """fp32 array holding dtype-rounded values -> the bit pattern torch stores for that dtype."""
    if dtype == "float32":
        return up_f32
    if dtype == "float16":
        return up_f32.astype(np.float16).view(np.uint16)
    return (up_f32.view(np.uint32) >> 16).astype(np.uint16)      # exact: values are already bf16


@pytest.mark.parametrize("name", golden_names())
def test_upsample_bits(oracle, name):
    g = load_golden(name)
    up = oracle.upsample(g["logits"], tuple(g["size"]), DT[g["logits_dtype"]])
    got = _up_bits(up, g["logits_dtype"])
    assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == str(g["up_sha256"])
    if "up" in g:
        assert np.array_equal(got.view(np.uint8), g["up"].view(np.uint8))


@pytest.mark.parametrize("name", golden_names())
def test_argmax_and_hist_exact(oracle, name):
    g = load_golden(name)
    f = oracle.path_forward(g["logits"], g["labels"], 255, g["class_w"], oracle.OHEM_THRESH, DT[g["logits_dtype"]])
    assert np.array_equal(f["arg_px"].reshape(g["argmax"].shape), g["argmax"].astype(np.int32))
    assert np.array_equal(f["hist"], g["mat_i64"])
    assert np.array_equal(f["hist"].astype(np.float32), g["hist_f32"])
    assert int(f["sums"][2]) == int(g["n_valid"])
    # ConfusionMatrix.update twin (utils.py:99-109)
    mat = oracle.hist_pairs(g["labels"].reshape(-1).astype(np.int64), g["argmax"].reshape(-1).astype(np.int64),
                            g["logits"].shape[1])
    assert np.array_equal(mat, g["mat_i64"])


@pytest.mark.parametrize("name", golden_names())
def test_loss_px_and_scalar(oracle, name):
    g = load_golden(name)
    dt = DT[g["logits_dtype"]]
    f = oracle.path_forward(g["logits"], g["labels"], 255, g["class_w"], oracle.OHEM_THRESH, dt,
                            want_arg=False, want_hist=False)
    np.testing.assert_allclose(f["loss_px"].reshape(g["loss_px"].shape), g["loss_px"], rtol=2e-6, atol=2e-6)
    if g["loss_kind"] == "ce":
        loss, _ = oracle.ce_loss_and_grad(g["logits"], g["labels"], 255, g["class_w"], dt, want_grad=False)
    else:
        loss, _, info = oracle.ohem_loss_and_grad(g["logits"], g["labels"], 255, g["class_w"], ldtype=dt, want_grad=False)
        assert abs(info["n_hard"] - int(g["n_hard"])) <= 2          # pixels within an ulp of the threshold
        assert info["branch"] == int(int(g["n_hard"]) < int(g["n_valid"]) // 16)
    if np.isnan(g["loss"]):
        assert np.isnan(loss)                                       # all-ignored batch
    else:
        assert abs(loss - float(g["loss"])) <= 1e-6 * abs(float(g["loss"])) + 1e-7


@pytest.mark.parametrize("name", [n for n in golden_names() if "all_ignored" not in n])
def test_gradient(oracle, name):
    g = load_golden(name)
    dt = DT[g["logits_dtype"]]
    if g["loss_kind"] == "ce":
        _, grad = oracle.ce_loss_and_grad(g["logits"], g["labels"], 255, g["class_w"], dt)
    else:
        _, grad, _ = oracle.ohem_loss_and_grad(g["logits"], g["labels"], 255, g["class_w"], ldtype=dt)
    ref = g["grad"]
    scale = np.abs(ref).max()
    # fp32 reference grads carry fp32 accumulation noise; low-precision ones were rounded to bf16/fp16
    tol = 1e-5 if g["logits_dtype"] == "float32" else 1e-2
    assert np.abs(grad - ref).max() <= tol * scale, (np.abs(grad - ref).max(), scale)


@pytest.mark.parametrize("name", focal_names())
def test_focal_loss_and_gradient(oracle, name):
    """FocalLoss epilogue against the reference's own FocalLoss."""
    g = load_focal(name)
    dt = DT[g["logits_dtype"]]
    loss, grad = oracle.focal_loss_and_grad(g["logits"], g["labels"], float(g["alpha"]), float(g["gamma"]),
                                            bool(g["size_average"]), 255, dt)
    assert abs(loss - float(g["loss"])) <= 2e-6 * abs(float(g["loss"]))
    tol = 1e-5 if g["logits_dtype"] == "float32" else 1e-2
    assert np.abs(grad - g["grad"]).max() <= tol * np.abs(g["grad"]).max()


@pytest.mark.parametrize("name", criterion_names())
def test_criterion_loss_and_gradient(oracle, name):
    """engine.criterion = CE + multiclass dice loss against the reference's own function."""
    g = load_criterion(name)
    dt = DT[g["logits_dtype"]]
    ign = int(g["ignore_index"])
    loss, grad, info = oracle.criterion_loss_and_grad(g["logits"], g["labels"], g["class_w"], bool(g["dice"]), ign, dt,
                                                      want_grad=bool(np.isfinite(g["loss"])))
    if not np.isfinite(g["loss"]):                                   # every pixel ignored: mean over nothing, like the reference
        assert np.isnan(loss) and np.isnan(g["loss"])
        return
    assert abs(loss - float(g["loss"])) <= 1e-5 * abs(float(g["loss"])), (loss, float(g["loss"]))
    tol = 1e-5 if g["logits_dtype"] == "float32" else 1e-2
    assert np.abs(grad - g["grad"]).max() <= tol * np.abs(g["grad"]).max(), (np.abs(grad - g["grad"]).max(), np.abs(g["grad"]).max())


def test_criterion_oracle_gradient_is_the_derivative_of_its_loss(oracle):
    """Independent of the goldens: central differences of the oracle's own criterion value reproduce its analytic gradient
    (dice through softmax and the transposed bilinear resize), including a class that is absent from one image."""
    rng = np.random.default_rng(7)
    B, C, h, w, H, W = 2, 4, 5, 6, 11, 13
    low = rng.normal(0, 2, (B, C, h, w)).astype(np.float32)
    lab = rng.integers(0, C, (B, H, W)).astype(np.int64)
    lab[0][lab[0] == 3] = 1                      # class 3 absent in image 0
    lab[1, :3, :] = 255
    cw = (rng.random(C) + 0.5).astype(np.float32)
    _, grad, _ = oracle.criterion_loss_and_grad(low, lab, cw, True, 255)
    eps = 1e-2
    for _ in range(12):
        idx = tuple(int(rng.integers(0, n)) for n in low.shape)
        lp, lm = low.copy(), low.copy()
        lp[idx] += eps
        lm[idx] -= eps
        fp, _, _ = oracle.criterion_loss_and_grad(lp, lab, cw, True, 255, want_grad=False)
        fm, _, _ = oracle.criterion_loss_and_grad(lm, lab, cw, True, 255, want_grad=False)
        fd = (fp - fm) / (float(lp[idx]) - float(lm[idx]))
        assert abs(fd - grad[idx]) <= 2e-3 * max(abs(grad[idx]), 1e-3) + 2e-5, (idx, fd, grad[idx])
