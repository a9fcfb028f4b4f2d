"""Parity of the CUDA path (through the Python mirror -> C ABI) with the oracle and the golden vectors.
Run on the B200 box: python -m pytest tests -m gpu"""
import os

import numpy as np
import pytest
import torch

from conftest import focal_names, golden_names, load_focal, load_golden

pytestmark = pytest.mark.gpu

TDT = {"float32": torch.float32, "bfloat16": torch.bfloat16, "float16": torch.float16}
ODT = {"float32": 0, "bfloat16": 1, "float16": 2}


@pytest.fixture(scope="module")
def sf():
    import segmentation_factory_b200 as m
    from segmentation_factory_b200 import _lib
    _lib.lib()                        # fail loudly if the extension is missing
    assert torch.cuda.is_available()
    return m


def _golden_tensors(g, dev="cuda"):
    if g["logits_dtype"] == "float32":
        low = torch.from_numpy(g["logits"])
    else:
        low = torch.from_numpy(g["logits"].view(np.int16)).view(TDT[g["logits_dtype"]])
    lab = torch.from_numpy(g["labels"])
    cw = None if g["class_w"] is None else torch.from_numpy(g["class_w"]).to(dev)
    return low.to(dev), lab.to(dev), cw


def _crit(sf, g, cw):
    return sf.OhemCrossEntropy(255, cw) if g["loss_kind"] == "ohem" else sf.CrossEntropy(255, cw)


@pytest.mark.parametrize("no_tma", ["0", "1"])
@pytest.mark.parametrize("name", golden_names())
def test_golden_hist_bit_exact(sf, name, no_tma, monkeypatch):
    monkeypatch.setenv("SF_NO_TMA", no_tma)
    g = load_golden(name)
    low, lab, _ = _golden_tensors(g)
    C = low.shape[1]
    m = sf.Metrics(C, 255, "cuda")
    cm = sf.ConfusionMatrix(C)
    sf.fused_eval_update(low, lab, m, cm)
    assert np.array_equal(cm.mat.cpu().numpy(), g["mat_i64"])
    assert np.array_equal(m.hist.cpu().numpy(), g["hist_f32"])
    # derived numbers (util/metrics.py:30-49, util/utils.py:115-123) follow bit-exactly
    iou, miou = m.compute_iou()
    assert np.array_equal(np.array(iou), g["iou"], equal_nan=True)
    assert miou == float(g["miou"]) or (np.isnan(miou) and np.isnan(g["miou"]))
    f1, mf1 = m.compute_f1()
    assert np.array_equal(np.array(f1), g["f1"], equal_nan=True)
    acc, macc = m.compute_pixel_acc()
    assert np.array_equal(np.array(acc), g["acc"], equal_nan=True)
    ag, ac, iu = cm.compute()
    assert np.array_equal(iu.cpu().numpy(), g["cm_iu"], equal_nan=True)
    assert np.array_equal(ac.cpu().numpy(), g["cm_acc"], equal_nan=True)
    assert np.array_equal(ag.cpu().numpy(), g["cm_acc_global"], equal_nan=True)
    # the unmodified engine.py:90 call: flattened labels + already-argmaxed predictions
    cm2 = sf.ConfusionMatrix(C)
    cm2.update(lab.flatten(), torch.from_numpy(g["argmax"].astype(np.int64)).cuda().flatten())
    assert np.array_equal(cm2.mat.cpu().numpy(), g["mat_i64"])


@pytest.mark.parametrize("no_tma", ["0", "1"])
@pytest.mark.parametrize("name", golden_names())
def test_golden_loss_and_grad(sf, name, no_tma, monkeypatch):
    monkeypatch.setenv("SF_NO_TMA", no_tma)
    g = load_golden(name)
    low, lab, cw = _golden_tensors(g)
    C = low.shape[1]
    x = low.clone().requires_grad_(True)
    hist = torch.zeros(C, C, dtype=torch.int64, device="cuda")
    thr = float(-torch.log(torch.tensor(0.7, dtype=torch.float))) if g["loss_kind"] == "ohem" else None
    loss = sf.fused_cross_entropy(x, lab, cw, 255, thr, hist)
    assert loss.dtype == torch.float32 and loss.dim() == 0
    assert np.array_equal(hist.cpu().numpy(), g["mat_i64"])          # same pass also feeds the matrix
    ref = float(g["loss"])
    if np.isnan(ref):
        assert torch.isnan(loss).item()                               # all-ignored batch -> NaN (engine.py:45 exits)
        return
    tol = 1e-5                                                        # fp32 arithmetic on identical (rounded) logits
    assert abs(loss.item() - ref) <= tol * abs(ref), (loss.item(), ref)
    loss.backward()
    got = x.grad.float().cpu().numpy()
    scale = np.abs(g["grad"]).max()
    gtol = 1e-5 if g["logits_dtype"] == "float32" else 2e-2
    assert np.abs(got - g["grad"]).max() <= gtol * scale, (np.abs(got - g["grad"]).max(), scale)
    # module API returns the same value
    loss2 = _crit(sf, g, cw)(low, lab)
    assert loss2.item() == loss.item()


def _synth(seed, B, C, h, w, H, W, kind, dtype=torch.float32, ldt=torch.int64, blk=32):
    g = torch.Generator().manual_seed(seed)
    coarse = torch.randint(0, C, (B, (H + blk - 1) // blk, (W + blk - 1) // blk), generator=g)
    coarse[torch.rand(coarse.shape, generator=g) < 0.1] = 255
    lab = coarse.repeat_interleave(blk, 1).repeat_interleave(blk, 2)[:, :H, :W].contiguous().to(ldt)
    if kind == "early":
        low = 3.0 * torch.randn((B, C, h, w), generator=g)
    else:
        ys = ((torch.arange(h) + 0.5) * H / h).long().clamp(max=H - 1)
        xs = ((torch.arange(w) + 0.5) * W / w).long().clamp(max=W - 1)
        l2 = lab[:, ys][:, :, xs].long()
        l2 = torch.where(l2 == 255, torch.zeros_like(l2), l2)
        flip = torch.rand(l2.shape, generator=g) < 0.02
        l2 = torch.where(flip, torch.randint(0, C, l2.shape, generator=g), l2)
        low = torch.randn((B, C, h, w), generator=g) + 12.0 * torch.nn.functional.one_hot(l2, C).permute(0, 3, 1, 2).float()
    return low.to(dtype).contiguous(), lab


def _bits(t):
    return t.numpy() if t.dtype == torch.float32 else t.view(torch.int16).numpy().view(np.uint16)


CASES = [
    # name, seed, B, C, h, w, H, W, kind, dtype, label dtype, loss kind, weighted
    ("city_b1_early", 1234, 1, 19, 256, 512, 1024, 2048, "early", torch.float32, torch.int64, "ohem", False),
    ("city_b1_late", 1235, 1, 19, 256, 512, 1024, 2048, "late", torch.float32, torch.int64, "ohem", False),
    ("city_u8_ce", 1236, 1, 19, 128, 256, 512, 1024, "early", torch.float32, torch.uint8, "ce", False),
    ("ade_ce", 2345, 2, 150, 128, 128, 512, 512, "early", torch.float32, torch.int64, "ce", True),
    ("coco_bf16_ohem", 3456, 2, 171, 128, 128, 512, 512, "late", torch.bfloat16, torch.int64, "ohem", False),
    ("cfg1_ohem", 0, 2, 150, 128, 128, 512, 512, "early", torch.float32, torch.int64, "ohem", False),
    ("voc_fp16_x8", 77, 2, 21, 40, 56, 320, 448, "late", torch.float16, torch.uint8, "ohem", False),
    ("odd_w_notma", 78, 1, 9, 50, 67, 200, 268, "early", torch.float32, torch.int64, "ohem", True),
    ("x16_ce", 79, 1, 5, 24, 32, 384, 512, "early", torch.float32, torch.int64, "ce", False),
    ("nonint_up", 80, 2, 19, 100, 130, 333, 517, "late", torch.float32, torch.int64, "ohem", False),
    ("downscale", 81, 1, 7, 300, 260, 200, 180, "early", torch.float32, torch.int64, "ce", False),
    ("fullres_identity", 82, 1, 19, 256, 320, 256, 320, "early", torch.float32, torch.int64, "ohem", False),
    ("fullres_bf16_c150", 83, 2, 150, 128, 160, 128, 160, "early", torch.bfloat16, torch.uint8, "ce", True),
    ("fullres_fp16_late", 84, 2, 21, 192, 256, 192, 256, "late", torch.float16, torch.int64, "ohem", False),
]


@pytest.mark.parametrize("no_fast", ["0", "1"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_seeded_vs_oracle(sf, oracle, case, no_fast, monkeypatch):
    monkeypatch.setenv("SF_NO_FAST", no_fast)                      # 1: force the generic kernels on fast-path shapes
    name, seed, B, C, h, w, H, W, kind, dtype, ldt, lk, weighted = case
    low, lab = _synth(seed, B, C, h, w, H, W, kind, dtype, ldt)
    cw = (torch.rand(C, generator=torch.Generator().manual_seed(7)) + 0.5) if weighted else None
    od = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[dtype]
    lown, labn = _bits(low), lab.numpy()
    cwn = None if cw is None else cw.numpy()

    x = low.cuda().requires_grad_(True)
    y = lab.cuda()
    hist = torch.zeros(C, C, dtype=torch.int64, device="cuda")
    thr = oracle.OHEM_THRESH if lk == "ohem" else None
    loss = sf.fused_cross_entropy(x, y, None if cw is None else cw.cuda(), 255, thr, hist)
    (loss * 1024.0).backward()                                     # a GradScaler-style scaled backward
    got = x.grad.float().cpu().numpy() / 1024.0

    ref_hist = oracle.path_forward(lown, labn, 255, None, 0.0, od, want_loss=False, want_lse=False, want_arg=False)["hist"]
    assert np.array_equal(hist.cpu().numpy(), ref_hist)            # bit-exact confusion matrix
    if lk == "ohem":
        ref_loss, ref_grad, info = oracle.ohem_loss_and_grad(lown, labn, 255, cwn, ldtype=od)
        # make sure the synthetic "late" sets really exercise the top-k branch
        if name in ("city_b1_late", "coco_bf16_ohem", "voc_fp16_x8"):
            assert info["branch"] == 1, info
    else:
        ref_loss, ref_grad = oracle.ce_loss_and_grad(lown, labn, 255, cwn, od)
    assert abs(loss.item() - ref_loss) <= 1e-5 * abs(ref_loss), (loss.item(), ref_loss)
    scale = np.abs(ref_grad).max()
    gtol = 1e-5 if dtype == torch.float32 else 2e-2
    err = np.abs(got - ref_grad).max()
    assert err <= gtol * scale, (err, scale)
    # eval-only kernel gives the same matrix
    hist2 = torch.zeros(C, C, dtype=torch.int64, device="cuda")
    sf.eval_hist(x.detach(), y, 255, hist2)
    assert np.array_equal(hist2.cpu().numpy(), ref_hist)


def test_full_size_cityscapes_properties(sf):
    """BASELINE config 2 at full size (8x19x256x512 -> 1024x2048): bit-exact matrix vs torch CPU (image by image),
    loss/grad vs the torch-eager restatement on the same GPU, and size-independent invariants."""
    from oracle import torch_restatement as tr
    B, C, h, w, H, W = 8, 19, 256, 512, 1024, 2048
    low, lab = _synth(1234, B, C, h, w, H, W, "early")
    x = low.cuda().requires_grad_(True)
    y = lab.cuda()
    hist = torch.zeros(C, C, dtype=torch.int64, device="cuda")
    thr = float(-torch.log(torch.tensor(0.7, dtype=torch.float)))
    loss = sf.fused_cross_entropy(x, y, None, 255, thr, hist)
    loss.backward()
    n_valid = int((lab != 255).sum())
    assert int(hist.sum()) == n_valid
    rows = torch.bincount(lab[lab != 255].flatten(), minlength=C)
    assert torch.equal(hist.sum(1).cpu(), rows)                     # row sums = label counts
    # softmax - onehot sums to zero over classes and the bilinear weights sum to one
    assert abs(x.grad.double().sum().item()) <= 1e-4 * x.grad.double().abs().sum().item()
    ref = torch.zeros(C, C, dtype=torch.int64)
    torch.set_num_threads(os.cpu_count() or 1)
    for b in range(B):
        ref += tr.confusion_counts(tr.upsample(low[b:b + 1], (H, W)), lab[b:b + 1], C)
    assert torch.equal(hist.cpu(), ref)
    rl, rg = tr.loss_and_grad(low.cuda(), y, "ohem")
    assert abs(loss.item() - rl.item()) <= 1e-5 * abs(rl.item())
    assert (x.grad - rg).abs().max().item() <= 1e-5 * rg.abs().max().item()
    # bitwise reproducible (no float atomics anywhere on the path): the same call twice gives the same bits
    x2 = low.cuda().requires_grad_(True)
    loss2 = sf.fused_cross_entropy(x2, y, None, 255, thr, torch.zeros_like(hist))
    loss2.backward()
    assert loss2.item() == loss.item() and torch.equal(x2.grad, x.grad)
    # without the matrix a different kernel instantiation runs: same values up to float association
    x3 = low.cuda().requires_grad_(True)
    loss3 = sf.fused_cross_entropy(x3, y, None, 255, thr, None)
    loss3.backward()
    assert abs(loss3.item() - loss.item()) <= 1e-6 * abs(loss.item())
    assert (x3.grad - x.grad).abs().max().item() <= 1e-6 * x.grad.abs().max().item()


def test_tuple_preds_and_aux_weights(sf):
    low, lab = _synth(5, 2, 19, 32, 64, 128, 256, "early")
    x, y = low.cuda(), lab.cuda()
    ce, oh = sf.CrossEntropy(), sf.OhemCrossEntropy()
    assert abs(ce((x, x, x), y).item() - 1.8 * ce(x, y).item()) <= 1e-5 * ce(x, y).item() * 1.8
    assert abs(oh((x, x), y).item() - 2.0 * oh(x, y).item()) <= 1e-6
    assert sf.get_loss("OhemCrossEntropy", 255, None)(x, y).item() == oh(x, y).item()


def test_small_valid_count_edges(sf, oracle):
    """n_valid < 16 -> n_min = 0: mean over pixels above the threshold only; none above -> NaN."""
    low, lab = _synth(6, 1, 19, 32, 64, 128, 256, "early")
    lab[:] = 255
    lab[0, 5, 7:17] = 3
    x = low.cuda().requires_grad_(True)
    loss = sf.OhemCrossEntropy()(x, lab.cuda())
    ref, grad, info = oracle.ohem_loss_and_grad(low.numpy(), lab.numpy())
    assert abs(loss.item() - ref) <= 1e-5 * abs(ref)
    loss.backward()
    assert np.abs(x.grad.cpu().numpy() - grad).max() <= 1e-5 * np.abs(grad).max()
    low2 = torch.zeros_like(low)
    low2[:, 3] = 30.0                                              # every valid pixel is confidently right
    loss = sf.OhemCrossEntropy()(low2.cuda(), lab.cuda())
    assert torch.isnan(loss).item()


def test_nan_and_inf_logits_follow_torch_argmax(sf):
    from oracle import torch_restatement as tr
    low, lab = _synth(8, 1, 19, 32, 64, 128, 256, "early")
    low[0, 4, 10, 20] = float("nan")
    low[0, 9, 10, 21] = float("nan")
    low[0, 2, 3, 3] = float("inf")
    hist = torch.zeros(19, 19, dtype=torch.int64, device="cuda")
    sf.eval_hist(low.cuda(), lab.cuda(), 255, hist)
    ref = tr.confusion_counts(tr.upsample(low, (128, 256)), lab, 19)
    assert torch.equal(hist.cpu(), ref)
    assert torch.isnan(sf.CrossEntropy()(low.cuda(), lab.cuda())).item()


def test_same_size_nan_argmax_and_error_codes(sf):
    """Full-resolution logits (the unmodified engine.py call): first-NaN argmax like torch; C*C histogram too large for
    shared memory is reported, not silently mis-counted."""
    from oracle import torch_restatement as tr
    from segmentation_factory_b200._lib import SegFactoryError
    low, lab = _synth(11, 1, 19, 128, 256, 128, 256, "early")
    low[0, 7, 5, 9] = float("nan")
    low[0, 3, 5, 9] = float("nan")
    low[0, 2, 60, 60] = float("inf")
    hist = torch.zeros(19, 19, dtype=torch.int64, device="cuda")
    sf.eval_hist(low.cuda(), lab.cuda(), 255, hist)
    assert torch.equal(hist.cpu(), tr.confusion_counts(low, lab, 19))
    big = torch.randn(1, 300, 32, 64, device="cuda")
    lab2 = torch.zeros(1, 128, 256, dtype=torch.long, device="cuda")
    with pytest.raises(SegFactoryError):
        sf.eval_hist(big, lab2, 255, torch.zeros(300, 300, dtype=torch.int64, device="cuda"))
    loss = sf.CrossEntropy()(big, lab2)                              # loss only (no matrix): any C works
    ref = torch.nn.functional.cross_entropy(torch.nn.functional.interpolate(big, size=(128, 256), mode="bilinear", align_corners=False), lab2)
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())


@pytest.mark.parametrize("name", focal_names())
def test_focal_loss_golden(sf, name):
    g = load_focal(name)
    if g["logits_dtype"] == "float32":
        low = torch.from_numpy(g["logits"])
    else:
        low = torch.from_numpy(g["logits"].view(np.int16)).view(TDT[g["logits_dtype"]])
    x = low.cuda().requires_grad_(True)
    crit = sf.FocalLoss(float(g["alpha"]), float(g["gamma"]), bool(g["size_average"]), 255)
    loss = crit(x, torch.from_numpy(g["labels"]).cuda())
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    loss.backward()
    tol = 1e-5 if g["logits_dtype"] == "float32" else 2e-2
    assert np.abs(x.grad.float().cpu().numpy() - g["grad"]).max() <= tol * np.abs(g["grad"]).max()


def test_focal_loss_fast_paths_vs_oracle(sf, oracle):
    """FocalLoss on the Cityscapes-shaped fast kernels and on same-size logits."""
    for (B, C, h, w, H, W) in [(1, 19, 128, 256, 512, 1024), (1, 19, 128, 256, 128, 256)]:
        low, lab = _synth(91, B, C, h, w, H, W, "early")
        x = low.cuda().requires_grad_(True)
        loss = sf.FocalLoss(0.25, 2.0)(x, lab.cuda())
        loss.backward()
        ref, grad = oracle.focal_loss_and_grad(low.numpy(), lab.numpy(), 0.25, 2.0, True)
        assert abs(loss.item() - ref) <= 1e-5 * abs(ref)
        assert np.abs(x.grad.cpu().numpy() - grad).max() <= 1e-5 * np.abs(grad).max()


def test_c_abi_direct_call(sf):
    """Call the C ABI without the Python mirror: raw pointers, explicit stream, error codes."""
    import ctypes
    from segmentation_factory_b200 import _lib
    L = _lib.lib()
    low, lab = _synth(9, 1, 19, 32, 64, 128, 256, "early")
    x, y = low.cuda(), lab.cuda()
    hist = torch.zeros(19, 19, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rc = L.sf_eval_hist(x.data_ptr(), 0, y.data_ptr(), 0, 1, 19, 32, 64, 128, 256, 255, hist.data_ptr(), st)
    assert rc == 0
    assert int(hist.sum()) == int((lab != 255).sum())
    assert L.sf_eval_hist(None, 0, y.data_ptr(), 0, 1, 19, 32, 64, 128, 256, 255, hist.data_ptr(), st) == -1
    assert L.sf_eval_hist(x.data_ptr(), 7, y.data_ptr(), 0, 1, 19, 32, 64, 128, 256, 255, hist.data_ptr(), st) == -1
    ws = torch.empty(16, dtype=torch.uint8, device="cuda")
    rc = L.sf_ce_forward(x.data_ptr(), 0, y.data_ptr(), 0, 1, 19, 32, 64, 128, 256, 255, None, ctypes.c_float(-1.0),
                         None, None, None, ws.data_ptr(), 16, st)
    assert rc == -4 and b"workspace" in L.sf_strerror(rc)


@pytest.mark.parametrize("C,dtype", [(19, torch.float32), (21, torch.bfloat16), (2, torch.float32), (9, torch.float16)])
def test_eval_kernel_special_values_follow_torch_argmax(sf, C, dtype):
    """The eval-only kernel takes its maximum NaN-propagating and sends NaN / +-Inf / +-0 maxima to an exact path: first NaN
    wins, -0 == +0 (lowest index), ties go to the lowest index, -inf never wins, an all -inf pixel gives index 0."""
    from oracle import torch_restatement as tr
    low, lab = _synth(71, 1, C, 32, 64, 128, 256, "early")
    low = ((low * 2).round() / 2).to(dtype).float()                           # ties; values exactly representable in `dtype`
    low[0, C - 1, 10, 20] = float("nan")
    low[0, 0, 10, 21] = float("nan")
    low[0, 1, 10, 21] = float("nan")
    low[0, 1, 3, 3] = float("inf")
    low[0, 0, 20, 40] = float("-inf")
    low[0, :, 22:24, 50:52] = float("-inf")                                   # every class -inf: NaN after interpolation or -inf
    low[0, :, 25:30, 10:30] = 0.0                                             # every class ties at +0
    low[0, C // 2, 25:30, 10:30] = -0.0
    low[0, :, 16:20, 40:60] = -1.0
    low[0, C - 1, 16:20, 40:60] = -0.0                                        # maximum is -0 ...
    low[0, 0, 16:20, 44:48] = 0.0                                             # ... and +0 at a lower index in part of the area
    x = low.to(dtype)
    hist = torch.zeros(C, C, dtype=torch.int64, device="cuda")
    sf.eval_hist(x.cuda(), lab.cuda(), 255, hist)
    ref = tr.confusion_counts(tr.upsample(x, (128, 256)), lab, C)
    assert torch.equal(hist.cpu(), ref)
