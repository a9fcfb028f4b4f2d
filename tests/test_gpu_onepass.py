"""The one-pass forward+backward walk (sf_ce_onepass / sf_ce_onepass_backward, csrc/sf_onepass.cu) against the oracle,
against the forward / finalize / backward kernel route on the same inputs, and through the raw C ABI.
Run on the B200 box: python -m pytest tests -m gpu"""
import ctypes

import numpy as np
import pytest
import torch

from test_gpu_parity import _bits, _synth

pytestmark = pytest.mark.gpu

THR = float(-torch.log(torch.tensor(0.7, dtype=torch.float)))


@pytest.fixture(scope="module")
def sf():
    import segmentation_factory_b200 as m
    from segmentation_factory_b200 import _lib
    _lib.lib()
    assert torch.cuda.is_available()
    return m


def _run(sf, low, lab, cw, thr, want_hist=True, gscale=1.0):
    C = low.shape[1]
    x = low.cuda().requires_grad_(True)
    hist = torch.zeros(C, C, dtype=torch.int64, device="cuda") if want_hist else None
    loss = sf.fused_cross_entropy(x, lab.cuda(), None if cw is None else cw.cuda(), 255, thr, hist)
    (loss * gscale).backward()
    return loss.detach(), x.grad.detach(), hist


# name, seed, B, C, h, w, H, W, kind, dtype, label dtype, loss kind, weighted
ONEPASS_CASES = [
    ("c19_x4_early_ohem", 11, 2, 19, 64, 128, 256, 512, "early", torch.float32, torch.int64, "ohem", False),
    ("c19_x4_late_ohem", 12, 2, 19, 64, 128, 256, 512, "late", torch.float32, torch.int64, "ohem", False),
    ("c19_x4_ce_weighted_u8", 13, 1, 19, 72, 100, 288, 400, "early", torch.float32, torch.uint8, "ce", True),
    ("c19_x4_ohem_weighted", 14, 1, 19, 40, 72, 160, 288, "early", torch.float32, torch.int64, "ohem", True),
    ("c21_x8_bf16_late", 15, 2, 21, 24, 48, 192, 384, "late", torch.bfloat16, torch.int64, "ohem", False),
    ("c21_x4_fp16_early", 16, 1, 21, 33, 72, 132, 288, "early", torch.float16, torch.uint8, "ohem", False),
    ("c19_x8_ce", 17, 1, 19, 20, 36, 160, 288, "early", torch.float32, torch.int64, "ce", False),
    ("c19_x4_bf16_ce", 18, 3, 19, 10, 40, 40, 160, "early", torch.bfloat16, torch.int64, "ce", False),
    ("c19_x4_ragged_tiles", 19, 1, 19, 27, 92, 108, 368, "late", torch.float32, torch.int64, "ohem", False),
    ("c2_x4_ce_weighted_u8", 31, 2, 2, 40, 64, 160, 256, "early", torch.float32, torch.uint8, "ce", True),
    ("c2_x8_bf16_ohem", 32, 1, 2, 20, 48, 160, 384, "early", torch.bfloat16, torch.int64, "ohem", False),
    ("c9_x4_late_ohem", 33, 2, 9, 36, 72, 144, 288, "late", torch.float32, torch.int64, "ohem", False),
    ("c9_x8_fp16_ce", 34, 1, 9, 18, 40, 144, 320, "early", torch.float16, torch.int64, "ce", False),
    ("c19_x4_tall_tiles", 35, 6, 19, 96, 160, 384, 640, "early", torch.float32, torch.int64, "ohem", False),
]


def _supported(low, lab):
    from segmentation_factory_b200 import _lib
    code = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[low.dtype]
    x = low.cuda()
    B, C, h, w = low.shape
    return _lib.lib().sf_ce_onepass_supported(code, x.data_ptr(), C, h, w, lab.shape[1], lab.shape[2]) == 1


@pytest.mark.parametrize("case", ONEPASS_CASES, ids=[c[0] for c in ONEPASS_CASES])
def test_onepass_vs_oracle_and_split_route(sf, oracle, case, monkeypatch):
    name, seed, B, C, h, w, H, W, kind, dtype, ldt, lk, weighted = case
    low, lab = _synth(seed, B, C, h, w, H, W, kind, dtype, ldt, blk=16)
    assert _supported(low, lab), "case must exercise the one-pass kernels"
    cw = (torch.rand(C, generator=torch.Generator().manual_seed(7)) + 0.5) if weighted else None
    thr = THR if lk == "ohem" else None
    od = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[dtype]
    lown, labn = _bits(low), lab.numpy()
    cwn = None if cw is None else cw.numpy()

    monkeypatch.setenv("SF_NO_ONEPASS", "0")
    loss, grad, hist = _run(sf, low, lab, cw, thr, gscale=512.0)
    monkeypatch.setenv("SF_NO_ONEPASS", "1")
    loss_s, grad_s, hist_s = _run(sf, low, lab, cw, thr, gscale=512.0)

    ref_hist = oracle.path_forward(lown, labn, 255, None, 0.0, od, want_loss=False, want_lse=False, want_arg=False)["hist"]
    assert np.array_equal(hist.cpu().numpy(), ref_hist)                       # bit-exact confusion matrix
    assert torch.equal(hist, hist_s)
    if lk == "ohem":
        ref_loss, ref_grad, info = oracle.ohem_loss_and_grad(lown, labn, 255, cwn, ldtype=od)
        if name in ("c19_x4_late_ohem", "c19_x4_ragged_tiles", "c9_x4_late_ohem"):
            assert info["branch"] == 1, info                                   # oracle numbering: 1 = top-k fallback
    else:
        ref_loss, ref_grad = oracle.ce_loss_and_grad(lown, labn, 255, cwn, od)
    assert abs(loss.item() - ref_loss) <= 1e-5 * abs(ref_loss), (loss.item(), ref_loss)
    assert abs(loss.item() - loss_s.item()) <= 2e-6 * abs(loss_s.item())
    gtol = 1e-5 if dtype == torch.float32 else 2e-2
    scale = np.abs(ref_grad).max()
    got = grad.float().cpu().numpy() / 512.0
    assert np.abs(got - ref_grad).max() <= gtol * scale, (np.abs(got - ref_grad).max(), scale)
    assert (grad.float() - grad_s.float()).abs().max().item() <= gtol * 512.0 * scale


def test_onepass_without_matrix_and_repeated_backward(sf, oracle):
    low, lab = _synth(21, 2, 19, 48, 96, 192, 384, "early", blk=16)
    ref_loss, ref_grad, _ = oracle.ohem_loss_and_grad(low.numpy(), lab.numpy(), 255, None)
    x = low.cuda().requires_grad_(True)
    loss = sf.OhemCrossEntropy()(x, lab.cuda())                               # module API, no matrix requested
    loss.backward(retain_graph=True)
    g1 = x.grad.clone()
    x.grad = None
    loss.backward()                                                           # the saved sums must survive the first backward
    assert torch.equal(g1, x.grad)
    assert abs(loss.item() - ref_loss) <= 1e-5 * abs(ref_loss)
    assert np.abs(g1.cpu().numpy() - ref_grad).max() <= 1e-5 * np.abs(ref_grad).max()
    # second-order differentiation is refused loudly instead of returning a gradient without a graph
    x2 = low.cuda().requires_grad_(True)
    l2 = sf.OhemCrossEntropy()(x2, lab.cuda())
    (g,) = torch.autograd.grad(l2, x2, create_graph=True)
    with pytest.raises(RuntimeError):
        g.sum().backward()


def test_onepass_bitwise_reproducible(sf):
    low, lab = _synth(22, 2, 19, 64, 128, 256, 512, "early", blk=16)
    l1, g1, h1 = _run(sf, low, lab, None, THR)
    l2, g2, h2 = _run(sf, low, lab, None, THR)
    assert l1.item() == l2.item() and torch.equal(g1, g2) and torch.equal(h1, h2)


def test_onepass_argmax_special_values(sf):
    """NaN (first NaN wins), +-inf, exact ties (lowest index) and +-0 maxima follow torch.argmax bit-exactly."""
    from oracle import torch_restatement as tr
    low, lab = _synth(23, 1, 19, 32, 64, 128, 256, "early", blk=16)
    low = (low * 2).round() / 2                                               # ties
    low[0, 4, 10, 20] = float("nan")
    low[0, 9, 10, 21] = float("nan")
    low[0, 2, 3, 3] = float("inf")
    low[0, 5, 20, 40] = float("-inf")
    low[0, :, 25:30, 10:30] = 0.0                                             # all-zero logits: every class ties at +0
    low[0, 3, 25:30, 10:30] = -0.0
    low[0, :, 16:20, 40:60] = -1.0
    low[0, 7, 16:20, 40:60] = -0.0                                            # maximum is -0
    low[0, 11, 16:20, 40:60] = 0.0                                            # ... and +0 at a higher index: -0 == +0, 7 wins
    x = low.cuda().requires_grad_(True)
    hist = torch.zeros(19, 19, dtype=torch.int64, device="cuda")
    loss = sf.fused_cross_entropy(x, lab.cuda(), None, 255, None, hist)
    ref = tr.confusion_counts(tr.upsample(low, (128, 256)), lab, 19)
    assert torch.equal(hist.cpu(), ref)
    assert torch.isnan(loss).item()
    hist2 = torch.zeros(19, 19, dtype=torch.int64, device="cuda")
    loss2 = sf.fused_cross_entropy(x, lab.cuda(), None, 255, THR, hist2)
    assert torch.equal(hist2.cpu(), ref)


def test_onepass_c_abi_direct(sf):
    """Raw C-ABI sequence sf_ce_onepass -> sf_loss_finalize -> sf_ce_onepass_backward, and its error codes."""
    from segmentation_factory_b200 import _lib
    L = _lib.lib()
    B, C, h, w, H, W = 2, 19, 32, 64, 128, 256
    low, lab = _synth(24, B, C, h, w, H, W, "early", blk=16)
    x, y = low.cuda(), lab.cuda()
    st = torch.cuda.current_stream().cuda_stream
    wsb = L.sf_workspace_bytes()
    ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
    n_px = B * H * W
    loss_px = torch.empty(n_px, device="cuda")
    acc = torch.empty(B * C * h * w, device="cuda")
    grad = torch.empty_like(x)
    out = torch.empty((), device="cuda")
    hist = torch.zeros(C * C, dtype=torch.int64, device="cuda")
    assert L.sf_ce_onepass_supported(0, x.data_ptr(), C, h, w, H, W) == 1
    geo = (B, C, h, w, H, W, 255)
    rc = L.sf_ce_onepass(x.data_ptr(), 0, y.data_ptr(), 0, *geo, None, ctypes.c_float(THR), loss_px.data_ptr(),
                         hist.data_ptr(), acc.data_ptr(), ws.data_ptr(), wsb, st)
    assert rc == 0
    assert L.sf_loss_finalize(loss_px.data_ptr(), n_px, ctypes.c_float(THR), ws.data_ptr(), wsb, out.data_ptr(), st) == 0
    rc = L.sf_ce_onepass_backward(x.data_ptr(), 0, y.data_ptr(), 0, *geo, None, loss_px.data_ptr(), acc.data_ptr(), None,
                                  ws.data_ptr(), grad.data_ptr(), st)
    assert rc == 0
    # the per-pixel loss vector equals the one sf_ce_forward writes (same arithmetic up to the exp-sum association)
    loss_px2 = torch.empty(n_px, device="cuda")
    lse = torch.empty(n_px, device="cuda")
    ws2 = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
    assert L.sf_ce_forward(x.data_ptr(), 0, y.data_ptr(), 0, *geo, None, ctypes.c_float(THR), loss_px2.data_ptr(),
                           lse.data_ptr(), None, ws2.data_ptr(), wsb, st) == 0
    assert (loss_px - loss_px2).abs().max().item() <= 2e-6 * loss_px2.abs().max().item()
    assert int(hist.sum()) == int((lab != 255).sum())
    x2 = low.cuda().requires_grad_(True)
    ref = sf.fused_cross_entropy(x2, y, None, 255, THR, None)
    ref.backward()
    assert out.item() == ref.item() and torch.equal(grad, x2.grad)
    # error codes: shapes outside the kernels, missing buffers
    assert L.sf_ce_onepass_supported(0, x.data_ptr(), 150, h, w, H, W) == 0
    rc = L.sf_ce_onepass(x.data_ptr(), 0, y.data_ptr(), 0, B, C, h, w, 3 * h, 3 * w, 255, None, ctypes.c_float(THR),
                         loss_px.data_ptr(), None, acc.data_ptr(), ws.data_ptr(), wsb, st)
    assert rc == -7 and b"one-pass" in L.sf_strerror(rc)
    rc = L.sf_ce_onepass(x.data_ptr(), 0, y.data_ptr(), 0, *geo, None, ctypes.c_float(THR), None, None, acc.data_ptr(),
                         ws.data_ptr(), wsb, st)
    assert rc == -1
    rc = L.sf_ce_onepass(x.data_ptr(), 0, y.data_ptr(), 0, *geo, None, ctypes.c_float(THR), loss_px.data_ptr(), None,
                         acc.data_ptr(), ws.data_ptr(), 16, st)
    assert rc == -4


def test_onepass_full_size_cityscapes_late_branch(sf):
    """BASELINE config 2 at full size on the top-k (late) branch: the speculative sums are void and the second walk must
    reproduce the torch-eager gradient."""
    from oracle import torch_restatement as tr
    B, C, h, w, H, W = 4, 19, 256, 512, 1024, 2048
    low, lab = _synth(1235, B, C, h, w, H, W, "late")
    x = low.cuda().requires_grad_(True)
    y = lab.cuda()
    loss = sf.fused_cross_entropy(x, y, None, 255, THR, None)
    loss.backward()
    rl, rg = tr.loss_and_grad(low.cuda(), y, "ohem")
    assert abs(loss.item() - rl.item()) <= 1e-5 * abs(rl.item())
    # OHEM is discontinuous at the k-th value: a pixel whose loss lies within float rounding of it may be kept by one
    # implementation and dropped by the other (its whole 1/K contribution moves: up to 4 cells x C channels).  Allow a
    # handful of such pixels out of 8.4 M; everything else must agree to 1e-5.
    bad = ((x.grad - rg).abs() > 1e-5 * rg.abs().max()).sum().item()
    assert bad <= 4 * 4 * C, bad


def test_branch_hint_switches_route_without_changing_results(sf, monkeypatch):
    """After a step that ended on the top-k branch the next forward takes the forward / finalize / backward kernels (the
    hint is read without a host sync, so synchronise first to make it visible); values equal the forced split route bit for bit."""
    from segmentation_factory_b200 import functional as fn
    low, lab = _synth(12, 2, 19, 64, 128, 256, 512, "late", blk=16)
    monkeypatch.setenv("SF_NO_ONEPASS", "1")
    l_ref, g_ref, _ = _run(sf, low, lab, None, THR, want_hist=False)
    monkeypatch.setenv("SF_NO_ONEPASS", "0")
    fn._BranchHint._by_device.clear()
    l1, g1, _ = _run(sf, low, lab, None, THR, want_hist=False)                  # speculates, re-walks
    torch.cuda.synchronize()
    assert fn._BranchHint.last_branch(torch.device("cuda", torch.cuda.current_device())) == 2
    l2, g2, _ = _run(sf, low, lab, None, THR, want_hist=False)                  # hinted: split route
    assert l2.item() == l_ref.item() and torch.equal(g2, g_ref)
    assert abs(l1.item() - l_ref.item()) <= 2e-6 * abs(l_ref.item())
    assert (g1 - g_ref).abs().max().item() <= 1e-5 * g_ref.abs().max().item()
    early, lab2 = _synth(11, 2, 19, 64, 128, 256, 512, "early", blk=16)
    _run(sf, early, lab2, None, THR, want_hist=False)                          # threshold branch again -> hint resets
    torch.cuda.synchronize()
    assert fn._BranchHint.last_branch(torch.device("cuda", torch.cuda.current_device())) == 1
    fn._BranchHint._by_device.clear()


def test_onepass_boundary_row_exchange_matches_halo_recompute(sf, monkeypatch):
    """Tiles hand the lower part of their last row group to the tile below through two float atomics on a zeroed row
    (h a multiple of the tile height); SF_NO_XCHG=1 recomputes that row group in the lower tile instead.  Same loss and matrix,
    gradients equal to rounding, and the exchange itself is bitwise reproducible."""
    low, lab = _synth(61, 2, 19, 64, 128, 256, 512, "early", blk=16)          # 8-row tiles: 8 boundary rows per plane
    l1, g1, h1 = _run(sf, low, lab, None, THR)
    l1b, g1b, _ = _run(sf, low, lab, None, THR)
    monkeypatch.setenv("SF_NO_XCHG", "1")
    l2, g2, h2 = _run(sf, low, lab, None, THR)
    assert torch.equal(g1, g1b) and l1.item() == l1b.item()
    assert l1.item() == l2.item() and torch.equal(h1, h2)
    assert (g1 - g2).abs().max().item() <= 1e-6 * g2.abs().max().item()
    assert not torch.equal(g1, torch.zeros_like(g1))
