"""The many-class one-pass walk (csrc/sf_onepass_chunk.cuh: 24 < C <= 255, 16-class chunks, two passes per row group; opt-in
with SF_ONEPASS_CHUNK=1 because the forward / backward kernel pair is faster at the BASELINE shapes) against the oracle and
against the forward / finalize / backward kernel route on the same inputs.
Run on the B200 box: python -m pytest tests -m gpu"""
import numpy as np
import pytest
import torch

from test_gpu_onepass import THR, _run, _supported
from test_gpu_parity import _bits, _synth

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _opt_in(monkeypatch):
    monkeypatch.setenv("SF_ONEPASS_CHUNK", "1")


@pytest.fixture(scope="module")
def sf():
    import segmentation_factory_b200 as m
    from segmentation_factory_b200 import _lib
    _lib.lib()
    assert torch.cuda.is_available()
    return m


# name, seed, B, C, h, w, H, W, kind, dtype, label dtype, loss kind, weighted, forced tile width, forced tile height
CHUNK_CASES = [
    ("c150_x4_ce", 41, 2, 150, 32, 48, 128, 192, "early", torch.float32, torch.int64, "ce", False, 0, 0),
    ("c150_x4_ce_wide_tiles", 42, 2, 150, 32, 64, 128, 256, "early", torch.float32, torch.int64, "ce", False, 192, 0),
    ("c150_x4_ohem_early", 43, 1, 150, 40, 72, 160, 288, "early", torch.float32, torch.int64, "ohem", False, 0, 0),
    ("c150_x4_ohem_late_topk", 44, 1, 150, 40, 72, 160, 288, "late", torch.float32, torch.int64, "ohem", False, 0, 0),
    ("c171_x4_bf16_ohem", 45, 2, 171, 32, 64, 128, 256, "early", torch.bfloat16, torch.int64, "ohem", False, 0, 0),
    ("c171_x4_bf16_ce_wide", 46, 1, 171, 24, 64, 96, 256, "early", torch.bfloat16, torch.uint8, "ce", True, 192, 5),
    ("c150_x8_fp16_ce_weighted_u8", 47, 1, 150, 18, 40, 144, 320, "early", torch.float16, torch.uint8, "ce", True, 0, 0),
    ("c32_x4_no_tail_chunk", 48, 1, 32, 27, 92, 108, 368, "early", torch.float32, torch.int64, "ohem", True, 0, 3),
    ("c25_x4_tail_of_nine", 49, 2, 25, 16, 36, 64, 144, "late", torch.float32, torch.int64, "ohem", False, 128, 2),
    ("c255_x4_most_classes", 50, 1, 255, 12, 52, 48, 208, "early", torch.float32, torch.int64, "ce", False, 0, 0),
    ("c150_x8_bf16_late", 51, 1, 150, 16, 48, 128, 384, "late", torch.bfloat16, torch.int64, "ohem", False, 0, 0),
    ("c150_x4_one_tall_tile", 52, 1, 150, 33, 36, 132, 144, "early", torch.float32, torch.int64, "ce", False, 128, 60),
]


@pytest.mark.parametrize("case", CHUNK_CASES, ids=[c[0] for c in CHUNK_CASES])
def test_chunked_onepass_vs_oracle_and_split_route(sf, oracle, case, monkeypatch):
    name, seed, B, C, h, w, H, W, kind, dtype, ldt, lk, weighted, nt, th = case
    low, lab = _synth(seed, B, C, h, w, H, W, kind, dtype, ldt, blk=16)
    assert _supported(low, lab), "case must exercise the one-pass kernels"
    cw = (torch.rand(C, generator=torch.Generator().manual_seed(7)) + 0.5) if weighted else None
    thr = THR if lk == "ohem" else None
    od = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[dtype]
    lown, labn = _bits(low), lab.numpy()
    cwn = None if cw is None else cw.numpy()
    if nt:
        monkeypatch.setenv("SF_DEBUG_OC_NT", str(nt))
    if th:
        monkeypatch.setenv("SF_DEBUG_OC_TH", str(th))

    monkeypatch.setenv("SF_NO_ONEPASS", "0")
    loss, grad, hist = _run(sf, low, lab, cw, thr, gscale=512.0)
    monkeypatch.setenv("SF_NO_ONEPASS", "1")
    split_hist = C * C * 4 <= 160 * 1024                                      # the kernel pair keeps its matrix in shared memory
    loss_s, grad_s, hist_s = _run(sf, low, lab, cw, thr, want_hist=split_hist, gscale=512.0)

    ref_hist = oracle.path_forward(lown, labn, 255, None, 0.0, od, want_loss=False, want_lse=False, want_arg=False)["hist"]
    assert np.array_equal(hist.cpu().numpy(), ref_hist)                       # bit-exact confusion matrix
    assert hist_s is None or torch.equal(hist, hist_s)
    if lk == "ohem":
        ref_loss, ref_grad, info = oracle.ohem_loss_and_grad(lown, labn, 255, cwn, ldtype=od)
        if name in ("c150_x4_ohem_late_topk", "c25_x4_tail_of_nine"):
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


def test_chunked_onepass_without_matrix_bitwise_reproducible(sf, oracle):
    low, lab = _synth(53, 2, 150, 32, 64, 128, 256, "early", blk=16)
    ref_loss, ref_grad = oracle.ce_loss_and_grad(low.numpy(), lab.numpy(), 255, None, 0)
    l1, g1, _ = _run(sf, low, lab, None, None, want_hist=False)
    l2, g2, _ = _run(sf, low, lab, None, None, want_hist=False)
    assert l1.item() == l2.item() and torch.equal(g1, g2)
    assert abs(l1.item() - ref_loss) <= 1e-5 * abs(ref_loss)
    assert np.abs(g1.cpu().numpy() - ref_grad).max() <= 1e-5 * np.abs(ref_grad).max()


def test_chunked_onepass_argmax_special_values(sf):
    """NaN (first NaN wins, also across chunks), +-inf, exact ties (lowest index) and +-0 maxima follow torch.argmax."""
    from oracle import torch_restatement as tr
    C = 150
    low, lab = _synth(54, 1, C, 32, 64, 128, 256, "early", blk=16)
    low = (low * 2).round() / 2                                               # ties, within and across chunks
    low[0, 40, 10, 20] = float("nan")
    low[0, 9, 10, 21] = float("nan")
    low[0, 130, 10, 21] = float("nan")
    low[0, 2, 3, 3] = float("inf")
    low[0, 100, 3, 5] = float("inf")
    low[0, 5, 20, 40] = float("-inf")
    low[0, :, 25:30, 10:30] = 0.0                                             # all-zero logits: every class ties at +0
    low[0, 3, 25:30, 10:30] = -0.0
    low[0, :, 16:20, 40:60] = -1.0
    low[0, 77, 16:20, 40:60] = -0.0                                           # maximum is -0
    low[0, 111, 16:20, 40:60] = 0.0                                           # ... and +0 in a later chunk: -0 == +0, 77 wins
    x = low.cuda().requires_grad_(True)
    hist = torch.zeros(C, C, dtype=torch.int64, device="cuda")
    loss = sf.fused_cross_entropy(x, lab.cuda(), None, 255, None, hist)
    ref = tr.confusion_counts(tr.upsample(low, (128, 256)), lab, C)
    assert torch.equal(hist.cpu(), ref)
    assert torch.isnan(loss).item()
    hist2 = torch.zeros(C, C, dtype=torch.int64, device="cuda")
    sf.fused_cross_entropy(x, lab.cuda(), None, 255, THR, hist2)
    assert torch.equal(hist2.cpu(), ref)


def test_chunked_onepass_full_size_cfg3(sf):
    """BASELINE configs[2] at full size (16 x 150 x 128 x 128 -> 512 x 512, plain CE): the one-pass walk against the kernel route."""
    import os
    low, lab = _synth(55, 16, 150, 128, 128, 512, 512, "early", blk=32)
    os.environ["SF_NO_ONEPASS"] = "0"
    l1, g1, h1 = _run(sf, low, lab, None, None)
    os.environ["SF_NO_ONEPASS"] = "1"
    try:
        l2, g2, h2 = _run(sf, low, lab, None, None)
    finally:
        os.environ["SF_NO_ONEPASS"] = "0"
    assert torch.equal(h1, h2) and int(h1.sum()) == int((lab != 255).sum())
    assert abs(l1.item() - l2.item()) <= 2e-6 * abs(l2.item())
    assert (g1 - g2).abs().max().item() <= 1e-5 * g2.abs().max().item()
