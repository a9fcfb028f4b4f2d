"""engine.criterion (CE + multiclass dice loss) on the GPU: goldens made by the reference's own
function, seeded cases against the oracle, the per-image sums through the raw C ABI, and size-independent properties at
the full Cityscapes shape.  Run on the B200 box: python -m pytest tests -m gpu"""
import numpy as np
import pytest
import torch

from conftest import criterion_names, load_criterion
from test_gpu_parity import TDT, _bits, _synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sf():
    import segmentation_factory_b200 as m
    from segmentation_factory_b200 import _lib
    _lib.lib()
    assert torch.cuda.is_available()
    return m


@pytest.mark.parametrize("no_tma", ["0", "1"])
@pytest.mark.parametrize("name", criterion_names())
def test_criterion_golden(sf, name, no_tma, monkeypatch):
    monkeypatch.setenv("SF_NO_TMA", no_tma)
    g = load_criterion(name)
    if g["logits_dtype"] == "float32":
        low = torch.from_numpy(g["logits"])
    else:
        low = torch.from_numpy(g["logits"].view(np.int16)).view(TDT[g["logits_dtype"]])
    C = low.shape[1]
    x = low.cuda().requires_grad_(True)
    cw = None if g["class_w"] is None else torch.from_numpy(g["class_w"]).cuda()
    loss = sf.criterion(x, torch.from_numpy(g["labels"]).cuda(), cw, num_classes=C, dice=bool(g["dice"]),
                        ignore_index=int(g["ignore_index"]))
    if not np.isfinite(g["loss"]):                                   # every pixel ignored: nan, like the reference
        assert np.isnan(loss.item())
        return
    assert abs(loss.item() - float(g["loss"])) <= 2e-5 * abs(float(g["loss"])), (loss.item(), float(g["loss"]))
    loss.backward()
    tol = 2e-5 if g["logits_dtype"] == "float32" else 2e-2
    err = np.abs(x.grad.float().cpu().numpy() - g["grad"]).max()
    assert err <= tol * np.abs(g["grad"]).max(), (err, np.abs(g["grad"]).max())


DICE_CASES = [
    # name, seed, B, C, h, w, H, W, kind, dtype, label dtype, weighted
    ("city_early", 101, 2, 19, 64, 128, 256, 512, "early", torch.float32, torch.int64, False),
    ("city_late_u8", 102, 2, 19, 64, 128, 256, 512, "late", torch.float32, torch.uint8, False),
    ("ade_c150_weighted", 103, 2, 150, 32, 32, 128, 128, "late", torch.float32, torch.int64, True),
    ("coco_bf16_c171", 104, 1, 171, 32, 32, 128, 128, "early", torch.bfloat16, torch.int64, False),
    ("voc_fp16_x8", 105, 2, 21, 20, 28, 160, 224, "late", torch.float16, torch.uint8, False),
    ("nonint_odd_w", 106, 1, 9, 50, 67, 167, 223, "early", torch.float32, torch.int64, True),
    ("downscale", 107, 1, 7, 150, 130, 100, 90, "early", torch.float32, torch.int64, False),
    ("identity", 108, 1, 19, 96, 128, 96, 128, "late", torch.float32, torch.int64, False),
    ("binary", 109, 3, 2, 32, 64, 128, 256, "early", torch.float32, torch.int64, True),
    ("fullres_bf16_c150_u8", 110, 2, 150, 64, 80, 64, 80, "early", torch.bfloat16, torch.uint8, True),
    ("fullres_fp16_late", 112, 3, 21, 96, 128, 96, 128, "late", torch.float16, torch.int64, False),
    ("fullres_fp32_weighted", 113, 2, 19, 128, 160, 128, 160, "late", torch.float32, torch.int64, True),
    ("fullres_odd_plane", 114, 1, 5, 37, 51, 37, 51, "early", torch.float32, torch.int64, False),     # plane % 4 != 0: generic identity walk
    ("x2_c24", 115, 1, 24, 40, 64, 80, 128, "late", torch.float32, torch.uint8, False),                # scale 2, C = 24: generic, one chunk
]


@pytest.mark.parametrize("no_fast", ["0", "1"])
@pytest.mark.parametrize("case", DICE_CASES, ids=[c[0] for c in DICE_CASES])
def test_criterion_seeded_vs_oracle(sf, oracle, case, no_fast, monkeypatch):
    monkeypatch.setenv("SF_NO_FAST", no_fast)                      # 1: force the generic kernels on fast-path shapes
    name, seed, B, C, h, w, H, W, kind, dtype, ldt, weighted = case
    low, lab = _synth(seed, B, C, h, w, H, W, kind, dtype, ldt, blk=16)
    cw = (torch.rand(C, generator=torch.Generator().manual_seed(7)) + 0.5) if weighted else None
    od = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[dtype]
    ref_loss, ref_grad, info = oracle.criterion_loss_and_grad(_bits(low), lab.numpy(), None if cw is None else cw.numpy(),
                                                              True, 255, od)
    x = low.cuda().requires_grad_(True)
    loss = sf.criterion(x, lab.cuda(), None if cw is None else cw.cuda(), num_classes=C, dice=True, ignore_index=255)
    (loss * 512.0).backward()
    got = x.grad.float().cpu().numpy() / 512.0
    assert abs(loss.item() - ref_loss) <= 2e-5 * abs(ref_loss), (loss.item(), ref_loss, info["ce"], info["dice"])
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    err = np.abs(got - ref_grad).max()
    assert err <= tol * np.abs(ref_grad).max(), (err, np.abs(ref_grad).max())
    # run-to-run bitwise reproducible (integer accumulation of the sums, gather-form backward)
    x2 = low.cuda().requires_grad_(True)
    loss2 = sf.criterion(x2, lab.cuda(), None if cw is None else cw.cuda(), num_classes=C, dice=True, ignore_index=255)
    (loss2 * 512.0).backward()
    assert loss2.item() == loss.item() and torch.equal(x2.grad, x.grad)


def _abi_stats(low, lab, ignore=255):
    """sf_ce_forward -> sf_loss_finalize -> sf_dice_stats -> sf_dice_finalize through raw pointers."""
    from segmentation_factory_b200 import _lib
    L = _lib.lib()
    B, C, h, w = low.shape
    _, H, W = lab.shape
    x, y = low.cuda(), lab.cuda()
    lc = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[low.dtype]
    yc = {torch.int64: 0, torch.uint8: 1}[lab.dtype]
    st = torch.cuda.current_stream().cuda_stream
    wsb, dwb = L.sf_workspace_bytes(), L.sf_dice_workspace_bytes(B, C)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    dws = torch.empty(dwb, dtype=torch.uint8, device="cuda")
    lse = torch.empty(B * H * W, dtype=torch.float32, device="cuda")
    out = torch.empty((), dtype=torch.float32, device="cuda")
    geo = (B, C, h, w, H, W, ignore)
    assert L.sf_ce_forward(x.data_ptr(), lc, y.data_ptr(), yc, *geo, None, -1.0, None, lse.data_ptr(), None, ws.data_ptr(), wsb, st) == 0
    assert L.sf_loss_finalize(None, B * H * W, -1.0, ws.data_ptr(), wsb, None, st) == 0
    assert L.sf_dice_stats(x.data_ptr(), lc, y.data_ptr(), yc, *geo, lse.data_ptr(), dws.data_ptr(), dwb, st) == 0
    assert L.sf_dice_finalize(B, C, 1e-6, ws.data_ptr(), dws.data_ptr(), dwb, out.data_ptr(), st) == 0
    raw = dws.cpu().numpy()
    stats = raw[64:64 + B * C * 24].view(np.uint64).reshape(B, C, 3).astype(np.float64)
    stats[..., 0] /= 2.0 ** 40
    stats[..., 1] /= 2.0 ** 26
    head = raw[:16].view(np.float32)
    return stats, float(head[0]), float(head[1]), float(out.item()), (L, x, y, lse, ws, dws, geo, lc, yc, st)


def test_dice_sums_c_abi_vs_oracle(sf, oracle):
    low, lab = _synth(111, 2, 19, 32, 64, 128, 256, "late", blk=16)
    lab[1, :, :100] = 255
    stats, dice, total, out, (L, x, y, lse, ws, dws, geo, lc, yc, st) = _abi_stats(low, lab)
    f = oracle.path_forward(low.numpy(), lab.numpy(), 255, None, 0.0, 0, want_arg=False, want_hist=False)
    ref = oracle.dice_stats(low.numpy(), lab.numpy(), f["lse_px"], 255, 0)
    assert np.array_equal(stats[..., 2], ref[..., 2])                                   # label counts: exact
    assert np.abs(stats[..., 0] - ref[..., 0]).max() <= 2e-6 * ref[..., 0].max() + 1e-9
    assert np.abs(stats[..., 1] - ref[..., 1]).max() <= 2e-6 * max(ref[..., 1].max(), 1.0)
    ref_dice, _, _ = oracle.dice_from_stats(ref)
    assert abs(dice - ref_dice) <= 1e-6 and out == total
    # error codes of the new entry points
    assert L.sf_dice_stats(x.data_ptr(), lc, y.data_ptr(), yc, *geo, None, dws.data_ptr(), dws.numel(), st) == -1
    assert L.sf_dice_stats(x.data_ptr(), lc, y.data_ptr(), yc, *geo, lse.data_ptr(), dws.data_ptr(), 8, st) == -4
    assert L.sf_dice_finalize(2, 19, 1e-6, ws.data_ptr(), dws.data_ptr(), 8, None, st) == -4


def test_dice_full_size_cityscapes_properties(sf):
    """BASELINE config 2 shape (8x19x256x512 -> 1024x2048): properties that hold at any size."""
    B, C = 8, 19
    low, lab = _synth(1234, B, C, 256, 512, 1024, 2048, "late")
    lab[3] = 255                                                                          # one image without a roi
    stats, dice, total, out, _ = _abi_stats(low, lab)
    n_valid = (lab != 255).flatten(1).sum(1).numpy().astype(np.float64)
    assert np.array_equal(stats[..., 2].sum(1), n_valid)                                  # every roi pixel counted once
    cnt = np.stack([(lab == c).flatten(1).sum(1).numpy() for c in range(C)], 1)
    assert np.array_equal(stats[..., 2], cnt.astype(np.float64))
    assert np.abs(stats[..., 0].sum(1) - n_valid).max() <= 2e-6 * n_valid.max()           # softmax sums to one per pixel
    assert (stats[..., 1] <= stats[..., 0] + 1e-6).all() and (stats[..., 1] <= stats[..., 2] + 1e-6).all()
    assert stats[3].sum() == 0.0                                                          # dead image: D = 1 for every class
    assert 0.0 <= dice <= 1.0
    x = low.cuda().requires_grad_(True)
    loss = sf.criterion(x, lab.cuda(), None, num_classes=C, dice=True, ignore_index=255)
    assert abs(loss.item() - total) <= 1e-6 * abs(total)       # criterion takes the fused CE + stats forward; _abi_stats the two launches
    loss.backward()
    g = x.grad
    assert torch.isfinite(g).all() and float(g[3].abs().max()) == 0.0
    # both terms are functions of softmax: the class-sum of the gradient vanishes at every low-res position
    assert float(g.sum(1).abs().max()) <= 1e-5 * float(g.abs().max())
    # dice off == the plain CE path
    x2 = low.cuda().requires_grad_(True)
    l2 = sf.criterion(x2, lab.cuda(), None, num_classes=C, dice=False, ignore_index=255)
    l3 = sf.CrossEntropy(255)(low.cuda(), lab.cuda())
    # l2 wants a gradient (one-pass walk), l3 does not (forward kernel): same per-pixel terms, different summation order
    assert abs(l2.item() - l3.item()) <= 2e-6 * abs(l3.item()) and abs((total - dice) - l3.item()) <= 1e-5 * abs(l3.item())


def test_fused_ce_dice_forward_matches_two_launch_route(sf, oracle):
    """sf_ce_dice_forward (one launch on integer-scale shapes: the log-sum-exp's exponentials feed the softmax sums) against
    sf_ce_forward + sf_dice_stats and the oracle."""
    from segmentation_factory_b200 import _lib
    L = _lib.lib()
    for (B, C, h, w, H, W, dt) in [(2, 19, 64, 128, 256, 512, torch.float32), (2, 21, 24, 64, 192, 512, torch.bfloat16),
                                   (1, 7, 50, 67, 167, 223, torch.float32)]:
        low, lab = _synth(120, B, C, h, w, H, W, "late", dt, blk=16)
        stats2, dice2, total2, _, (_, x, y, lse2, ws2, dws2, geo, lc, yc, st) = _abi_stats(low, lab)
        wsb, dwb = L.sf_workspace_bytes(), L.sf_dice_workspace_bytes(B, C)
        ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
        dws = torch.empty(dwb, dtype=torch.uint8, device="cuda")
        lse = torch.empty(B * H * W, dtype=torch.float32, device="cuda")
        out = torch.empty((), dtype=torch.float32, device="cuda")
        assert L.sf_ce_dice_forward(x.data_ptr(), lc, y.data_ptr(), yc, *geo, None, lse.data_ptr(), ws.data_ptr(), wsb,
                                    dws.data_ptr(), dwb, st) == 0
        assert L.sf_loss_finalize(None, B * H * W, -1.0, ws.data_ptr(), wsb, None, st) == 0
        assert L.sf_dice_finalize(B, C, 1e-6, ws.data_ptr(), dws.data_ptr(), dwb, out.data_ptr(), st) == 0
        raw = dws.cpu().numpy()
        stats = raw[64:64 + B * C * 24].view(np.uint64).reshape(B, C, 3).astype(np.float64)
        stats[..., 0] /= 2.0 ** 40
        stats[..., 1] /= 2.0 ** 26
        assert torch.equal(lse, lse2)                                                     # same lse bits as the plain CE forward
        assert np.array_equal(stats[..., 2], stats2[..., 2])
        assert np.abs(stats[..., 0] - stats2[..., 0]).max() <= 2e-6 * stats2[..., 0].max() + 1e-9
        assert np.abs(stats[..., 1] - stats2[..., 1]).max() <= 2e-6 * max(stats2[..., 1].max(), 1.0)
        assert abs(out.item() - total2) <= 1e-6 * abs(total2)
