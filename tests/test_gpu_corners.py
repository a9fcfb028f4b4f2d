"""Corners the round-1 review asked for: every class count train_gpu.py:48-49 lists ({2, 9, 19, 21, 151, 172}) at x4 / x8,
the literal engine.py:90-91 call pair on full-resolution logits with flattened labels, the documented deviations pinned
by tests, and the path's one collective over NCCL.  Run on the B200 box: python -m pytest tests -m gpu"""
import os
import socket

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


# name, seed, B, C, h, w, H, W, kind, dtype, label dtype, loss kind, weighted
CLASS_CASES = [
    ("kvasir_c2_x4_u8", 101, 2, 2, 48, 64, 192, 256, "early", torch.float32, torch.uint8, "ce", True),
    ("kvasir_c2_x8_ohem", 102, 1, 2, 24, 40, 192, 320, "early", torch.float32, torch.int64, "ohem", False),
    ("synapse_c9_x4_ohem", 103, 2, 9, 56, 56, 224, 224, "late", torch.float32, torch.int64, "ohem", False),
    ("synapse_c9_x8_bf16", 104, 1, 9, 28, 40, 224, 320, "early", torch.bfloat16, torch.int64, "ce", False),
    ("ade_c151_x4_ce", 105, 2, 151, 32, 48, 128, 192, "early", torch.float32, torch.int64, "ce", False),
    ("ade_c151_x8_ohem", 106, 1, 151, 24, 32, 192, 256, "late", torch.float32, torch.int64, "ohem", True),
    ("coco_c172_x4_bf16", 107, 2, 172, 32, 48, 128, 192, "early", torch.bfloat16, torch.int64, "ohem", False),
    ("coco_c172_x8_ce", 108, 1, 172, 24, 32, 192, 256, "early", torch.float32, torch.uint8, "ce", False),
]


@pytest.mark.parametrize("case", CLASS_CASES, ids=[c[0] for c in CLASS_CASES])
def test_reference_class_counts_vs_oracle(sf, oracle, case):
    name, seed, B, C, h, w, H, W, kind, dtype, ldt, lk, weighted = case
    low, lab = _synth(seed, B, C, h, w, H, W, kind, dtype, ldt, blk=16)
    cw = (torch.rand(C, generator=torch.Generator().manual_seed(7)) + 0.5) if weighted else None
    od = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[dtype]
    lown, labn = _bits(low), lab.numpy()
    cwn = None if cw is None else cw.numpy()
    x = low.cuda().requires_grad_(True)
    y = lab.cuda()
    hist = torch.zeros(C, C, dtype=torch.int64, device="cuda")          # C = 172: the 118 KB shared-memory matrix
    loss = sf.fused_cross_entropy(x, y, None if cw is None else cw.cuda(), 255, THR if lk == "ohem" else None, hist)
    loss.backward()
    ref_hist = oracle.path_forward(lown, labn, 255, None, 0.0, od, want_loss=False, want_lse=False, want_arg=False)["hist"]
    assert np.array_equal(hist.cpu().numpy(), ref_hist)
    if lk == "ohem":
        ref_loss, ref_grad, _ = oracle.ohem_loss_and_grad(lown, labn, 255, cwn, ldtype=od)
    else:
        ref_loss, ref_grad = oracle.ce_loss_and_grad(lown, labn, 255, cwn, od)
    assert abs(loss.item() - ref_loss) <= 1e-5 * abs(ref_loss), (loss.item(), ref_loss)
    gtol = 1e-5 if dtype == torch.float32 else 2e-2
    err = np.abs(x.grad.float().cpu().numpy() - ref_grad).max()
    assert err <= gtol * np.abs(ref_grad).max(), (err, np.abs(ref_grad).max())
    hist2 = torch.zeros(C, C, dtype=torch.int64, device="cuda")
    sf.eval_hist(x.detach(), y, 255, hist2)                               # eval-only kernel: same matrix
    assert np.array_equal(hist2.cpu().numpy(), ref_hist)


def test_engine_evaluate_call_pair_on_full_res_logits(sf):
    """engine.py:90-91 verbatim: confmat.update(labels.flatten(), outputs.argmax(1).flatten());
    metric.update(outputs, labels.flatten()) with outputs at label resolution (the unmodified model)."""
    from oracle import torch_restatement as tr
    B, C, H, W = 2, 19, 96, 160
    low, lab = _synth(201, B, C, H, W, H, W, "early", blk=16)
    outputs, labels = low.cuda(), lab.cuda()
    metric = sf.Metrics(C, 255, "cuda")
    confmat = sf.ConfusionMatrix(C)
    confmat.update(labels.flatten(), outputs.argmax(1).flatten())
    metric.update(outputs, labels.flatten())
    ref = tr.confusion_counts(low, lab, C)
    assert torch.equal(metric.hist.cpu(), ref.float())
    # ConfusionMatrix keeps every label in [0, C): 255 is outside, so both masks coincide here (util/utils.py:106)
    assert torch.equal(confmat.mat.cpu(), ref)
    iou, miou = metric.compute_iou()
    h = ref.float()
    ious = h.diag() / (h.sum(0) + h.sum(1) - h.diag())
    assert miou == round(ious[~ious.isnan()].mean().item() * 100, 2)
    # an in-range ignore label makes the two reference accumulators differ: the fused form refuses it
    with pytest.raises(ValueError):
        sf.fused_eval_update(outputs, labels, sf.Metrics(C, 0, "cuda"), sf.ConfusionMatrix(C))


def test_documented_deviations_are_pinned(sf):
    """(1) int64 labels are read as their low 32 bits; (2) labels outside [0, C) other than the ignore
    value are skipped (torch device-asserts); (3) a C x C int32 matrix that does not fit shared memory is an error,
    while the loss alone works for any C."""
    from segmentation_factory_b200._lib import SegFactoryError
    B, C, h, w, H, W = 1, 19, 32, 64, 128, 256
    low, lab = _synth(202, B, C, h, w, H, W, "early", blk=16)
    x, y = low.cuda(), lab.cuda()
    base = torch.zeros(C, C, dtype=torch.int64, device="cuda")
    sf.eval_hist(x, y, 255, base)
    # (1) 2**32 + k is counted as class k
    y_hi = y.clone()
    sel = y_hi == 3
    y_hi[sel] = 2 ** 32 + 3
    h1 = torch.zeros_like(base)
    sf.eval_hist(x, y_hi, 255, h1)
    assert torch.equal(h1, base)
    # (2) out-of-range labels (here 100 and -7) take no part in matrix, loss or gradient
    y_bad = y.clone()
    y_bad[sel] = 100
    y_bad[y == 5] = -7
    h2 = torch.zeros_like(base)
    x2 = low.cuda().requires_grad_(True)
    l2 = sf.fused_cross_entropy(x2, y_bad, None, 255, None, h2)
    l2.backward()
    y_ign = y.clone()
    y_ign[sel] = 255
    y_ign[y == 5] = 255
    h3 = torch.zeros_like(base)
    x3 = low.cuda().requires_grad_(True)
    l3 = sf.fused_cross_entropy(x3, y_ign, None, 255, None, h3)
    l3.backward()
    assert torch.equal(h2, h3) and l2.item() == l3.item() and torch.equal(x2.grad, x3.grad)
    # (3) too many classes for the shared-memory matrix
    big = torch.randn(1, 260, 16, 32, device="cuda")
    yb = torch.zeros(1, 64, 128, dtype=torch.long, device="cuda")
    with pytest.raises(SegFactoryError, match="shared memory"):
        sf.eval_hist(big, yb, 255, torch.zeros(260, 260, dtype=torch.int64, device="cuda"))
    assert torch.isfinite(sf.CrossEntropy()(big, yb)).item()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    import segmentation_factory_b200 as sf
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    C = 19
    low, lab = _synth(300 + rank, 2, C, 32, 64, 128, 256, "early", blk=16)      # this rank's shard of the validation set
    m, cm = sf.Metrics(C, 255, f"cuda:{rank}"), sf.ConfusionMatrix(C)
    sf.fused_eval_update(low.cuda(), lab.cuda(), m, cm)
    m2, cm2 = sf.Metrics(C, 255, f"cuda:{rank}"), sf.ConfusionMatrix(C)
    sf.fused_eval_update(low.cuda(), lab.cuda(), m2, cm2)
    cm.reduce_from_all_processes()                   # engine.py:100-101: two collectives
    m.reduce_from_all_processes()
    sf.reduce_from_all_processes(m2, cm2)            # the packed form: one collective
    torch.cuda.synchronize()
    q.put((rank, cm.mat.cpu().numpy(), m.hist.cpu().numpy(), cm2.mat.cpu().numpy(), m2.hist.cpu().numpy(), m.compute_iou()[1]))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run with gpurun --gpus 2)")
def test_nccl_reduce_from_all_processes_two_ranks(sf):
    import torch.multiprocessing as mp
    from oracle import torch_restatement as tr
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    total = torch.zeros(19, 19, dtype=torch.int64)
    for r in range(2):
        low, lab = _synth(300 + r, 2, 19, 32, 64, 128, 256, "early", blk=16)
        total += tr.confusion_counts(tr.upsample(low, (128, 256)), lab, 19)
    for rank, mat, hist, mat2, hist2, miou in res:
        assert np.array_equal(mat, total.numpy()) and np.array_equal(mat2, total.numpy())
        assert np.array_equal(hist, total.float().numpy()) and np.array_equal(hist2, total.float().numpy())
    assert res[0][5] == res[1][5]
