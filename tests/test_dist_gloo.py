"""World-size-2 CPU (gloo) test of the path's only collective: the int64 matrix all-reduce behind
Metrics / ConfusionMatrix.reduce_from_all_processes (util/metrics.py:108-114, util/utils.py:125-131).
Each rank accumulates its own shard of per-call counts; after the reduce both hold the global matrix and
the derived mIoU / F1 / acc equal a single-process run over all shards."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, counts, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import segmentation_factory_b200 as sf
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    C = counts.shape[-1]
    m = sf.Metrics(C, 255, "cpu")
    cm = sf.ConfusionMatrix(C)
    cm._ensure(torch.device("cpu"))
    for c in counts[rank::world]:                      # this rank's shard of the evaluation batches
        m._accumulate(torch.from_numpy(c))
        cm.mat += torch.from_numpy(c)
    cm.reduce_from_all_processes()
    m.reduce_from_all_processes()
    q.put((rank, cm.mat.numpy().copy(), m.hist.numpy().copy(), m.hist_i64.numpy().copy(),
           m.compute_iou(), m.compute_f1(), m.compute_pixel_acc(), cm.compute()[2].numpy().copy()))
    dist.destroy_process_group()


def test_matrix_all_reduce_world2():
    import segmentation_factory_b200 as sf
    g = np.random.default_rng(3)
    C = 19
    counts = g.integers(0, 50_000, size=(6, C, C)).astype(np.int64)
    counts[:, 7, :] = 0
    counts[:, :, 7] = 0                                # an absent class: NaN IoU must be filtered identically
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, counts, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    total = counts.sum(0)
    single = sf.Metrics(C, 255, "cpu")
    for c in counts:
        single._accumulate(torch.from_numpy(c))
    cm1 = sf.ConfusionMatrix(C)
    cm1.mat = torch.from_numpy(total)
    for rank, mat, hist, hist_i64, iou, f1, acc, iu in res:
        assert np.array_equal(mat, total)
        assert np.array_equal(hist_i64, total)
        assert np.array_equal(hist, total.astype(np.float32))       # exact regime (< 2**24 per cell)
        for got, ref in ((iou, single.compute_iou()), (f1, single.compute_f1()), (acc, single.compute_pixel_acc())):
            assert np.array_equal(np.array(got[0]), np.array(ref[0]), equal_nan=True) and got[1] == ref[1]
        assert np.array_equal(iu, cm1.compute()[2].numpy(), equal_nan=True)


def test_reduce_is_a_noop_without_process_group():
    import segmentation_factory_b200 as sf
    m = sf.Metrics(4, 255, "cpu")
    m._accumulate(torch.ones(4, 4, dtype=torch.int64))
    m.reduce_from_all_processes()
    cm = sf.ConfusionMatrix(4)
    cm.reduce_from_all_processes()                      # mat is None: must not raise (util/utils.py:125-129 returns early)
    assert float(m.hist.sum()) == 16.0
