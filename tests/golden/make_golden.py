"""Generate tests/golden/*.npz by running the REFERENCE's own code.

Run in the dev container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports util.losses / util.metrics / util.utils from the reference checkout and torch's
F.interpolate (models/build_models.py:65), feeds them seeded synthetic inputs and
stores inputs + outputs.  The committed .npz files are what pins oracle/segpath_oracle.c; the
reference ships no tests for this path.  All spatial sizes keep H_out + W_out > 128 so torch takes
its generic NCHW bilinear kernel (the one real workloads hit; see oracle header).
"""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("SEGFACTORY_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
from util.losses import CrossEntropy, FocalLoss, OhemCrossEntropy  # noqa: E402
from util.metrics import Metrics  # noqa: E402
from util.utils import ConfusionMatrix  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(4)


def make_labels(g, B, C, H, W, blk=8, p_ignore=0.1, dtype=torch.int64):
    hb, wb = (H + blk - 1) // blk, (W + blk - 1) // blk
    coarse = torch.randint(0, C, (B, hb, wb), generator=g)
    coarse[torch.rand((B, hb, wb), generator=g) < p_ignore] = 255
    lab = coarse.repeat_interleave(blk, 1).repeat_interleave(blk, 2)[:, :H, :W].contiguous()
    return lab.to(dtype)


def make_logits(g, kind, B, C, h, w, labels, flip_p=0.03):
    if kind == "early":
        return 3.0 * torch.randn((B, C, h, w), generator=g)
    if kind == "ties":
        return torch.round(2.0 * torch.randn((B, C, h, w), generator=g)) * 0.5
    if kind == "late":
        H, W = labels.shape[1:]
        ys = ((torch.arange(h) + 0.5) * H / h).long().clamp(max=H - 1)
        xs = ((torch.arange(w) + 0.5) * W / w).long().clamp(max=W - 1)
        lab = labels[:, ys][:, :, xs].long()
        lab = torch.where(lab == 255, torch.zeros_like(lab), lab)
        flip = torch.rand(lab.shape, generator=g) < flip_p
        lab = torch.where(flip, torch.randint(0, C, lab.shape, generator=g), lab)
        return torch.randn((B, C, h, w), generator=g) + 12.0 * F.one_hot(lab, C).permute(0, 3, 1, 2).float()
    raise ValueError(kind)


def bits(t: torch.Tensor) -> np.ndarray:
    if t.dtype == torch.float32:
        return t.detach().numpy()
    return t.detach().view(torch.int16).numpy().view(np.uint16)


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_case(name, seed, B, C, hw, HW, kind, dtype=torch.float32, ldt=torch.int64, loss="ohem",
             weighted=False, store_up=False, all_ignored=False, flip_p=0.03, low_override=None):
    g = torch.Generator().manual_seed(seed)
    (h, w), (H, W) = hw, HW
    labels = make_labels(g, B, C, H, W, dtype=ldt)
    if all_ignored:
        labels.fill_(255)
    low = make_logits(g, kind, B, C, h, w, labels, flip_p).to(dtype) if low_override is None else low_override
    cw = torch.rand(C, generator=torch.Generator().manual_seed(7)) + 0.5 if weighted else None

    lowg = low.clone().requires_grad_(True)
    up = F.interpolate(lowg, size=(H, W), mode="bilinear", align_corners=False)      # build_models.py:65
    upf = up.float()                                                                 # autocast: CE runs in fp32
    crit = OhemCrossEntropy(255, cw) if loss == "ohem" else CrossEntropy(255, cw)
    val = crit(upf, labels.long())
    grad = None
    if torch.isfinite(val):
        val.backward()
        grad = lowg.grad.float().numpy()

    with torch.no_grad():
        m = Metrics(C, 255, "cpu")
        m.update(upf, labels.long().flatten())                                       # engine.py:91
        cm = ConfusionMatrix(C)
        cm.update(labels.long().flatten(), upf.argmax(1).flatten())                  # engine.py:90
        loss_px = F.cross_entropy(upf, labels.long(), weight=cw, ignore_index=255, reduction="none")
        n_valid = int((labels != 255).sum())
        n_hard = int((loss_px > crit.thresh).sum()) if loss == "ohem" else -1
        iou, f1, acc = m.compute_iou(), m.compute_f1(), m.compute_pixel_acc()
        acc_g, acc_c, iu = cm.compute()

    d = dict(
        logits=bits(low), logits_dtype=str(dtype).replace("torch.", ""), labels=labels.numpy(),
        size=np.array([H, W]), kind=kind, loss_kind=loss,
        class_w=(cw.numpy() if weighted else np.zeros(0, np.float32)),
        loss=np.float32(val.item()), grad=(grad if grad is not None else np.zeros(0, np.float32)),
        argmax=upf.argmax(1).numpy().astype(np.uint8 if C <= 256 else np.int32),
        hist_f32=m.hist.numpy(), mat_i64=cm.mat.numpy(),
        loss_px=loss_px.numpy().astype(np.float32), n_valid=n_valid, n_hard=n_hard,
        up_sha256=digest(bits(up)),
        iou=np.array(iou[0]), miou=iou[1], f1=np.array(f1[0]), mf1=f1[1], acc=np.array(acc[0]), macc=acc[1],
        cm_acc_global=acc_g.numpy(), cm_acc=acc_c.numpy(), cm_iu=iu.numpy(),
    )
    if store_up:
        d["up"] = bits(up)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(f"{name:28s} loss={val.item():.6f} n_valid={n_valid} n_hard={n_hard} n_min={n_valid // 16} "
          f"miou={iou[1]} up={d['up_sha256'][:12]}")


def focal_case(name, seed, B, C, hw, HW, kind, alpha, gamma, size_average, dtype=torch.float32):
    """FocalLoss (util/losses.py:9-25) through the upsample, value and gradient."""
    g = torch.Generator().manual_seed(seed)
    (h, w), (H, W) = hw, HW
    labels = make_labels(g, B, C, H, W)
    low = make_logits(g, kind, B, C, h, w, labels).to(dtype)
    lowg = low.clone().requires_grad_(True)
    up = F.interpolate(lowg, size=(H, W), mode="bilinear", align_corners=False).float()
    val = FocalLoss(alpha, gamma, size_average, 255)(up, labels)
    val.backward()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), logits=bits(low), logits_dtype=str(dtype).replace("torch.", ""),
                        labels=labels.numpy(), size=np.array([H, W]), alpha=np.float32(alpha), gamma=np.float32(gamma),
                        size_average=np.int32(size_average), loss=np.float32(val.item()), grad=lowg.grad.float().numpy())
    print(f"{name:28s} focal={val.item():.6f}")


def criterion_case(name, seed, B, C, hw, HW, kind, weighted=False, dice=True, dtype=torch.float32, dead_image=False,
                   ignore_index=255, all_ignored=False):
    """engine.criterion (engine.py:10-15): CE + multiclass dice loss through the upsample."""
    from engine import criterion                                                    # the reference's own function
    g = torch.Generator().manual_seed(seed)
    (h, w), (H, W) = hw, HW
    labels = make_labels(g, B, C, H, W, p_ignore=0.1 if ignore_index == 255 else 0.0)
    if dead_image:
        labels[0].fill_(255)                                                        # empty roi -> sets_sum == 0 branch
    if all_ignored:
        labels.fill_(255)
    low = make_logits(g, kind, B, C, h, w, labels).to(dtype)
    cw = torch.rand(C, generator=torch.Generator().manual_seed(7)) + 0.5 if weighted else None
    lowg = low.clone().requires_grad_(True)
    up = F.interpolate(lowg, size=(H, W), mode="bilinear", align_corners=False).float()
    val = criterion(up, labels, cw, num_classes=C, dice=dice, ignore_index=ignore_index)
    if torch.isfinite(val):
        val.backward()
    else:
        lowg.grad = torch.zeros_like(lowg)
    with torch.no_grad():
        ce = F.cross_entropy(up, labels, ignore_index=ignore_index, weight=cw)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), logits=bits(low), logits_dtype=str(dtype).replace("torch.", ""),
                        labels=labels.numpy(), size=np.array([H, W]), dice=np.int32(dice), ignore_index=np.int32(ignore_index),
                        class_w=(cw.numpy() if weighted else np.zeros(0, np.float32)),
                        loss=np.float32(val.item()), ce=np.float32(ce.item()), grad=lowg.grad.float().numpy())
    print(f"{name:28s} criterion={val.item():.6f} ce={ce.item():.6f}")


def criterion_cases():
    criterion_case("crit_city_early", 61, 2, 19, (16, 32), (64, 128), "early")
    criterion_case("crit_city_late", 62, 2, 19, (16, 32), (64, 128), "late")
    criterion_case("crit_binary_weighted", 63, 2, 2, (16, 32), (64, 128), "early", weighted=True)
    criterion_case("crit_ade_nonint", 64, 1, 150, (18, 23), (64, 80), "late")
    criterion_case("crit_voc_bf16", 65, 1, 21, (30, 41), (100, 135), "early", dtype=torch.bfloat16)
    criterion_case("crit_dead_image", 66, 2, 19, (16, 32), (64, 128), "early", dead_image=True)
    criterion_case("crit_nodice_weighted", 67, 1, 19, (16, 32), (64, 128), "early", weighted=True, dice=False)
    criterion_case("crit_default_ignore_m100", 68, 2, 19, (16, 32), (64, 128), "late", ignore_index=-100)   # engine.py:10 default
    criterion_case("crit_all_ignored", 69, 1, 19, (16, 32), (64, 128), "early", all_ignored=True)            # CE = nan -> nan


def label_lut_case():
    """Label ingest: the datasets' own encode functions on random raw ids."""
    import importlib.util

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m
    cs = load("ref_cityscapes", "datasets/cityscapes.py").Cityscapes
    coco_cls = load("ref_coco_stuff", "datasets/coco_stuff.py").COCOStuff
    g = np.random.default_rng(5)
    raw_cs = g.integers(0, 34, (3, 67, 131), dtype=np.uint8)            # Cityscapes label ids 0..33
    enc_cs = cs.encode_target(raw_cs)                                   # cityscapes.py:130-131
    coco = coco_cls.__new__(coco_cls)                                   # __init__ needs the image folder; its table code:
    coco.use_id = [i for i in range(182) if i not in coco.UNUSEID]      # coco_stuff.py:66
    coco.label_map = np.arange(256)                                     # :67
    for i, k in enumerate(coco.use_id):                                 # :68-69
        coco.label_map[k] = i
    raw_coco = g.integers(0, 256, (2, 50, 77), dtype=np.uint8)
    enc_coco = coco.encode(raw_coco).numpy()                            # :98-100
    np.savez_compressed(os.path.join(OUT, "label_lut.npz"), cs_table=np.asarray(cs.id_to_train_id), cs_raw=raw_cs,
                        cs_enc=np.asarray(enc_cs), coco_table=coco.label_map, coco_raw=raw_coco, coco_enc=enc_coco)
    print("label_lut", np.asarray(cs.id_to_train_id).shape, enc_cs.dtype, enc_coco.dtype)


def hist_f32_case():
    """Metrics.hist float32 accumulation past 2**24 and compute_* on a big matrix."""
    g = torch.Generator().manual_seed(99)
    C = 19
    m = Metrics(C, 255, "cpu")
    counts = []
    for _ in range(6):
        c = torch.randint(0, 9_000_000, (C, C), generator=g, dtype=torch.int64)
        c[torch.rand((C, C), generator=g) < 0.5] //= 1000
        c[5, :] = 0
        c[:, 5] = 0                                       # an absent class -> NaN IoU, filtered
        counts.append(c)
        m.hist += c                                        # metrics.py:27 (float32 += int64)
    iou, f1, acc = m.compute_iou(), m.compute_f1(), m.compute_pixel_acc()
    cm = ConfusionMatrix(C)
    cm.mat = torch.stack(counts).sum(0)
    acc_g, acc_c, iu = cm.compute()
    np.savez_compressed(os.path.join(OUT, "hist_f32_accum.npz"), counts=torch.stack(counts).numpy(),
                        hist_f32=m.hist.numpy(), iou=np.array(iou[0]), miou=iou[1], f1=np.array(f1[0]), mf1=f1[1],
                        acc=np.array(acc[0]), macc=acc[1], cm_acc_global=acc_g.numpy(), cm_acc=acc_c.numpy(),
                        cm_iu=iu.numpy(), cm_str=str(cm))
    print("hist_f32_accum", iou[1], f1[1], acc[1])


if __name__ == "__main__":
    if sys.argv[1:] == ["lut"]:
        label_lut_case()
        sys.exit(0)
    if sys.argv[1:] == ["crit"]:
        criterion_cases()
        sys.exit(0)
    run_case("city_early_ohem", 1234, 2, 19, (16, 32), (64, 128), "early")
    run_case("city_late_ohem", 1235, 2, 19, (16, 32), (64, 128), "late", flip_p=0.01)
    run_case("city_early_ce_u8", 1236, 1, 19, (16, 32), (64, 128), "early", ldt=torch.uint8, loss="ce")
    run_case("ade_ce_weighted", 2345, 1, 150, (16, 20), (64, 80), "early", loss="ce", weighted=True)
    run_case("ade_ohem_weighted", 2346, 1, 150, (16, 20), (64, 80), "late", loss="ohem", weighted=True)
    run_case("nonint_ce", 11, 1, 5, (40, 53), (133, 151), "early", loss="ce", store_up=True)
    run_case("nonint_ohem_x2", 12, 2, 9, (33, 40), (66, 80), "late", loss="ohem")
    run_case("downscale_ce", 13, 1, 4, (90, 120), (60, 100), "early", loss="ce")
    run_case("coco_bf16_ohem", 3456, 1, 171, (16, 20), (64, 80), "early", dtype=torch.bfloat16)
    run_case("voc_bf16_nonint_ce", 3457, 1, 21, (30, 41), (100, 135), "early", dtype=torch.bfloat16, loss="ce",
             ldt=torch.uint8, store_up=True)
    run_case("city_fp16_ohem", 3458, 1, 19, (16, 32), (64, 128), "late", dtype=torch.float16, flip_p=0.005)
    run_case("fp16_nonint_ce", 3459, 1, 7, (30, 41), (100, 135), "early", dtype=torch.float16, loss="ce")
    run_case("ties_ce", 21, 1, 19, (16, 32), (64, 128), "ties", loss="ce")
    run_case("ties_bf16_ohem", 22, 1, 19, (16, 32), (64, 128), "ties", dtype=torch.bfloat16)
    run_case("identity_ohem", 31, 1, 19, (64, 96), (64, 96), "early")
    run_case("identity_w_only_ce", 32, 1, 6, (40, 100), (80, 100), "early", loss="ce")
    run_case("binary_ce_weighted", 33, 2, 2, (16, 32), (64, 128), "early", loss="ce", weighted=True)
    run_case("all_ignored_ohem", 41, 1, 19, (16, 32), (64, 128), "early", all_ignored=True)
    run_case("all_ignored_ce", 42, 1, 19, (16, 32), (64, 128), "early", loss="ce", all_ignored=True)
    hist_f32_case()
    focal_case("focal_g2_mean", 51, 2, 19, (16, 32), (64, 128), "early", 0.25, 2.0, True)
    focal_case("focal_default", 52, 1, 19, (16, 32), (64, 128), "late", 1, 0, True)
    focal_case("focal_g15_sum_bf16", 53, 1, 21, (30, 41), (100, 135), "early", 0.5, 1.5, False, torch.bfloat16)
    criterion_cases()
    label_lut_case()
