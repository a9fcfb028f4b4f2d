"""Integration on the GPU: (1) the ops are CUDA-graph capturable (no host sync, caller's stream, caller-owned memory) and a
replayed graph reproduces the eager results bit for bit; (2) a tiny training loop through the drop-in API follows the same
parameter trajectory as the reference formulation (F.interpolate + torch losses), also under autocast + GradScaler as in
engine.py:41-52.  Run on the B200 box: python -m pytest tests -m gpu"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from test_gpu_parity import _synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sf():
    import segmentation_factory_b200 as m
    from segmentation_factory_b200 import _lib
    _lib.lib()
    assert torch.cuda.is_available()
    return m


def test_cuda_graph_capture_and_replay(sf, monkeypatch):
    # eager steps would switch kernels after a top-k step (branch hint); a captured graph replays one fixed route
    monkeypatch.setenv("SF_NO_BRANCH_HINT", "1")
    B, C = 2, 19
    sets = [_synth(300 + i, B, C, 64, 128, 256, 512, "early" if i % 2 == 0 else "late") for i in range(3)]
    x_static = sets[0][0].cuda().clone().requires_grad_(True)
    y_static = sets[0][1].cuda().clone()
    hist = torch.zeros(C, C, dtype=torch.int64, device="cuda")
    crit = sf.OhemCrossEntropy(255)

    def step():
        x_static.grad = None
        loss = sf.fused_cross_entropy(x_static, y_static, None, 255, float(crit.thresh), hist)
        loss.backward()
        dl = sf.criterion(x_static.detach(), y_static, None, num_classes=C, dice=True, ignore_index=255)
        return loss.detach(), dl.detach(), x_static.grad

    # eager results for every input set
    eager = []
    for lo, la in sets:
        with torch.no_grad():
            x_static.copy_(lo.cuda()); y_static.copy_(la.cuda())
        hist.zero_()
        l, d, g = step()
        eager.append((l.item(), d.item(), g.clone(), hist.clone()))

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                       # warm-up on the capture stream (lazy module loads, TMA encoder)
        step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        l_g, d_g, g_g = step()
    for (lo, la), (l, d, g, h) in zip(sets, eager):
        with torch.no_grad():
            x_static.copy_(lo.cuda()); y_static.copy_(la.cuda())
        hist.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert l_g.item() == l and d_g.item() == d                      # OHEM branch (thresh vs top-k) is decided on the device
        assert torch.equal(g_g, g) and torch.equal(hist, h)


class TinyHead(torch.nn.Module):
    """A stand-in for backbone + decode head: features -> low-resolution logits (1/4 scale)."""

    def __init__(self, C):
        super().__init__()
        self.body = torch.nn.Conv2d(3, 16, 3, stride=4, padding=1)
        self.cls = torch.nn.Conv2d(16, C, 1)

    def forward(self, img):
        return self.cls(F.relu(self.body(img)))


def _reference_criterion(logits_low, labels, C, dice):
    """engine.py:10-15 on upsampled logits, written with the same torch calls as the reference."""
    up = F.interpolate(logits_low, size=labels.shape[-2:], mode="bilinear", align_corners=False)
    loss = F.cross_entropy(up.float(), labels, ignore_index=255)
    if dice:
        x = F.softmax(up.float(), dim=1)
        roi = (labels != 255)
        onehot = F.one_hot(torch.where(roi, labels, torch.zeros_like(labels)), C).permute(0, 3, 1, 2).float()
        m = roi.unsqueeze(1).float()
        inter = (x * onehot * m).flatten(2).sum(2)
        sets = (x * m).flatten(2).sum(2) + (onehot * m).flatten(2).sum(2)
        d = (2 * inter + 1e-6) / (sets + 1e-6)                            # no empty roi in this test
        loss = loss + 1 - d.mean()
    return loss


@pytest.mark.parametrize("mode", ["ohem_fp32", "criterion_dice_fp32", "criterion_dice_autocast"])
def test_tiny_training_loop_follows_reference(sf, mode):
    torch.manual_seed(0)
    C, B = 19, 2
    g = torch.Generator().manual_seed(5)
    imgs = [torch.randn((B, 3, 128, 256), generator=g).cuda() for _ in range(4)]
    labs = [_synth(400 + i, B, C, 32, 64, 128, 256, "early", blk=16)[1].cuda() for i in range(4)]
    ours, ref = TinyHead(C).cuda(), TinyHead(C).cuda()
    ref.load_state_dict(ours.state_dict())
    opt_o = torch.optim.SGD(ours.parameters(), lr=0.05, momentum=0.9)
    opt_r = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9)
    amp = mode.endswith("autocast")
    sc_o = torch.amp.GradScaler("cuda", enabled=amp, init_scale=1024.0)
    sc_r = torch.amp.GradScaler("cuda", enabled=amp, init_scale=1024.0)
    ohem_ref = None
    for step in range(8):
        img, lab = imgs[step % 4], labs[step % 4]
        opt_o.zero_grad(); opt_r.zero_grad()
        with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
            lo, lr_ = ours(img), ref(img)
            if mode == "ohem_fp32":
                l_o = sf.OhemCrossEntropy(255)(lo, lab)
                up = F.interpolate(lr_, size=lab.shape[-2:], mode="bilinear", align_corners=False)
                px = F.cross_entropy(up, lab, ignore_index=255, reduction="none").view(-1)
                hard = px[px > 0.35667494]
                n_min = int((lab != 255).sum()) // 16
                if hard.numel() < n_min:
                    hard, _ = px.topk(n_min)
                l_r = hard.mean()
            else:
                l_o = sf.criterion(lo, lab, None, num_classes=C, dice=True, ignore_index=255)
                l_r = _reference_criterion(lr_, lab, C, True)
        tol = 2e-3 if amp else 2e-5
        assert abs(l_o.item() - l_r.item()) <= tol * abs(l_r.item()), (step, l_o.item(), l_r.item())
        sc_o.scale(l_o).backward(); sc_r.scale(l_r).backward()
        sc_o.step(opt_o); sc_r.step(opt_r)
        sc_o.update(); sc_r.update()
    for (n, po), pr in zip(ours.named_parameters(), ref.parameters()):
        err = (po - pr).abs().max().item()
        assert err <= (5e-3 if amp else 2e-4) * max(pr.abs().max().item(), 1e-3), (n, err)
