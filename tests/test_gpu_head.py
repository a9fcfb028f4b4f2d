"""The heads' 1x1 classifier on tcgen05 (sf_head_classifier) against F.conv2d, the reference's own operator
(models/heads/segformer.py:39,57; upernet.py:32,50; fpn.py:23,37).  bf16 / fp16 tensor-core tolerance: the products are
exact, the fp32 accumulation order differs from cuDNN's, and the result is rounded once to the output dtype.
Run on the B200 box: python -m pytest tests -m gpu"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sf():
    import segmentation_factory_b200 as m
    from segmentation_factory_b200 import _lib
    _lib.lib()
    assert torch.cuda.is_available()
    return m


CASES = [
    # name, B, K, C, h, w, in dtype, out dtype, bias
    ("segformer_b0_city", 2, 256, 19, 64, 128, torch.bfloat16, torch.bfloat16, True),
    ("segformer_768_ade", 1, 768, 150, 32, 48, torch.bfloat16, torch.float32, True),
    ("uper_512_coco_fp16", 1, 512, 171, 24, 40, torch.float16, torch.float16, True),
    ("fpn_128_voc_nobias", 2, 128, 21, 40, 56, torch.float16, torch.float32, False),
    ("ragged_tail_tile", 1, 200, 9, 25, 24, torch.bfloat16, torch.float32, True),      # h*w = 600: not a multiple of 128; K % 64 != 0
    ("binary_many_tiles", 3, 64, 2, 96, 160, torch.bfloat16, torch.bfloat16, True),
    ("c172_two_accumulators", 2, 256, 172, 48, 64, torch.bfloat16, torch.bfloat16, True),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_head_classifier_vs_conv2d(sf, case):
    name, B, K, C, h, w, dt, odt, has_bias = case
    g = torch.Generator().manual_seed(hash(name) % 1000)
    x = torch.randn((B, K, h, w), generator=g).to(dt).cuda()
    wt = (torch.randn((C, K, 1, 1), generator=g) / K ** 0.5).to(dt).cuda()
    bias = torch.randn(C, generator=g).cuda() if has_bias else None
    with torch.no_grad():
        got = sf.head_classifier(x, wt, bias, odt)
        ref = F.conv2d(x.float(), wt.float(), bias)                 # fp32 math on the same bf16 / fp16 values
    assert got.shape == (B, C, h, w) and got.dtype == odt
    err = (got.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = 2e-5 if odt == torch.float32 else (8e-3 if odt == torch.bfloat16 else 1e-3)     # fp32 accumulate; one output rounding
    assert err <= tol * scale, (err, scale)


def test_head_classifier_feeds_the_loss_path(sf, oracle):
    """features -> classifier -> OhemCrossEntropy / confusion matrix without touching the logits in between; the module
    loads a Conv2d state_dict unchanged."""
    g = torch.Generator().manual_seed(5)
    B, K, C, h, w = 2, 256, 19, 32, 64
    conv = torch.nn.Conv2d(K, C, 1)
    head = sf.HeadClassifier(K, C)
    head.load_state_dict(conv.state_dict())
    head = head.cuda()
    x = torch.randn((B, K, h, w), generator=g).to(torch.bfloat16).cuda()
    logits = head(x)
    assert logits.dtype == torch.bfloat16 and logits.shape == (B, C, h, w)
    lab = torch.randint(0, C, (B, 4 * h, 4 * w), generator=g).cuda()
    hist = torch.zeros(C, C, dtype=torch.int64, device="cuda")
    loss = sf.fused_cross_entropy(logits, lab, None, 255, float(oracle.OHEM_THRESH), hist)
    lown = logits.view(torch.int16).cpu().numpy().view(np.uint16)
    ref_loss, _, _ = oracle.ohem_loss_and_grad(lown, lab.cpu().numpy(), 255, None, ldtype=1)
    assert abs(loss.item() - ref_loss) <= 1e-5 * abs(ref_loss)
    assert int(hist.sum()) == lab.numel()
    with pytest.raises(TypeError):
        sf.head_classifier(x.float(), conv.weight.cuda(), None)
    with pytest.raises(RuntimeError):
        sf.head_classifier(x.requires_grad_(True), conv.weight.cuda(), None)
