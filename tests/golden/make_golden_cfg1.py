"""BASELINE config 1 on REAL decode-head output: MiT-B0 + SegFormerHead(150) from the reference checkout
(models/backbones/mit.py:192-218, models/heads/segformer.py:42-58), random-init weights (seed 0), eval mode, then the
reference's OhemCrossEntropy / Metrics / ConfusionMatrix on its logits.  Reduced input (1x3x160x160 -> logits
1x150x40x40, labels 160x160) so the fixture stays under 1 MB; SURVEY 8(c) describes the loading trick: `import models`
pulls timm / fvcore, so the two files are loaded by path with a stub `models.layers` exposing DropPath / trunc_normal_.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_cfg1.py        (dev container only)
"""
import importlib.util
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (imports the reference's util package)

REF = mg.REF


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def build_logits():
    drop = _load("_ref_drop_path", os.path.join(REF, "models", "layers", "drop_path.py"))
    init = _load("_ref_initialize", os.path.join(REF, "models", "layers", "initialize.py"))
    stub = types.ModuleType("models.layers")
    stub.DropPath = drop.DropPath
    stub.trunc_normal_ = init.trunc_normal_
    pkg = types.ModuleType("models")
    pkg.__path__ = []
    sys.modules["models"] = pkg
    sys.modules["models.layers"] = stub
    mit = _load("_ref_mit", os.path.join(REF, "models", "backbones", "mit.py"))
    seg = _load("_ref_segformer", os.path.join(REF, "models", "heads", "segformer.py"))
    torch.manual_seed(0)
    backbone = mit.MiT("B0").eval()
    head = seg.SegFormerHead([32, 64, 160, 256], 256, 150).eval()
    x = torch.randn(1, 3, 160, 160, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        low = head(backbone(x))
    assert low.shape == (1, 150, 40, 40), low.shape
    # random-init logits are tiny (|z| ~ 0.1): scale so argmax margins and the OHEM threshold are exercised like trained output
    return (low * 20.0).contiguous()


if __name__ == "__main__":
    low = build_logits()
    mg.run_case("cfg1_real_mit_b0_ohem", 0, 1, 150, (40, 40), (160, 160), "early", loss="ohem", low_override=low)
