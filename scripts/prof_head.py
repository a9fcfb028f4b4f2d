"""A few sf_head_classifier launches (Cityscapes-shaped SegFormer features) for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import segmentation_factory_b200 as sf
x = [torch.randn((8, 256, 256, 512), device="cuda", dtype=torch.bfloat16) for _ in range(2)]
w = (torch.randn((19, 256), device="cuda") / 16).to(torch.bfloat16)
b = torch.randn(19, device="cuda")
with torch.no_grad():
    for i in range(4):
        o = sf.head_classifier(x[i % 2], w, b)
torch.cuda.synchronize()
print(float(o.float().abs().mean()))
