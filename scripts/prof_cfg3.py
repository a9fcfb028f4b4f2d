"""cfg-3 (16 x 150 x 128 x 128 -> 512 x 512, plain CE + matrix) forward + backward through the Python API, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import segmentation_factory_b200 as sf
from bench import synth_batch
lo, la = synth_batch(torch, 2345, "early", b=16, c=150, lh=128, lw=128, h=512, w=512)
x = lo.cuda().requires_grad_(True); y = la.cuda()
hist = torch.zeros(150, 150, dtype=torch.int64, device="cuda")
for _ in range(3):
    x.grad = None
    sf.fused_cross_entropy(x, y, None, 255, None, hist).backward()
torch.cuda.synchronize()
