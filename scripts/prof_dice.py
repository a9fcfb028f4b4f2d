"""One criterion(dice) fwd+bwd at cfg-2 and cfg-3 shapes plus the label LUT, for ncu captures of the dice / LUT kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
import segmentation_factory_b200 as sf
from bench_configs import synth

for (B, C, h, w, H, W) in [(8, 19, 256, 512, 1024, 2048), (16, 150, 128, 128, 512, 512)]:
    lo, la = synth(B, C, h, w, H, W, torch.float32, 1)
    for _ in range(2):
        x = lo.detach().requires_grad_(True)
        sf.criterion(x, la, None, num_classes=C, dice=True, ignore_index=255).backward()
lut = sf.LabelLUT(list(range(19)) + [255] * 16)
raw = torch.randint(0, 35, (64, 1024, 2048), dtype=torch.uint8, device="cuda")
for _ in range(2):
    lut(raw); lut(raw, torch.int64)
torch.cuda.synchronize()
