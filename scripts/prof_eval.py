"""A few eval-only launches (cfg-5 shape) for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
import segmentation_factory_b200 as sf
from bench_configs import synth
lo, la = synth(8, 19, 256, 512, 1024, 2048, torch.float32, 4567)
hist = torch.zeros(19, 19, dtype=torch.int64, device="cuda")
for _ in range(4):
    sf.eval_hist(lo, la, 255, hist)
torch.cuda.synchronize()
