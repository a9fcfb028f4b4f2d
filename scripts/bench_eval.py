"""Eval-only kernel (cfg-5 shape: 8 x 19 x 256 x 512 -> 1024 x 2048) and the split-route forward, CUDA-event timed."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import segmentation_factory_b200 as sf
from bench import synth_batch
dev = torch.device("cuda:0")
kind = sys.argv[1] if len(sys.argv) > 1 else "early"
sets = [tuple(t.to(dev) for t in synth_batch(torch, 4567 + i, kind)) for i in range(4)]
hist = torch.zeros(19, 19, dtype=torch.int64, device=dev)
def t(fn, n=30):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
thr = float(-torch.log(torch.tensor(0.7)))
ev = t(lambda i: sf.eval_hist(sets[i % 4][0], sets[i % 4][1], 255, hist))
with torch.no_grad():
    fw = t(lambda i: sf.fused_cross_entropy(sets[i % 4][0], sets[i % 4][1], None, 255, thr, hist))
print(json.dumps({"kind": kind, "eval_hist_ms": ev, "forward_ohem_hist_nograd_ms": fw}))
