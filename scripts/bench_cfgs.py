"""cfg-3 / cfg-4 through the Python API with per-kernel CUDA-event times (forward+finalize vs backward)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import segmentation_factory_b200 as sf
from bench import synth_batch
dev = torch.device("cuda:0")
thr = float(-torch.log(torch.tensor(0.7)))
for key, c, dt, ohem in (("cfg3_ade_ce_f32", 150, torch.float32, False), ("cfg4_coco_ce_bf16", 171, torch.bfloat16, False), ("cfg4_coco_ohem_bf16", 171, torch.bfloat16, True)):
    sets = []
    for i in range(2):
        lo, la = synth_batch(torch, 2345 + i, "early", b=16, c=c, lh=128, lw=128, h=512, w=512, dtype=dt)
        sets.append((lo.to(dev).requires_grad_(True), la.to(dev)))
    hist = torch.zeros(c, c, dtype=torch.int64, device=dev)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(13)]
    for i in range(13):
        x, y = sets[i % 2]
        x.grad = None
        ev[i][0].record()
        loss = sf.fused_cross_entropy(x, y, None, 255, thr if ohem else None, hist)
        ev[i][1].record()
        loss.backward()
        ev[i][2].record()
    torch.cuda.synchronize()
    f = sum(e[0].elapsed_time(e[1]) for e in ev[3:]) / 10
    b = sum(e[1].elapsed_time(e[2]) for e in ev[3:]) / 10
    print(json.dumps({"cfg": key, "fwd_ms": f, "bwd_ms": b, "total_ms": f + b}))
