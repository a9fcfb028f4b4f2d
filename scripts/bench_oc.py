"""Many-class one-pass walk (sf_onepass_chunk.cuh) at cfg-3 / cfg-4: forward-call time per forced tile width / height."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SF_ONEPASS_CHUNK"] = "1"
import torch
import segmentation_factory_b200 as sf
from bench import synth_batch
dev = torch.device("cuda:0")
for key, c, dt in (("cfg3_ade_ce_f32", 150, torch.float32), ("cfg4_coco_ce_bf16", 171, torch.bfloat16)):
    sets = []
    for i in range(2):
        lo, la = synth_batch(torch, 2345 + i, "early", b=16, c=c, lh=128, lw=128, h=512, w=512, dtype=dt)
        sets.append((lo.to(dev).requires_grad_(True), la.to(dev)))
    hist = torch.zeros(c, c, dtype=torch.int64, device=dev)
    for nt, th in ((0, 0), (128, 0), (192, 0), (128, 13), (128, 9), (128, 7), (192, 11), (192, 8), (192, 16)):
        os.environ["SF_DEBUG_OC_NT"] = str(nt); os.environ["SF_DEBUG_OC_TH"] = str(th)
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(9)]
        for i in range(9):
            x, y = sets[i % 2]
            x.grad = None
            ev[i][0].record()
            loss = sf.fused_cross_entropy(x, y, None, 255, None, hist)
            ev[i][1].record()
            loss.backward()
            ev[i][2].record()
        torch.cuda.synchronize()
        f = sum(e[0].elapsed_time(e[1]) for e in ev[3:]) / 6
        b = sum(e[1].elapsed_time(e[2]) for e in ev[3:]) / 6
        print(json.dumps({"cfg": key, "nt": nt, "th": th, "fwd_ms": round(f, 4), "bwd_ms": round(b, 4)}), flush=True)
