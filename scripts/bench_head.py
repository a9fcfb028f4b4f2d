"""Time sf_head_classifier on Cityscapes-shaped features (8 x K x 256 x 512 bf16 -> 8 x 19 x 256 x 512) against cuDNN's conv."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
import segmentation_factory_b200 as sf
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = torch.device("cuda:0")
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (B, K, C, h, w, odt) in [(8, 256, 19, 256, 512, torch.bfloat16), (8, 256, 19, 256, 512, torch.float32), (8, 768, 19, 256, 512, torch.bfloat16), (16, 512, 150, 128, 128, torch.bfloat16)]:
    xs = [torch.randn((B, K, h, w), device=dev, dtype=torch.bfloat16) for _ in range(2)]     # 2 x 0.5-1.6 GB > L2
    wt = (torch.randn((C, K, 1, 1), device=dev) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(C, device=dev)
    it = [0]
    def ours():
        it[0] += 1
        return sf.head_classifier(xs[it[0] % 2], wt, bias, odt)
    def cudnn():
        it[0] += 1
        return F.conv2d(xs[it[0] % 2], wt, bias.to(torch.bfloat16))
    with torch.no_grad():
        a, b = t(ours), t(cudnn)
    nbytes = B * K * h * w * 2 + B * C * h * w * (4 if odt == torch.float32 else 2)
    print(json.dumps({"shape": [B, K, C, h, w], "out": str(odt), "ms": a, "GBps": nbytes / a / 1e6, "frac_of_hbm": nbytes / a / 1e6 / peak,
                      "torch_conv2d_bf16_ms": b}))
    del xs
