"""Small shapes through every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import segmentation_factory_b200 as sf

def run(B, C, h, w, H, W, dtype=torch.float32, kind="ohem"):
    g = torch.Generator().manual_seed(1)
    lab = torch.randint(0, C, (B, H, W), generator=g)
    lab[torch.rand((B, H, W), generator=g) < 0.1] = 255
    x = (3 * torch.randn((B, C, h, w), generator=g)).to(dtype).cuda().requires_grad_(True)
    hist = torch.zeros(C, C, dtype=torch.int64, device="cuda")
    thr = 0.35667494 if kind == "ohem" else None
    loss = sf.fused_cross_entropy(x, lab.cuda(), None, 255, thr, hist)
    loss.backward()
    sf.eval_hist(x.detach(), lab.cuda(), 255, hist)
    f = sf.FocalLoss(0.25, 2.0)(x, lab.cuda()); f.backward()
    d = sf.criterion(x, lab.cuda(), None, num_classes=C, dice=True, ignore_index=255); d.backward()   # dice walks + backward
    raw = torch.randint(0, 35, (B, H, W), generator=g, dtype=torch.uint8).cuda()
    lut = sf.LabelLUT(list(range(19)) + [255] * 16)
    lut(raw); lut(raw, torch.int64); lut(raw.flatten()[3:-5]); lut(raw.flatten()[1:-2], torch.int64)
    torch.cuda.synchronize()
    print("ok", B, C, h, w, H, W, dtype, kind, float(loss), int(hist.sum()))

run(1, 19, 32, 80, 128, 320)                         # fast kernels (S=4)
run(1, 21, 24, 64, 192, 512, torch.bfloat16)         # fast kernels (S=8, bf16)
run(1, 40, 32, 80, 128, 320, torch.float32, "ce")    # chunk kernels
run(1, 7, 20, 33, 61, 100)                           # generic, non-integer, no TMA (odd pitch)
run(1, 5, 24, 32, 50, 70, torch.float16)             # generic with TMA
run(1, 19, 64, 96, 64, 96)                           # same-size streaming kernels
os.environ["SF_NO_FAST"] = "1"
run(1, 19, 32, 80, 128, 320)                         # generic kernels on a fast-path shape (top-k select may fire)
