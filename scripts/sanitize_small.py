"""Small shapes through every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import segmentation_factory_b200 as sf

def run(B, C, h, w, H, W, dtype=torch.float32, kind="ohem", late=False):
    g = torch.Generator().manual_seed(1)
    lab = torch.randint(0, C, (B, (H + 7) // 8, (W + 7) // 8), generator=g).repeat_interleave(8, 1).repeat_interleave(8, 2)[:, :H, :W].contiguous()
    lab[torch.rand((B, H, W), generator=g) < 0.1] = 255
    low = 3 * torch.randn((B, C, h, w), generator=g)
    if late:                                             # confident logits: the OHEM top-k branch (one-pass re-walk / radix select)
        ys = ((torch.arange(h) + 0.5) * H / h).long().clamp(max=H - 1); xs = ((torch.arange(w) + 0.5) * W / w).long().clamp(max=W - 1)
        l2 = lab[:, ys][:, :, xs]; l2 = torch.where(l2 == 255, torch.zeros_like(l2), l2)
        low = torch.randn((B, C, h, w), generator=g) + 12.0 * torch.nn.functional.one_hot(l2, C).permute(0, 3, 1, 2).float()
    x = low.to(dtype).cuda().requires_grad_(True)
    hist = torch.zeros(C, C, dtype=torch.int64, device="cuda")
    thr = 0.35667494 if kind == "ohem" else None
    loss = sf.fused_cross_entropy(x, lab.cuda(), None, 255, thr, hist)
    loss.backward()
    sf.eval_hist(x.detach(), lab.cuda(), 255, hist)
    met, conf = sf.Metrics(C, 255, "cuda"), sf.ConfusionMatrix(C)
    sf.fused_eval_update(x.detach(), lab.cuda(), met, conf)       # sf_eval_update: last-CTA accumulation where the fast kernel runs
    f = sf.FocalLoss(0.25, 2.0)(x, lab.cuda()); f.backward()
    d = sf.criterion(x, lab.cuda(), None, num_classes=C, dice=True, ignore_index=255); d.backward()   # dice walks + backward
    raw = torch.randint(0, 35, (B, H, W), generator=g, dtype=torch.uint8).cuda()
    lut = sf.LabelLUT(list(range(19)) + [255] * 16)
    lut(raw); lut(raw, torch.int64); lut(raw.flatten()[3:-5]); lut(raw.flatten()[1:-2], torch.int64)
    torch.cuda.synchronize()
    print("ok", B, C, h, w, H, W, dtype, kind, float(loss), int(hist.sum()))

run(1, 19, 32, 80, 128, 320)                         # one-pass walk (S=4, 16-row tiles) + fast forward kernels
run(1, 19, 32, 80, 128, 320, late=True)              # one-pass walk whose speculation fails: radix select + backward re-walk
run(1, 9, 20, 40, 160, 320, torch.float16)           # one-pass walk, C = 9, S = 8
run(2, 2, 12, 40, 48, 160, torch.bfloat16, "ce")     # one-pass walk, C = 2, plain CE (scale kernel)
run(1, 21, 24, 64, 192, 512, torch.bfloat16)         # fast kernels (S=8, bf16)
run(1, 40, 32, 80, 128, 320, torch.float32, "ce")    # chunk kernels
run(1, 7, 20, 33, 61, 100)                           # generic, non-integer, no TMA (odd pitch)
run(1, 5, 24, 32, 50, 70, torch.float16)             # generic with TMA
run(1, 19, 64, 96, 64, 96)                           # same-size streaming kernels
os.environ["SF_ONEPASS_CHUNK"] = "1"
run(1, 40, 24, 72, 96, 288, torch.float32, "ce")     # many-class one-pass walk (opt-in): tail chunk of 8, narrow tiles
run(1, 150, 12, 64, 48, 256, torch.bfloat16)         # ... OHEM, bf16, wide tiles
run(1, 40, 24, 72, 96, 288, late=True)               # ... speculation fails: two-pass re-walk in backward mode
os.environ["SF_ONEPASS_CHUNK"] = "0"
os.environ["SF_NO_FAST"] = "1"
run(1, 19, 32, 80, 128, 320)                         # generic kernels on a fast-path shape (top-k select may fire)
os.environ["SF_NO_FAST"] = "0"
os.environ["SF_NO_ONEPASS"] = "1"
run(1, 19, 32, 80, 128, 320)                         # forward / finalize / backward route on a one-pass shape
xh = torch.randn(2, 200, 25, 24, device="cuda", dtype=torch.bfloat16)
wh = torch.randn(19, 200, device="cuda", dtype=torch.bfloat16)
with torch.no_grad():
    o = sf.head_classifier(xh, wh, torch.randn(19, device="cuda"), torch.float32)          # tcgen05 classifier: ragged tile + K tail
    o2 = sf.head_classifier(torch.randn(1, 512, 16, 24, device="cuda", dtype=torch.float16), torch.randn(171, 512, device="cuda", dtype=torch.float16), None)
torch.cuda.synchronize()
print("ok head", float(o.abs().sum()), float(o2.float().abs().sum()))
