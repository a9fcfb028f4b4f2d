"""Device-time the BASELINE.json configs other than the headline (parity cases, not bench lines):
cfg1 (2x150x128^2->512^2 OHEM fp32), cfg3 (16x150, plain CE), cfg4 (16x171 bf16 CE/OHEM), cfg5 (eval-only sweep slice),
plus torch-eager on the same GPU for cfg2/cfg5.  Prints one JSON object per config."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
import segmentation_factory_b200 as sf

dev = torch.device("cuda")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
THR = float(-torch.log(torch.tensor(0.7, dtype=torch.float)))


def synth(B, C, h, w, H, W, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    coarse = torch.randint(0, C, (B, H // 32, W // 32), generator=g)
    coarse[torch.rand(coarse.shape, generator=g) < 0.1] = 255
    lab = coarse.repeat_interleave(32, 1).repeat_interleave(32, 2).contiguous()
    low = (3.0 * torch.randn((B, C, h, w), generator=g)).to(dtype)
    return low.to(dev), lab.to(dev)


def timeit(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run(name, B, C, h, w, H, W, dtype, kind, seed, sets=3):
    data = [synth(B, C, h, w, H, W, dtype, seed + i) for i in range(sets)]
    esz = 4 if dtype == torch.float32 else 2
    npx = B * H * W
    it = [0]
    hist = torch.zeros(C, C, dtype=torch.int64, device=dev)
    if kind == "eval":
        def step():
            lo, la = data[it[0] % sets]; it[0] += 1
            sf.eval_hist(lo, la, 255, hist)
        bpp = C * esz * (h * w) / (H * W) + 8
    else:
        def step():
            lo, la = data[it[0] % sets]; it[0] += 1
            x = lo.detach().requires_grad_(True)
            loss = sf.fused_cross_entropy(x, la, None, 255, THR if kind == "ohem" else None, hist)
            loss.backward()
        bpp = 3 * C * esz * (h * w) / (H * W) + 16 + (12 if kind == "ohem" else 0)
    ms = timeit(step)
    out = {"config": name, "ms_per_step": ms, "Gpx_s": npx / ms / 1e6, "alg_bytes_per_px": bpp,
           "frac_of_measured_hbm": bpp * npx / (ms * 1e-3) / 1e9 / PEAK}
    print(json.dumps(out), flush=True)
    return out


def torch_eager(name, B, C, h, w, H, W, kind):
    lo, la = synth(B, C, h, w, H, W, torch.float32, 99)
    def step():
        x = lo.detach().requires_grad_(True)
        up = F.interpolate(x, size=(H, W), mode="bilinear", align_corners=False)
        if kind == "eval":
            with torch.no_grad():
                pred = up.argmax(1).flatten(); t = la.flatten(); keep = t != 255
                torch.bincount(t[keep] * C + pred[keep], minlength=C * C)
            return
        n_min = la[la != 255].numel() // 16
        l = F.cross_entropy(up, la, ignore_index=255, reduction="none").view(-1)
        hard = l[l > THR]
        if hard.numel() < n_min:
            hard, _ = l.topk(n_min)
        hard.mean().backward()
        with torch.no_grad():
            pred = up.argmax(1).flatten(); t = la.flatten(); keep = t != 255
            torch.bincount(t[keep] * C + pred[keep], minlength=C * C)
    ms = timeit(step, n=10, warm=3)
    print(json.dumps({"config": name + " [torch eager, same GPU]", "ms_per_step": ms, "Gpx_s": B * H * W / ms / 1e6}), flush=True)


def label_lut(name, B, H, W, dtype):
    """sf_label_lut: HBM-bound streaming map, 2 B/px (uint8 out) or 9 B/px (int64 out)."""
    lut = sf.LabelLUT(list(range(19)) + [255] * 16)
    raws = [torch.randint(0, 35, (B, H, W), dtype=torch.uint8, device=dev) for _ in range(8)]   # 8 x 16.8 MB in, rotating
    it = [0]
    def step():
        lut(raws[it[0] % 8], dtype); it[0] += 1
    ms = timeit(step, n=200, warm=20)
    bpp = 2 if dtype == torch.uint8 else 9
    print(json.dumps({"config": name, "ms_per_step": ms, "Gpx_s": B * H * W / ms / 1e6, "alg_bytes_per_px": bpp,
                      "GB_s": bpp * B * H * W / (ms * 1e-3) / 1e9,
                      "frac_of_measured_hbm": bpp * B * H * W / (ms * 1e-3) / 1e9 / PEAK}), flush=True)


if __name__ == "__main__":
    if sys.argv[1:] == ["lut"]:
        label_lut("label LUT 8x1024x2048 uint8 -> uint8 (incl. python + output alloc)", 8, 1024, 2048, torch.uint8)
        label_lut("label LUT 8x1024x2048 uint8 -> int64", 8, 1024, 2048, torch.int64)
        label_lut("label LUT 64x1024x2048 uint8 -> uint8", 64, 1024, 2048, torch.uint8)
        label_lut("label LUT 64x1024x2048 uint8 -> int64", 64, 1024, 2048, torch.int64)
        sys.exit(0)
    run("cfg2 cityscapes 8x19 OHEM fp32 (python API incl. allocs)", 8, 19, 256, 512, 1024, 2048, torch.float32, "ohem", 1234)
    run("cfg1 2x150x128^2->512^2 OHEM fp32", 2, 150, 128, 128, 512, 512, torch.float32, "ohem", 0)
    run("cfg3 16x150x128^2->512^2 CE fp32", 16, 150, 128, 128, 512, 512, torch.float32, "ce", 2345)
    run("cfg4 16x171 bf16 CE", 16, 171, 128, 128, 512, 512, torch.bfloat16, "ce", 3456)
    run("cfg4 16x171 bf16 OHEM", 16, 171, 128, 128, 512, 512, torch.bfloat16, "ohem", 3456)
    run("cfg5 eval-only 8x19x256x512->1024x2048 fp32", 8, 19, 256, 512, 1024, 2048, torch.float32, "eval", 4567)
    run("cfg2' bf16 logits OHEM", 8, 19, 256, 512, 1024, 2048, torch.bfloat16, "ohem", 1234)
    torch_eager("cfg2 cityscapes 8x19 OHEM fp32", 8, 19, 256, 512, 1024, 2048, "ohem")
    torch_eager("cfg5 eval-only 8x19", 8, 19, 256, 512, 1024, 2048, "eval")
