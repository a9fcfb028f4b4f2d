"""Device-time engine.criterion (CE + multiclass dice) through the Python mirror, next to a torch-eager
restatement of the reference's function (engine.py:10-15, util/losses.py:126-177: same ATen calls, incl. the B x C Python
loop with its host syncs) on the same GPU.  Prints one JSON object per line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
import segmentation_factory_b200 as sf
from bench_configs import synth, timeit


def eager_criterion(inputs, target, num_classes, ignore_index=255, eps=1e-6):
    loss = F.cross_entropy(inputs, target, ignore_index=ignore_index)
    dt = target.clone()
    mask = target == ignore_index
    dt[mask] = 0
    dt = F.one_hot(dt, num_classes).float()
    dt[mask] = ignore_index
    dt = dt.permute(0, 3, 1, 2)
    x = F.softmax(inputs, dim=1)
    dice = 0.
    for c in range(x.shape[1]):
        d = 0.
        for i in range(x.shape[0]):
            x_i = x[i, c].reshape(-1); t_i = dt[i, c].reshape(-1)
            roi = t_i != ignore_index
            x_i = x_i[roi]; t_i = t_i[roi]
            inter = torch.dot(x_i, t_i)
            sets = x_i.sum() + t_i.sum()
            if sets == 0:
                sets = 2 * inter
            d += (2 * inter + eps) / (sets + eps)
        dice += d / x.shape[0]
    return loss + 1 - dice / x.shape[1]


def run(name, B, C, h, w, H, W, dtype, seed, eager=True):
    data = [synth(B, C, h, w, H, W, dtype, seed + i) for i in range(3)]
    it = [0]
    def step():
        lo, la = data[it[0] % 3]; it[0] += 1
        x = lo.detach().requires_grad_(True)
        sf.criterion(x, la, None, num_classes=C, dice=True, ignore_index=255).backward()
    def step_ce():
        lo, la = data[it[0] % 3]; it[0] += 1
        x = lo.detach().requires_grad_(True)
        sf.criterion(x, la, None, num_classes=C, dice=False, ignore_index=255).backward()
    ms, ms_ce = timeit(step), timeit(step_ce)
    out = {"config": name, "criterion_dice_ms": ms, "criterion_ce_only_ms": ms_ce, "Gpx_s": B * H * W / ms / 1e6}
    if eager:
        lo, la = data[0]
        def step_e():
            x = lo.detach().float().requires_grad_(True)
            up = F.interpolate(x, size=(H, W), mode="bilinear", align_corners=False)
            eager_criterion(up, la, C).backward()
        out["torch_eager_ms"] = timeit(step_e, n=3, warm=1)
        out["speedup_vs_eager"] = out["torch_eager_ms"] / ms
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    run("cfg2 cityscapes 8x19x256x512->1024x2048 fp32 criterion(dice)", 8, 19, 256, 512, 1024, 2048, torch.float32, 1234)
    run("cfg3 16x150x128^2->512^2 fp32 criterion(dice)", 16, 150, 128, 128, 512, 512, torch.float32, 2345)
    run("cfg4 16x171 bf16 criterion(dice)", 16, 171, 128, 128, 512, 512, torch.bfloat16, 3456, eager=False)
    # what an UNMODIFIED engine.py passes: the model has already upsampled (build_models.py:65), logits at label size
    run("same-size 8x19x1024x2048 fp32 criterion(dice)", 8, 19, 1024, 2048, 1024, 2048, torch.float32, 1234, eager=False)
