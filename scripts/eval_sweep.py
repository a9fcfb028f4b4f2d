"""BASELINE.json configs[4]: validation sweep -- 500 Cityscapes-shaped images (1x19x256x512 logits -> 1024x2048 labels),
eval-only argmax + confusion matrix through Metrics / ConfusionMatrix, one int64 all-reduce at the end
(engine.evaluate, engine.py:74-104).  Run with python (1 GPU) or torch.distributed.run (N GPUs).  Prints one JSON line.
Every rank also recomputes the global matrix of a few images with torch ops to check the reduced result."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import segmentation_factory_b200 as sf

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    sys.stdout.flush(); fd = os.dup(1); os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev); dist.barrier(); torch.cuda.synchronize()
    import ctypes; ctypes.CDLL(None).fflush(None); os.dup2(fd, 1); os.close(fd)
N_IMG, C, h, w, H, W = 500, 19, 256, 512, 1024, 2048
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
mine = list(range(rank, N_IMG, world))                       # this rank's images (DistributedSampler without padding)
# a pool of 16 distinct synthetic images per rank, cycled (image i uses pool[i % 16]); generated on the device
g = torch.Generator(device=dev).manual_seed(4567 + rank)
pool = []
for i in range(16):
    coarse = torch.randint(0, C, (1, H // 32, W // 32), generator=g, device=dev)
    coarse[torch.rand(coarse.shape, generator=g, device=dev) < 0.1] = 255
    lab = coarse.repeat_interleave(32, 1).repeat_interleave(32, 2).contiguous()
    l2 = torch.where(coarse == 255, torch.zeros_like(coarse), coarse).repeat_interleave(8, 1).repeat_interleave(8, 2)
    low = torch.randn((1, C, h, w), generator=g, device=dev) + 4.0 * torch.nn.functional.one_hot(l2, C).permute(0, 3, 1, 2).float()
    pool.append((low.contiguous(), lab))
batches = []
for k in range(0, len(mine), bs):
    ids = mine[k:k + bs]
    batches.append((torch.cat([pool[i % 16][0] for i in ids]), torch.cat([pool[i % 16][1] for i in ids])))

phase = {}
def sweep():
    t = time.perf_counter()
    metric = sf.Metrics(C, 255, dev); confmat = sf.ConfusionMatrix(C)
    for lo, la in batches:
        sf.fused_eval_update(lo, la, metric, confmat)
    phase["loop_launch_ms"] = (time.perf_counter() - t) * 1e3
    torch.cuda.synchronize(); phase["loop_done_ms"] = (time.perf_counter() - t) * 1e3
    confmat.reduce_from_all_processes(); metric.reduce_from_all_processes()
    torch.cuda.synchronize(); phase["reduced_ms"] = (time.perf_counter() - t) * 1e3
    return metric, confmat

_m, _c = sweep(); _c.compute()[2].mean().item(); _m.compute_iou(); torch.cuda.synchronize()   # warm-up incl. the compute ops
if world > 1: dist.barrier()
t0 = time.perf_counter()
metric, confmat = sweep()
miou = confmat.compute()[2].mean().item() * 100          # train_gpu.py:313 (forces the sync, like the reference)
dt = time.perf_counter() - t0
t = torch.tensor([dt], dtype=torch.float64, device=dev)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
# check: global pixel count and (rank-local) torch recomputation of this rank's share
tot = int(confmat.mat.sum())
ref = torch.zeros(C, C, dtype=torch.int64, device=dev)
for lo, la in batches[:4]:
    up = torch.nn.functional.interpolate(lo, size=(H, W), mode="bilinear", align_corners=False)
    pred = up.argmax(1).flatten(); tg = la.flatten(); keep = tg != 255
    ref += torch.bincount(tg[keep] * C + pred[keep], minlength=C * C).view(C, C)
m2 = sf.ConfusionMatrix(C)
for lo, la in batches[:4]:
    m2.update_from_logits(la, lo)
ok_local = bool((m2.mat - ref).abs().sum().item() <= 4)   # torch CUDA's bilinear may round a handful of near-ties differently
if rank == 0:
    print(json.dumps({"config": "cfg5 eval sweep", "images": N_IMG, "n_gpus": world, "batch": bs, "ms_total": t.item() * 1e3,
                      "Gpx_s": N_IMG * H * W / t.item() / 1e9, "mIoU": round(miou, 4), "hist_f32_miou": metric.compute_iou()[1],
                      "pixels_counted": tot, "phases": phase, "local_check_vs_torch_cuda": ok_local}))
if world > 1: dist.destroy_process_group()
