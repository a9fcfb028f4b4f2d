import sys, time, torch
sys.path.insert(0, '.')
import segmentation_factory_b200 as sf
dev = torch.device('cuda'); C,h,w,H,W = 19,256,512,1024,2048
g = torch.Generator(device=dev).manual_seed(1)
def mk(B):
    coarse = torch.randint(0, C, (B, H // 32, W // 32), generator=g, device=dev)
    coarse[torch.rand(coarse.shape, generator=g, device=dev) < 0.1] = 255
    lab = coarse.repeat_interleave(32, 1).repeat_interleave(32, 2).contiguous()
    l2 = torch.where(coarse == 255, torch.zeros_like(coarse), coarse).repeat_interleave(8, 1).repeat_interleave(8, 2)
    low = torch.randn((B, C, h, w), generator=g, device=dev) + 4.0 * torch.nn.functional.one_hot(l2, C).permute(0, 3, 1, 2).float()
    return low.contiguous(), lab
for B in (1, 8):
    lo, la = mk(B)
    hist = torch.zeros(C, C, dtype=torch.int64, device=dev)
    m = sf.Metrics(C, 255, dev); cm = sf.ConfusionMatrix(C)
    for name, fn in (("eval_hist", lambda: sf.eval_hist(lo, la, 255, hist)), ("fused_eval_update", lambda: sf.fused_eval_update(lo, la, m, cm))):
        for _ in range(5): fn()
        torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(50): fn()
        e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
        print(B, name, 'gpu us/call', e0.elapsed_time(e1) * 1000 / 50, 'wall us/call', (t1 - t0) * 1e6 / 50)
