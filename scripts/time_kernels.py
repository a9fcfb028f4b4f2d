"""Event-time fwd / finalize / bwd separately for one config through the C ABI (like bench.py does for cfg2)."""
import ctypes, json, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from segmentation_factory_b200 import _lib
L = _lib.lib()
B, C, h, w, H, W = [int(x) for x in sys.argv[1:7]]
dt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[sys.argv[7]]
kind = sys.argv[8]
code = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[dt]
dev = torch.device("cuda"); g = torch.Generator().manual_seed(0)
coarse = torch.randint(0, C, (B, H // 32, W // 32), generator=g); coarse[torch.rand(coarse.shape, generator=g) < 0.1] = 255
lab = coarse.repeat_interleave(32, 1).repeat_interleave(32, 2).contiguous().to(dev)
sets = [(3.0 * torch.randn((B, C, h, w), generator=g)).to(dt).to(dev) for _ in range(3)]
n = B * H * W
ws = torch.zeros(L.sf_workspace_bytes(), dtype=torch.uint8, device=dev)
lpx = torch.empty(n, device=dev); lse = torch.empty(n, device=dev); grad = torch.empty_like(sets[0]); out = torch.empty((), device=dev)
hist = torch.zeros(C * C, dtype=torch.int64, device=dev)
HP = None if os.environ.get('NO_HIST') else hist.data_ptr()
LP = None if os.environ.get('NO_LOSS') else lpx.data_ptr()
SP = None if os.environ.get('NO_LSE') else lse.data_ptr()
thr = 0.35667494 if kind == "ohem" else -1.0
st = torch.cuda.current_stream().cuda_stream
def step(i, ev=None):
    lo = sets[i % 3]
    if ev: ev[0].record()
    rc = L.sf_ce_forward(lo.data_ptr(), code, lab.data_ptr(), 0, B, C, h, w, H, W, 255, None, ctypes.c_float(thr), LP, SP, HP, ws.data_ptr(), ws.numel(), st)
    if ev: ev[1].record()
    rc |= L.sf_loss_finalize(lpx.data_ptr(), n, ctypes.c_float(thr), ws.data_ptr(), ws.numel(), out.data_ptr(), st)
    if ev: ev[2].record()
    if os.environ.get('NO_LSE') or os.environ.get('NO_LOSS'):
        if ev: ev[3].record()
        assert rc == 0, rc
        return
    rc |= L.sf_ce_backward(lo.data_ptr(), code, lab.data_ptr(), 0, B, C, h, w, H, W, 255, None, lpx.data_ptr(), lse.data_ptr(), None, ws.data_ptr(), grad.data_ptr(), st)
    if ev: ev[3].record()
    assert rc == 0, rc
for i in range(5): step(i)
torch.cuda.synchronize()
evs = [[torch.cuda.Event(True) for _ in range(4)] for _ in range(20)]
for i in range(20): step(i, evs[i])
torch.cuda.synchronize()
avg = lambda a, b: sum(e[a].elapsed_time(e[b]) for e in evs) / len(evs)
print(json.dumps({"shape": sys.argv[1:9], "fwd_ms": avg(0, 1), "finalize_ms": avg(1, 2), "bwd_ms": avg(2, 3), "Mpx": n / 1e6}))
