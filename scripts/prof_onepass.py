"""A few one-pass steps (configs[1] shape) through the C ABI, for ncu.  Usage: prof_onepass.py [early|late] [steps]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import synth_batch, B, C, LH, LW, H, W
from segmentation_factory_b200 import _lib
kind = sys.argv[1] if len(sys.argv) > 1 else "early"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
L = _lib.lib()
dev = torch.device("cuda:0")
sets = [tuple(t.to(dev) for t in synth_batch(torch, 1234 + i, kind)) for i in range(2)]
n_px = B * H * W
wsb = L.sf_workspace_bytes()
ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
loss_px = torch.empty(n_px, device=dev)
acc = torch.empty((B, C, LH, LW), device=dev)
grad = torch.empty((B, C, LH, LW), device=dev)
out = torch.empty((), device=dev)
hist = torch.zeros(C * C, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
thr = ctypes.c_float(float(-torch.log(torch.tensor(0.7))))
geo = (B, C, LH, LW, H, W, 255)
for i in range(steps):
    lo, la = sets[i % 2]
    rc = L.sf_ce_onepass(lo.data_ptr(), 0, la.data_ptr(), 0, *geo, None, thr, loss_px.data_ptr(), hist.data_ptr(), acc.data_ptr(), ws.data_ptr(), wsb, st)
    rc |= L.sf_loss_finalize(loss_px.data_ptr(), n_px, thr, ws.data_ptr(), wsb, out.data_ptr(), st)
    rc |= L.sf_ce_onepass_backward(lo.data_ptr(), 0, la.data_ptr(), 0, *geo, None, loss_px.data_ptr(), acc.data_ptr(), None, ws.data_ptr(), grad.data_ptr(), st)
    assert rc == 0, rc
torch.cuda.synchronize()
print("loss", out.item())
