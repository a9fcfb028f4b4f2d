"""Time the one-pass step (configs[1]) kernel by kernel for one library / variant.  Prints one JSON line.
Usage: [SEGFACTORY_LIB=...] [SF_DEBUG_OP_VARIANT=n] python scripts/bench_walk.py [early|late] [steps]"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import synth_batch, B, C, LH, LW, H, W
from segmentation_factory_b200 import _lib
kind = sys.argv[1] if len(sys.argv) > 1 else "early"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
L = _lib.lib()
dev = torch.device("cuda:0")
sets = [tuple(t.to(dev) for t in synth_batch(torch, 1234 + i, kind)) for i in range(4)]
n_px = B * H * W
wsb = L.sf_workspace_bytes()
ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
loss_px = torch.empty(n_px, device=dev)
acc = torch.empty((B, C, LH, LW), device=dev)
grad = torch.empty((B, C, LH, LW), device=dev)
out = torch.empty((), device=dev)
hist = torch.zeros(C * C, dtype=torch.int64, device=dev)
stream = torch.cuda.current_stream()
st = stream.cuda_stream
thr = ctypes.c_float(float(-torch.log(torch.tensor(0.7))))
geo = (B, C, LH, LW, H, W, 255)
def step(i, ev=None):
    lo, la = sets[i % 4]
    if ev: ev[0].record(stream)
    rc = L.sf_ce_onepass(lo.data_ptr(), 0, la.data_ptr(), 0, *geo, None, thr, loss_px.data_ptr(), hist.data_ptr(), acc.data_ptr(), ws.data_ptr(), wsb, st)
    if ev: ev[1].record(stream)
    rc |= L.sf_loss_finalize(loss_px.data_ptr(), n_px, thr, ws.data_ptr(), wsb, out.data_ptr(), st)
    if ev: ev[2].record(stream)
    rc |= L.sf_ce_onepass_backward(lo.data_ptr(), 0, la.data_ptr(), 0, *geo, None, loss_px.data_ptr(), acc.data_ptr(), None, ws.data_ptr(), grad.data_ptr(), st)
    if ev: ev[3].record(stream)
    assert rc == 0, rc
for i in range(5): step(i)
torch.cuda.synchronize()
evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
for i in range(steps): step(i, evs[i])
torch.cuda.synchronize()
avg = lambda a, b: sum(e[a].elapsed_time(e[b]) for e in evs) / steps
print(json.dumps({"lib": os.path.basename(_lib.LIB_PATH), "variant": os.environ.get("SF_DEBUG_OP_VARIANT", "0"), "kind": kind,
                  "walk_ms": avg(0, 1), "finalize_ms": avg(1, 2), "backward_ms": avg(2, 3), "step_ms": avg(0, 3), "loss": out.item(),
                  "hist_sum": int(hist.sum().item()), "grad_abs_sum": grad.double().abs().sum().item()}))
