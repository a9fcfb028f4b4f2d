"""Host-side cost per call of the public API (tiny shapes: the GPU is never the limiter), wall clock per call and a cProfile
breakdown.  Usage: host_overhead.py [n]"""
import cProfile, ctypes, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import segmentation_factory_b200 as sf
from segmentation_factory_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
dev = torch.device("cuda:0")
C = 19
lo = torch.randn(1, C, 64, 128, device=dev)
la = torch.randint(0, C, (1, 256, 512), device=dev)
m, cm = sf.Metrics(C, 255, dev), sf.ConfusionMatrix(C)
L = _lib.lib()
hist = torch.zeros(C, C, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
def wall(fn, n=n):
    for _ in range(50): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / n * 1e6
raw = lambda: L.sf_eval_hist(lo.data_ptr(), 0, la.data_ptr(), 0, 1, C, 64, 128, 256, 512, 255, hist.data_ptr(), st)
print("raw C-ABI sf_eval_hist          us/call", round(wall(raw), 2))
print("sf.eval_hist                    us/call", round(wall(lambda: sf.eval_hist(lo, la, 255, hist)), 2))
print("sf.fused_eval_update            us/call", round(wall(lambda: sf.fused_eval_update(lo, la, m, cm)), 2))
x = lo.clone().requires_grad_(True)
thr = float(-torch.log(torch.tensor(0.7)))
def train():
    x.grad = None
    sf.fused_cross_entropy(x, la, None, 255, thr, None).backward()
print("OhemCE fwd+bwd (one-pass route)  us/call", round(wall(train, n // 4), 2))
pr = cProfile.Profile(); pr.enable()
for _ in range(500): sf.fused_eval_update(lo, la, m, cm)
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(14); print(s.getvalue()[:2600])
