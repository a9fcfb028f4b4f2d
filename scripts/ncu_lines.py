"""Aggregate the CUDA-source page of an .ncu-rep by source line: stall samples and executed instructions per line (top N)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
for k, r in enumerate(rows):
    if "Source" in r and any("Sampl" in c for c in r):
        h, body = r, rows[k + 1:]
        break
if h is None:
    print("no source page"); sys.exit(0)
ix = {c: i for i, c in enumerate(h)}
def f(r, c):
    try:
        return float(r[ix[c]] or 0)
    except Exception:
        return 0.0
tot = sum(f(r, "Warp Stall Sampling (All Samples)") for r in body) or 1.0
ninst = sum(f(r, "Instructions Executed") for r in body) or 1.0
print(f"samples {tot:.0f}  warp instructions {ninst / 1e6:.1f} M")
ln = ix.get("#", 0)
out = []
for r in body:
    if len(r) <= ix["Source"]:
        continue
    out.append((f(r, "Warp Stall Sampling (All Samples)"), f(r, "Instructions Executed"), r[ln], r[ix["Source"]].strip()[:110]))
print("%7s %7s %9s  %5s  %s" % ("smpl%", "inst%", "inst(M)", "line", "source"))
for s, n, l, t in sorted(out, key=lambda x: -x[0])[:top]:
    print("%7.2f %7.2f %9.2f  %5s  %s" % (100 * s / tot, 100 * n / ninst, n / 1e6, l, t))
