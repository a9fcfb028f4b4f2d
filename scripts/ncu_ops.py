"""Aggregate the SASS source page of an .ncu-rep by opcode: share of stall samples, wait / scoreboard parts, executed count."""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
for k, r in enumerate(rows):
    if "Source" in r and any("Sampl" in c for c in r):
        h, body = r, rows[k + 1:]
        break
ix = {c: i for i, c in enumerate(h)}
def f(r, c):
    try:
        return float(r[ix[c]] or 0)
    except Exception:
        return 0.0
tot = sum(f(r, "Warp Stall Sampling (All Samples)") for r in body)
ninst = sum(f(r, "Instructions Executed") for r in body)
by = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
for r in body:
    s = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", s)
    op = m.group(2) if m else s[:10]
    v = by[op]
    v[0] += f(r, "Warp Stall Sampling (All Samples)"); v[1] += f(r, "stall_wait"); v[2] += f(r, "Instructions Executed")
    v[3] += f(r, "stall_short_sb") + f(r, "stall_long_sb"); v[4] += f(r, "stall_barrier") + f(r, "stall_branch_resolving")
print(f"samples {tot:.0f}  warp instructions {ninst / 1e6:.1f} M")
print("%-10s %8s %8s %8s %8s %9s %7s" % ("op", "samples%", "wait%", "sb%", "bar/br%", "inst(M)", "inst%"))
for op, v in sorted(by.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 28]:
    print("%-10s %8.2f %8.2f %8.2f %8.2f %9.2f %7.2f" % (op, 100 * v[0] / tot, 100 * v[1] / tot, 100 * v[3] / tot, 100 * v[4] / tot, v[2] / 1e6, 100 * v[2] / ninst))
