"""Join an `ncu --page source --csv` SASS dump with nvdisasm line info: executed warp-instructions per CUDA line.
usage: ncu_by_line.py <ncu_source.csv> <cubin> <mangled-kernel-substring> [topN]"""
import collections, csv, re, subprocess, sys
src_csv, cubin, sub = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# locate the function section
lines = []; cur = None; infunc = False
for ln in dis:
    if ln.startswith(".text."):
        infunc = sub in ln
        continue
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln): lines.append(cur)
print("sass instrs: ncu", len(data), "nvdisasm", len(lines))
agg = collections.Counter(); samp = collections.Counter()
for r, l in zip(data, lines):
    agg[l] += int(r[idx["Instructions Executed"]]); samp[l] += int(r[idx["# Samples"]])
tot = sum(agg.values()); ts = sum(samp.values())
srcs = {}
for (f, n), c in agg.most_common(top):
    if f not in srcs:
        try: srcs[f] = open("/root/repo/segmentation_factory_b200/csrc/" + f).read().splitlines()
        except OSError: srcs[f] = []
    text = srcs[f][n - 1].strip()[:100] if n - 1 < len(srcs[f]) else ""
    print(f"{c:11d} {100*c/tot:5.1f}% samp {100*samp[(f,n)]/ts:5.1f}% {f}:{n:<4d} {text}")
