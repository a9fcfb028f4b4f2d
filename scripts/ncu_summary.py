"""Summarise an .ncu-rep: key raw metrics per kernel + top stall lines from the source page.
Usage: ncu_summary.py file.ncu-rep [n_source_lines]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
nsrc = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active", "lts__t_sector_hit_rate.pct"]
stall = [h for h in hdr if "smsp__average_warp" in h and "issue_stalled" in h and h.endswith("_not_issued.ratio") is False and "per_issue_active" in h]
for d in data:
    print("=" * 100)
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:75s} {d[i]} {units[i]}")
    st = sorted(((float(d[hdr.index(s)].replace(',', '') or 0), s) for s in stall), reverse=True)
    print("-- warp stall (cycles per issued instruction) --")
    for v, s in st[:9]:
        print(f"   {v:7.3f}  {s.split('issue_stalled_')[1].split('_per')[0]}")
if nsrc:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    # find header row
    for k, r in enumerate(rows):
        if "Source" in r and any("Sampl" in c for c in r):
            h = r; body = rows[k + 1:]; break
    else:
        print("no source page"); sys.exit(0)
    si = h.index("Source"); ci = [i for i, c in enumerate(h) if c.startswith("Warp Stall Sampling (All")]
    ai = h.index("Address") if "Address" in h else 0
    ei = [i for i, c in enumerate(h) if c == "Instructions Executed"]
    if ci:
        c = ci[0]
        tot = sum(float(r[c] or 0) for r in body if len(r) > c)
        top = sorted((r for r in body if len(r) > c), key=lambda r: -float(r[c] or 0))[:nsrc]
        print(f"-- top {nsrc} SASS lines by stall samples (total {tot:.0f}) --")
        for r in top:
            print(f"   {float(r[c] or 0) / tot * 100:5.2f}%  {r[ai][-6:]}  {r[si][:90]}")
