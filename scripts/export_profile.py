"""Summarise an .ncu-rep (ncu --set full) into profiles/<name>.md + <name>.csv: the metrics the roofline uses.
usage: python scripts/export_profile.py gpurun_out/prof.ncu-rep profiles/r01_v4_full"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
keys = [k for k in KEYS if k in idx]
with open(out + ".csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(keys); w.writerow([units[idx[k]] for k in keys])
    for r in rows[2:]:
        w.writerow([r[idx[k]] for k in keys])
traffic = {}
with open(out + ".md", "w") as f:
    f.write(f"# ncu summary of `{rep.split('/')[-1]}` (ncu --set full --clock-control none)\n\n")
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        f.write(f"## {name[:110]}\n\n| metric | value | unit |\n|---|---|---|\n")
        for k in keys[1:]:
            f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")
        def num(k):
            return float(r[idx[k]].replace(",", ""))
        def to_bytes(k):
            u = units[idx[k]].lower(); v = num(k)
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        tb = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
        f.write(f"| dram read+write per launch | {tb/1e6:.1f} | MB |\n\n")
        op_mode = name.split("<")[1].split(">")[0].split(",")[6].strip() if "onepass_kernel<" in name else ""   # 7th template argument: OP_FUSED / OP_BWD
        short = ("walk" if op_mode == "0" else "backward" if op_mode == "1" else "head_gemm" if "head_gemm" in name else
                 "fwd" if "fwd" in name else "bwd" if "bwd" in name else "finalize" if "finalize" in name else name[:20])
        traffic[short] = tb
if "walk" in traffic:                                   # the bench step: stamp the kernel sources the capture ran on (bench.py checks it)
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import src_sha16
    traffic["src_sha16"] = src_sha16()
    traffic["source"] = "ncu --set full --clock-control none, one step of scripts/prof_onepass.py early"
json.dump(traffic, open(out + "_traffic.json", "w"), indent=1)
print("wrote", out + ".md/.csv/_traffic.json", traffic)
