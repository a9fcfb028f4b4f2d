#!/bin/bash
# Run on the B200 box (through gpurun): the evidence set profiles/README.md lists, into gpurun_out/$1 (default r02h).
O=gpurun_out/${1:-r02h}; mkdir -p $O
prof() { name=$1; sel=$2; skip=$3; cnt=$4; shift 4
  timeout 300 ncu --set full --clock-control none --import-source on -k "$sel" -s $skip -c $cnt -o /tmp/$name -f "$@" > $O/$name.log 2>&1
  python scripts/export_profile.py /tmp/$name.ncu-rep $O/${name}_full > /dev/null 2>&1
  python scripts/ncu_summary.py /tmp/$name.ncu-rep 0 > $O/${name}_summary.txt 2>&1
  python scripts/ncu_ops.py /tmp/$name.ncu-rep 30 > $O/${name}_ops.txt 2>&1
  rm -f /tmp/$name.ncu-rep; }
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/ncu_bench.log 2>&1
prof step "regex:onepass_kernel|finalize_kernel" 3 3 python scripts/prof_onepass.py early 3
cp $O/step_full_traffic.json profiles/traffic_latest.json      # so that the bench lines below quote the traffic of THIS source tree
prof step_late "regex:onepass_kernel|finalize_kernel" 3 3 python scripts/prof_onepass.py late 3
prof head "regex:head_gemm_kernel" 2 1 python scripts/prof_head.py
prof eval "regex:fwd_fast_kernel" 2 1 python scripts/prof_eval.py
timeout 500 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --kind late --no-extras > $O/bench_late.json 2> $O/bench_late.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 200 python scripts/bench_cfgs.py > $O/cfgs.jsonl 2>&1
timeout 200 python scripts/bench_head.py 2>/dev/null | grep '^{' > $O/head.jsonl
timeout 100 python scripts/bench_eval.py early 2>/dev/null | tail -1 > $O/eval.jsonl; timeout 100 python scripts/bench_eval.py late 2>/dev/null | tail -1 >> $O/eval.jsonl
cp profiles/traffic_latest.json $O/traffic_latest.json
python - <<PY
import json
for n in ("bench", "bench_late", "bench_reference"):
    try:
        d = json.loads(open("$O/%s.json" % n).read().strip().splitlines()[-1])
        print(n, d.get("ms_per_step"), d.get("value"), (d.get("roofline") or {}).get("traffic"), (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(n, "unreadable", e)
PY
