#!/bin/bash
# final evidence: bench (headline + configs3 + host arrays + cpu baseline), other workloads, production launch list, small runs
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err
echo "bench exit $?"
( timeout 600 python bench.py --walls --steps 40 --no-cpu-baseline --no-host-arrays --no-configs3 ) > gpurun_out/y_bench_walls.json 2>> gpurun_out/y_bench.err
PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 > gpurun_out/y_c4prod_probe.json
PROBE_STEPS=28 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/y_c4prod_launches.csv python scripts/c4_production_probe.py > gpurun_out/y_c4prod_ncu.log 2>&1
python scripts/summarise_launches.py gpurun_out/y_c4prod_launches.csv 128 | grep -v "^at::\|at_cuda\|unnamed" | head -24 > gpurun_out/y_c4prod_summary.txt
timeout 600 python scripts/small_run_latency.py > gpurun_out/small_run_latency.jsonl 2>/dev/null
timeout 600 python scripts/longrun_configs4.py --n 1e5 --out gpurun_out/r2_configs4_history_n1e5.json > gpurun_out/fin_longrun_1e5.log 2>&1; tail -1 gpurun_out/fin_longrun_1e5.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/y_smoke.log 2>&1; tail -1 gpurun_out/y_smoke.log
