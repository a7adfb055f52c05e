#!/bin/bash
# round-2 evidence refresh with the bulk-copy mover: tests, bench (both arms), other workloads, ncu list + full capture
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 ) > gpurun_out/y_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/y_pytest.log; tail -5 gpurun_out/y_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err
echo "bench exit $?"
( timeout 600 python bench.py --impl reference --steps 8 --warmup 1 ) > gpurun_out/y_bench_reference.json 2> gpurun_out/y_bench_reference.err
echo "reference arm exit $?"
( timeout 600 python bench.py --cells 65536 --steps 72 --no-cpu-baseline --no-host-arrays --no-configs3 ) > gpurun_out/y_bench_65536.json 2>> gpurun_out/y_bench.err
( timeout 600 python bench.py --walls --steps 40 --no-cpu-baseline --no-host-arrays --no-configs3 ) > gpurun_out/y_bench_walls.json 2>> gpurun_out/y_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-host-arrays --no-configs3 > gpurun_out/y_ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_move --launch-skip 8 --launch-count 1 -o gpurun_out/r2_mover python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-host-arrays --no-configs3 > gpurun_out/y_ncu_mover.log 2>&1
echo "ncu mover exit $?"
PROBE_STEPS=28 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/y_c4prod_launches.csv python scripts/c4_production_probe.py > gpurun_out/y_c4prod_ncu.log 2>&1
python scripts/summarise_launches.py gpurun_out/y_c4prod_launches.csv 128 | grep -v "^at::\|at_cuda\|unnamed" | head -24 > gpurun_out/y_c4prod_summary.txt
PROBE_STEPS=20 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_move_win --launch-skip 212 --launch-count 2 -o gpurun_out/r2_move_win python scripts/c4_production_probe.py > gpurun_out/y_ncu_win.log 2>&1
echo "ncu win exit $?"
PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 > gpurun_out/y_c4prod_probe.json
timeout 600 python scripts/small_run_latency.py > gpurun_out/small_run_latency.jsonl 2>/dev/null
( timeout 600 python bench.py --cells 24576 --steps 40 --no-cpu-baseline --no-host-arrays --no-configs3 ) > gpurun_out/y_bench_24576.json 2>> gpurun_out/y_bench.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/y_smoke.log 2>&1; tail -1 gpurun_out/y_smoke.log
