#!/bin/bash
# evidence for profiles/: plain bench first, then the ncu captures of the same commands
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err
echo "bench exit $?"
( timeout 600 python bench.py --impl reference --steps 8 --warmup 1 ) > gpurun_out/t_bench_reference.json 2> gpurun_out/t_bench_reference.err
echo "reference arm exit $?"
timeout 600 python scripts/longrun_configs4.py --n 1e6 --out gpurun_out/r2_configs4_history_n1e6.json > gpurun_out/t_longrun_1e6.log 2>&1; tail -2 gpurun_out/t_longrun_1e6.log | cut -c1-600
timeout 600 python scripts/longrun_configs4.py --n 1e5 --out gpurun_out/r2_configs4_history_n1e5.json > gpurun_out/t_longrun_1e5.log 2>&1; tail -1 gpurun_out/t_longrun_1e5.log | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-host-arrays --no-configs3 > gpurun_out/t_ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_move --launch-skip 8 --launch-count 1 -o gpurun_out/r2_mover python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-host-arrays --no-configs3 > gpurun_out/t_ncu_mover.log 2>&1
echo "ncu mover exit $?"
PROBE_STEPS=20 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_move_win --launch-skip 212 --launch-count 2 -o gpurun_out/r2_move_win python scripts/c4_production_probe.py > gpurun_out/t_ncu_win.log 2>&1
echo "ncu win exit $?"
