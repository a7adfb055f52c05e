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
