#!/bin/bash
# GPU pass: full GPU test suite, default bench line, configs[3] launch list (timing only)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=10 ) > gpurun_out/b_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/b_pytest.log
tail -4 gpurun_out/b_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
echo "bench exit $?"
tail -c 400 gpurun_out/b_bench.err
PROBE_STEPS=64 timeout 600 python scripts/c4_production_probe.py > gpurun_out/b_c4prod.log 2>&1
echo "c4 probe exit $?"; tail -2 gpurun_out/b_c4prod.log
PROBE_STEPS=40 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b_c4prod_launches.csv python scripts/c4_production_probe.py > gpurun_out/b_c4prod_ncu.log 2>&1
echo "ncu exit $?"
