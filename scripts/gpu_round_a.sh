#!/bin/bash
# first GPU pass of round 2: full GPU test suite, default bench line, configs[3] launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 ) > gpurun_out/a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench exit $?"
tail -c 600 gpurun_out/a_bench.err
