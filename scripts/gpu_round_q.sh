#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/q_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/q_pytest.log
tail -4 gpurun_out/q_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
echo "bench exit $?"
L=gpurun_out/q_probe.log
echo "== production-like 65536-cell run" > $L
PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 >> $L
echo "== periodic 65536-cell bench" >> $L
timeout 300 python bench.py --cells 65536 --steps 72 --no-cpu-baseline --no-host-arrays --no-configs3 2>&1 | tail -1 >> $L
echo "== walls 4096-cell bench" >> $L
timeout 300 python bench.py --walls --steps 40 --no-cpu-baseline --no-host-arrays --no-configs3 2>&1 | tail -1 >> $L
echo "== 24576-cell periodic bench (window covers the grid)" >> $L
timeout 300 python bench.py --cells 24576 --steps 40 --no-cpu-baseline --no-host-arrays --no-configs3 2>&1 | tail -1 >> $L
cut -c1-400 $L
PROBE_STEPS=28 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/q_c4prod_launches.csv python scripts/c4_production_probe.py > gpurun_out/q_c4prod_ncu.log 2>&1
echo "ncu list exit $?"
