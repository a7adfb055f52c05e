#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_loop.py tests/test_gpu_sort.py -m gpu -q -x ) > gpurun_out/g_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/g_pytest.log
tail -4 gpurun_out/g_pytest.log
L=gpurun_out/g_probe.log
echo "== production-like 65536-cell run, per-lane stray patches: segments per CTA" > $L
for s in 1 2; do echo -n "segments $s: " >> $L; ESPIC_WINDOW_SEGMENTS=$s PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-90 >> $L; done
echo "== electron sort interval (segments default)" >> $L
for k in 20 36; do echo -n "interval $k: " >> $L; PROBE_SORT_E=$k PROBE_STEPS=$((k*4)) timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-90 >> $L; done
echo "== periodic 65536-cell bench" >> $L
timeout 300 python bench.py --cells 65536 --steps 72 --no-cpu-baseline --no-host-arrays --no-configs3 2>&1 | tail -1 | cut -c1-300 >> $L
cat $L
PROBE_STEPS=20 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_move_win --launch-skip 212 --launch-count 2 -o gpurun_out/g_move_win python scripts/c4_production_probe.py > gpurun_out/g_ncu.log 2>&1
echo "ncu exit $?"
PROBE_STEPS=28 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/g_c4prod_launches.csv python scripts/c4_production_probe.py > gpurun_out/g_c4prod_ncu.log 2>&1
echo "ncu list exit $?"
