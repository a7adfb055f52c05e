#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/d_probe.log
echo "== k_move periodic, current build" > $L
timeout 300 python scripts/mover_micro_probe.py 2>&1 | tail -1 >> $L
echo "== production-like 65536-cell run: segments per CTA" >> $L
for s in 2 3 4 6; do echo -n "segments $s: " >> $L; ESPIC_WINDOW_SEGMENTS=$s PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-90 >> $L; done
echo "== production-like 65536-cell run: electron sort interval" >> $L
for k in 20 28 36 44 56; do echo -n "interval $k: " >> $L; PROBE_SORT_E=$k PROBE_STEPS=$((k*3)) timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-90 >> $L; done
cat $L
( time timeout 1500 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/d_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/d_pytest.log
tail -4 gpurun_out/d_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
echo "bench exit $?"
PROBE_STEPS=20 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_move_win --launch-skip 212 --launch-count 2 -o gpurun_out/d_move_win python scripts/c4_production_probe.py > gpurun_out/d_ncu.log 2>&1
echo "ncu exit $?"
