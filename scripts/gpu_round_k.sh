#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_loop.py tests/test_gpu_boundary.py -m gpu -q -x ) > gpurun_out/k_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/k_pytest.log
tail -3 gpurun_out/k_pytest.log
L=gpurun_out/k_probe.log
echo "== production-like 65536-cell run, interleaved free slots: segments per CTA" > $L
for s in 1 2 4; do echo -n "segments $s: " >> $L; ESPIC_WINDOW_SEGMENTS=$s PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-90 >> $L; done
echo "== electron sort interval (segments default)" >> $L
for k in 20 36 44; do echo -n "interval $k: " >> $L; PROBE_SORT_E=$k PROBE_STEPS=$((k*4)) timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-90 >> $L; done
python scripts/segment_content_probe.py 2>&1 | tail -10 >> $L
cat $L
