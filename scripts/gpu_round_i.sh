#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_loop.py -m gpu -q -x ) > gpurun_out/i_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/i_pytest.log
tail -3 gpurun_out/i_pytest.log
L=gpurun_out/i_probe.log
echo "== per-segment phases (debug build), segments 4" > $L
ESPIC_LIB_PATH=$PWD/variants/libespic_dbg.so ESPIC_WINDOW_SEGMENTS=4 PROBE_STEPS=28 python scripts/cta_cycles_probe.py 2>&1 | tail -12 >> $L
echo "== production-like 65536-cell run: segments per CTA" >> $L
for s in 1 2 4 8; do echo -n "segments $s: " >> $L; ESPIC_WINDOW_SEGMENTS=$s PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-90 >> $L; done
echo "== periodic 65536-cell bench, segments 1 2 4" >> $L
for s in 1 2 4; do ESPIC_WINDOW_SEGMENTS=$s timeout 300 python bench.py --cells 65536 --steps 72 --no-cpu-baseline --no-host-arrays --no-configs3 2>&1 | tail -1 | cut -c1-200 >> $L; done
cat $L
