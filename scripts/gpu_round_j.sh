#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/j_probe.log
echo "== production-like 65536-cell run: per-segment variants x segments per CTA (ms/step)" > $L
for v in wprev s0t0 s0 t0 cur; do
  if [ $v = cur ]; then unset ESPIC_LIB_PATH; else export ESPIC_LIB_PATH=$PWD/variants/libespic_$v.so; fi
  for s in 1 2 4; do echo -n "$v segments $s: " >> $L; ESPIC_WINDOW_SEGMENTS=$s PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-45 >> $L; done
done
unset ESPIC_LIB_PATH
cat $L
