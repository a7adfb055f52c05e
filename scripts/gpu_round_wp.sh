#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/wp_variants.log
echo "== window mover, register loads one unit ahead (wpNNN = CTA width), production-like 65536-cell run (ms/step); cur = loads at use, 896" > $L
for v in cur wp896 wp768 wp640 cur; do
  if [ $v = cur ]; then unset ESPIC_LIB_PATH; else export ESPIC_LIB_PATH=$PWD/variants/libespic_$v.so; fi
  for pf in 1 2; do
    echo -n "$v prefetch $pf: " >> $L; ESPIC_PREFETCH_UNITS=$pf PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-100 >> $L
  done
done
unset ESPIC_LIB_PATH
cat $L
