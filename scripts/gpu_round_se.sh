#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/se_variants.log
echo "== electron sort interval (auto = 28), 65536-cell production run through run(), 8 intervals timed (ms/step)" > $L
for se in 28 30 32 34 36 26; do
  echo -n "sort every $se: " >> $L; PROBE_SORT_E=$se PROBE_RUN=1 PROBE_STEPS=$((8*se)) timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-100 >> $L
done
cat $L
