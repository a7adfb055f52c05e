#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/pf_variants.log
echo "== window mover L2 prefetch distance, production-like 65536-cell run (ms/step); default 1 unit" > $L
for pf in 1 2 3 4 1; do
  echo -n "prefetch $pf: " >> $L; ESPIC_PREFETCH_UNITS=$pf PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-100 >> $L
done
cat $L
