#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run_variants.log
echo "== 65536-cell production run, 224 timed steps (8 electron sort intervals): Simulation.step() vs Simulation.run() (graph replay)" > $L
for m in 0 1 0 1; do
  echo -n "PROBE_RUN=$m: " >> $L; PROBE_RUN=$m PROBE_STEPS=224 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-100 >> $L
done
cat $L
