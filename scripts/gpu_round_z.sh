#!/bin/bash
# injection kernels: warp-queue sampler vs one particle per thread, commit CTAs per SM
mkdir -p gpurun_out
L=gpurun_out/z_variants.log
timeout 600 python -m pytest tests -m gpu -q -x -k "inject or sampl or nonperiodic or walls or longrun or native" 2>&1 | tail -3 > $L
echo "== production-like 65536-cell run (ms/step): oldsample = one particle per thread" >> $L
for v in cur oldsample cur; do
  if [ $v = cur ]; then unset ESPIC_LIB_PATH; else export ESPIC_LIB_PATH=$PWD/variants/libespic_$v.so; fi
  echo -n "$v: " >> $L; PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-100 >> $L
done
unset ESPIC_LIB_PATH
for m in 2 4; do
  echo -n "cur, commit CTAs per SM $m: " >> $L; ESPIC_COMMIT_CTAS=$m PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-100 >> $L
done
cat $L
PROBE_STEPS=28 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/z_c4prod_launches.csv python scripts/c4_production_probe.py > gpurun_out/z_c4prod_ncu.log 2>&1
python scripts/summarise_launches.py gpurun_out/z_c4prod_launches.csv 128 | grep -v "^at::\|at_cuda" | head -24 | tee gpurun_out/z_c4prod_summary.txt
