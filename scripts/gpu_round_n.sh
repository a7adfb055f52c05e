#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/n_probe.log
echo "== production-like 65536-cell run: k_move_win CTA width" > $L
for v in w768 w896 cur; do
  if [ $v = cur ]; then unset ESPIC_LIB_PATH; else export ESPIC_LIB_PATH=$PWD/variants/libespic_$v.so; fi
  echo -n "$v: " >> $L; PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-90 >> $L
done
unset ESPIC_LIB_PATH
echo "== headline bench with 896-thread k_move" >> $L
timeout 600 python bench.py --no-cpu-baseline --no-host-arrays --no-configs3 2>&1 | tail -1 | cut -c1-1200 >> $L
cat $L
