#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/s_probe.log
echo "== gather from global (55296-cell plane): production-like run vs sort interval" > $L
export ESPIC_LIB_PATH=$PWD/variants/libespic_gg.so
for k in auto 28 44 60 80; do
  if [ $k = auto ]; then unset PROBE_SORT_E; steps=120; else export PROBE_SORT_E=$k; steps=$((k*3)); fi
  echo -n "gg interval $k: " >> $L; PROBE_STEPS=$steps timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -2 | cut -c1-100 | tr '\n' ' ' >> $L; echo >> $L
done
unset ESPIC_LIB_PATH PROBE_SORT_E
cat $L
