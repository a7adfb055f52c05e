#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/x_variants.log
echo "== k_move walls, ms per 1e8-particle launch (cur = 896 threads, register loads)" > $L
for v in cur w512s w640s w512 cur w512s; do
  if [ $v = cur ]; then unset ESPIC_LIB_PATH; else export ESPIC_LIB_PATH=$PWD/variants/libespic_$v.so; fi
  echo -n "$v: " >> $L; PROBE_WALLS=1 timeout 300 python scripts/mover_micro_probe.py 2>&1 | tail -1 >> $L
done
echo "== window mover CTA width, production-like 65536-cell run (ms/step; cur = 896)" >> $L
for v in cur win768 win640 win512 cur; do
  if [ $v = cur ]; then unset ESPIC_LIB_PATH; else export ESPIC_LIB_PATH=$PWD/variants/libespic_$v.so; fi
  echo -n "$v: " >> $L; PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-100 >> $L
done
unset ESPIC_LIB_PATH
cat $L
