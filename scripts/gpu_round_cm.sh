#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/cm_variants.log
( timeout 1500 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -4 > $L
for v in cur oldcommit cur oldcommit; do
  if [ $v = cur ]; then unset ESPIC_LIB_PATH; else export ESPIC_LIB_PATH=$PWD/variants/libespic_$v.so; fi
  echo -n "$v production 65536: " >> $L; PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-100 >> $L
done
unset ESPIC_LIB_PATH
for m in 2 4; do
  echo -n "cur, commit CTAs per SM $m: " >> $L; ESPIC_COMMIT_CTAS=$m PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-100 >> $L
done
cat $L
