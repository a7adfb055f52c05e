#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/o_probe.log
echo "== production-like 65536-cell run: default (896-thread window CTAs) vs 1024" > $L
for v in cur w1024 cur; do
  if [ $v = cur ]; then unset ESPIC_LIB_PATH; else export ESPIC_LIB_PATH=$PWD/variants/libespic_$v.so; fi
  echo -n "$v: " >> $L; PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-90 >> $L
done
unset ESPIC_LIB_PATH
cat $L
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/o_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/o_pytest.log
tail -4 gpurun_out/o_pytest.log
