#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/t16_variants.log
echo "== k_poisson cluster of 16 CTAs (non-portable) on the 65537-point grid, production-like run (ms/step)" > $L
for v in cur team16 cur team16; do
  if [ $v = cur ]; then unset ESPIC_LIB_PATH; else export ESPIC_LIB_PATH=$PWD/variants/libespic_$v.so; fi
  echo -n "$v: " >> $L; PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-100 >> $L
done
export ESPIC_LIB_PATH=$PWD/variants/libespic_team16.so
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "poisson or solve or field" 2>&1 | tail -2 >> $L
unset ESPIC_LIB_PATH
cat $L
