#!/bin/bash
mkdir -p gpurun_out
echo "== k_move periodic variants (ms per 1e8-particle launch)" > gpurun_out/c_variants.log
for v in v0 v1 v2 v3 v4 cur; do
  if [ $v = cur ]; then unset ESPIC_LIB_PATH; else export ESPIC_LIB_PATH=$PWD/variants/libespic_$v.so; fi
  echo -n "$v: " >> gpurun_out/c_variants.log
  timeout 300 python scripts/mover_micro_probe.py 2>&1 | tail -1 >> gpurun_out/c_variants.log
done
echo "== window variants, production-like 65536-cell run (ms/step)" >> gpurun_out/c_variants.log
for v in w00 w01 w10 cur; do
  if [ $v = cur ]; then unset ESPIC_LIB_PATH; else export ESPIC_LIB_PATH=$PWD/variants/libespic_$v.so; fi
  echo -n "$v: " >> gpurun_out/c_variants.log
  PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-120 >> gpurun_out/c_variants.log
done
unset ESPIC_LIB_PATH
cat gpurun_out/c_variants.log
( time timeout 1500 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/c_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c_pytest.log
tail -4 gpurun_out/c_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
echo "bench exit $?"
timeout 600 python scripts/host_chunk_probe.py > gpurun_out/c_hostchunk.log 2>&1; cat gpurun_out/c_hostchunk.log
