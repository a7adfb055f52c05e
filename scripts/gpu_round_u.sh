#!/bin/bash
# small-run latency (stepwise / native / graph) and the k_move bulk-copy (TMA) A/B
mkdir -p gpurun_out
timeout 900 python scripts/small_run_latency.py > gpurun_out/small_run_latency.jsonl 2> gpurun_out/small_run_latency.err
cat gpurun_out/small_run_latency.jsonl; tail -3 gpurun_out/small_run_latency.err
echo "== k_move periodic: cp.async ring (cur) vs bulk copies (tma), ms per 1e8-particle launch" > gpurun_out/u_variants.log
for v in cur tma cur tma; do
  if [ $v = cur ]; then unset ESPIC_LIB_PATH; else export ESPIC_LIB_PATH=$PWD/variants/libespic_$v.so; fi
  echo -n "$v: " >> gpurun_out/u_variants.log
  timeout 300 python scripts/mover_micro_probe.py 2>&1 | tail -1 >> gpurun_out/u_variants.log
done
export ESPIC_LIB_PATH=$PWD/variants/libespic_tma.so
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "mover or step or periodic" 2>&1 | tail -3 >> gpurun_out/u_variants.log
unset ESPIC_LIB_PATH
cat gpurun_out/u_variants.log
