#!/bin/bash
mkdir -p gpurun_out
for kv in "ESPIC_STAGING=0" "ESPIC_STRIDED=0"; do
  echo "== $kv"
  env $kv timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_loop.py tests/test_gpu_run.py tests/test_gpu_boundary.py -m gpu -q -x 2>&1 | tail -45
done > gpurun_out/knobs.log 2>&1
