#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -3
( time timeout 900 python bench.py ) > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err
echo "bench exit $?"; tail -2 gpurun_out/y_bench.err
for i in 1 2; do PROBE_STEPS=64 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 > gpurun_out/y_c4prod_probe.json; cut -c1-100 gpurun_out/y_c4prod_probe.json; done
PROBE_RUN=1 PROBE_STEPS=256 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 > gpurun_out/y_c4prod_probe_run.json; cut -c1-100 gpurun_out/y_c4prod_probe_run.json
