#!/bin/bash
mkdir -p gpurun_out
PROBE_WARM_RUN=1 PROBE_RUN=1 PROBE_STEPS=256 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 > gpurun_out/y_c4prod_probe_run.json
cut -c1-100 gpurun_out/y_c4prod_probe_run.json
