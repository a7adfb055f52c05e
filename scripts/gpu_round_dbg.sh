#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -3
for w in 0 1; do
  echo "== warm-up through run(): $w"; PROBE_WARM_RUN=$w PROBE_RUN=1 PROBE_STEPS=256 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-100
done
