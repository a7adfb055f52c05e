#!/bin/bash
mkdir -p gpurun_out
PROBE_STEPS=4 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_sample_classify|k_inject_commit" --launch-skip 402 --launch-count 4 -o gpurun_out/r2_inject python scripts/c4_production_probe.py > gpurun_out/inj_ncu.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/inj_ncu.log | cut -c1-200
