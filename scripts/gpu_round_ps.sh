#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_poisson --launch-skip 5 --launch-count 12 -o gpurun_out/r2_poisson2 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-host-arrays --no-configs3 > gpurun_out/ps_ncu.log 2>&1
echo "ncu exit $?"
