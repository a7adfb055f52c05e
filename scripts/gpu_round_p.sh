#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_sort_scatter --launch-skip 2 --launch-count 1 -o gpurun_out/p_sort python scripts/sort_only.py > gpurun_out/p_sort.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/p_sort.log
