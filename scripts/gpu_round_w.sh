#!/bin/bash
mkdir -p gpurun_out
for p in 1 0; do
  PROBE_PERIODIC=$p timeout 300 python scripts/small_kernels_probe.py > gpurun_out/w_small_$p.log 2>&1
  PROBE_PERIODIC=$p timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/w_small_launches_$p.csv python scripts/small_kernels_probe.py > gpurun_out/w_small_ncu_$p.log 2>&1
  python scripts/summarise_launches.py gpurun_out/w_small_launches_$p.csv 50 | tee gpurun_out/w_small_summary_$p.txt
done
PROBE_N=1000000 PROBE_CELLS=4096 PROBE_PERIODIC=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/w_1e6_launches.csv python scripts/small_kernels_probe.py > gpurun_out/w_1e6_ncu.log 2>&1
python scripts/summarise_launches.py gpurun_out/w_1e6_launches.csv 50 | tee gpurun_out/w_1e6_summary.txt
