#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/ctl_variants.log
: > $L
for m in 0 1 0 1; do
  echo -n "production 65536, PROBE_RUN=$m: " >> $L; PROBE_RUN=$m PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-100 >> $L
done
cat $L
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/ctl_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/ctl_pytest.log; tail -6 gpurun_out/ctl_pytest.log
timeout 600 python scripts/longrun_configs4.py --n 1e5 --out gpurun_out/r2_configs4_history_n1e5.json > gpurun_out/ctl_longrun_1e5.log 2>&1; grep -o '"seconds": [0-9.]*' gpurun_out/r2_configs4_history_n1e5.json | head -1
timeout 900 python scripts/longrun_configs4.py --n 1e6 --out gpurun_out/r2_configs4_history_n1e6.json > gpurun_out/ctl_longrun_1e6.log 2>&1; grep -o '"seconds": [0-9.]*' gpurun_out/r2_configs4_history_n1e6.json | head -1; tail -1 gpurun_out/ctl_longrun_1e6.log | cut -c1-300
