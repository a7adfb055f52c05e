#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_loop.py tests/test_gpu_boundary.py -m gpu -q -x ) > gpurun_out/r_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r_pytest.log
tail -3 gpurun_out/r_pytest.log
timeout 600 python scripts/host_chunk_probe.py > gpurun_out/r_hostchunk.log 2>&1; cat gpurun_out/r_hostchunk.log
PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-100
PROBE_STEPS=28 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r_c4prod_launches.csv python scripts/c4_production_probe.py > gpurun_out/r_c4prod_ncu.log 2>&1
echo "ncu list exit $?"
