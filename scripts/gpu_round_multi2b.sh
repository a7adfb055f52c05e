#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x ) > gpurun_out/m2_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/m2_pytest.log
tail -3 gpurun_out/m2_pytest.log
grep "ORACLE\|MULTI_GPU" gpurun_out/multi_gpu_check.log
