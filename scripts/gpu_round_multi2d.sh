#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -3
grep "MULTI_GPU" gpurun_out/multi_gpu_check.log
