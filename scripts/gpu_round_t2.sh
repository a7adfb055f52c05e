#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_run.py -m gpu -q -x 2>&1 | tail -30 | tee gpurun_out/t2_tests.log
