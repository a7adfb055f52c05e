#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/ep_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/ep_pytest.log; tail -25 gpurun_out/ep_pytest.log
ESPIC_FUSED_EPILOGUE=0 timeout 900 python -m pytest tests/test_gpu_loop.py -m gpu -q -x -k "fused_step_equals" 2>&1 | tail -3
