#!/bin/bash
# native multi-step run: parity tests, latency probe, then the regular suites
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_run.py -x -q 2>&1 | tail -15 > gpurun_out/native_tests.log
cat gpurun_out/native_tests.log
timeout 600 python scripts/small_run_latency.py > gpurun_out/small_run_latency.jsonl 2> gpurun_out/small_run_latency.err
cat gpurun_out/small_run_latency.jsonl; tail -5 gpurun_out/small_run_latency.err
timeout 900 python -m pytest tests/test_gpu_longrun.py tests/test_gpu_boundary.py -x -q 2>&1 | tail -8 | tee gpurun_out/native_longrun.log
