#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/m2_smi.txt
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_boundary.py -m gpu -q -x ) > gpurun_out/m2_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/m2_pytest.log
tail -5 gpurun_out/m2_pytest.log
cat gpurun_out/multi_gpu_check.log | head -30
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/m2_bench.json 2> gpurun_out/m2_bench.err
echo "bench exit $?"; tail -c 1500 gpurun_out/m2_bench.json
