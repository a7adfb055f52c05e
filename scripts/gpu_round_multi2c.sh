#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -4
cp gpurun_out/multi_gpu_check.log gpurun_out/multi_gpu_check_2gpu.log 2>/dev/null
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/scale2_n2.json 2> gpurun_out/scale2_n2.err
tail -1 gpurun_out/scale2_n2.json | cut -c1-200
