#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/scale_ngpu.txt
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  echo "N=$n exit $?"; cut -c1-160 gpurun_out/scale_n$n.json
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-host-arrays --no-configs3 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
echo "N=1 exit $?"; cut -c1-160 gpurun_out/scale_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 tests/multi_gpu_check.py > gpurun_out/multi_gpu_check_8gpu.log 2>&1
echo "multi check exit $?"; grep "ORACLE\|MULTI_GPU\|cells=" gpurun_out/multi_gpu_check_8gpu.log | cut -c1-330
