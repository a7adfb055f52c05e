#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/gc2_variants.log
echo "cur = __grid_constant__ on k_poisson / k_move_win / k_inject_commit, k_move by value; gcall = k_move too; nogc = none" > $L
for v in cur gcall nogc cur gcall; do
  if [ $v = cur ]; then unset ESPIC_LIB_PATH; else export ESPIC_LIB_PATH=$PWD/variants/libespic_$v.so; fi
  echo -n "$v: " >> $L
  timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-host-arrays --no-configs3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('ms/step %.4f mover %.4f rest %.2f us' % (d['ms_per_step'], r['launch_ms_mean'], 1e3*(d['ms_per_step']-2*r['launch_ms_mean'])))" >> $L
done
for v in cur gcall; do
  if [ $v = cur ]; then unset ESPIC_LIB_PATH; else export ESPIC_LIB_PATH=$PWD/variants/libespic_$v.so; fi
  echo -n "$v walls 4096: " >> $L
  timeout 300 python bench.py --walls --steps 40 --warmup 5 --no-cpu-baseline --no-host-arrays --no-configs3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('ms/step %.4f mover %.4f rest %.2f us' % (d['ms_per_step'], r['launch_ms_mean'], 1e3*(d['ms_per_step']-2*r['launch_ms_mean'])))" >> $L
done
unset ESPIC_LIB_PATH
cat $L
