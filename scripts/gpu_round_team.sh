#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/team_variants.log
: > $L
for t in 8 4 2 1 8; do
  echo -n "ESPIC_SOLVE_TEAM=$t: " >> $L
  ESPIC_SOLVE_TEAM=$t timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-host-arrays --no-configs3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('ms/step %.4f mover %.4f rest %.4f us' % (d['ms_per_step'], r['launch_ms_mean'], 1e3*(d['ms_per_step']-2*r['launch_ms_mean'])))" >> $L
done
cat $L
