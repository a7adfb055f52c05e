#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/wt_variants.log
echo "== window mover: register loads, 27648-cell window, 896 threads (cur) vs bulk-copy ring, 20480-cell window, 512 threads (wintma)" > $L
export ESPIC_LIB_PATH=$PWD/variants/libespic_wintma.so
timeout 600 python -m pytest tests -m gpu -q -x -k "large_grid or window or sort or multi_segment" 2>&1 | tail -3 >> $L
for v in wintma cur wintma cur; do
  if [ $v = cur ]; then unset ESPIC_LIB_PATH; else export ESPIC_LIB_PATH=$PWD/variants/libespic_$v.so; fi
  echo -n "$v: " >> $L; PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-100 >> $L
done
export ESPIC_LIB_PATH=$PWD/variants/libespic_wintma.so
echo -n "wintma, periodic 65536: " >> $L
timeout 300 python bench.py --cells 65536 --steps 72 --no-cpu-baseline --no-host-arrays --no-configs3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('ms/step %.4f mover %.4f frac %.3f' % (d['ms_per_step'], r['launch_ms_mean'], r['frac']))" >> $L
unset ESPIC_LIB_PATH
cat $L
