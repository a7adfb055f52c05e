#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/m_probe.log
echo "== per-segment cycles, segments 1 and 2" > $L
ESPIC_WINDOW_SEGMENTS=1 PROBE_STEPS=28 python scripts/cta_cycles_probe.py 2>&1 | tail -4 >> $L
ESPIC_WINDOW_SEGMENTS=2 PROBE_STEPS=28 python scripts/cta_cycles_probe.py 2>&1 | tail -4 >> $L
echo "== k_move periodic / walls CTA width variants (ms per 1e8-particle launch)" >> $L
for v in m512 m640 m768 m896 m896c cur; do
  if [ $v = cur ]; then unset ESPIC_LIB_PATH ESPIC_MOVE_THREADS; else export ESPIC_LIB_PATH=$PWD/variants/libespic_$v.so; t=${v#m}; export ESPIC_MOVE_THREADS=${t%c}; fi
  echo -n "$v periodic: " >> $L; timeout 300 python scripts/mover_micro_probe.py 2>&1 | tail -1 >> $L
  echo -n "$v walls: " >> $L; PROBE_WALLS=1 timeout 300 python scripts/mover_micro_probe.py 2>&1 | tail -1 >> $L
done
unset ESPIC_LIB_PATH ESPIC_MOVE_THREADS
cat $L
