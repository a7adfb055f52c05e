#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/l_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/l_pytest.log
tail -4 gpurun_out/l_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err
echo "bench exit $?"
L=gpurun_out/l_probe.log
echo "== k_move periodic CTA width variants" > $L
for v in m512 m640 m768 m832 m896 m960 m896c m512c cur; do
  if [ $v = cur ]; then unset ESPIC_LIB_PATH ESPIC_MOVE_THREADS; else export ESPIC_LIB_PATH=$PWD/variants/libespic_$v.so; t=${v#m}; export ESPIC_MOVE_THREADS=${t%c}; fi
  echo -n "$v: " >> $L; timeout 300 python scripts/mover_micro_probe.py 2>&1 | tail -1 >> $L
done
unset ESPIC_LIB_PATH ESPIC_MOVE_THREADS
echo "== k_move walls (4096 cells)" >> $L
PROBE_WALLS=1 timeout 300 python scripts/mover_micro_probe.py 2>&1 | tail -1 >> $L
cat $L
PROBE_STEPS=28 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_c4prod_launches.csv python scripts/c4_production_probe.py > gpurun_out/l_c4prod_ncu.log 2>&1
echo "ncu list exit $?"
