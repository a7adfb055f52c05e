#!/bin/bash
# bulk-copy staging A/B (stages, walls), window CTA width, then the regular suites and the bench
mkdir -p gpurun_out
L=gpurun_out/v_variants.log
echo "== k_move periodic, ms per 1e8-particle launch" > $L
for v in cur notma tma3 cur tma3; do
  if [ $v = cur ]; then unset ESPIC_LIB_PATH; else export ESPIC_LIB_PATH=$PWD/variants/libespic_$v.so; fi
  echo -n "$v: " >> $L; timeout 300 python scripts/mover_micro_probe.py 2>&1 | tail -1 >> $L
done
echo "== k_move walls, ms per 1e8-particle launch" >> $L
for v in cur wst wst3 cur wst; do
  if [ $v = cur ]; then unset ESPIC_LIB_PATH; else export ESPIC_LIB_PATH=$PWD/variants/libespic_$v.so; fi
  echo -n "$v: " >> $L; PROBE_WALLS=1 timeout 300 python scripts/mover_micro_probe.py 2>&1 | tail -1 >> $L
done
echo "== window mover CTA width, production-like 65536-cell run (ms/step)" >> $L
for v in cur win1024; do
  if [ $v = cur ]; then unset ESPIC_LIB_PATH; else export ESPIC_LIB_PATH=$PWD/variants/libespic_$v.so; fi
  echo -n "$v: " >> $L; PROBE_STEPS=56 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 | cut -c1-160 >> $L
done
unset ESPIC_LIB_PATH
cat $L
timeout 600 python scripts/small_run_latency.py > gpurun_out/small_run_latency.jsonl 2> gpurun_out/small_run_latency.err
cat gpurun_out/small_run_latency.jsonl | cut -c1-170; tail -3 gpurun_out/small_run_latency.err
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 ) > gpurun_out/v_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/v_pytest.log
tail -12 gpurun_out/v_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err
echo "bench exit $?"; tail -1 gpurun_out/v_bench.json | cut -c1-600
