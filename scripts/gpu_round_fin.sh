#!/bin/bash
# final evidence: tests, bench (headline + configs3 + host arrays + cpu baseline), reference arm, probes, smoke
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/fin_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/fin_pytest.log; tail -4 gpurun_out/fin_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err
echo "bench exit $?"
( timeout 600 python bench.py --impl reference --steps 8 --warmup 1 ) > gpurun_out/y_bench_reference.json 2> gpurun_out/y_bench_reference.err
PROBE_STEPS=128 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 > gpurun_out/y_c4prod_probe.json
PROBE_RUN=1 PROBE_STEPS=256 timeout 300 python scripts/c4_production_probe.py 2>&1 | tail -1 > gpurun_out/y_c4prod_probe_run.json
PROBE_STEPS=28 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/y_c4prod_launches.csv python scripts/c4_production_probe.py > gpurun_out/y_c4prod_ncu.log 2>&1
python scripts/summarise_launches.py gpurun_out/y_c4prod_launches.csv 128 | grep -v "^at::\|at_cuda\|unnamed" | head -24 > gpurun_out/y_c4prod_summary.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/y_smoke.log 2>&1; tail -1 gpurun_out/y_smoke.log
