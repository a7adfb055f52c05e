#!/bin/bash
# final evidence: tests, bench (headline + configs3 + host arrays + cpu baseline), reference arm, small runs, smoke
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/fin_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/fin_pytest.log; tail -4 gpurun_out/fin_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err
echo "bench exit $?"
( timeout 600 python bench.py --impl reference --steps 8 --warmup 1 ) > gpurun_out/y_bench_reference.json 2> gpurun_out/y_bench_reference.err
timeout 600 python scripts/small_run_latency.py > gpurun_out/small_run_latency.jsonl 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/y_smoke.log 2>&1; tail -1 gpurun_out/y_smoke.log
