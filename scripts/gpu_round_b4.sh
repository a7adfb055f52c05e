#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -3
( timeout 900 python bench.py --no-host-arrays --no-cpu-baseline ) 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c3=d['configs3']
print('headline %.4g %.4f frac %.4f | configs3 %.5g  %.4f ms (instrumented %.4f) frac %.4f launches %d ok %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], c3['value'], c3['ms_per_step'], c3['ms_per_step_instrumented'], c3['roofline']['frac'], c3['gpu_launches'], c3['checks']['ok']))"
timeout 300 python scripts/small_run_latency.py 2>/dev/null > gpurun_out/small_run_latency.jsonl; python -c "
import json
for l in open('gpurun_out/small_run_latency.jsonl'):
    x=json.loads(l); print(x['n_particles'], x['periodic'], x['mode'], x['us_per_step'], x['kernels_per_step'])"
( timeout 600 python bench.py --walls --steps 40 --no-cpu-baseline --no-host-arrays --no-configs3 ) 2>/dev/null > gpurun_out/y_bench_walls.json; python -c "
import json
d=json.loads([x for x in open('gpurun_out/y_bench_walls.json') if x.startswith('{')][-1]); print('walls 4096: %.4f ms, %d launches' % (d['ms_per_step'], d['gpu_launches']))"
