#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -3
for i in 1 2; do
( timeout 900 python bench.py --no-host-arrays --no-cpu-baseline ) 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c3=d['configs3']
print('headline %.4g %.4f frac %.4f | configs3 %.5g  %.4f ms (instrumented %.4f) frac %.4f launches %d ok %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], c3['value'], c3['ms_per_step'], c3['ms_per_step_instrumented'], c3['roofline']['frac'], c3['gpu_launches'], c3['checks']['ok']))"
done
