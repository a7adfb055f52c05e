#!/bin/bash
# mover time vs grid size (development probe)
for c in "$@"; do
  python bench.py --steps 10 --cells $c --no-cpu-baseline --no-host-arrays 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('cells', $c, 'value', d['value'], 'ms/step', d['ms_per_step'], 'mover ms', d['roofline']['launch_ms_mean'], 'frac', d['roofline']['frac'])"
done
