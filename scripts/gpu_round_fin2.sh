#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/fin2_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/fin2_pytest.log; tail -5 gpurun_out/fin2_pytest.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 --no-host-arrays --no-configs3 2>/dev/null | tail -1 | cut -c1-220
