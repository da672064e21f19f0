#!/bin/bash
set -u
O=gpurun_out/r2t
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
grep -E "^E  |passed|failed|Error|rc=" $O/pytest_gpu.log | head -30; tail -2 $O/smoke.log
