#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest41.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest41.log
tail -5 gpurun_out/r02_pytest41.log
for K in 2 8 16 32; do
  INIT=blocks$K CHAINS=1184 M=300 timeout 600 python profiles/phase_breakdown.py 2>&1 | grep -E "kernel_ms|D score|A total"
done
timeout 900 python bench.py --extra none --steps 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('C3', d['value'], d['ms_per_step'], d['roofline']['frac'])
"
CHAINS=1184 THREADS=0 MOVES=1024 timeout 900 python profiles/phase_breakdown_c4.py 2>&1 | grep -E "chains|D score"
NC=8,2 MOVES=1000 timeout 600 python profiles/phase_breakdown_c5.py 2>&1 | grep -E "^NC"
