#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cluster_sweep.py tests/test_chain_statistics.py -m gpu -q -x > gpurun_out/r02_pytest13.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest13.log
tail -4 gpurun_out/r02_pytest13.log
HIRM_DEBUG_CLUSTERS=1 timeout 1200 python bench.py --extra c5 --no-cpu-baseline > gpurun_out/r02_bench13.json 2> gpurun_out/r02_bench13.err; echo "bench exit $?"
grep hirm gpurun_out/r02_bench13.err | sort | uniq -c | head
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench13.json') if l.startswith('{')][-1])
v=d['config']['extra_workloads']['c5']
print({a:b for a,b in v.items() if a in ('value','ms_per_step','entity_step_ms','relation_step_ms','split_s')})
"
