#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_abi.py -q 2>&1 | tail -1
timeout 1200 python bench.py --extra c5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench54.json 2> gpurun_out/r02_bench54.err; echo "bench exit $?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench54.json') if l.startswith('{')][-1])
c5=d['config']['extra_workloads']['c5']
print({k:c5[k] for k in ('ms_per_step','entity_step_ms','relation_step_ms','steps','warmup','irms_total','relations_moved')})
"
