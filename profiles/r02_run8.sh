#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest8.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest8.log
tail -6 gpurun_out/r02_pytest8.log
NL=1e8 timeout 900 python profiles/ingest_throughput.py > gpurun_out/r02_ingest.txt 2>&1; echo "ingest exit $?"
tail -6 gpurun_out/r02_ingest.txt
timeout 1200 python bench.py --extra c5 > gpurun_out/r02_bench8.json 2> gpurun_out/r02_bench8.err; echo "bench exit $?"
tail -c 400 gpurun_out/r02_bench8.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench8.json'))
v=d['config']['extra_workloads']['c5']
print({a:b for a,b in v.items() if a in ('value','ms_per_step','entity_step_ms','relation_step_ms','split_s','events','irms_total','relations_moved')})
"
