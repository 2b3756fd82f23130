#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest12.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest12.log
tail -4 gpurun_out/r02_pytest12.log
NC=8,2 MOVES=1000 timeout 600 python profiles/phase_breakdown_c5.py > gpurun_out/r02_phase_c5_g.txt 2>&1; echo "phase c5 exit $?"
tail -18 gpurun_out/r02_phase_c5_g.txt
timeout 1500 python bench.py > gpurun_out/r02_bench12.json 2> gpurun_out/r02_bench12.err; echo "bench exit $?"
tail -c 300 gpurun_out/r02_bench12.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench12.json') if l.startswith('{')][-1])
print('C3', d['value'], d['n_gpus'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
for k,v in d['config']['extra_workloads'].items():
    print(k, {a:b for a,b in v.items() if a in ('value','n_gpus','ms_per_step','entity_step_ms','relation_step_ms','split_s','events','irms_total','collectives')})
"
