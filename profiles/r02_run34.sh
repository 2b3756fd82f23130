#!/bin/bash
# candidate run at N=1 with the pre-grouping path: whole GPU suite, smoke, default bench with all sub-records, then ncu --set full of
# the pre-grouping + sequential kernels on config 4 (592 chains x 256 moves per domain)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest34.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest34.log
tail -4 gpurun_out/r02_pytest34.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02_smoke34.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r02_smoke34.log
timeout 1500 python bench.py > gpurun_out/r02_bench34.json 2> gpurun_out/r02_bench34.err; echo "bench exit $?"
tail -c 300 gpurun_out/r02_bench34.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench34.json') if l.startswith('{')][-1])
print('C3', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'] if d['cpu_baseline'] else None)
for k,v in d['config']['extra_workloads'].items():
    print(k, {a:b for a,b in v.items() if a in ('value','value_per_gpu','ms_per_step','entity_step_ms','relation_step_ms','split_s','ms_per_sweep','rounds_per_step','rows_grouped_ahead')})
"
CHAINS=592 MOVES=256 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pregroup_rows|generic_sweep" -s 6 -c 2 -f -o gpurun_out/r02_pregroup_c4_final \
    python profiles/phase_breakdown_c4.py > gpurun_out/r02_ncu_pg_final.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/r02_ncu_pg_final.log
