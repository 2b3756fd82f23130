#!/bin/bash
# end-of-round candidate run at N=1: whole GPU suite, smoke, default bench with all sub-records
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest55.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest55.log
tail -4 gpurun_out/r02_pytest55.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02_smoke55.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r02_smoke55.log
timeout 1500 python bench.py > gpurun_out/r02_bench55.json 2> gpurun_out/r02_bench55.err; echo "bench exit $?"
tail -c 300 gpurun_out/r02_bench55.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench55.json') if l.startswith('{')][-1])
print('C3', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'] if d['cpu_baseline'] else None)
for k,v in d['config']['extra_workloads'].items():
    print(k, {a:b for a,b in v.items() if a in ('value','value_per_gpu','ms_per_step','entity_step_ms','relation_step_ms','split_s','ms_per_sweep','rounds_per_step','rows_grouped_ahead')})
    if k == 'dense_k_sweep':
        for r in v['rows']: print('   K', r['planted_tables'], '%.1f M moves/s' % (r['moves_per_s']/1e6), '%.3f' % r['frac_of_hbm_peak'])
"
