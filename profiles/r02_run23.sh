#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest23.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest23.log
tail -4 gpurun_out/r02_pytest23.log
NC=8,2,1 MOVES=1000 timeout 600 python profiles/phase_breakdown_c5.py > gpurun_out/r02_phase_c5_m.txt 2>&1; echo "phase c5 exit $?"
grep -E "^NC" gpurun_out/r02_phase_c5_m.txt
timeout 1500 python bench.py > gpurun_out/r02_bench23.json 2> gpurun_out/r02_bench23.err; echo "bench exit $?"
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench23.json'):
    l=l.strip()
    if not l.startswith('{'): continue
    d=json.loads(l)
    print('C3', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
    for k in ('c4','c5','c3_at_reference_sample_size'):
        print(k, d.get('extra',{}).get(k) or d.get(k))
PY
