#!/bin/bash
# C3 headline with the four combinations of (ln by table + short log1p series) x (candidate sums by one lane each)
mkdir -p gpurun_out
cp hierarchical-irm_b200/libhirm_engine.so /tmp/lib_new_new.so
for v in oldlog_oldsum newlog_oldsum oldlog_newsum new_new; do
  if [ $v = new_new ]; then cp /tmp/lib_new_new.so hierarchical-irm_b200/libhirm_engine.so; else cp gpurun_variants/lib_$v.so hierarchical-irm_b200/libhirm_engine.so; fi
  timeout 900 python bench.py --extra none --steps 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$v C3', d['value'], d['ms_per_step'], d['roofline']['frac'])
"
done
cp /tmp/lib_new_new.so hierarchical-irm_b200/libhirm_engine.so
