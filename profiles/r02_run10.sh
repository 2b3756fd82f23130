#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest10.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest10.log
tail -4 gpurun_out/r02_pytest10.log
HIRM_DEBUG_CLUSTERS=1 NC=8,2 MOVES=1000 timeout 600 python profiles/phase_breakdown_c5.py > gpurun_out/r02_phase_c5_f.txt 2>&1; echo "phase c5 exit $?"
tail -18 gpurun_out/r02_phase_c5_f.txt
CHAINS=592 MOVES=512 timeout 600 python profiles/phase_breakdown_c4.py > gpurun_out/r02_phase_c4_f.txt 2>&1; echo "phase c4 exit $?"
tail -8 gpurun_out/r02_phase_c4_f.txt
NE=20000 NU=200 NB=8 BOBS=1000000 NC=8 KCAP=32 CHAINS=256 ITERS=3 PLANTED=1 timeout 900 python profiles/hirm_scale.py > gpurun_out/r02_hirm_scaled_c256_n1.json 2> gpurun_out/r02_hirm_scaled.err; echo "hirm_scale exit $?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_hirm_scaled_c256_n1.json') if l.startswith('{')][-1])
print({k:v for k,v in d.items() if k in ('entity_moves_per_s','relation_step_s','timers','setup_s')})
"
