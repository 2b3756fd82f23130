#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pregroup.py tests/test_cluster_sweep.py tests/test_hashed_tables.py -m gpu -q -x 2>&1 | tail -3
CHAINS=592,1184 THREADS=0 MOVES=1024 timeout 900 python profiles/phase_breakdown_c4.py > gpurun_out/r02_phase_c4_n.txt 2>&1; echo "phase c4 exit $?"
cat gpurun_out/r02_phase_c4_n.txt
NC=8,2,1 MOVES=1000 timeout 600 python profiles/phase_breakdown_c5.py > gpurun_out/r02_phase_c5_n.txt 2>&1; echo "phase c5 exit $?"
grep -E "^NC" gpurun_out/r02_phase_c5_n.txt
