#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pregroup.py tests/test_cluster_sweep.py -m gpu -q -x 2>&1 | tail -3
CHAINS=592,1184 THREADS=0 MOVES=1024 timeout 900 python profiles/phase_breakdown_c4.py > gpurun_out/r02_phase_c4_p.txt 2>&1; echo "phase c4 exit $?"
cat gpurun_out/r02_phase_c4_p.txt
