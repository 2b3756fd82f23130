#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest30.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest30.log
tail -4 gpurun_out/r02_pytest30.log
CHAINS=592,1184 THREADS=0 MOVES=1024 timeout 900 python profiles/phase_breakdown_c4.py > gpurun_out/r02_phase_c4_o.txt 2>&1; echo "phase c4 exit $?"
cat gpurun_out/r02_phase_c4_o.txt
NC=8,2,1 MOVES=1000 timeout 600 python profiles/phase_breakdown_c5.py > gpurun_out/r02_phase_c5_o.txt 2>&1; echo "phase c5 exit $?"
grep -E "^NC|cycles" gpurun_out/r02_phase_c5_o.txt
