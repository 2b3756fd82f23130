#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest21.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest21.log
tail -4 gpurun_out/r02_pytest21.log
CHAINS=592,1184 MOVES=512 timeout 600 python profiles/phase_breakdown_c4.py > gpurun_out/r02_phase_c4_k.txt 2>&1; echo "phase c4 exit $?"
grep chains gpurun_out/r02_phase_c4_k.txt; tail -7 gpurun_out/r02_phase_c4_k.txt
NC=8,2 MOVES=1000 timeout 600 python profiles/phase_breakdown_c5.py > gpurun_out/r02_phase_c5_k.txt 2>&1; echo "phase c5 exit $?"
tail -16 gpurun_out/r02_phase_c5_k.txt
