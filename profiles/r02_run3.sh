#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest3.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest3.log
tail -8 gpurun_out/r02_pytest3.log
NC=8,2 MOVES=1000 timeout 600 python profiles/phase_breakdown_c5.py > gpurun_out/r02_phase_c5_b.txt 2>&1; echo "phase c5 exit $?"
tail -20 gpurun_out/r02_phase_c5_b.txt
CHAINS=592 MOVES=512 timeout 600 python profiles/phase_breakdown_c4.py > gpurun_out/r02_phase_c4_c.txt 2>&1; echo "phase c4 exit $?"
tail -8 gpurun_out/r02_phase_c4_c.txt
