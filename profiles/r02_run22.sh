#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest22.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest22.log
tail -4 gpurun_out/r02_pytest22.log
NC=8,2,1 MOVES=1000 timeout 600 python profiles/phase_breakdown_c5.py > gpurun_out/r02_phase_c5_l.txt 2>&1; echo "phase c5 exit $?"
tail -24 gpurun_out/r02_phase_c5_l.txt
