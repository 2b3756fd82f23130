#!/bin/bash
# round 2, GPU call 2: GPU test suite after the accumulator / hashed-table changes, phase breakdowns (C4, C5 shape)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest2.log
tail -5 gpurun_out/r02_pytest2.log
CHAINS=148,592,1184 MOVES=512 timeout 600 python profiles/phase_breakdown_c4.py > gpurun_out/r02_phase_c4_b.txt 2>&1; echo "phase c4 exit $?"
tail -24 gpurun_out/r02_phase_c4_b.txt
NC=8,2 MOVES=1000 timeout 600 python profiles/phase_breakdown_c5.py > gpurun_out/r02_phase_c5.txt 2>&1; echo "phase c5 exit $?"
tail -20 gpurun_out/r02_phase_c5.txt
