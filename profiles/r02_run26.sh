#!/bin/bash
mkdir -p gpurun_out
CHAINS=592,1184 THREADS=0 MOVES=1024 timeout 900 python profiles/phase_breakdown_c4.py > gpurun_out/r02_phase_c4_l.txt 2>&1; echo "phase c4 exit $?"
cat gpurun_out/r02_phase_c4_l.txt
HIRM_PREGROUP=0 CHAINS=592 THREADS=0 MOVES=1024 timeout 900 python profiles/phase_breakdown_c4.py > gpurun_out/r02_phase_c4_l0.txt 2>&1; echo "phase c4 exit $?"
cat gpurun_out/r02_phase_c4_l0.txt
