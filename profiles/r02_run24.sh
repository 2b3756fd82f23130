#!/bin/bash
mkdir -p gpurun_out
CHAINS=148,296,592 THREADS=256,512,1024 MOVES=1024 timeout 900 python profiles/phase_breakdown_c4.py > gpurun_out/r02_phase_c4_k.txt 2>&1; echo "phase c4 exit $?"
cat gpurun_out/r02_phase_c4_k.txt
