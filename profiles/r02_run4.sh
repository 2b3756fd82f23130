#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest4.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest4.log
tail -12 gpurun_out/r02_pytest4.log
NC=8,2 MOVES=1000 timeout 600 python profiles/phase_breakdown_c5.py > gpurun_out/r02_phase_c5_c.txt 2>&1; echo "phase c5 exit $?"
tail -20 gpurun_out/r02_phase_c5_c.txt
HIRM_CLUSTER_SIZE=1 NC=8,2 MOVES=1000 timeout 600 python profiles/phase_breakdown_c5.py > gpurun_out/r02_phase_c5_c1.txt 2>&1; echo "phase c5 (no cluster) exit $?"
tail -20 gpurun_out/r02_phase_c5_c1.txt
