#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest6.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest6.log
tail -6 gpurun_out/r02_pytest6.log
NC=8,2 MOVES=1000 timeout 600 python profiles/phase_breakdown_c5.py > gpurun_out/r02_phase_c5_d.txt 2>&1; echo "phase c5 exit $?"
tail -16 gpurun_out/r02_phase_c5_d.txt
HIRM_CLUSTER_SIZE=1 NC=8 MOVES=1000 timeout 600 python profiles/phase_breakdown_c5.py > gpurun_out/r02_phase_c5_d1.txt 2>&1; echo "phase c5 (no cluster) exit $?"
tail -8 gpurun_out/r02_phase_c5_d1.txt
CHAINS=148,592 MOVES=512 timeout 600 python profiles/phase_breakdown_c4.py > gpurun_out/r02_phase_c4_d.txt 2>&1; echo "phase c4 exit $?"
tail -16 gpurun_out/r02_phase_c4_d.txt
