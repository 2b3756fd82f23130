#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest17.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest17.log
tail -4 gpurun_out/r02_pytest17.log
CHAINS=592 MOVES=512 timeout 600 python profiles/phase_breakdown_c4.py > gpurun_out/r02_phase_c4_i.txt 2>&1; echo "phase c4 exit $?"
tail -8 gpurun_out/r02_phase_c4_i.txt
NC=8,2 MOVES=1000 timeout 600 python profiles/phase_breakdown_c5.py > gpurun_out/r02_phase_c5_i.txt 2>&1; echo "phase c5 exit $?"
tail -18 gpurun_out/r02_phase_c5_i.txt
CHAINS=592 MOVES=256 timeout 900 ncu --set full --clock-control none --import-source on -k regex:generic_sweep -s 1 -c 1 -f -o gpurun_out/r02_generic_c4 \
    python profiles/phase_breakdown_c4.py > gpurun_out/r02_ncu_c4.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/r02_ncu_c4.log
