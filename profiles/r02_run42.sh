#!/bin/bash
mkdir -p gpurun_out
NOPHASE=1 INIT=blocks32 CHAINS=444 M=200 timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_ree -s 5 -c 1 -f -o gpurun_out/r02_dense_k32 \
    python profiles/phase_breakdown.py > gpurun_out/r02_ncu_dense_k32.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/r02_ncu_dense_k32.log
