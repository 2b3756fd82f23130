#!/bin/bash
mkdir -p gpurun_out
NC=1 MOVES=300 timeout 900 ncu --set full --clock-control none --import-source on -k regex:generic_sweep -s 1 -c 1 -f -o gpurun_out/r02_generic_c5_big \
    python profiles/phase_breakdown_c5.py > gpurun_out/r02_ncu_c5_big.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/r02_ncu_c5_big.log
