#!/bin/bash
mkdir -p gpurun_out
HIRM_CLUSTER_SIZE=16 ZTABLES=47 NC=2 MOVES=200 timeout 900 ncu --set full --clock-control none --import-source on -k regex:generic_sweep -s 1 -c 1 -f -o gpurun_out/r02_generic_c5_k47 \
    python profiles/phase_breakdown_c5.py > gpurun_out/r02_ncu_c5_k47.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/r02_ncu_c5_k47.log
