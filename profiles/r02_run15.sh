#!/bin/bash
mkdir -p gpurun_out
HIRM_CLUSTER_SIZE=1 NC=8 MOVES=1000 timeout 600 python profiles/phase_breakdown_c5.py > gpurun_out/r02_phase_c5_dbg1.txt 2>&1; tail -9 gpurun_out/r02_phase_c5_dbg1.txt
NC=8 MOVES=1000 timeout 600 python profiles/phase_breakdown_c5.py > gpurun_out/r02_phase_c5_dbg8.txt 2>&1; tail -9 gpurun_out/r02_phase_c5_dbg8.txt
