#!/bin/bash
# round 2, GPU call 1: full GPU test suite, default bench (with the C4 / C5 sub-records), C4 phase breakdown
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest1.log
tail -5 gpurun_out/r02_pytest1.log
timeout 900 python bench.py > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; echo "bench exit $?"
tail -c 600 gpurun_out/r02_bench1.err
CHAINS=148,592,1184 MOVES=512 timeout 600 python profiles/phase_breakdown_c4.py > gpurun_out/r02_phase_c4.txt 2>&1; echo "phase exit $?"
cat gpurun_out/r02_phase_c4.txt | tail -40
