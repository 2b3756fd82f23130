#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest7.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest7.log
tail -6 gpurun_out/r02_pytest7.log
NC=8,2 MOVES=1000 timeout 600 python profiles/phase_breakdown_c5.py > gpurun_out/r02_phase_c5_e.txt 2>&1; echo "phase c5 exit $?"
tail -16 gpurun_out/r02_phase_c5_e.txt
timeout 1200 python bench.py > gpurun_out/r02_bench7.json 2> gpurun_out/r02_bench7.err; echo "bench exit $?"
tail -c 400 gpurun_out/r02_bench7.err
