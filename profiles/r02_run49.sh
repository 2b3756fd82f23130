#!/bin/bash
mkdir -p gpurun_out
HIRM_DEBUG_CLUSTERS=1 timeout 1200 python bench.py --extra c5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench49.json 2> gpurun_out/r02_bench49.err; echo "bench exit $?"
grep "class" gpurun_out/r02_bench49.err | tail -40
