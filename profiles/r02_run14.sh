#!/bin/bash
mkdir -p gpurun_out
HIRM_DEBUG_CLUSTERS=1 timeout 1200 python bench.py --extra c5 --no-cpu-baseline --steps 1 --warmup 1 --burn-in 0 --chains 148 > gpurun_out/r02_bench14.json 2> gpurun_out/r02_bench14.err; echo "bench exit $?"
grep "hirm\]" gpurun_out/r02_bench14.err | tail -40
