#!/bin/bash
mkdir -p gpurun_out
for K in 4 8 16 32; do
  INIT=blocks$K CHAINS=1184 M=300 timeout 600 python profiles/phase_breakdown.py 2>&1 | tail -16
done > gpurun_out/r02_phase_dense_k.txt 2>&1
cat gpurun_out/r02_phase_dense_k.txt
