#!/bin/bash
mkdir -p gpurun_out
for t in 0 256; do
  if [ $t -gt 0 ]; then export HIRM_GENERIC_THREADS=$t; fi
  NE=20000 NU=200 NB=8 BOBS=1000000 NC=8 KCAP=32 CHAINS=256 timeout 600 python profiles/phase_breakdown_generic.py > gpurun_out/r02_phase_scaled_c5_t$t.txt 2>&1; echo "phase scaled c5 threads=$t exit $?"
  tail -9 gpurun_out/r02_phase_scaled_c5_t$t.txt
done
unset HIRM_GENERIC_THREADS
NE=20000 NU=200 NB=8 BOBS=1000000 NC=8 KCAP=32 CHAINS=64 timeout 600 python profiles/phase_breakdown_generic.py > gpurun_out/r02_phase_scaled_c5_c64.txt 2>&1
tail -9 gpurun_out/r02_phase_scaled_c5_c64.txt
