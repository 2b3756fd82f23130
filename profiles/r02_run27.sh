#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pregroup.py -m gpu -q -x 2>&1 | tail -3
CHAINS=592,1184 THREADS=0 MOVES=1024 timeout 900 python profiles/phase_breakdown_c4.py > gpurun_out/r02_phase_c4_m.txt 2>&1; echo "phase c4 exit $?"
cat gpurun_out/r02_phase_c4_m.txt
CHAINS=592 MOVES=256 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pregroup_rows|generic_sweep" -s 6 -c 2 -f -o gpurun_out/r02_pregroup_c4 \
    python profiles/phase_breakdown_c4.py > gpurun_out/r02_ncu_pg.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/r02_ncu_pg.log
