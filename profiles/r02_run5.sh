#!/bin/bash
# ncu source-level profile of the generic kernel on config-5-shaped IRMs (NC=8: 8 IRMs of 250 unary + 6 binary), one CTA per IRM
mkdir -p gpurun_out
export HIRM_CLUSTER_SIZE=1 NC=8 MOVES=300
python profiles/phase_breakdown_c5.py > gpurun_out/r02_ncu_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:generic_sweep -s 1 -c 1 -f -o gpurun_out/r02_generic_c5 \
    python profiles/phase_breakdown_c5.py > gpurun_out/r02_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/r02_ncu.log; ls -la gpurun_out/*.ncu-rep
