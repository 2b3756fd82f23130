#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python profiles/c5_irm_costs.py > gpurun_out/r02_c5_irm_costs.txt 2>&1; echo "exit $?"
head -30 gpurun_out/r02_c5_irm_costs.txt; tail -5 gpurun_out/r02_c5_irm_costs.txt
