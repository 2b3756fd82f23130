#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cluster_sweep.py tests/test_engine_golden.py tests/test_pregroup.py tests/test_hashed_tables.py -m gpu -q -x 2>&1 | tail -5
timeout 1200 python profiles/c5_irm_costs.py > gpurun_out/r02_c5_irm_costs_b.txt 2>&1; echo "exit $?"
head -12 gpurun_out/r02_c5_irm_costs_b.txt
