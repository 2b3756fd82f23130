#!/bin/bash
timeout 900 python -m pytest tests/test_cluster_sweep.py tests/test_engine_golden.py tests/test_pregroup.py tests/test_hashed_tables.py tests/test_full_size.py -m gpu -q -x 2>&1 | tail -3
for K in 10 47; do
  echo "ZTABLES $K cluster 16"; HIRM_CLUSTER_SIZE=16 ZTABLES=$K NC=2 MOVES=500 timeout 600 python profiles/phase_breakdown_c5.py 2>&1 | grep -E "^NC|D score"
done
