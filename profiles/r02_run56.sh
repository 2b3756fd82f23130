#!/bin/bash
for K in 10 47; do
  echo "ZTABLES $K cluster 16"; HIRM_CLUSTER_SIZE=16 ZTABLES=$K NC=2 MOVES=500 timeout 600 python profiles/phase_breakdown_c5.py 2>&1 | grep -E "^NC|cycles"
done
