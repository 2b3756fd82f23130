#!/bin/bash
for CH in 1 3 5; do
  echo "CHAINS $CH cluster 8"; HIRM_DEBUG_CLUSTERS=1 HIRM_CLUSTER_SIZE=8 CHAINS=$CH NC=2 MOVES=500 timeout 600 python profiles/phase_breakdown_c5.py 2>&1 | grep -E "^NC|cycles|class|halving" | tail -9
done
