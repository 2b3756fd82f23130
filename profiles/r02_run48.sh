#!/bin/bash
for T in 256 512 1024; do
  echo "threads $T"; HIRM_GENERIC_THREADS=$T NC=8,1 MOVES=1000 timeout 600 python profiles/phase_breakdown_c5.py 2>&1 | grep -E "^NC"
done
