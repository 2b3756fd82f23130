#!/bin/bash
mkdir -p gpurun_out
CHAINS=888,1184 THREADS=128,192,256 MOVES=1024 timeout 900 python profiles/phase_breakdown_c4.py 2>&1 | grep -E "chains|D score|A group"
