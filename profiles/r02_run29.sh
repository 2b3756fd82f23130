#!/bin/bash
mkdir -p gpurun_out
CHAINS=592 THREADS=0 MOVES=512 timeout 900 python profiles/phase_breakdown_c4.py 2>&1 | tail -9
