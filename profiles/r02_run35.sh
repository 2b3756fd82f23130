#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pregroup.py tests/test_full_size.py -m gpu -q -x 2>&1 | tail -3
CHAINS=592,1184,2368 THREADS=0 MOVES=1024 timeout 900 python profiles/phase_breakdown_c4.py 2>&1 | grep chains
