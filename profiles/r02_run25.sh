#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pregroup.py -m gpu -q -x > gpurun_out/r02_pytest25.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest25.log
tail -30 gpurun_out/r02_pytest25.log
