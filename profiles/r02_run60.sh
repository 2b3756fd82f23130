#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest60.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest60.log
tail -3 gpurun_out/r02_pytest60.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02_smoke60.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r02_smoke60.log
