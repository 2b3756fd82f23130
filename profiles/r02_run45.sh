#!/bin/bash
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02_smoke45.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r02_smoke45.log
