#!/bin/bash
# 2 GPUs: the tests that need two devices (NCCL world 2 == world 1), smoke (runs the NCCL test), default bench on 2 ranks
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_sharded_hirm.py -m gpu -q > gpurun_out/r02_pytest59_n2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest59_n2.log
tail -4 gpurun_out/r02_pytest59_n2.log
timeout 900 python __graft_entry__.py smoke > gpurun_out/r02_smoke59_n2.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r02_smoke59_n2.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02_bench59_n2.json 2> gpurun_out/r02_bench59_n2.err; echo "bench exit $?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench59_n2.json') if l.startswith('{')][-1])
print('C3', d['value'], d['n_gpus'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
for k,v in d['config']['extra_workloads'].items():
    print(k, {a:b for a,b in v.items() if a in ('value','value_per_gpu','ms_per_step','entity_step_ms','relation_step_ms','events','n_gpus')})
"
