#!/bin/bash
# 2 GPUs: NCCL equality test of the sharded HIRM, bench.py at N=2 (C3 replicas, C4 replicas, C5 sharded)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_sharded_hirm.py -m gpu -q -k "nccl" > gpurun_out/r02_pytest9_nccl.log 2>&1; echo "nccl test exit $?"
tail -4 gpurun_out/r02_pytest9_nccl.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench9_n2.json 2> gpurun_out/r02_bench9_n2.err; echo "bench n2 exit $?"
tail -c 600 gpurun_out/r02_bench9_n2.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench9_n2.json'))
print('C3', d['value'], d['n_gpus'], d['ms_per_step'])
for k,v in d['config']['extra_workloads'].items():
    print(k, {a:b for a,b in v.items() if a in ('value','n_gpus','ms_per_step','entity_step_ms','relation_step_ms','split_s','events','irms_total','collectives')})
"
