#!/bin/bash
# 8 GPUs: bench.py as the driver launches it (C3 / C4 replicas, C5 sharded over 8 ranks)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 3 --extra c4,c5 > gpurun_out/r02_bench20_n8.json 2> gpurun_out/r02_bench20_n8.err; echo "bench n8 exit $?"
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r02_bench20_n8.err | tail -5
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench20_n8.json') if l.startswith('{')][-1])
print('C3', d['value'], d['n_gpus'], d['ms_per_step'], d['e2e']['value'])
for k,v in d['config']['extra_workloads'].items():
    print(k, {a:b for a,b in v.items() if a in ('value','n_gpus','ms_per_step','entity_step_ms','relation_step_ms','split_s','events','irms_total','collectives')})
"
