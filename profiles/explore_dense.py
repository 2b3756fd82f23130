"""Tuning sweep for the dense R(e,e) kernel: moves/s and algorithmic GB/s for several
(chains per SM, chunk bytes) settings, at burn-in (CRP-prior start) and in steady state.
Run on the GPU box:  python profiles/explore_dense.py  > gpurun_out/explore.txt"""
import os
import sys
import time

sys.path.insert(0, 'hierarchical-irm_b200'); sys.path.insert(0, '.')
import numpy as np
import torch
import bench
from hirm import engine as E

n = int(os.environ.get('N', 20000)); kcap = int(os.environ.get('KCAP', 64)); M = int(os.environ.get('M', 2000))
configs = os.environ.get('CONFIGS', '1:2048,2:2048,2:4096,2:1024,3:2048').split(',')
pitch = (n + 15) // 16 * 16
x = torch.full((n, pitch), 0xFF, dtype=torch.uint8).pin_memory()
bench.planted_ree_into(x.numpy(), n, 0)
eng = E.Engine(0)
sms = torch.cuda.get_device_properties(0).multi_processor_count
maxc = max(int(c.split(':')[0]) for c in configs) * sms
h0 = eng.irm_create([n], [[0, 0]], max_tables=kcap)
eng.set_observations_dense_host_ptr(h0, 0, x.data_ptr(), pitch); eng.finalize(h0); eng.synchronize()
hs = [h0]
for c in range(1, maxc):
    h = eng.irm_create([n], [[0, 0]], max_tables=kcap); eng.share_observations(h, h0); hs.append(h)
rng = np.random.default_rng(1)


def init(kind):
    for h in hs:
        if kind == 'crp':
            z = bench.crp_prior_labels_fast(n, 1.0, rng)
        elif kind == 'rand64':
            z = rng.integers(0, min(kcap - 1, 63), n).astype(np.int32)
        elif kind.startswith('truth') and len(kind) > 5:      # truthK: the two planted blocks plus K-2 small tables of 7 entities
            z = (np.arange(n) >= n // 2).astype(np.int32)
            for t in range(2, int(kind[5:])):
                z[rng.choice(n, 7, replace=False)] = t
        else:
            z = (np.arange(n) >= n // 2).astype(np.int32)
        eng.set_assignments(h, 0, z)
    eng.get_counts(hs[0], 0)


def run(hsub, Mm, seed):
    moves = [(np.zeros(Mm, np.int32), rng.permutation(n).astype(np.int32)[:Mm]) for _ in hsub]
    eng.sweep(hsub, moves, seed=seed, streams=np.arange(len(hsub), dtype=np.uint64))
    return eng.last_sweep_info()['kernel_ms']


for kind in os.environ.get('INITS', 'truth,crp').split(','):
    for cfg in configs:
        cps, chunk = cfg.split(':')
        os.environ['HIRM_DENSE_CHUNK'] = chunk
        os.environ['HIRM_DENSE_CTAS_PER_SM'] = cps
        hsub = hs[:int(cps) * sms]
        init(kind)
        run(hsub, 50, 1)
        ms = run(hsub, M, 2)
        moves = len(hsub) * M
        tabs = [len(np.unique(eng.get_assignments(h, 0))) for h in hsub[:4]]
        print('init=%-6s chains/SM=%s chunk=%-5s  %8.3f ms  %7.2f M moves/s  %7.1f GB/s (%.1f%% of 6532)  us/move/chain=%.2f tables=%s plan=%s'
              % (kind, cps, chunk, ms, moves / ms / 1e3, moves * 2.0 * n / ms / 1e6, moves * 2.0 * n / ms / 1e6 / 65.32,
                 ms * 1e3 / M, tabs, eng.last_dense_plan()), flush=True)
