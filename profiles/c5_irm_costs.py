"""Config 5 at its stated size, 8 chains from the CRP-prior seating of the relations: after ITER HIRM iterations, every IRM swept ALONE
(MOVES entity moves, cluster size as the launcher would give it: HIRM_CLUSTER_SIZE unset) -- time per move against the IRM's relations,
live tables and cluster size.  Shows what the entity step's slowest IRM is made of.
    python profiles/c5_irm_costs.py        env: ITER (2), MOVES (1000), CHAINS (8)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hierarchical-irm_b200")); sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench_workloads as W
from hirm.sharded import ChainSet

dev = torch.device("cuda", 0)
n, nu, nb, bobs = 200000, 2000, 50, 10_000_000
domains, schema, obs, truth = W.c5_hirm(dev, n=n, n_unary=nu, n_binary=nb, bobs=bobs, seed=0)
h = ChainSet(domains, schema, obs, int(os.environ.get("CHAINS", 8)), seed=0, max_tables=64, device=0)
h.iterate(int(os.environ.get("ITER", 2)))
eng = h.backend.eng
moves = int(os.environ.get("MOVES", 1000))
rng = np.random.default_rng(1)
rows = []
for c, ch in enumerate(h.chains):
    for t in sorted(ch.irms):
        info = ch.irms[t]
        hd = info["handle"]
        rels = info["rels"]
        n_un = sum(1 for r in rels if len(schema[ch.rnames[r]]) == 1)
        mv = [(np.zeros(moves, np.int32), rng.permutation(n)[:moves].astype(np.int32))]
        eng.sweep([hd], mv, seed=3, flags=1)                 # warm-up, score only (state unchanged)
        eng.sweep([hd], mv, seed=3, flags=1)
        inf = eng.last_sweep_info()
        tabs = len(np.unique(eng.get_assignments(hd, 0)))
        rows.append((inf["kernel_ms"] * 1e3 / moves, c, t, n_un, len(rels) - n_un, tabs, inf["cluster_size"], info.get("steps", -1)))
rows.sort(reverse=True)
print("us/move  chain table  unary binary  tables  cluster  sweeps_done")
for r in rows:
    print("%7.1f  %5d %5d  %5d %6d  %6d  %7d  %6d" % r)
