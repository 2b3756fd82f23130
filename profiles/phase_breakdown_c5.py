"""Per-phase SM cycles of the generic sweep kernel on IRMs of BASELINE config 5's shape at full size (200 000 entities;
NU unary + NB binary relations split over NC planted relation clusters => one IRM per cluster).
    python profiles/phase_breakdown_c5.py     env: NE NU NB BOBS NC (comma list) MOVES KCAP"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hierarchical-irm_b200")); sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench_workloads as W
from hirm.sharded import ChainSet


def env(k, d):
    return int(float(os.environ.get(k, d)))


dev = torch.device("cuda", 0)
ne, nu, nb, bobs, moves, kcap = env("NE", 200000), env("NU", 2000), env("NB", 50), env("BOBS", 1e7), env("MOVES", 2000), env("KCAP", 64)
names = ["A group (CSR row, z gather, smem atomics)", "B1 bitmap scan (+C candidates)", "B2 group decode", "D score", "E sample", "F apply"]
for nc in [int(x) for x in os.environ.get("NC", "8,2").split(",")]:
    domains, schema, obs, truth = W.c5_hirm(dev, n=ne, n_unary=nu, n_binary=nb, bobs=bobs, n_clusters=nc, seed=0)
    nchains = env("CHAINS", 1)
    h = ChainSet(domains, schema, obs, nchains, seed=0, max_tables=kcap, relation_tables={r: truth[r] for r in schema}, device=0)
    eng = h.backend.eng
    hs = [ch.irms[t]["handle"] for ch in h.chains for t in sorted(ch.irms)]
    rng = np.random.default_rng(1)
    if os.environ.get("ZTABLES"):                         # every IRM re-seated into ZTABLES tables of random sizes (a fresh IRM's CRP-prior-like state)
        kt = int(os.environ["ZTABLES"])
        for hh in hs:
            w = rng.dirichlet(np.ones(kt) * 0.7)
            eng.set_assignments(hh, 0, rng.choice(kt, size=ne, p=w).astype(np.int32))
    mv = [(np.zeros(moves, np.int32), rng.permutation(ne)[:moves].astype(np.int32)) for _ in hs]
    eng.sweep(hs, mv, seed=1)
    buf = torch.zeros(len(hs) * 32, dtype=torch.int64, device=dev)
    eng.debug_phase_cycles(buf.data_ptr())
    eng.sweep(hs, mv, seed=2)
    info = eng.last_sweep_info()
    eng.debug_phase_cycles(None)
    ph = buf.cpu().numpy().reshape(len(hs), 32)[:, :6].astype(np.float64) / moves
    tot = ph.mean(axis=0).sum()
    tabs = len(np.unique(eng.get_assignments(hs[0], 0)))
    print("NC %d: %d IRMs of %d unary + %d binary relations, %d moves each: kernel %.1f ms = %.1f us per move per IRM, tables %d"
          % (nc, len(hs), nu // nc, nb // nc, moves, info["kernel_ms"], info["kernel_ms"] * 1e3 / moves, tabs), flush=True)
    for i, nm in enumerate(names):
        print("    %-46s %9.0f cycles/move  %5.1f%%" % (nm, ph[:, i].mean(), 100 * ph[:, i].mean() / tot))
    del h, obs
    torch.cuda.empty_cache()
