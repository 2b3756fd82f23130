"""Throughput of the generic (CSR) sweep kernel on a scaled BASELINE config 4 shape: sparse ternary
'bernoulli R a b c', planted 8 tables per domain, `chains` independent chains sharing the observations.
    python profiles/explore_generic.py   (env: NE entities per domain, NOBS cells, CHAINS, SWEEPS, KCAP)"""
import os
import sys
import time

sys.path.insert(0, 'hierarchical-irm_b200'); sys.path.insert(0, '.')
import numpy as np
from hirm import engine as E

ne = int(os.environ.get('NE', 20000)); nobs = int(os.environ.get('NOBS', 2000000)); chains = int(os.environ.get('CHAINS', 148))
sweeps = int(os.environ.get('SWEEPS', 2)); kcap = int(os.environ.get('KCAP', 32))
rng = np.random.default_rng(0)
cells = np.unique(rng.integers(0, ne ** 3, size=int(nobs * 1.02), dtype=np.int64))[:nobs]
ids = np.stack([cells // (ne * ne), (cells // ne) % ne, cells % ne], axis=1).astype(np.int32)
zt = [rng.integers(0, 8, ne) for _ in range(3)]
theta = rng.beta(0.5, 0.5, size=(8, 8, 8))
val = (rng.random(len(ids)) < theta[zt[0][ids[:, 0]], zt[1][ids[:, 1]], zt[2][ids[:, 2]]]).astype(np.uint8)
eng = E.Engine(0)
t0 = time.perf_counter()
h0 = eng.irm_create([ne, ne, ne], [[0, 1, 2]], max_tables=kcap)
eng.set_observations(h0, 0, ids, val); eng.finalize(h0)
hs = [h0]
for c in range(1, chains):
    h = eng.irm_create([ne, ne, ne], [[0, 1, 2]], max_tables=kcap); eng.share_observations(h, h0); hs.append(h)
for c, h in enumerate(hs):
    for d in range(3):
        eng.set_assignments(h, d, rng.integers(0, 12, ne).astype(np.int32))
    eng.set_rng(h, c, 0)
eng.logp_scores(hs)
print('setup %.1f s' % (time.perf_counter() - t0), flush=True)
for s in range(sweeps + 1):
    t0 = time.perf_counter()
    nm = eng.sweep_all(hs, n_sweeps=1, seed=3)
    dt = time.perf_counter() - t0
    info = eng.last_sweep_info()
    deg = 3.0 * len(ids) / (3 * ne)
    print('sweep %d: %d moves in %.3f s (kernel %.1f ms) -> %.2f M moves/s, %.1f GB/s CSR bytes (deg %.0f, 9 B/visit), tables %s'
          % (s, nm, dt, info['kernel_ms'], nm / dt / 1e6, nm * deg * 9 / dt / 1e9, deg,
             [len(np.unique(eng.get_assignments(hs[0], d))) for d in range(3)]), flush=True)
