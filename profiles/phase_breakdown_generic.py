"""Where a move of the generic (CSR) sweep kernel spends its time: thread 0's SM cycles per phase
(hirm_engine_debug_phase_cycles) on the IRMs of a scaled config 5 (profiles/hirm_scale.py generator).
    python profiles/phase_breakdown_generic.py     env: NE NU NB BOBS NC KCAP CHAINS"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hierarchical-irm_b200")); sys.path.insert(0, os.path.join(ROOT, "profiles"))
import numpy as np
import torch

import hirm_scale as HS
from hirm.sharded import ChainSet

env = HS.env
ne, nu, nb, bobs, nc = env("NE", 20000), env("NU", 200), env("NB", 8), env("BOBS", 1000000), env("NC", 8)
domains, schema, obs, truth = HS.make(ne, nu, nb, bobs, nc, 0)
h = ChainSet(domains, schema, obs, env("CHAINS", 8), seed=0, max_tables=env("KCAP", 32), relation_tables={r: truth[r] for r in schema})
h.transition_irms(); h.transition_irms()
n_irms = sum(len(c.irms) for c in h.chains)
buf = torch.zeros(n_irms * 32, dtype=torch.int64, device="cuda")
h.backend.eng.debug_phase_cycles(buf.data_ptr())
h.transition_irms()
torch.cuda.synchronize()
info = h.backend.eng.last_sweep_info()
h.backend.eng.debug_phase_cycles(None)
ph = buf.cpu().numpy().reshape(n_irms, 32)[:, :6].astype(np.float64) / ne
names = ["A remove+group (CSR row, z gather, atomics, fence)", "B1 bitmap scan (+C candidates)", "B2 group decode", "D score", "E sample", "F apply (+fence)"]
print("%d IRMs, %d moves each, kernel %.1f ms = %.2f us per move per IRM" % (n_irms, ne, info["kernel_ms"], info["kernel_ms"] * 1e3 / ne))
tot = ph.mean(axis=0).sum()
for i, nm in enumerate(names):
    print("  %-52s %8.0f cycles/move  %5.1f%%" % (nm, ph[:, i].mean(), 100 * ph[:, i].mean() / tot))
print("  total %.0f cycles = %.2f us at 1.965 GHz" % (tot, tot / 1965.0))
