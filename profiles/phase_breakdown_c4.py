"""Per-phase SM cycles of the generic (CSR) sweep kernel on BASELINE config 4 at full size (thread 0's clock64 per phase,
hirm_engine_debug_phase_cycles), for several chain counts / CTA widths.
    python profiles/phase_breakdown_c4.py     env: CHAINS (comma list) MOVES (per domain per chain) THREADS (comma list, 0 = engine's choice)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hierarchical-irm_b200")); sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench_workloads as W
from hirm import engine as E

dev = torch.device("cuda", 0)
n, nobs = int(os.environ.get("NE", 100000)), int(float(os.environ.get("NOBS", 1e8)))
ids, val, _ = W.c4_ternary(dev, n=n, nobs=nobs, seed=0)
eng = E.Engine(0)
m = int(os.environ.get("MOVES", 1024))
names = ["A group (CSR row, z gather, smem atomics)", "B1 bitmap scan (+C candidates)", "B2 group decode", "D score", "E sample", "F apply"]
h0 = eng.irm_create([n, n, n], [[0, 1, 2]], max_tables=64)
eng.set_observations_device(h0, 0, val.numel(), ids.data_ptr(), val.data_ptr())
eng.finalize(h0)
made = [h0]
for chains in [int(c) for c in os.environ.get("CHAINS", "148,592,1184").split(",")]:
    while len(made) < chains:
        h = eng.irm_create([n, n, n], [[0, 1, 2]], max_tables=64)
        eng.share_observations(h, h0)
        made.append(h)
    hs = made[:chains]
    for c, h in enumerate(hs):
        eng.set_rng(h, c, 0)
    eng.seat_prior(hs, seed=1)
    eng.logp_scores(hs)
    for thr in [int(t) for t in os.environ.get("THREADS", "0").split(",")]:
        if thr:
            os.environ["HIRM_GENERIC_THREADS"] = str(thr)
        else:
            os.environ.pop("HIRM_GENERIC_THREADS", None)
        g = torch.Generator(device=dev); g.manual_seed(5)
        per = 3 * m
        d_off = torch.arange(chains + 1, dtype=torch.int64, device=dev) * per
        d_dom = torch.arange(3, dtype=torch.int32, device=dev).repeat_interleave(m).repeat(chains).contiguous()
        d_stream = torch.arange(chains, dtype=torch.int64, device=dev)
        irm_ids = np.asarray(hs, dtype=np.int32)
        for rep in range(2):
            items = torch.randint(0, n, (chains * per,), device=dev, generator=g, dtype=torch.int32)
            ctr = torch.full((chains,), rep * per, dtype=torch.int64, device=dev)
            buf = torch.zeros(chains * 32, dtype=torch.int64, device=dev)
            eng.debug_phase_cycles(buf.data_ptr() if rep else None)
            eng.sweep_device(irm_ids, d_off.data_ptr(), d_dom.data_ptr(), items.data_ptr(), 3, d_stream.data_ptr(), ctr.data_ptr())
            torch.cuda.synchronize()
        info = eng.last_sweep_info()
        eng.debug_phase_cycles(None)
        ph = buf.cpu().numpy().reshape(chains, 32)[:, :6].astype(np.float64) / per
        tot = ph.mean(axis=0).sum()
        rate = chains * per / (info["kernel_ms"] * 1e-3)
        tabs = [len(np.unique(eng.get_assignments(hs[0], d))) for d in range(3)]
        print("chains %d threads %s: kernel %.1f ms, %.2f M moves/s = %.1f GB/s CSR bytes, %.1f us per move per chain, tables %s"
              % (chains, thr or "auto", info["kernel_ms"], rate / 1e6, rate * (nobs / n) * 9 / 1e9, info["kernel_ms"] * 1e3 / per, tabs), flush=True)
        for i, nm in enumerate(names):
            print("    %-46s %9.0f cycles/move  %5.1f%%" % (nm, ph[:, i].mean(), 100 * ph[:, i].mean() / tot))
        sub = buf.cpu().numpy().reshape(chains, 32)[:, 8:24].astype(np.float64) / per      # SUBT sub-timers (temporary instrumentation)
        if sub.any():
            print("    sub-timers [8..23]: " + " ".join("%d:%.0f" % (8 + j, sub[:, j].mean()) for j in range(16) if sub[:, j].any()))
