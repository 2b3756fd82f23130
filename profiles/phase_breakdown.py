import sys, os, time, json
sys.path.insert(0, 'hierarchical-irm_b200'); sys.path.insert(0, '.')
import numpy as np, torch
import bench
from hirm import engine as E
n = int(os.environ.get('N', 20000)); chains = int(os.environ.get('CHAINS', 148)); M = int(os.environ.get('M', 500)); kcap = int(os.environ.get('KCAP', 64))
burn = int(os.environ.get('BURN', 0))
pitch = (n + 15)//16*16
x = torch.full((n, pitch), 0xFF, dtype=torch.uint8).pin_memory()
bench.planted_ree_into(x.numpy(), n, 0)
eng = E.Engine(0)
h0 = eng.irm_create([n], [[0,0]], max_tables=kcap)
eng.set_observations_dense_host_ptr(h0, 0, x.data_ptr(), pitch); eng.finalize(h0); eng.synchronize()
hs=[h0]
for c in range(1, chains):
    h = eng.irm_create([n], [[0,0]], max_tables=kcap); eng.share_observations(h, h0); hs.append(h)
rng = np.random.default_rng(1)
for h in hs:
    if os.environ.get('INIT', 'crp') == 'crp': z = bench.crp_prior_labels_fast(n, 1.0, rng)
    elif os.environ.get('INIT', 'crp').startswith('blocks'): z = (np.arange(n) * int(os.environ['INIT'][6:]) // n).astype(np.int32)
    else: z = (np.arange(n) >= n//2).astype(np.int32)
    eng.set_assignments(h, 0, z)
buf = torch.zeros(chains*32, dtype=torch.int64, device='cuda')
def run(Mm, seedoff):
    moves = [(np.zeros(Mm, np.int32), rng.permutation(n).astype(np.int32)[:Mm]) for _ in hs]
    t0 = time.perf_counter(); eng.sweep(hs, moves, seed=seedoff, streams=np.arange(chains, dtype=np.uint64)); t1 = time.perf_counter()
    return eng.last_sweep_info()['kernel_ms']
if burn: run(burn, 1)
run(50, 2)
if not os.environ.get('NOPHASE'): eng.debug_phase_cycles(buf.data_ptr())
ms = run(M, 3)
ph = buf.cpu().numpy().reshape(chains, 32).astype(np.float64) / M
names = ['P total','P wait-empty','A total','A wait-go','-','-','-','-','D wait-ready','D fix+go','D score','D sample','D apply','D loop-top','-','-','-','-','-','D cand-setup']
print('n=%d chains=%d M=%d kernel_ms=%.3f us/move=%.3f' % (n, chains, M, ms, ms*1e3/M))
print('cycles/move (mean over chains):')
for i, nm in enumerate(names):
    if nm != '-': print('  %-14s %9.0f' % (nm, ph[:, i].mean()))
print('tables alive:', [len(np.unique(eng.get_assignments(h,0))) for h in hs[:6]])
