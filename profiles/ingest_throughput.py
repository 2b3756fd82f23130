"""Throughput of the GPU .obs reader (csrc/ingest_gpu.cu) on a synthetic file of BASELINE config 4's shape: NL lines
'<0|1> R a<i> b<j> c<k>' over 3 x 100 000 entities; the host reader on a 1e6-line sample of the same file for the codes.
    python profiles/ingest_throughput.py     env: NL (default 1e8) DIR (default /dev/shm or /tmp)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hierarchical-irm_b200"))
import numpy as np

from hirm import engine as E

nl = int(float(os.environ.get("NL", 1e8)))
d = os.environ.get("DIR", "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp")
base = os.path.join(d, "hirm_ingest_bench")
import subprocess
t0 = time.perf_counter()
with open(base + ".schema", "w") as f:
    f.write("bernoulli R a b c\n")
gen = os.path.join(d, "hirm_gen_obs")
subprocess.check_call(["gcc", "-O2", "-o", gen, os.path.join(ROOT, "profiles", "microbench", "gen_obs.c")])
subprocess.check_call([gen, base + ".obs", str(nl)])          # distinct cells: a strided walk through the 1e15-cell cube
size = os.path.getsize(base + ".obs")
print("wrote %d lines, %.2f GB in %.0f s" % (nl, size / 1e9, time.perf_counter() - t0), flush=True)
for rep in range(2):
    t0 = time.perf_counter()
    dom, schema, obs, names, times = E.load_dataset_gpu(base + ".schema", base + ".obs", with_names=False)
    dt = time.perf_counter() - t0
    print("GPU reader run %d: %.2f s wall for %d lines (%.1f M lines/s, %.2f GB/s of text); stages %s; domains %s; nobs %d"
          % (rep, dt, nl, nl / dt / 1e6, size / dt / 1e9, {k: round(v, 3) for k, v in times.items()}, dom, obs["R"][1].numel()), flush=True)
# codes against the host reader on the first 1e6 lines (first-appearance codes of a prefix are a prefix of the codes)
sample = base + "_sample.obs"
with open(base + ".obs", "rb") as f, open(sample, "wb") as g:
    n = 0
    for line in f:
        g.write(line); n += 1
        if n >= 1_000_000:
            break
t0 = time.perf_counter()
dh, sh, oh, nh = E.load_dataset(base + ".schema", sample)
th = time.perf_counter() - t0
dg, sg, og, ng, _ = E.load_dataset_gpu(base + ".schema", sample)
assert dh == dg and nh == ng
a = np.concatenate([oh["R"][0], oh["R"][1][:, None].astype(np.int32)], 1)
b = np.concatenate([og["R"][0].cpu().numpy(), og["R"][1].cpu().numpy()[:, None].astype(np.int32)], 1)
assert np.array_equal(a[np.lexsort(a.T[::-1])], b[np.lexsort(b.T[::-1])])
print("1e6-line sample: host reader %.2f s (%.2f M lines/s); GPU reader gives identical domains, names, codes and observations" % (th, 1e6 / th / 1e6))
for p in (base + ".obs", base + ".schema", sample):
    os.remove(p)
