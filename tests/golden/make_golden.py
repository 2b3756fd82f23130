"""Regenerate tests/golden/*.json.gz from the UNMODIFIED reference.

Runs oracle/_ref/ref_trace (oracle/ref_trace.cc compiled against the reference
headers in /root/reference/cxx, see oracle/Makefile) on the reference's own
datasets and on small synthetic datasets of the BASELINE.json config shapes.
Only runs where /root/reference exists (the build container); the fixtures it
writes are committed so that the GPU box never needs the reference tree.

    python tests/golden/make_golden.py
"""
import gzip
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("HIRM_REFERENCE", "/root/reference")
TRACE = os.path.join(ROOT, "oracle", "_ref", "ref_trace")


def trace(name, mode, base, seed, n, *flags):
    out = subprocess.check_output([TRACE, mode, base, str(seed), str(n), *flags])
    path = os.path.join(HERE, name + ".json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(out)
    print("%-28s %8d bytes (%d raw)" % (name, os.path.getsize(path), len(out)))


def write_dataset(tmp, name, schema_lines, obs_lines):
    base = os.path.join(tmp, name)
    with open(base + ".schema", "w") as f:
        f.write("\n".join(schema_lines) + "\n")
    with open(base + ".obs", "w") as f:
        f.write("\n".join(obs_lines) + "\n")
    return base


def synth_ree(tmp, n=40, seed=0):
    """BASELINE config 3 shape: 'bernoulli R e e', dense, two planted blocks, p = .9/.1 incl. diagonal."""
    rng = np.random.default_rng(seed)
    zs = (np.arange(n) >= n // 2).astype(int)
    p = np.where(zs[:, None] != zs[None, :], 0.9, 0.1)
    x = (rng.random((n, n)) < p).astype(int)
    obs = ["%d R e%d e%d" % (x[i, j], i, j) for i in range(n) for j in range(n)]
    return write_dataset(tmp, "ree%d" % n, ["bernoulli R e e"], obs)


def synth_ternary(tmp, na=12, nb=10, nc=9, nobs=420, seed=1):
    """BASELINE config 4 shape: sparse 'bernoulli R a b c' over three distinct domains."""
    rng = np.random.default_rng(seed)
    cells = rng.choice(na * nb * nc, size=nobs, replace=False)
    za, zb, zc = rng.integers(0, 3, na), rng.integers(0, 3, nb), rng.integers(0, 2, nc)
    theta = rng.beta(0.5, 0.5, size=(3, 3, 2))
    obs = []
    for c in cells:
        a, b, cc = c // (nb * nc), (c // nc) % nb, c % nc
        v = int(rng.random() < theta[za[a], zb[b], zc[cc]])
        obs.append("%d R a%d b%d c%d" % (v, a, b, cc))
    return write_dataset(tmp, "tern", ["bernoulli R a b c"], obs)


def synth_mixed(tmp, n=24, m=15, seed=2):
    """Config 5 shape in miniature: unary + binary (e x e, sparse) + binary (e x f) + ternary (f e e)."""
    rng = np.random.default_rng(seed)
    schema = ["bernoulli U%d e" % k for k in range(4)] + \
             ["bernoulli B e e", "bernoulli C e f", "bernoulli T f e e"]
    obs = []
    for k in range(4):
        for i in range(n):
            if rng.random() < 0.9:
                obs.append("%d U%d e%d" % (rng.integers(0, 2), k, i))
    for i in range(n):
        for j in range(n):
            if rng.random() < 0.5:
                obs.append("%d B e%d e%d" % (rng.integers(0, 2), i, j))
    for i in range(n):
        for j in range(m):
            if rng.random() < 0.7:
                obs.append("%d C e%d f%d" % (rng.integers(0, 2), i, j))
    for f in range(m):
        for i in range(n):
            for j in range(n):
                if rng.random() < 0.06:
                    obs.append("%d T f%d e%d e%d" % (rng.integers(0, 2), f, i, j))
    return write_dataset(tmp, "mixed", schema, obs)


def main():
    if not os.path.exists(TRACE):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    assets = os.path.join(REF, "cxx", "assets")
    with tempfile.TemporaryDirectory() as tmp:
        trace("animals_unary_irm", "irm", os.path.join(assets, "animals.unary"), 3, 2)
        trace("nations_binary_irm", "irm", os.path.join(assets, "nations.binary"), 12, 2)
        trace("two_relations_irm", "irm", os.path.join(assets, "two_relations"), 1, 2, "shuffle")
        trace("animals_unary_hirm", "hirm", os.path.join(assets, "animals.unary"), 10, 3)
        trace("synth_ree40_irm", "irm", synth_ree(tmp), 5, 3, "shuffle")
        trace("synth_ternary_irm", "irm", synth_ternary(tmp), 7, 3)
        trace("synth_mixed_irm", "irm", synth_mixed(tmp), 11, 2, "shuffle")


if __name__ == "__main__":
    sys.exit(main())
