"""Golden chain statistics for BASELINE config 3's shape, `bernoulli R e e` (a REPEATED domain: SURVEY.md 8a-Q1.5).

For every seed two chains of the reference on the same dense two-block R(e, e) data set (n = 60, p = .9 / .1, diagonal
included -- the generator of SURVEY.md 8d in miniature):
  * `ref_bf_logp_score`: oracle/_ref/ref_chain_bf -- the reference's own inference_irm iteration with the candidate
    log-probabilities taken from the brute force over its own scoring (re-seat, difference Relation::logp_score) instead of
    Relation::logp_gibbs_exact.  This is the exact Gibbs conditional; the engine and the oracle must match it.
  * `ref_logp_score`: the UNMODIFIED reference CLI (`hirm.out --mode irm --verbose`), kept beside it to show the size of
    the reference's repeated-domain bias (cxx/hirm.hh:421) on this shape -- not asserted against.
Writes tests/golden/chain_stats_ree.json.gz (observations included).

    python tests/golden/make_chain_stats_ree.py
"""
import gzip
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref")
SEEDS, ITERS, N = 64, 40, 60


def main():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    rng = np.random.default_rng(21)
    zs = np.arange(N) >= N // 2
    x = (rng.random((N, N)) < np.where(zs[:, None] != zs[None, :], 0.9, 0.1)).astype(int)
    with tempfile.TemporaryDirectory() as tmp:
        base = os.path.join(tmp, "ree%d" % N)
        with open(base + ".schema", "w") as f:
            f.write("bernoulli R e e\n")
        with open(base + ".obs", "w") as f:
            for i in range(N):
                for j in range(N):
                    f.write("%d R e%d e%d\n" % (x[i, j], i, j))
        bf, bk, plain = [], [], []
        for seed in range(1, SEEDS + 1):
            o = subprocess.check_output([os.path.join(REF, "ref_chain_bf"), base, str(seed), str(ITERS)], text=True)
            rows = [l.split() for l in o.splitlines() if l.startswith("iter")]
            assert len(rows) == ITERS
            bf.append([float(r[2]) for r in rows]); bk.append([int(r[3]) for r in rows])
            o = subprocess.check_output([os.path.join(REF, "hirm.out"), "--mode", "irm", "--seed", str(seed), "--iters", str(ITERS), "--verbose", base], text=True)
            vals = [float(l.split()[1]) for l in o.splitlines() if len(l.split()) == 2 and l[0].isdigit()]
            per = N + 1
            assert len(vals) == per * ITERS
            plain.append([vals[(k + 1) * per - 1] for k in range(ITERS)])
    obs = [[i, j, int(x[i, j])] for i in range(N) for j in range(N)]       # codes = first appearance = e0, e1, ... in file order
    out = {"seeds": SEEDS, "iters": ITERS, "generator": "tests/golden/make_chain_stats_ree.py",
           "datasets": {"ree60": {"domains": {"e": N}, "relations": [{"name": "R", "domains": ["e", "e"], "obs": obs}],
                                  "ref_bf_logp_score": bf, "ref_bf_tables": bk, "ref_logp_score": plain}}}
    with gzip.GzipFile(os.path.join(ROOT, "tests", "golden", "chain_stats_ree.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(out, separators=(",", ":")).encode())
    a, b = np.asarray(bf)[:, ITERS // 2:], np.asarray(plain)[:, ITERS // 2:]
    print("ree%d: brute-force reference logp_score after burn-in %.2f (sd of chain means %.2f, tables %.2f); unmodified reference %.2f (sd %.2f)"
          % (N, a.mean(), a.mean(axis=1).std(), np.asarray(bk)[:, ITERS // 2:].mean(), b.mean(), b.mean(axis=1).std()))


if __name__ == "__main__":
    sys.exit(main())
