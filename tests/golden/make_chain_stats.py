"""Golden chain statistics from the UNMODIFIED reference CLI (oracle/_ref/hirm.out, built by oracle/Makefile from
/root/reference/cxx): for every seed, `hirm.out --mode irm --seed S --iters T --verbose` -- the reference's own
mt19937, sweep order and alpha moves (cxx/hirm.cc:39-59) -- and the IRM::logp_score it prints (REPORT_SCORE,
cxx/hirm.cc:26-37) at the end of every iteration.  Datasets in which every domain occurs at most once per relation, so
that the reference's conditional is exact:
  * animals.unary as ONE IRM (cxx/assets/animals.unary.{schema,obs}: 50 animals x 85 unary relations);
  * a synthetic `bernoulli R a b` + `bernoulli U a` with planted blocks.
  * animals.unary as an HIRM (config 1), through oracle/_ref/ref_chain.
Writes tests/golden/chain_stats.json.gz (observations included: /root/reference does not exist on the GPU box).

    python tests/golden/make_chain_stats.py
"""
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HIRM_OUT = os.path.join(ROOT, "oracle", "_ref", "hirm.out")
SEEDS, ITERS = 64, 40


def read_dataset(base):
    """(.schema, .obs) -> domain names, relations [(name, [domains])], coded observations (first-appearance codes per
    domain, cxx/util_io.cc:63-96)."""
    rels = []
    for line in open(base + ".schema"):
        f = line.split()
        if f:
            assert f[0] == "bernoulli"
            rels.append((f[1], f[2:]))
    rdom = dict(rels)
    codes, obs = {}, {r: [] for r, _ in rels}
    for line in open(base + ".obs"):
        f = line.split()
        if not f:
            continue
        v, r, items = int(float(f[0])), f[1], f[2:]
        row = []
        for d, it in zip(rdom[r], items):
            c = codes.setdefault(d, {})
            row.append(c.setdefault(it, len(c)))
        obs[r].append(row + [v])
    return {d: len(c) for d, c in codes.items()}, rels, obs


def run_reference(base, n_moves_per_iter, n_domains):
    scores = []
    for seed in range(1, SEEDS + 1):
        out = subprocess.check_output([HIRM_OUT, "--mode", "irm", "--seed", str(seed), "--iters", str(ITERS), "--verbose", base],
                                      text=True)
        vals = [float(l.split()[1]) for l in out.splitlines() if len(l.split()) == 2 and l[0].isdigit()]
        per = n_moves_per_iter + n_domains
        assert len(vals) == per * ITERS, (len(vals), per, ITERS)
        scores.append([vals[(k + 1) * per - 1] for k in range(ITERS)])
    return scores


def synth(tmp):
    rng = np.random.default_rng(7)
    na, nb = 30, 24
    za, zb = rng.integers(0, 3, na), rng.integers(0, 2, nb)
    theta = np.array([[0.9, 0.1], [0.2, 0.8], [0.5, 0.05]])
    tu = np.array([0.85, 0.1, 0.5])
    base = os.path.join(tmp, "synth_ab")
    with open(base + ".schema", "w") as f:
        f.write("bernoulli R a b\nbernoulli U a\n")
    with open(base + ".obs", "w") as f:
        for i in range(na):
            f.write("%d U a%d\n" % (rng.random() < tu[za[i]], i))
            for j in range(nb):
                if rng.random() < 0.8:
                    f.write("%d R a%d b%d\n" % (rng.random() < theta[za[i], zb[j]], i, j))
    return base


def main():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    tmp = tempfile.mkdtemp()
    out = {"seeds": SEEDS, "iters": ITERS, "generator": "tests/golden/make_chain_stats.py", "datasets": {}}
    for name, base in (("animals_unary_irm", None), ("synth_ab", None)):
        if name == "animals_unary_irm":
            for ext in (".schema", ".obs"):
                shutil.copy("/root/reference/cxx/assets/animals.unary" + ext, os.path.join(tmp, "animals.unary" + ext))
            base = os.path.join(tmp, "animals.unary")
        else:
            base = synth(tmp)
        n_ent, rels, obs = read_dataset(base)
        scores = run_reference(base, sum(n_ent.values()), len(n_ent))
        out["datasets"][name] = {"domains": n_ent, "relations": [{"name": r, "domains": ds, "obs": obs[r]} for r, ds in rels],
                                 "ref_logp_score": scores}
        s = np.asarray(scores)[:, ITERS // 2:]
        print(name, "reference logp_score after burn-in: mean %.2f, sd of chain means %.2f" % (s.mean(), s.mean(axis=1).std()))
    # HIRM chains: oracle/_ref/ref_chain (oracle/ref_chain.cc) runs the reference's own inference_hirm iteration and
    # prints HIRM::logp_score and the number of IRMs at the end of every iteration
    base = os.path.join(tmp, "animals.unary")
    n_ent, rels, obs = read_dataset(base)
    hs, hk = [], []
    for seed in range(1, SEEDS + 1):
        o = subprocess.check_output([os.path.join(ROOT, "oracle", "_ref", "ref_chain"), base, str(seed), str(ITERS)], text=True)
        rows = [l.split() for l in o.splitlines() if l.startswith("iter")]
        assert len(rows) == ITERS
        hs.append([float(r[2]) for r in rows]); hk.append([int(r[3]) for r in rows])
    out["datasets"]["animals_unary_hirm"] = {"domains": n_ent, "relations": [{"name": r, "domains": ds, "obs": obs[r]} for r, ds in rels],
                                             "ref_logp_score": hs, "ref_n_irms": hk}
    s = np.asarray(hs)[:, ITERS // 2:]
    print("animals_unary_hirm reference HIRM logp_score after burn-in: mean %.2f, sd of chain means %.2f, IRMs %.2f"
          % (s.mean(), s.mean(axis=1).std(), np.asarray(hk)[:, ITERS // 2:].mean()))
    with gzip.open(os.path.join(ROOT, "tests", "golden", "chain_stats.json.gz"), "wt") as f:
        json.dump(out, f)
    shutil.rmtree(tmp)


if __name__ == "__main__":
    sys.exit(main())
