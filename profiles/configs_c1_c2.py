"""BASELINE configs 1 and 2 end to end through the facade, with the unmodified reference CLI timed beside it on the
same host (oracle/_ref/hirm.out; the datasets are rebuilt from the committed golden fixtures, so the reference
tree itself is not needed).   python profiles/configs_c1_c2.py > gpurun_out/configs_c1_c2.txt"""
import os
import random
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hierarchical-irm_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import golden_util as G
from hirm import HIRM, IRM

HIRM_OUT = os.path.join(ROOT, "oracle", "_ref", "hirm.out")


def dataset(case):
    fx = G.load(case)
    rels = {}
    for irm in fx["irms"]:
        for r in irm["relations"]:
            a = np.asarray(r["obs"], dtype=np.int64).reshape(-1, len(r["domains"]) + 1)
            rels[r["name"]] = (list(r["domains"]), a)
    return rels


def write_files(rels, base):
    with open(base + ".schema", "w") as f:
        for r, (ds, _) in sorted(rels.items()):
            f.write("bernoulli %s %s\n" % (r, " ".join(ds)))
    with open(base + ".obs", "w") as f:
        for r, (ds, a) in sorted(rels.items()):
            for row in a:
                f.write("%d %s %s\n" % (row[-1], r, " ".join("%s%d" % (d[0], x) for d, x in zip(ds, row[:-1]))))


def time_reference(rels, mode, iters):
    if not os.path.exists(HIRM_OUT):
        return None
    with tempfile.TemporaryDirectory() as tmp:
        base = os.path.join(tmp, "data")
        write_files(rels, base)
        t0 = time.perf_counter()
        subprocess.check_call([HIRM_OUT, "--mode", mode, "--iters", str(iters), "--seed", "10", base], stdout=subprocess.DEVNULL)
        return time.perf_counter() - t0


def main():
    iters = 10
    # ---- config 1: animals.unary HIRM (50 animals x 85 unary relations) ----
    rels = dataset("animals_unary_hirm")
    h = HIRM({r: ds for r, (ds, _) in rels.items()}, prng=random.Random(10))
    for r, (ds, a) in rels.items():
        for row in a:
            h.incorporate(r, tuple(int(x) for x in row[:-1]), int(row[-1]))
    h.transition_irms()                                   # builds the device state (not timed: the reference's load is not either)
    t0 = time.perf_counter()
    entity_moves = relation_moves = 0
    for _ in range(iters):
        h.transition_cluster_assignments(); relation_moves += len(h.schema)
        h.transition_irms(); entity_moves += sum(len(d.items) for irm in h.irms.values() for d in irm.domains.values())
    dt = time.perf_counter() - t0
    ref = time_reference(rels, "hirm", iters)
    print("C1 animals.unary HIRM, %d iterations: engine %.3f s (%d entity moves, %d relation moves, %d IRMs, logp_score %.2f); "
          "reference hirm.out (1 core, whole run incl. load) %s" % (iters, dt, entity_moves, relation_moves, len(h.irms), h.logp_score(),
                                                                   "%.3f s" % ref if ref else "n/a"))
    # the same model on the sharded-HIRM path (one score matrix per relation sweep), one chain and 64 chains in one launch
    from hirm.sharded import ChainSet
    domains = {"animal": 1 + max(int(a[:, :-1].max()) for _, a in rels.values())}
    schema = {r: ds for r, (ds, _) in rels.items()}
    obs = {r: (a[:, :-1].astype(np.int32), a[:, -1].astype(np.uint8)) for r, (ds, a) in rels.items()}
    for n_chains in (1, 64):
        cs = ChainSet(domains, schema, obs, n_chains, seed=10, max_tables=48)
        cs.iterate(1)
        t0 = time.perf_counter()
        cs.iterate(iters)
        dt = time.perf_counter() - t0
        print("C1 animals.unary HIRM on hirm.sharded, %d chain(s), %d iterations: %.3f s = %.1f ms per chain-iteration (%.1f IRMs per chain)"
              % (n_chains, iters, dt, 1e3 * dt / (iters * n_chains), np.mean([len(h.irms) for h in cs.chains])))
    # ---- config 2: nations.binary IRM (14 countries x 56 predicates x 111 features) ----
    rels = dataset("nations_binary_irm")
    irm = IRM({r: ds for r, (ds, _) in rels.items()}, prng=random.Random(12))
    for r, (ds, a) in rels.items():
        for row in a:
            irm.incorporate(r, tuple(int(x) for x in row[:-1]), int(row[-1]))
    irm.transition_cluster_assignments()
    t0 = time.perf_counter()
    moves = 0
    for _ in range(iters):
        irm.transition_cluster_assignments(); moves += sum(len(d.items) for d in irm.domains.values())
        irm.transition_crp_alphas()
    dt = time.perf_counter() - t0
    ref = time_reference(rels, "irm", iters)
    print("C2 nations.binary IRM, %d sweeps: engine %.3f s (%d entity moves = %.0f moves/s, logp_score %.2f); "
          "reference hirm.out (1 core, whole run incl. load) %s" % (iters, dt, moves, moves / dt, irm.logp_score(), "%.3f s" % ref if ref else "n/a"))


if __name__ == "__main__":
    main()
