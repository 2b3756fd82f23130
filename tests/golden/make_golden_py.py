"""Regenerate tests/golden/nations_binary_py.json.gz from the UNMODIFIED Python backend of the reference.

BASELINE.json config 2: "nations binary IRM (examples/nations_binary_irm.py: 14 countries x 56 binary relations),
checked step-by-step against the Python backend".  This script stages /root/reference/src as the package `hirm` in a
temporary directory (nothing is copied into the repository), imports it, and drives the reference's OWN public members
one entity move at a time (SURVEY.md 8c recipe):

    Domain.tables_weights_gibbs           (src/hirm.py:150-152 -> CRP.tables_weights_gibbs :101-108)
    Relation.logp_gibbs_exact             (src/hirm.py:293-309)
    Relation.set_cluster_assignment_gibbs (src/hirm.py:352-369)
    Domain.set_cluster_assignment_gibbs   (src/hirm.py:153-157)
    Relation.logp_score / CRP.logp_score  (src/hirm.py:351, :92-99)
    CRP.transition_alpha's grid           (src/hirm.py:109-116; src/util_math.py:14-17: 30 points)

exactly as IRM.transition_cluster_assignment_item (src/hirm.py:401-421) strings them together, in an explicit
(domain, entity) order -- IRM.transition_cluster_assignments iterates a Python set of strings, whose order depends on
PYTHONHASHSEED, so the order is fixed here: domains in sorted order, entities by first-appearance code.

The model is built the way examples/nations_binary_irm.py:26-33 builds it: `prng = LoggedRandom(12)` (a random.Random
subclass that logs every random() it hands out), `IRM(schema, prng=prng)`, `incorporate` for every line of the .obs file
-- so the initial seating consumes the variates of random.Random(12) exactly as the example does.  Every move then takes
ONE variate from the same generator (as log_choices -> Random.choices would, src/util_math.py:24-28) and is recorded with

  labels   candidate tables, ascending (the engine's order)
  ref_lp   log(crp weight) + sum of Relation.logp_gibbs_exact over the relations with an observation of the entity:
           the Python backend's own candidate log-probabilities (exact where the domain occurs once per relation, Q1)
  bf_lp    brute force from the backend's own scoring: really re-seat the entity, difference Relation.logp_score
  u        the logged variate;   choice: the engine's inversion rule on bf_lp, then forced through
           set_cluster_assignment_gibbs
  py_choice  what Random.choices' own rule (bisect_right over cumulative weights in dict order) picks from ref_lp and
           the same u -- recorded for reference, not asserted (candidate order and inversion rule differ by design)

plus a state checkpoint (assignments, count tables, logp_score) at the start and after every sweep, and one
CRP.transition_alpha trace per domain and sweep (30-point grid, per-point CRP.logp_score, same-u choice).
The output has the format of the C++ traces (tests/golden_util.py replays both).

    PYTHONHASHSEED=0 python tests/golden/make_golden_py.py
"""
import bisect
import gzip
import itertools
import json
import math
import os
import random
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("HIRM_REFERENCE", "/root/reference")
SWEEPS = 10          # examples/nations_binary_irm.py:32
SEED = 12            # examples/nations_binary_irm.py:27


def stage_reference(tmp):
    """/root/reference/src -> <tmp>/hirm (the layout setup.py installs); plotting is not imported."""
    pkg = os.path.join(tmp, "hirm")
    shutil.copytree(os.path.join(REF, "src"), pkg)
    sys.path.insert(0, tmp)


class LoggedRandom(random.Random):
    """random.Random(12) that remembers every variate it hands out (`Random` subclass logging random())."""

    def __init__(self, seed):
        super().__init__(seed)
        self.log = []

    def random(self):
        u = super().random()
        self.log.append(u)
        return u


def engine_choose(lp, u):
    """DESIGN.md sampling rule: first index whose running sum of exp(lp - logsumexp) exceeds u * total."""
    m = max(lp)
    Z = m + math.log(sum(math.exp(x - m) for x in lp))
    p = [math.exp(x - Z) for x in lp]
    total = 0.0
    for v in p:
        total += v
    x, c = u * total, 0.0
    for i, v in enumerate(p):
        c += v
        if c > x:
            return i
    return len(p) - 1


def python_choose(population, lp, u):
    """Random.choices(population, weights) with one variate u (CPython Lib/random.py): bisect over cumulative weights."""
    m = max(lp)
    Z = m + math.log(sum(math.exp(x - m) for x in lp))
    cum = list(itertools.accumulate(math.exp(x - Z) for x in lp))
    return population[bisect.bisect(cum, u * cum[-1], 0, len(population) - 1)]


def main():
    with tempfile.TemporaryDirectory() as tmp:
        stage_reference(tmp)
        from hirm import IRM
        from hirm.util_io import load_observations, load_schema
        from hirm.util_math import log_linspace

        base = os.path.join(REF, "examples", "datasets", "nations.binary")
        schema = load_schema(base + ".schema")
        data = load_observations(base + ".obs")
        prng = LoggedRandom(SEED)
        irm = IRM(schema, prng=prng)
        for relation, items, value in data:
            irm.incorporate(relation, items, value)
        n_seat_draws = len(prng.log)

        # entity codes by first appearance in the .obs file (cxx/util_io.cc:63-96: what the engine's ingest makes)
        codes = {d: {} for d in irm.domains}
        for relation, items, value in data:
            for d, it in zip(schema[relation], items):
                codes[d].setdefault(it, len(codes[d]))
        dnames = sorted(irm.domains)
        rnames = list(schema)

        def state():
            z = {}
            for d in dnames:
                arr = [-1] * len(codes[d])
                for it, t in irm.domains[d].crp.assignments.items():
                    arr[codes[d][it]] = t
                z[d] = arr
            counts = {}
            for r in rnames:
                cells = sorted((list(k) + [int(c.N), int(c.s)]) for k, c in irm.relations[r].clusters.items())
                counts[r] = [v for row in cells for v in row]
            return dict(z=z, counts=counts, logp_score=float(irm.logp_score()),
                        alpha={d: float(irm.domains[d].crp.alpha) for d in dnames})

        def relation_scores(d):
            return sum(float(irm.relations[r].logp_score()) for r in irm.domain_to_relations[d])

        def reseat(d, item, table):
            domain = irm.domains[d]
            if domain.get_cluster_assignment(item) == table:
                return
            for r in irm.domain_to_relations[d]:
                rel = irm.relations[r]
                if rel.has_observation(domain, item):
                    rel.set_cluster_assignment_gibbs(domain, item, table)
            domain.set_cluster_assignment_gibbs(item, table)

        moves = []
        for sweep in range(SWEEPS):
            for d in dnames:
                domain = irm.domains[d]
                repeated = any(list(schema[r]).count(d) > 1 for r in irm.domain_to_relations[d])
                for item in sorted(codes[d], key=codes[d].get):
                    cur = domain.get_cluster_assignment(item)
                    crp_dist = domain.tables_weights_gibbs(item)              # the backend's own candidates
                    py_tables = list(crp_dist)
                    py_lp = [math.log(crp_dist[t]) for t in py_tables]
                    for r in irm.domain_to_relations[d]:                       # src/hirm.py:408-412
                        rel = irm.relations[r]
                        if rel.has_observation(domain, item):
                            py_lp = [x + y for x, y in zip(py_lp, rel.logp_gibbs_exact(domain, item, py_tables))]
                    labels = sorted(py_tables)
                    ref_lp = [py_lp[py_tables.index(t)] for t in labels]
                    base_score = relation_scores(d)
                    bf_lp = []
                    for t in labels:                                           # brute force: re-seat, difference logp_score
                        reseat(d, item, t)
                        bf_lp.append(math.log(crp_dist[t]) + relation_scores(d) - base_score)
                        reseat(d, item, cur)
                    u = prng.random()                                          # the variate log_choices would consume
                    choice = labels[engine_choose(bf_lp, u)]
                    moves.append(dict(d=d, i=codes[d][item], cur=cur, alpha=float(domain.crp.alpha), repeated=repeated,
                                      labels=labels, ref_lp=ref_lp, bf_lp=bf_lp, u=u, choice=choice,
                                      py_choice=python_choose(py_tables, py_lp, u)))
                    reseat(d, item, choice)
            for d in dnames:                                                   # CRP.transition_alpha (src/hirm.py:109-116)
                crp = irm.domains[d].crp
                grid = log_linspace(1 / crp.N, crp.N + 1, 30)
                keep = crp.alpha
                lp = []
                for g in grid:
                    crp.alpha = g
                    lp.append(float(crp.logp_score()))
                crp.alpha = keep
                u = prng.random()
                crp.alpha = grid[engine_choose(lp, u)]
                moves.append(dict(amove=d, ngrid=30, u=u, lp=lp, alpha=float(crp.alpha)))
            moves.append(dict(ckpt=sweep, state=state()))

        # the initial state must be captured before the moves: rebuild it from a second, identical run
        prng2 = LoggedRandom(SEED)
        irm2 = IRM(schema, prng=prng2)
        for relation, items, value in data:
            irm2.incorporate(relation, items, value)
        assert prng2.log == prng.log[:n_seat_draws]
        irm_moved, irm = irm, irm2
        init = state()
        irm = irm_moved

        relations = []
        for r in rnames:
            obs = []
            for items, v in irm.relations[r].data.items():
                obs += [codes[d][it] for d, it in zip(schema[r], items)] + [int(v)]
            relations.append(dict(name=r, domains=list(schema[r]), obs=obs))
        out = dict(mode="irm", backend="python (src/hirm.py)", seed=SEED, sweeps=SWEEPS, seat_draws=n_seat_draws,
                   pythonhashseed=os.environ.get("PYTHONHASHSEED"),
                   irms=[dict(domains={d: len(codes[d]) for d in dnames}, relations=relations, init=init, moves=moves)])
        path = os.path.join(HERE, "nations_binary_py.json.gz")
        raw = json.dumps(out, separators=(",", ":")).encode()
        with gzip.GzipFile(path, "wb", mtime=0) as f:
            f.write(raw)
        n_moves = sum(1 for m in moves if "d" in m)
        print("nations_binary_py  %d bytes (%d raw), %d moves, %d changed, %d seat draws"
              % (os.path.getsize(path), len(raw), n_moves, sum(1 for m in moves if "d" in m and m["choice"] != m["cur"]), n_seat_draws))


if __name__ == "__main__":
    sys.exit(main())
