"""Test infrastructure for hirm.sharded.ShardedHIRM: a CPU backend made of the plain-C oracle (entity moves, alpha
moves, IRM::logp_score) and numpy restatements of the relation-score kernels, plus a synthetic HIRM generator.
The backend mirrors hirm.sharded.EngineBackend call for call, so that (1) the host logic of the sharded HIRM can run
under gloo without a GPU and (2) on a GPU the engine can be compared with it move by move (same Philox positions =>
same assignments)."""
import math

import numpy as np

from hirm.sharded import NGRID_ALPHA, host_sweep_order
from oracle import hirm_oracle as O

ALPHA_TAG = 0x616C706861      # aux_kernels.cuh ALPHA_SEED_TAG
PRIOR_TAG = 0x7072696F72      # aux_kernels.cuh PRIOR_SEED_TAG
MAX_DOMAINS = 8
_LG = np.vectorize(math.lgamma, otypes=[np.float64])


def relation_score_numpy(ids, val, zs, kcap=None):
    """Relation::logp_score (cxx/hirm.hh:516-522) of COO observations under per-position label arrays zs."""
    if len(val) == 0:
        return 0.0
    ids = np.asarray(ids).reshape(len(val), -1)
    key = np.zeros(len(val), dtype=np.int64)
    for p in range(ids.shape[1]):
        _, inv = np.unique(zs[p], return_inverse=True)
        key = key * (inv.max() + 1) + inv[ids[:, p]]
    cells, inv = np.unique(key, return_inverse=True)
    N = np.bincount(inv, minlength=len(cells)).astype(np.float64)
    s = np.bincount(inv, weights=np.asarray(val, dtype=np.float64), minlength=len(cells))
    return float(np.sum(_LG(s + 1) + _LG(N - s + 1) - _LG(N + 2)))


def crp_prior_restated(n, alpha, seed, key):
    """aux_kernels.cuh crp_prior_kernel, sequentially: customer i opens a table when u (i + alpha) < alpha, else joins the
    table of customer floor(v i); labels in order of appearance."""
    z, tables = np.zeros(n, np.int32), 0
    for i in range(n):
        u = O.uniform(seed ^ PRIOR_TAG, key, 2 * i)
        v = O.uniform(seed ^ PRIOR_TAG, key, 2 * i + 1)
        if i == 0 or u * (i + alpha) < alpha:
            z[i] = tables
            tables += 1
        else:
            z[i] = z[min(int(v * i), i - 1)]
    return z


class OracleBackend:
    def __init__(self, n_entities, rel_domains, observations, max_tables, seed):
        self.n_entities, self.rel_domains = [int(n) for n in n_entities], [list(map(int, ds)) for ds in rel_domains]
        self.obs = [(np.ascontiguousarray(i, np.int32).reshape(-1, len(ds)), np.ascontiguousarray(v, np.uint8))
                    for ds, (i, v) in zip(self.rel_domains, observations)]
        self.max_tables, self.seed = max_tables, int(seed)
        self.irms, self._next = {}, 0

    def create_irm(self, rels, z, alpha, uid, steps):
        o = O.Oracle(self.n_entities, [self.rel_domains[r] for r in rels])
        for k, r in enumerate(rels):
            o.set_observations(k, *self.obs[r])
        for d in range(len(self.n_entities)):
            o.set_assignments(d, z[d])
            o.set_alpha(d, alpha[d])
        h = self._next
        self._next += 1
        self.irms[h] = dict(o=o, uid=uid, steps=steps, alpha=list(alpha), rels=list(rels))
        return h

    def destroy_irm(self, h):
        del self.irms[h]

    def get_state(self, h):
        m = self.irms[h]
        return [m["o"].get_assignments(d) for d in range(len(self.n_entities))], list(m["alpha"])

    def get_counts(self, h, k):
        return self.irms[h]["o"].get_counts(k)

    def step(self, handles):
        per = sum(self.n_entities)
        for h in handles:
            m = self.irms[h]
            md, mi = host_sweep_order(self.n_entities, self.seed, m["uid"], m["steps"])
            m["o"].sweep(md, mi, seed=self.seed, stream=m["uid"], offset=m["steps"] * per)
            for d in range(len(self.n_entities)):
                u = O.uniform(self.seed ^ ALPHA_TAG, m["uid"], m["steps"] * MAX_DOMAINS + d)
                m["alpha"][d] = m["o"].transition_alpha(d, u, NGRID_ALPHA)[0]
            m["steps"] += 1

    def logp_scores(self, handles):
        return np.array([self.irms[h]["o"].logp_score() for h in handles])

    def score_matrix(self, rels, handles):
        out = np.zeros((len(rels), len(handles)))
        for j, h in enumerate(handles):
            z = [self.irms[h]["o"].get_assignments(d) for d in range(len(self.n_entities))]
            for i, r in enumerate(rels):
                out[i, j] = relation_score_numpy(*self.obs[r], [z[d] for d in self.rel_domains[r]])
        return out

    def scores_fresh(self, rels, keys):
        out = np.zeros(len(rels))
        for i, r in enumerate(rels):
            z = {d: crp_prior_restated(self.n_entities[d], 1.0, self.seed, int(keys[i][d])) for d in set(self.rel_domains[r])}
            out[i] = relation_score_numpy(*self.obs[r], [z[d] for d in self.rel_domains[r]])
        return out

    def prior_draw(self, d, key):
        return crp_prior_restated(self.n_entities[d], 1.0, self.seed, int(key))


def synthetic_hirm(n_ent=(60, 40), n_unary=10, n_binary=4, n_clusters=3, seed=0, density=0.6, tables=3):
    """Relations in `n_clusters` planted relation clusters, each cluster with its own planted partition of every domain;
    unary relations over domain 0 (fully observed), binary ones over (d0, d1) and (d0, d0), a random subset observed."""
    rng = np.random.default_rng(seed)
    domains = {"d%d" % k: n for k, n in enumerate(n_ent)}
    zc = [[rng.integers(0, tables, n) for n in n_ent] for _ in range(n_clusters)]
    schema, obs, truth = {}, {}, {}
    for j in range(n_unary + n_binary):
        c = j % n_clusters
        if j < n_unary:
            name, doms = "u%02d" % j, [0]
            ids = np.arange(n_ent[0], dtype=np.int32)[:, None]
        else:
            doms = [0, 1] if (j - n_unary) % 2 == 0 and len(n_ent) > 1 else [0, 0]
            name = "b%02d" % j
            grid = np.stack(np.meshgrid(np.arange(n_ent[doms[0]]), np.arange(n_ent[doms[1]]), indexing="ij"), -1).reshape(-1, 2)
            ids = grid[rng.random(len(grid)) < density].astype(np.int32)
        theta = rng.beta(0.3, 0.3, size=[tables] * len(doms))
        p = theta[tuple(zc[c][d][ids[:, k]] for k, d in enumerate(doms))]
        schema[name] = ["d%d" % d for d in doms]
        obs[name] = (ids, (rng.random(len(ids)) < p).astype(np.uint8))
        truth[name] = c
    return domains, schema, obs, truth
