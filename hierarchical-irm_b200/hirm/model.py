"""Host-side mirror of the reference's Python interface for the entity-move path (src/hirm.py):
IRM / Relation / Domain / CRP with the same names, arguments and error behaviour, the state
living on the GPU engine (hirm_engine.h) between calls.

What runs where
  * IRM.incorporate (src/hirm.py:387) buffers the observation on the host and seats unseen entities
    by the sequential CRP-prior draw of Domain.incorporate (:130-136) with the caller's `prng`;
  * the first transition uploads observations + assignments (dense slabs for a single fully
    observed R(e, e) relation, per-entity CSR otherwise) and from then on
    transition_cluster_assignments / transition_cluster_assignment_item / transition_crp_alphas /
    logp_score are engine calls; there is no host implementation of the move;
  * crp.tables / crp.assignments / relation.clusters are read-back views, refreshed lazily.
Out of scope: IRM.logp (predictive), HIRM relation moves, plotting.
"""
import random

import numpy as np

from . import engine as E

_ENGINES = {}


def default_engine(device=-1):
    if device not in _ENGINES:
        _ENGINES[device] = E.Engine(device)
    return _ENGINES[device]


class BetaBernoulli:
    """Read-back view of one count-table cell (src/hirm.py:17-53)."""

    def __init__(self, N=0, s=0):
        self.alpha, self.beta, self.N, self.s = 1, 1, int(N), int(s)

    def __repr__(self):
        return "BetaBernoulli(N=%d, s=%d)" % (self.N, self.s)


class CRP:
    """View of a domain's clustering (src/hirm.py:55-122): alpha, N, tables, assignments."""

    def __init__(self, domain):
        self._d = domain

    @property
    def alpha(self):
        return self._d._irm._alpha[self._d.name]

    @alpha.setter
    def alpha(self, v):
        self._d._irm._set_alpha(self._d.name, float(v))

    @property
    def N(self):
        return len(self._d.items)

    @property
    def assignments(self):
        self._d._irm._pull()
        return dict(self._d._z)

    @property
    def tables(self):
        self._d._irm._pull()
        t = {}
        for item, tab in self._d._z.items():
            t.setdefault(tab, set()).add(item)
        return t

    def transition_alpha(self):
        self._d._irm.transition_crp_alphas([self._d.name])


class Domain:
    def __init__(self, name, irm):
        self.name, self._irm = name, irm
        self.items = set()
        self._codes = {}          # entity -> int code, first appearance order (cxx/util_io.cc:63-96)
        self._names = []
        self._z = {}              # entity -> table label (host mirror)
        self.crp = CRP(self)

    def incorporate(self, item, table=None):
        """Domain.incorporate (src/hirm.py:130-136): an unseen entity is seated by a CRP-prior draw."""
        if item in self.items:
            assert table is None
            return
        irm = self._irm
        irm._pull()
        if table is None:
            sizes = {}
            for t in self._z.values():
                sizes[t] = sizes.get(t, 0) + 1
            if not sizes:
                table = 0                                        # CRP.sample with N == 0 (src/hirm.py:95-97)
            else:
                tabs = list(sizes) + [max(sizes) + 1]
                w = [sizes[t] for t in sizes] + [irm._alpha[self.name]]
                table = irm.prng.choices(tabs, weights=w)[0]
        self.items.add(item)
        self._codes[item] = len(self._names)
        self._names.append(item)
        self._z[item] = int(table)
        irm._dirty = True

    def unincorporate(self, item):
        raise NotImplementedError()

    def get_cluster_assignment(self, item):
        assert item in self.items
        self._irm._pull()
        return self._z[item]

    def __repr__(self):
        return "Domain(name=%r)" % (self.name,)


class Relation:
    def __init__(self, name, domains, irm):
        self.name, self.domains, self._irm = name, tuple(domains), irm
        self.data = {}
        self.data_r = {d.name: {} for d in self.domains}

    def incorporate(self, items, value):
        """Relation.incorporate (src/hirm.py:176-188)."""
        items = tuple(items)
        assert items not in self.data
        assert len(items) == len(self.domains)
        if value not in (0, 1):
            raise ValueError("BetaBernoulli observations must be 0 or 1")
        self.data[items] = value
        for domain, item in zip(self.domains, items):
            domain.incorporate(item)
            self.data_r[domain.name].setdefault(item, set()).add(items)
        self._irm._dirty = True

    def unincorporate(self, items):
        raise NotImplementedError()

    def has_observation(self, domain, item):
        return item in self.data_r[domain.name]

    @property
    def clusters(self):
        """{cluster tuple: BetaBernoulli(N, s)} for the non-empty cells (src/hirm.py:169)."""
        irm = self._irm
        irm._push()
        cells = irm._eng.get_counts(irm._h, irm._rindex[self.name])
        a = len(self.domains)
        return {tuple(int(x) for x in c[:a]): BetaBernoulli(c[a], c[a + 1]) for c in cells}

    def logp_score(self):
        from math import lgamma
        return sum(lgamma(c.s + 1) + lgamma(c.N - c.s + 1) - lgamma(c.N + 2) for c in self.clusters.values())


class IRM:
    """src/hirm.py:378-455 on the GPU engine.  `max_tables` bounds the tables per domain kept on the device."""

    def __init__(self, schema, prng=None, engine=None, max_tables=None):
        self.schema, self.domains, self.relations, self.domain_to_relations = {}, {}, {}, {}
        self.prng = prng or random
        self._eng = engine
        self._max_tables = max_tables
        self._alpha = {}
        self._h = None
        self._dirty = True          # host observations / seats changed since the last upload
        self._stale = False         # device assignments changed since the last read-back
        self._seed = None
        self._moves = 0
        for relation, domains in schema.items():
            self.add_relation(relation, domains)

    # ---- schema -------------------------------------------------------------------------
    def add_relation(self, r, domains):
        assert r not in self.schema
        assert r not in self.relations
        self._pull()
        for d in domains:
            if d not in self.domains:
                self.domains[d] = Domain(d, self)
                self.domain_to_relations[d] = set()
                self._alpha[d] = 1.0
            self.domain_to_relations[d].add(r)
        self.relations[r] = Relation(r, [self.domains[d] for d in domains], self)
        self.schema[r] = domains
        self._dirty = True

    def remove_relation(self, r):
        self._pull()
        for d in {d.name for d in self.relations[r].domains}:
            self.domain_to_relations[d].discard(r)
            if len(self.domain_to_relations[d]) == 0:
                del self.domain_to_relations[d]
                del self.domains[d]
                del self._alpha[d]
        del self.relations[r]
        del self.schema[r]
        self._dirty = True

    def incorporate(self, r, items, value):
        self.relations[r].incorporate(items, value)

    def unincorporate(self, r, items):
        raise NotImplementedError()

    # ---- host <-> device ------------------------------------------------------------------
    def _pull(self):
        """Device assignments -> host mirror."""
        if not self._stale or self._h is None:
            return
        for k, d in enumerate(self._dorder):
            dom = self.domains[d]
            lab = self._eng.get_assignments(self._h, k)
            for code, item in enumerate(dom._names[:len(lab)]):
                dom._z[item] = int(lab[code])
            self._alpha[d] = self._eng.get_alpha(self._h, k)
        self._stale = False

    def _set_alpha(self, d, v):
        self._alpha[d] = v
        if self._h is not None and not self._dirty:
            self._eng.set_alpha(self._h, self._dorder.index(d), v)

    def _push(self):
        """Host observations + seats -> a fresh engine IRM (only after the host side changed)."""
        if not self._dirty and self._h is not None:
            return
        self._pull()
        if self._eng is None:
            self._eng = default_engine()
        if self._h is not None:
            self._eng.irm_destroy(self._h)
        self._dorder = list(self.domains)
        self._rorder = list(self.relations)
        self._rindex = {r: k for k, r in enumerate(self._rorder)}
        dindex = {d: k for k, d in enumerate(self._dorder)}
        n_ent = [len(self.domains[d]._names) for d in self._dorder]
        rel_domains = [[dindex[d.name] for d in self.relations[r].domains] for r in self._rorder]
        dense = self._dense_candidate()
        kcap = []
        for d in self._dorder:
            k_now = len(set(self.domains[d]._z.values()))
            cap = self._max_tables or max(32, 4 * k_now + 8)
            kcap.append(int(min(cap, 64 if dense else 255)))
        self._h = self._eng.irm_create(n_ent, rel_domains, max_tables=kcap)
        for k, r in enumerate(self._rorder):
            rel = self.relations[r]
            if dense:
                n = n_ent[0]
                x = np.zeros((n, n), dtype=np.uint8)
                codes = rel.domains[0]._codes
                for (i, j), v in rel.data.items():
                    x[codes[i], codes[j]] = v
                self._eng.set_observations_dense_host(self._h, k, x)
            else:
                ids = np.asarray([[dom._codes[it] for dom, it in zip(rel.domains, items)] for items in rel.data],
                                 dtype=np.int32).reshape(-1, len(rel.domains))
                val = np.asarray(list(rel.data.values()), dtype=np.uint8)
                self._eng.set_observations(self._h, k, ids, val)
        self._eng.finalize(self._h)
        for k, d in enumerate(self._dorder):
            dom = self.domains[d]
            self._eng.set_assignments(self._h, k, np.asarray([dom._z[it] for it in dom._names], dtype=np.int32))
            self._eng.set_alpha(self._h, k, self._alpha[d])
        if self._seed is None:
            self._seed = self.prng.getrandbits(62)
        self._eng.set_rng(self._h, self._seed & 0xFFFFFFFF, self._moves)
        self._dirty = False

    def _dense_candidate(self):
        """One domain, one binary relation R(e, e) with every cell observed: the dense slab layout."""
        if len(self.relations) != 1 or len(self.domains) != 1:
            return False
        rel = next(iter(self.relations.values()))
        n = len(rel.domains[0]._names)
        return len(rel.domains) == 2 and n >= 2 and len(rel.data) == n * n

    # ---- inference ---------------------------------------------------------------------------
    def transition_cluster_assignments(self, domains=None):
        """One sweep over the given domains (default: all), src/hirm.py:392-400.  The order of the moves is
        engine-defined (the reference's is the iteration order of a Python set)."""
        self._push()
        if domains is None:
            self._moves += self._eng.sweep_all([self._h], n_sweeps=1, seed=self._seed)
        else:
            md, mi = [], []
            for d in domains:
                k = self._dorder.index(d)
                order = list(range(len(self.domains[d]._names)))
                self.prng.shuffle(order)
                md += [k] * len(order)
                mi += order
            self._explicit(md, mi)
        self._stale = True

    def transition_cluster_assignment_item(self, d, item):
        """One move (src/hirm.py:401-421)."""
        self._push()
        self._explicit([self._dorder.index(d)], [self.domains[d]._codes[item]])
        self._stale = True

    def _explicit(self, md, mi):
        md, mi = np.asarray(md, dtype=np.int32), np.asarray(mi, dtype=np.int32)
        self._eng.sweep([self._h], [(md, mi)], seed=self._seed, streams=[self._seed & 0xFFFFFFFF], counters=[self._moves])
        self._moves += len(mi)
        self._eng.set_rng(self._h, self._seed & 0xFFFFFFFF, self._moves)

    def transition_crp_alphas(self, domains=None, ngrid=30):
        """CRP.transition_alpha for the given domains (src/hirm.py:109-116, :422-427; 30 grid points in the
        Python backend, 20 in cxx/hirm.hh:164)."""
        self._push()
        before = {d: self._eng.get_alpha(self._h, k) for k, d in enumerate(self._dorder)}
        self._eng.transition_alpha([self._h], ngrid=ngrid, seed=self._seed)
        for k, d in enumerate(self._dorder):
            if domains is not None and d not in domains:
                self._eng.set_alpha(self._h, k, before[d])
            self._alpha[d] = self._eng.get_alpha(self._h, k)

    def logp_score(self):
        self._push()
        return float(self._eng.logp_scores([self._h])[0])

    def logp(self, observations):
        raise NotImplementedError("predictive queries are outside the engine's scope (SURVEY.md section 2)")
