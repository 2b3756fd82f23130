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
  * HIRM (src/hirm.py:457-571) keeps the relation CRP on the host and gets every candidate table's data term
    from the device (hirm_engine_relation_scores).
  * Relation.logp / IRM.logp / HIRM.logp (predictive queries, src/hirm.py:310-350,428-430,553-562,573-628) are small host
    enumerations over the read-back count tables: off the hot path, kept so that the reference's own tests and
    examples run on the facade.
Out of scope: plotting.
"""
import itertools
import math
import random

import numpy as np

from . import engine as E

_ENGINES = {}


def default_engine(device=-1):
    if device not in _ENGINES:
        _ENGINES[device] = E.Engine(device)
    return _ENGINES[device]


class BetaBernoulli:
    """Read-back view of one count-table cell (src/hirm.py:17-53): alpha, beta, N, s and the two closed forms."""

    def __init__(self, alpha=1, beta=1, prng=None, N=0, s=0):
        self.alpha, self.beta, self.N, self.s = alpha, beta, int(N), int(s)
        self.prng = prng or random

    def logp(self, x):
        """Predictive log-probability of one more observation (src/hirm.py:35-42); -inf for a value outside {0, 1}."""
        if x not in (0, 1):
            return -math.inf
        num = self.s + self.alpha if x == 1 else self.N - self.s + self.beta
        return math.log(num) - math.log(self.N + self.alpha + self.beta)

    def logp_score(self):
        """log marginal likelihood of the cell (src/hirm.py:43-46): lbeta(s + alpha, N - s + beta) - lbeta(alpha, beta)."""
        def lbeta(a, b):
            return math.lgamma(a) + math.lgamma(b) - math.lgamma(a + b)
        return lbeta(self.s + self.alpha, self.N - self.s + self.beta) - lbeta(self.alpha, self.beta)

    def __repr__(self):
        return "BetaBernoulli(alpha=%f, beta=%f, N=%d, s=%d)" % (self.alpha, self.beta, self.N, self.s)


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

    def tables_weights(self):
        """{table: customers} plus the fresh table max + 1 at weight alpha (src/hirm.py:95-100)."""
        if self.N == 0:
            return {0: 1}
        w = {t: len(c) for t, c in self.tables.items()}
        w[max(w) + 1] = self.alpha
        return w

    def tables_weights_gibbs(self, table):
        """The same with one customer of `table` removed; a singleton keeps its table at weight alpha and no fresh table
        is offered (src/hirm.py:101-108)."""
        assert 0 < self.N
        w = self.tables_weights()
        w[table] -= 1
        if w[table] == 0:
            w[table] = self.alpha
            del w[max(w)]
        return w

    def logp_score(self):
        """src/hirm.py:92-99."""
        sizes = [len(c) for c in self.tables.values()]
        return (len(sizes) * math.log(self.alpha) + sum(math.lgamma(c) for c in sizes) + math.lgamma(self.alpha)
                - math.lgamma(self.N + self.alpha))


class CodeSpace:
    """entity -> int code by first appearance (cxx/util_io.cc:63-96).  One per domain name; an HIRM shares it
    between its IRMs so that a relation's observations can be scored under any IRM's partition."""

    def __init__(self):
        self.codes, self.names = {}, []

    def code(self, item):
        c = self.codes.get(item)
        if c is None:
            c = self.codes[item] = len(self.names)
            self.names.append(item)
        return c


class Domain:
    def __init__(self, name, irm, space=None):
        self.name, self._irm = name, irm
        self.items = set()
        self._space = space or CodeSpace()
        self._z = {}              # entity -> table label (host mirror)
        self.crp = CRP(self)

    @property
    def _codes(self):
        return self._space.codes

    @property
    def _names(self):
        return self._space.names

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
        grew = item not in self._space.codes
        code = self._space.code(item)
        self._z[item] = int(table)
        if irm._h is not None and not irm._dirty and not grew and self.name in irm._dorder and code < irm._n_ent[irm._dorder.index(self.name)]:
            # the entity has a code on the device already (absent so far): seat it there, the count tables stand
            irm._eng.seat_entities(irm._h, irm._dorder.index(self.name), [code], [int(table)])
        else:
            irm._dirty = True

    def unincorporate(self, item):
        raise NotImplementedError()

    def get_cluster_assignment(self, item):
        assert item in self.items
        self._irm._pull()
        return self._z[item]

    def tables_weights(self):
        return self.crp.tables_weights()

    def tables_weights_gibbs(self, item):
        return self.crp.tables_weights_gibbs(self.get_cluster_assignment(item))

    def _fresh_options(self, item):
        """(tables, log-weights) an entity can sit at in a predictive query: its own table if it is seated, else every
        table and the fresh one with weight / (1 + N) (src/hirm.py:322-333,589-599)."""
        if item in self.items:
            return (self.get_cluster_assignment(item),), (0.0,)
        w = self.tables_weights()
        z = math.log(1 + self.crp.N)
        return tuple(w), tuple(math.log(v) - z for v in w.values())

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
        return {tuple(int(x) for x in c[:a]): BetaBernoulli(N=c[a], s=c[a + 1]) for c in cells}

    def logp_score(self):
        from math import lgamma
        return sum(lgamma(c.s + 1) + lgamma(c.N - c.s + 1) - lgamma(c.N + 2) for c in self.clusters.values())

    def get_cluster_assignment(self, items):
        return tuple(d.get_cluster_assignment(i) for d, i in zip(self.domains, items))

    def logp(self, items, value):
        """Predictive log-probability of one observation (src/hirm.py:310-350): unseen entities are summed out over the
        tables of their domain and one fresh table, position by position."""
        assert len(items) == len(self.domains)
        options = [d._fresh_options(i) for d, i in zip(self.domains, items)]
        cells = self.clusters
        empty = BetaBernoulli()
        terms = []
        for pick in itertools.product(*[range(len(t)) for t, _ in options]):
            z = tuple(options[p][0][k] for p, k in enumerate(pick))
            terms.append(sum(options[p][1][k] for p, k in enumerate(pick)) + cells.get(z, empty).logp(value))
        from .util_math import logsumexp
        return logsumexp(terms)

    def __repr__(self):
        return "Relation(name=%s, domains=%r)" % (self.name, self.domains)


class IRM:
    """src/hirm.py:378-455 on the GPU engine.  `max_tables` bounds the tables per domain kept on the device."""

    def __init__(self, schema, prng=None, engine=None, max_tables=None, spaces=None):
        self.schema, self.domains, self.relations, self.domain_to_relations = {}, {}, {}, {}
        self.prng = prng or random
        self._spaces = spaces          # {domain name: CodeSpace} shared with other IRMs (HIRM), or None
        self._dorder, self._n_ent = [], []
        self._eng = engine
        self._max_tables = max_tables
        self._alpha = {}
        self._h = None
        self._dirty = True          # host observations / seats changed since the last upload
        self._stale = False         # device assignments changed since the last read-back
        self._seed = None
        self._moves = 0
        self._sweeps = self._alpha_moves = 0   # Philox positions of the device sweep order / alpha moves: they survive a re-upload
        for relation, domains in schema.items():
            self.add_relation(relation, domains)

    # ---- schema -------------------------------------------------------------------------
    def add_relation(self, r, domains):
        assert r not in self.schema
        assert r not in self.relations
        self._pull()
        for d in domains:
            if d not in self.domains:
                self.domains[d] = Domain(d, self, None if self._spaces is None else self._spaces.setdefault(d, CodeSpace()))
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
            for item in dom.items:
                dom._z[item] = int(lab[dom._codes[item]])
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
        n_ent = self._n_ent = [len(self.domains[d]._names) for d in self._dorder]
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
            self._eng.set_assignments(self._h, k, np.asarray([dom._z.get(it, -1) for it in dom._names], dtype=np.int32))
            self._eng.set_alpha(self._h, k, self._alpha[d])
        if self._seed is None:
            self._seed = self.prng.getrandbits(62)
        self._eng.set_rng(self._h, self._seed & 0xFFFFFFFF, self._moves)
        # a re-created engine IRM starts its sweep / alpha counters at 0: carry the mirror's on, or the next alpha move
        # and sweep order would reuse the variates of the first ones
        self._eng.set_rng_counters(self._h, self._sweeps, self._alpha_moves)
        self._dirty = False

    def _dense_candidate(self):
        """One domain, one binary relation R(e, e) with every cell observed: the dense slab layout."""
        if len(self.relations) != 1 or len(self.domains) != 1:
            return False
        rel = next(iter(self.relations.values()))
        n = len(rel.domains[0]._names)
        return len(rel.domains) == 2 and n >= 2 and len(rel.data) == n * n and len(rel.domains[0].items) == n

    # ---- inference ---------------------------------------------------------------------------
    def transition_cluster_assignments(self, domains=None):
        """One sweep over the given domains (default: all), src/hirm.py:392-400.  The order of the moves is
        engine-defined (the reference's is the iteration order of a Python set)."""
        self._push()
        if domains is None and any(len(self.domains[d].items) < n for d, n in zip(self._dorder, self._n_ent)):
            domains = list(self._dorder)                         # some codes are not seated here: explicit lists
        if domains is None:
            self._moves += self._eng.sweep_all([self._h], n_sweeps=1, seed=self._seed)
            self._sweeps += 1
        else:
            md, mi = [], []
            for d in domains:
                k = self._dorder.index(d)
                order = sorted(self.domains[d]._codes[it] for it in self.domains[d].items)
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
        self._alpha_moves += 1
        for k, d in enumerate(self._dorder):
            if domains is not None and d not in domains:
                self._eng.set_alpha(self._h, k, before[d])
            self._alpha[d] = self._eng.get_alpha(self._h, k)

    def logp_score(self):
        self._push()
        return float(self._eng.logp_scores([self._h])[0])

    def logp(self, observations):
        """Joint predictive log-probability of (relation name, items, value) triples (src/hirm.py:428-430)."""
        return logp_observations([(self.relations[r], tuple(i), v) for r, i, v in observations])


def logp_observations(observations):
    """src/hirm.py:573-628: log p of a list of (Relation, items, value) given the current state.  Every distinct unseen
    (domain, entity) is one latent seat ranging over the domain's tables and a fresh one (weight / (1 + N)); the sum over
    all seatings is enumerated on the host from the read-back count tables."""
    from .util_math import logsumexp
    seats, cells, seen = {}, {}, set()
    for rel, items, value in observations:
        assert (rel.name, items) not in seen
        seen.add((rel.name, items))
        assert len(items) == len(rel.domains)
        for d, it in zip(rel.domains, items):
            seats.setdefault((d.name, it), d._fresh_options(it))
        if rel.name not in cells:
            cells[rel.name] = rel.clusters
    keys = list(seats)
    empty = BetaBernoulli()
    terms = []
    for pick in itertools.product(*[range(len(seats[k][0])) for k in keys]):
        at = {k: seats[k][0][j] for k, j in zip(keys, pick)}
        lp = sum(seats[k][1][j] for k, j in zip(keys, pick))
        for rel, items, value in observations:
            z = tuple(at[(d.name, it)] for d, it in zip(rel.domains, items))
            lp += cells[rel.name].get(z, empty).logp(value)
        terms.append(lp)
    return logsumexp(terms)


def choose_index(logps, u):
    """The engine's inversion rule (cxx/util_math.cc:58-65 / random.Random.choices): first index whose running
    sum of exp(lp - max) exceeds u * total, sums in index order."""
    from math import exp
    m = max(logps)
    e = [exp(float(v) - m) for v in logps]                # plain floats: this runs once per relation move of every chain
    total = 0.0
    for v in e:
        total += v
    x, cum = u * total, 0.0
    for i, v in enumerate(e):
        cum += v
        if cum > x:
            return i
    return len(e) - 1


class RelationCRP:
    """The CRP over relations of an HIRM (src/hirm.py:459; cxx/hirm.hh:790): small, kept on the host."""

    def __init__(self, prng):
        self.alpha, self.prng = 1.0, prng
        self.assignments, self.tables = {}, {}

    @property
    def N(self):
        return len(self.assignments)

    def tables_weights(self):
        if not self.tables:
            return {0: 1.0}
        w = {t: float(len(c)) for t, c in self.tables.items()}
        w[max(self.tables) + 1] = self.alpha
        return w

    def tables_weights_gibbs(self, table):
        """src/hirm.py:101-108: the customer's own table loses one seat; a singleton keeps weight alpha and no
        fresh table is offered."""
        w = self.tables_weights()
        w[table] -= 1
        if w[table] == 0:
            w[table] = self.alpha
            del w[max(w)]
        return w

    def sample(self):
        w = self.tables_weights()
        return self.prng.choices(list(w), weights=list(w.values()))[0]

    def incorporate(self, item, table):
        self.tables.setdefault(table, set()).add(item)
        self.assignments[item] = table

    def unincorporate(self, item):
        t = self.assignments.pop(item)
        self.tables[t].discard(item)
        if not self.tables[t]:
            del self.tables[t]

    def logp_score(self):
        from math import lgamma, log
        return (len(self.tables) * log(self.alpha) + sum(lgamma(len(c)) for c in self.tables.values())
                + lgamma(self.alpha) - lgamma(self.N + self.alpha))


class HIRM:
    """src/hirm.py:457-571 / cxx/hirm.hh:784-1014 on the GPU engine.

    Every relation-cluster IRM is an engine IRM; the entity sweeps of all of them run concurrently in one launch
    (`transition_irms`).  The relation move (`transition_cluster_assignment_relation`) gets the data term of every
    candidate table -- Relation.logp_score of the relation under that IRM's partitions -- from the device
    (hirm_engine_relation_scores) without moving data; entities a candidate IRM has not seen are seated there by
    CRP-prior draws first and stay behind as the reference's do (its remove_relation leaves them in the domain)."""

    def __init__(self, schema, prng=None, engine=None, max_tables=None):
        self.prng = prng or random
        self.crp = RelationCRP(self.prng)
        self.schema, self.irms = {}, {}
        self._eng, self._max_tables = engine, max_tables
        self._spaces = {}
        self._seat_hook = None        # tests: {(table, domain): {entity: label}} -> seats of a fresh IRM
        for r, ds in schema.items():
            self.add_relation(r, ds)

    def _new_irm(self, schema):
        return IRM(schema, prng=self.prng, engine=self._eng, max_tables=self._max_tables, spaces=self._spaces)

    # ---- schema / data --------------------------------------------------------------------
    def add_relation(self, r, domains):
        assert r not in self.schema
        self.schema[r] = domains
        table = self.crp.sample()
        self.crp.incorporate(r, table)
        if table in self.irms:
            self.irms[table].add_relation(r, domains)
        else:
            self.irms[table] = self._new_irm({r: domains})

    def remove_relation(self, r):
        del self.schema[r]
        table = self.crp.assignments[r]
        singleton = len(self.crp.tables[table]) == 1
        self.crp.unincorporate(r)
        self.irms[table].remove_relation(r)
        if singleton:
            assert not self.irms[table].relations
            self._drop(table)

    def _drop(self, table):
        irm = self.irms.pop(table)
        if irm._h is not None:
            irm._eng.irm_destroy(irm._h)
            irm._h = None

    def incorporate(self, r, items, value):
        self.relation_to_irm(r).incorporate(r, items, value)

    def unincorporate(self, r, items):
        raise NotImplementedError()

    def relation_to_table(self, r):
        return self.crp.assignments[r]

    def relation_to_irm(self, r):
        return self.irms[self.crp.assignments[r]]

    def get_relation(self, r):
        return self.relation_to_irm(r).relations[r]

    relation = get_relation                                  # src/hirm.py:476-478 names it `relation`

    def transition_crp_alpha(self):
        """CRP.transition_alpha of the relation CRP (src/hirm.py:544-545): 30-point grid, drawn with the caller's prng."""
        from .util_math import log_choices, log_linspace
        n = self.crp.N
        grid = log_linspace(1 / n, n + 1, 30)
        keep, lps = self.crp.alpha, []
        for g in grid:
            self.crp.alpha = g
            lps.append(self.crp.logp_score())
        self.crp.alpha = keep
        self.crp.alpha = log_choices(grid, lps, prng=self.prng)[0]

    def logp(self, observations):
        """src/hirm.py:553-562: the observations of every relation cluster are scored by its IRM."""
        by_table = {}
        for r, items, v in observations:
            by_table.setdefault(self.crp.assignments[r], []).append((r, items, v))
        return sum(self.irms[t].logp(obs) for t, obs in by_table.items())

    # ---- inference ---------------------------------------------------------------------------
    def transition_cluster_assignments(self, rs=None):
        """Relation moves (src/hirm.py:479-481): every relation once, in a prng-shuffled order."""
        if rs is None:
            rs = list(self.schema)
            self.prng.shuffle(rs)
        for r in rs:
            self.transition_cluster_assignment_relation(r)

    def _candidate_scores(self, r, tables, table_current):
        """Relation.logp_score of r under every candidate table's IRM.  Returns (scores, {table: fresh IRM})."""
        cur = self.irms[table_current]
        rel = cur.relations[r]
        dnames = [d.name for d in rel.domains]
        cur._push()
        self._eng = cur._eng
        cands, fresh = [], {}
        for t in tables:
            if t == table_current:
                cands.append(cur)
                continue
            irm = self.irms.get(t)
            if irm is None:
                irm = fresh[t] = self._new_irm({})
            irm._eng = irm._eng or self._eng
            for d in dnames:                                   # the trial add_relation + incorporate of the reference:
                if d not in irm.domains:                       # it creates the domains and seats the unseen entities
                    irm.domains[d] = Domain(d, irm, self._spaces.setdefault(d, CodeSpace()))
                    irm.domain_to_relations[d] = set()
                    irm._alpha[d] = 1.0
                    irm._dirty = True
            for items in rel.data:
                for d, item in zip(dnames, items):
                    dom = irm.domains[d]
                    if item not in dom.items:
                        forced = None if self._seat_hook is None else self._seat_hook.get((t, d), {}).get(item)
                        dom.incorporate(item, forced)
            irm._push()
            cands.append(irm)
        doms = [[c._dorder.index(d) for d in dnames] for c in cands]
        scores = self._eng.relation_scores(cur._h, cur._rindex[r], [c._h for c in cands], doms)
        return [float(x) for x in scores], fresh

    def transition_cluster_assignment_relation(self, r, u=None):
        """src/hirm.py:482-519 / cxx/hirm.hh:833-906.  Returns (tables, log-probabilities, choice)."""
        from math import log
        table_current = self.crp.assignments[r]
        crp_dist = self.crp.tables_weights_gibbs(table_current)
        tables = sorted(crp_dist)
        scores, fresh = self._candidate_scores(r, tables, table_current)
        logps = [log(crp_dist[t]) + s for t, s in zip(tables, scores)]
        if u is None:
            u = self.prng.random()
        choice = tables[choose_index(logps, u)]
        self._move(r, table_current, choice, fresh)
        return tables, logps, choice

    def _move(self, r, table_current, choice, fresh):
        # candidates that were only tried keep the seats they were given (cxx/hirm.hh:766 TODO);
        # a fresh IRM that was not chosen is dropped (:892-898)
        for t, irm in fresh.items():
            if t != choice:
                if irm._h is not None:
                    irm._eng.irm_destroy(irm._h)
            else:
                self.irms[t] = irm
        if choice != table_current:
            cur = self.irms[table_current]
            rel = cur.relations[r]
            domains, data = [d.name for d in rel.domains], dict(rel.data)
            cur.remove_relation(r)
            if not cur.relations:
                self._drop(table_current)
            dst = self.irms[choice]
            dst.add_relation(r, domains)
            for items, value in data.items():
                dst.incorporate(r, items, value)
        self.crp.unincorporate(r)
        self.crp.incorporate(r, choice)
        assert set(self.irms) == set(self.crp.tables)

    def set_cluster_assignment_gibbs(self, r, table):
        """src/hirm.py:520-543: force relation r into `table`."""
        cur = self.crp.assignments[r]
        fresh = {}
        if table not in self.irms:
            fresh[table] = self._new_irm({})
        self._move(r, cur, table, fresh)

    def transition_irms(self, n_sweeps=1, ngrid=30):
        """The entity sweeps + alpha moves of every IRM (cxx/hirm.cc:73-88), all IRMs concurrently."""
        irms = list(self.irms.values())
        for irm in irms:
            irm._push()
        self._eng = irms[0]._eng
        batch = [irm for irm in irms if all(len(irm.domains[d].items) == n for d, n in zip(irm._dorder, irm._n_ent))]
        for _ in range(n_sweeps):
            if batch:
                seed = batch[0]._seed
                self._eng.sweep_all([irm._h for irm in batch], n_sweeps=1, seed=seed)
                for irm in batch:
                    irm._moves += sum(irm._n_ent)
                    irm._sweeps += 1
                    irm._stale = True
            for irm in irms:
                if irm not in batch:
                    irm.transition_cluster_assignments()
            self._eng.transition_alpha([irm._h for irm in irms], ngrid=ngrid, seed=irms[0]._seed)
            for irm in irms:
                irm._alpha_moves += 1
                irm._stale = True

    def logp_score(self):
        return self.crp.logp_score() + sum(irm.logp_score() for irm in self.irms.values())
