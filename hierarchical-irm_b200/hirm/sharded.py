"""HIRM (cxx/hirm.hh:784-1014, src/hirm.py:457-571) with its relation-cluster IRMs sharded over GPUs.

In an HIRM every relation cluster owns a private IRM (private domain partitions, cxx/hirm.hh:749-757,787), so the
entity sweeps of different IRMs touch disjoint state (cxx/hirm.cc:73-88): one process per GPU, every rank owns a subset
of the IRMs and sweeps them in one launch, no collective on that path.  The one exchange step of an HIRM iteration is
the relation move HIRM::transition_cluster_assignment_relation (:833-906).  Its data term for relation r and candidate
table k -- Relation::logp_score of r after re-incorporating its data into IRM k (:859-866) -- depends only on r's
observations and on k's domain partitions, never on which other relations sit in k.  So per iteration

  1. every rank computes the [relations x own IRMs] score matrix in one batched launch (hirm_engine_relation_score_matrix)
     plus the fresh-table score of its share of the relations (hirm_engine_relation_scores_fresh);
  2. ONE all-gather (NCCL over NVLink; gloo in the CPU tests) makes the full matrix known everywhere (R x K doubles);
  3. the sequential CRP re-seating of the relations is O(R K) scalar work, replicated on every rank from a shared seed;
     a relation that opens a fresh IRM adds one column for the relations still to move (one broadcast, rare);
  4. IRMs whose relation set changed are rebuilt on their owner (the relations' COO arrays are replicated in HBM on every
     rank, so nothing but bookkeeping moves); IRMs migrate between ranks (partitions only) to keep the load balanced.

Results do not depend on the number of ranks: every random draw is keyed by (seed, IRM uid / relation, step).

Differences from the reference, by design: every IRM carries every domain of the schema and seats every entity code of
a domain (entities an IRM has no observation of move under the prior alone, cxx/hirm.hh:618 -- the reference's "ghost"
entities, SURVEY 8a-Q7/Q8 -- and by the CRP's consistency they do not change the law of the others), so a relation can
be scored under any IRM without trial seating; the order of the relation moves is a seeded permutation (the reference's
is a hash-container order).
"""
import math

import numpy as np

from . import shard
from .model import choose_index

NGRID_ALPHA = 20           # CRP::transition_alpha grid of the C++ backend (cxx/hirm.hh:164)


class Comm:
    """The few collectives the relation step needs, over torch.distributed (nccl: one rank per GPU; gloo: CPU tests).
    Without an initialised process group it is a world of one."""

    def __init__(self, group=None):
        self.group, self.dist, self.world, self.rank = group, None, 1, 0
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.dist, self.world, self.rank = dist, dist.get_world_size(group), dist.get_rank(group)
        except ImportError:
            pass
        self.n_allgather = self.n_broadcast = 0

    def _tensor(self, a):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
        return t.cuda() if self.dist.get_backend(self.group) == "nccl" else t

    def all_gather_f64(self, a):
        """a: float64 array -- or a CUDA tensor, gathered where it is (ncclAllGather on the buffer the score kernels filled,
        one device-to-host copy of the result) -- of the same shape on every rank -> list of arrays, one per rank."""
        self.n_allgather += 1
        on_device = not isinstance(a, np.ndarray) and hasattr(a, "is_cuda") and a.is_cuda
        if self.world == 1:
            return [a.cpu().numpy() if on_device else np.asarray(a, dtype=np.float64)]
        import torch
        shape = tuple(a.shape) if on_device else np.shape(a)
        t = (a.contiguous() if on_device and self.dist.get_backend(self.group) == "nccl" else self._tensor(a.cpu().numpy() if on_device else a)).reshape(-1)
        out = torch.empty(self.world * t.numel(), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, t, group=self.group)
        out = out.cpu().numpy().reshape((self.world,) + tuple(shape))
        return [out[k] for k in range(self.world)]

    def broadcast_f64(self, a, src):
        self.n_broadcast += 1
        if self.world == 1:
            return np.asarray(a, dtype=np.float64)
        t = self._tensor(a)
        self.dist.broadcast(t, src=self.dist.get_global_rank(self.group, src) if self.group is not None else src, group=self.group)
        return t.cpu().numpy()

    def broadcast_obj(self, obj, src):
        if self.world == 1:
            return obj
        box = [obj]
        self.dist.broadcast_object_list(box, src=self.dist.get_global_rank(self.group, src) if self.group is not None else src,
                                        group=self.group)
        return box[0]

    def all_gather_obj(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out


class EngineBackend:
    """Device side of one rank: the relation-owned COO arrays (torch tensors as buffers, replicated on every rank) and
    the engine IRMs of the relation clusters this rank owns.  Every call is a C-ABI call (include/hirm_engine.h)."""

    def __init__(self, n_entities, rel_domains, observations, max_tables, seed, engine=None, device=None, host_order=False, unary_blocks=True):
        import torch
        from . import engine as E
        if not torch.cuda.is_available():
            raise E.EngineError("no CUDA device: the engine has no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.eng = engine or E.Engine(self.device)
        # the engine caches bit vectors of unary relations by COO buffer address: a reused engine may hold those of buffers
        # torch has freed and handed out again
        self.eng.clear_relation_cache()
        self.n_entities, self.rel_domains = [int(n) for n in n_entities], [list(map(int, ds)) for ds in rel_domains]
        self.max_tables, self.seed, self.host_order = int(max_tables), int(seed), host_order
        dev = torch.device("cuda", self.device)
        self._buf, self.rels = [], []
        for ds, (ids, val) in zip(self.rel_domains, observations):
            if torch.is_tensor(ids):
                # COO arrays already in this GPU's memory (int32 [nobs * arity] / uint8 [nobs], contiguous): used in place,
                # relations may share an ids buffer; hirm_irm_set_observations_device validates codes and values
                t_ids, t_val = ids.reshape(-1), val
                assert t_ids.is_cuda and t_ids.dtype == torch.int32 and t_val.dtype == torch.uint8 and t_ids.is_contiguous()
                assert t_ids.numel() == t_val.numel() * len(ds)
            else:
                ids = np.ascontiguousarray(ids, dtype=np.int32).reshape(-1, len(ds))
                val = np.ascontiguousarray(val, dtype=np.uint8)
                assert len(ids) == len(val)
                if len(val) and (val.max() > 1 or ids.min() < 0 or np.any(ids.max(axis=0) >= np.asarray([self.n_entities[d] for d in ds]))):
                    raise ValueError("observation values must be 0 or 1 and entity codes in range")
                t_ids, t_val = torch.from_numpy(ids.reshape(-1)).to(dev), torch.from_numpy(val).to(dev)
            self._buf.append((t_ids, t_val))
            self.rels.append((ds, int(t_val.numel()), t_ids.data_ptr() if t_val.numel() else None, t_val.data_ptr() if t_val.numel() else None))
        self._meta = {}
        # unary relations of a domain as ONE relation-owned block of entity-major bit rows (hirm_unary_block_create): an IRM
        # picks its relations by column, so re-seating a unary relation moves no data and builds no CSR
        self.ublock = {}                                   # relation index -> (block handle, column)
        if unary_blocks:
            by_dom = {}
            for r, ds in enumerate(self.rel_domains):
                if len(ds) == 1 and self.rels[r][1] > 0:
                    by_dom.setdefault(ds[0], []).append(r)
            for d, rs in by_dom.items():
                if len(rs) < 2:
                    continue
                blk = self.eng.unary_block_create(self.n_entities[d], [self.rels[r][1:] for r in rs])
                for col, r in enumerate(rs):
                    self.ublock[r] = (blk, col)

    # -- IRMs -----------------------------------------------------------------------------------
    def create_irm(self, rels, z, alpha, uid, steps):
        """rels: global relation indices; z: labels per domain (every entity seated); steps: entity steps already made."""
        eng = self.eng
        h = eng.irm_create(self.n_entities, [self.rel_domains[r] for r in rels], max_tables=self.max_tables)
        for k, r in enumerate(rels):
            ds, nobs, pi, pv = self.rels[r]
            if r in self.ublock:
                eng.set_observations_unary(h, k, *self.ublock[r])
            else:
                eng.set_observations_device(h, k, nobs, pi, pv)
        eng.finalize(h)
        for d in range(len(self.n_entities)):
            eng.set_assignments(h, d, z[d])
            eng.set_alpha(h, d, alpha[d])
        per = sum(self.n_entities)
        eng.set_rng(h, uid, steps * per)
        eng.set_rng_counters(h, steps, steps)
        self._meta[h] = dict(uid=uid, steps=steps)
        return h

    def destroy_irm(self, h):
        self.eng.irm_destroy(h)
        self._meta.pop(h)

    def get_state(self, h):
        nd = len(self.n_entities)
        return [self.eng.get_assignments(h, d).copy() for d in range(nd)], [self.eng.get_alpha(h, d) for d in range(nd)]

    def get_counts(self, h, k):
        return self.eng.get_counts(h, k)

    # -- entity step: IRM::transition_cluster_assignments_all + CRP::transition_alpha of every listed IRM ------------
    def step(self, handles):
        if not handles:
            return
        eng = self.eng
        if self.host_order:
            moves = [host_sweep_order(self.n_entities, self.seed, self._meta[h]["uid"], self._meta[h]["steps"]) for h in handles]
            per = sum(self.n_entities)
            eng.sweep(handles, moves, seed=self.seed, streams=[self._meta[h]["uid"] for h in handles],
                      counters=[self._meta[h]["steps"] * per for h in handles])
        else:
            eng.sweep_all(handles, n_sweeps=1, seed=self.seed)
        eng.transition_alpha(handles, ngrid=NGRID_ALPHA, seed=self.seed)
        for h in handles:
            m = self._meta[h]
            m["steps"] += 1
            if self.host_order:
                eng.set_rng(h, m["uid"], m["steps"] * sum(self.n_entities))

    def logp_scores(self, handles):
        return self.eng.logp_scores(handles) if handles else np.zeros(0)

    # -- relation step --------------------------------------------------------------------------------------------------
    def score_matrix(self, rels, handles):
        return self.eng.relation_score_matrix([self.rels[r] for r in rels], handles)

    def score_matrix_device(self, rels, handles):
        """The same matrix left in this GPU's memory (a torch tensor as the buffer)."""
        import torch
        out = torch.zeros((len(rels), len(handles)), dtype=torch.float64, device=torch.device("cuda", self.device))
        if len(rels) and len(handles):
            self.eng.relation_score_matrix([self.rels[r] for r in rels], handles, d_out_ptr=out.data_ptr())
        return out

    def scores_fresh(self, rels, keys, chunk_slots=1 << 26):
        out = np.zeros(len(rels), np.float64)
        i = 0
        while i < len(rels):                               # bounded scratch: the fresh partitions of a chunk live in HBM at once
            j, slots = i, 0
            while j < len(rels) and (j == i or slots + self._slots(rels[j]) <= chunk_slots):
                slots += self._slots(rels[j]); j += 1
            out[i:j] = self.eng.relation_scores_fresh([self.rels[r] for r in rels[i:j]], self.n_entities,
                                                      [self.max_tables] * len(self.n_entities), keys[i:j], self.seed)
            i = j
        return out

    def _slots(self, r):
        return sum(self.n_entities[d] for d in set(self.rel_domains[r]))

    def prior_draw(self, d, key):
        return self.eng.crp_prior_draw(self.n_entities[d], key, self.seed, alpha=1.0, max_tables=self.max_tables)


def host_sweep_order(n_entities, seed, uid, step):
    """Sweep order of one entity step when the order is made on the host (parity tests): every domain in index order, a
    seeded permutation of its entities."""
    md, mi = [], []
    for d, n in enumerate(n_entities):
        perm = np.random.default_rng([int(seed), int(uid), int(step), d, 0x6F72]).permutation(n).astype(np.int32)
        md.append(np.full(n, d, np.int32)); mi.append(perm)
    return np.concatenate(md), np.concatenate(mi)


class ShardedHIRM:
    """domains: {name: number of entity codes}; schema: {relation: [domain names]}; observations: {relation: (ids
    [nobs, arity] int32 entity codes, val [nobs] in {0, 1})}.  Entity codes are those of cxx/util_io.cc:63-96
    (hirm.util_io encodes the reference's .obs files the same way)."""

    def __init__(self, domains, schema, observations, seed=0, max_tables=64, comm=None, backend=None,
                 relation_tables=None, partitions=None, alphas=None, rebalance=True, host_order=False, device=None, chain=0):
        self.dnames = list(domains)
        self.n_ent = [int(domains[d]) for d in self.dnames]
        if len(self.dnames) > 8:
            raise ValueError("at most 8 domains (HIRM_MAX_DOMAINS)")
        dindex = {d: k for k, d in enumerate(self.dnames)}
        self.rnames = list(schema)
        self.rindex = {r: k for k, r in enumerate(self.rnames)}
        self.rel_domains = [[dindex[d] for d in schema[r]] for r in self.rnames]
        obs = [observations[r] for r in self.rnames]
        self.rel_weight = [max(1, len(o[1])) * len(ds) for o, ds in zip(obs, self.rel_domains)]   # CSR entries of the relation
        self.seed, self.max_tables, self.rebalance = int(seed), int(max_tables), rebalance
        self.comm = comm or Comm()
        self.backend = backend or EngineBackend(self.n_ent, self.rel_domains, obs, max_tables, seed, device=device, host_order=host_order)
        self.alpha = 1.0                                   # CRP over relations (cxx/hirm.hh:790); never transitioned by inference_hirm
        self.irms = {}                                     # table label -> dict(uid, owner, rels, handle, steps, changed)
        self.rel_table = {}
        self.chain = int(chain)                            # independent seeded chains: disjoint Philox streams and keys
        self._next_uid, self._next_key, self.iteration = self.chain << 24, (1 << 60) | (self.chain << 44), 0
        self.timers = dict(entity=0.0, score=0.0, gather=0.0, reseat=0.0, rebuild=0.0)
        self.stats = dict(relations_moved=0, fresh_irms=0, irms_deleted=0, irms_rebuilt=0, irms_migrated=0)
        R = len(self.rnames)
        rng = np.random.default_rng([self.seed, self.chain, 0x696E6974])
        if relation_tables is None:                        # HIRM::add_relation (:947-957): sequential CRP(alpha) seating
            relation_tables, sizes = {}, {}
            for r in self.rnames:
                if not sizes:
                    t = 0
                else:
                    tabs = sorted(sizes) + [max(sizes) + 1]
                    w = np.array([sizes[x] for x in sorted(sizes)] + [self.alpha], dtype=np.float64)
                    t = tabs[int(rng.choice(len(tabs), p=w / w.sum()))]
                sizes[t] = sizes.get(t, 0) + 1
                relation_tables[r] = t
        for r in self.rnames:
            self.rel_table[self.rindex[r]] = int(relation_tables[r])
        labels = sorted(set(self.rel_table.values()))
        members = {t: sorted(k for k, x in self.rel_table.items() if x == t) for t in labels}
        owner, _ = shard.pack_irms([self._weight(members[t]) for t in labels], self.comm.world)
        for t, own in zip(labels, owner):
            keys = [self._take_key() for _ in self.dnames]
            info = self.irms[t] = dict(uid=self._take_uid(), owner=own, rels=set(members[t]), handle=None, steps=0, changed=False)
            if own == self.comm.rank:
                if partitions is not None:
                    z = [np.asarray(partitions[t][d], dtype=np.int32) for d in self.dnames]
                else:
                    z = [self.backend.prior_draw(d, keys[d]) for d in range(len(self.dnames))]
                al = [1.0] * len(self.dnames) if alphas is None else [float(alphas[t][d]) for d in self.dnames]
                info["handle"] = self.backend.create_irm(sorted(info["rels"]), z, al, info["uid"], 0)
        assert R == len(self.rel_table)

    # ---- bookkeeping ----------------------------------------------------------------------------
    def _take_uid(self):
        self._next_uid += 1
        return self._next_uid - 1

    def _take_key(self):
        self._next_key += 1
        return self._next_key - 1

    def _weight(self, rels):
        return sum(self.rel_weight[r] for r in rels) + sum(self.n_ent)

    def _local(self):
        return [t for t in sorted(self.irms) if self.irms[t]["owner"] == self.comm.rank]

    def loads(self):
        load = [0] * self.comm.world
        for info in self.irms.values():
            load[info["owner"]] += self._weight(info["rels"])
        return load

    def relation_to_table(self, r):
        return self.rel_table[self.rindex[r]]

    # ---- entity step (cxx/hirm.cc:73-88): no collective ---------------------------------------------------------------
    def transition_irms(self, n_steps=1):
        import time
        t0 = time.perf_counter()
        hs = [self.irms[t]["handle"] for t in self._local()]
        for _ in range(n_steps):
            self.backend.step(hs)
            for info in self.irms.values():
                info["steps"] += 1
        self.timers["entity"] += time.perf_counter() - t0

    # ---- relation step (cxx/hirm.cc:66-71 -> cxx/hirm.hh:833-906) --------------------------------------------------------
    def _fresh_keys(self):
        """Philox streams of this relation sweep's fresh-table draws, [R, n_domains] (replicated bookkeeping)."""
        R, nd = len(self.rnames), len(self.dnames)
        base = self._next_key
        self._next_key += R * nd
        return (base + np.arange(R * nd, dtype=np.uint64)).reshape(R, nd)

    def score_columns(self):
        """Steps 1-2: {table label: [R] scores of every relation under that IRM}, [R] fresh-table scores, fresh keys."""
        return score_columns_batched([self])[0]

    def relation_candidates(self, r, cols, fresh_score):
        """CRP::tables_weights_gibbs (:147-161) + the data terms: (tables ascending, log-probabilities, fresh label or None)."""
        cur = self.rel_table[r]
        tables = sorted(self.irms)
        fresh_label = None
        lp = []
        for t in tables:
            n = len(self.irms[t]["rels"]) - (1 if t == cur else 0)
            lp.append((math.log(n) if n > 0 else math.log(self.alpha)) + float(cols[t][r]))   # singleton: own table at weight alpha (:152-159)
        if len(self.irms[cur]["rels"]) > 1:                # otherwise no fresh table is offered
            fresh_label = tables[-1] + 1
            tables = tables + [fresh_label]
            lp.append(math.log(self.alpha) + fresh_score)
        return tables, lp, fresh_label

    def apply_relation_choice(self, r, choice, cols, keys=None, parts=None):
        """Re-seat relation r at table `choice` (the tail of :877-905).  A fresh table becomes an IRM whose partitions are
        the keyed CRP-prior draws (or `parts`, {domain index: labels}, in the trace tests); its score column for the
        other relations is computed by its owner and broadcast."""
        cur = self.rel_table[r]
        if choice == cur:
            return
        if choice not in self.irms:
            load = self.loads()
            own = min(range(self.comm.world), key=lambda k: (load[k], k))
            info = self.irms[choice] = dict(uid=self._take_uid(), owner=own, rels=set(), handle=None, steps=0, changed=True)
            self.stats["fresh_irms"] += 1
            latent = [self._take_key() for _ in self.dnames]          # domains the relation does not use: prior draws all the same
            col = np.zeros(len(self.rnames))
            if own == self.comm.rank:
                z = []
                for d in range(len(self.dnames)):
                    if parts is not None and d in parts:
                        z.append(np.asarray(parts[d], dtype=np.int32))
                    elif keys is not None and d in self.rel_domains[r]:
                        z.append(self.backend.prior_draw(d, int(keys[r, d])))
                    else:
                        z.append(self.backend.prior_draw(d, latent[d]))
                info["handle"] = self.backend.create_irm([], z, [1.0] * len(self.dnames), info["uid"], 0)
                col = self.backend.score_matrix(list(range(len(self.rnames))), [info["handle"]])[:, 0]
            cols[choice] = self.comm.broadcast_f64(col, own)
        src, dst = self.irms[cur], self.irms[choice]
        src["rels"].discard(r); dst["rels"].add(r)
        src["changed"] = dst["changed"] = True
        self.rel_table[r] = choice
        if not src["rels"]:                                # the emptied IRM is deleted (:886-890)
            if src["handle"] is not None:
                self.backend.destroy_irm(src["handle"])
            del self.irms[cur]
            cols.pop(cur, None)
            self.stats["irms_deleted"] += 1

    def transition_relations(self, scored=None, native=True):
        """One sweep of relation moves: every relation once, in a seeded order.  `scored`: the (columns, fresh scores,
        keys) of score_columns_batched when several chains were scored in one launch.  The sequential CRP walk runs in
        hirm_relation_reseat (C++, csrc/relation_moves.cc) between two fresh tables; `native=False` keeps it in Python
        (relation_candidates / apply_relation_choice: the same arithmetic, used by the trace tests)."""
        import time
        cols, fresh, keys = scored if scored is not None else self.score_columns()
        t0 = time.perf_counter()
        rng = np.random.default_rng([self.seed, self.chain, self.iteration, 0x72656C])
        order = rng.permutation(len(self.rnames))
        us = rng.random(len(self.rnames))
        moved = 0
        if not native:
            for r, u in zip(order, us):
                r = int(r)
                tables, lp, fresh_label = self.relation_candidates(r, cols, float(fresh[r]))
                choice = tables[choose_index(lp, float(u))]
                moved += int(choice != self.rel_table[r])
                self.apply_relation_choice(r, choice, cols, keys=keys)
        else:
            from . import engine as E
            R = len(self.rnames)
            order32, us64 = np.ascontiguousarray(order, dtype=np.int32), np.ascontiguousarray(us, dtype=np.float64)
            fresh64 = np.asarray(fresh, dtype=np.float64)
            pos = 0
            while pos < R:
                labels = np.asarray(sorted(self.irms), dtype=np.int32)
                sizes = np.asarray([len(self.irms[int(t)]["rels"]) for t in labels], dtype=np.int32)
                rel_tab = np.asarray([self.rel_table[r] for r in range(R)], dtype=np.int32)
                before = rel_tab.copy()
                M = np.empty((R, len(labels) + 1), dtype=np.float64)
                for j, t in enumerate(labels):
                    M[:, j] = cols[int(t)]
                M[:, len(labels)] = fresh64
                status, pos, n_moved, fresh_label = E.relation_reseat(labels, sizes, rel_tab, M, len(labels), order32, us64, self.alpha, pos)
                moved += n_moved
                for r in np.flatnonzero(rel_tab != before):           # the bookkeeping of the moves just made
                    r, src, dst = int(r), int(before[r]), int(rel_tab[r])
                    self.irms[src]["rels"].discard(r); self.irms[dst]["rels"].add(r)
                    self.irms[src]["changed"] = self.irms[dst]["changed"] = True
                    self.rel_table[r] = dst
                for t in [t for t, info in self.irms.items() if not info["rels"]]:   # emptied IRMs are deleted (:886-890)
                    if self.irms[t]["handle"] is not None:
                        self.backend.destroy_irm(self.irms[t]["handle"])
                    del self.irms[t]
                    cols.pop(t, None)
                    self.stats["irms_deleted"] += 1
                if status == 1:                                       # the relation at order[pos] opens a fresh table
                    self.apply_relation_choice(int(order32[pos]), int(fresh_label), cols, keys=keys)
                    moved += 1
                    pos += 1
        t1 = time.perf_counter()
        self.stats["relations_moved"] += moved
        self.materialize()
        self.timers["reseat"] += t1 - t0
        self.timers["rebuild"] += time.perf_counter() - t1
        self.iteration += 1
        return moved

    def materialize(self):
        """Step 4: bring the engine IRMs in line with the bookkeeping -- rebuild those whose relation set changed, move
        those whose owner changed (partitions, alphas and Philox positions travel; observations are everywhere)."""
        labels = sorted(self.irms)
        new_owner = {t: self.irms[t]["owner"] for t in labels}
        if self.rebalance and self.comm.world > 1:
            load = [0.0] * self.comm.world
            for t in sorted(labels, key=lambda t: (-self._weight(self.irms[t]["rels"]), t)):
                w, cur = self._weight(self.irms[t]["rels"]), self.irms[t]["owner"]
                k = min(range(self.comm.world), key=lambda k: (load[k] - (0.25 * w if k == cur else 0.0), k))   # sticky LPT
                new_owner[t] = k
                load[k] += w
        for t in labels:
            info = self.irms[t]
            old, new = info["owner"], new_owner[t]
            if not info["changed"] and old == new:
                continue
            self.stats["irms_rebuilt"] += 1
            self.stats["irms_migrated"] += int(old != new)
            state = None
            if old == self.comm.rank:
                state = self.backend.get_state(info["handle"])
                self.backend.destroy_irm(info["handle"])
                info["handle"] = None
            if old != new:
                state = self.comm.broadcast_obj(state, old)
            if new == self.comm.rank:
                info["handle"] = self.backend.create_irm(sorted(info["rels"]), state[0], state[1], info["uid"], info["steps"])
            info["owner"], info["changed"] = new, False

    # ---- whole iterations / read-back --------------------------------------------------------------------------------
    def iterate(self, n=1):
        """inference_hirm (cxx/hirm.cc:61-90): relation moves, then the entity sweeps and alpha moves of every IRM."""
        for _ in range(n):
            self.transition_relations()
            self.transition_irms()

    def crp_logp_score(self):
        sizes = [len(info["rels"]) for info in self.irms.values()]
        n = sum(sizes)
        return (len(sizes) * math.log(self.alpha) + sum(math.lgamma(c) for c in sizes) + math.lgamma(self.alpha)
                - math.lgamma(n + self.alpha))

    def logp_score(self):
        """HIRM::logp_score (cxx/hirm.hh:998-1005); per-IRM scores gathered and summed in table order on every rank.
        An IRM's score holds the CRP term of the domains its relations use, as the reference's does (its IRMs hold no
        other domain); the partitions this model keeps of the remaining domains are latent prior draws and add nothing."""
        local = self._local()
        sc = self.backend.logp_scores([self.irms[t]["handle"] for t in local])
        merged = {}
        for part in self.comm.all_gather_obj({t: float(s) for t, s in zip(local, sc)}):
            merged.update(part)
        return self.crp_logp_score() + sum(merged[t] for t in sorted(merged))

    def state(self):
        """{table: dict(relations=[names], z={domain: labels}, alpha={domain: a})}, the same on every rank."""
        mine = {}
        for t in self._local():
            z, al = self.backend.get_state(self.irms[t]["handle"])
            mine[t] = dict(relations=[self.rnames[r] for r in sorted(self.irms[t]["rels"])],
                           z={d: z[k] for k, d in enumerate(self.dnames)}, alpha={d: al[k] for k, d in enumerate(self.dnames)})
        merged = {}
        for part in self.comm.all_gather_obj(mine):
            merged.update(part)
        return {t: merged[t] for t in sorted(merged)}


def score_columns_batched(chains):
    """Steps 1-2 of the relation move for several chains over the same data at once: ONE score-matrix launch over every
    chain's local IRMs, ONE fresh-table launch, ONE all-gather.  Returns [(columns, fresh scores, keys)] per chain."""
    import time
    t0 = time.perf_counter()
    h0 = chains[0]
    comm, backend = h0.comm, h0.backend
    R, world, rank = len(h0.rnames), comm.world, comm.rank
    mine = [r for r in range(R) if r % world == rank]
    locals_ = [h._local() for h in chains]
    handles = [h.irms[t]["handle"] for h, loc in zip(chains, locals_) for t in loc]
    on_device = hasattr(backend, "score_matrix_device")
    m_all = backend.score_matrix_device(list(range(R)), handles) if on_device else backend.score_matrix(list(range(R)), handles)
    keys = [h._fresh_keys() for h in chains]
    f_all = backend.scores_fresh(mine * len(chains), np.concatenate([k[mine] for k in keys])) if mine else np.zeros(0)
    t1 = time.perf_counter()
    per_rank = [[[t for t in sorted(h.irms) if h.irms[t]["owner"] == k] for k in range(world)] for h in chains]
    kmax = max(len(p) for pr in per_rank for p in pr)
    if on_device:                                          # the block stays in HBM until the gathered matrix comes back
        import torch
        buf = torch.zeros((len(chains), R, kmax + 1), dtype=torch.float64, device=m_all.device)
        f_dev = torch.from_numpy(np.ascontiguousarray(f_all, dtype=np.float64)).to(m_all.device)
        mine_dev = torch.as_tensor(mine, dtype=torch.int64, device=m_all.device)
    else:
        buf = np.zeros((len(chains), R, kmax + 1), dtype=np.float64)
    off = 0
    for c, loc in enumerate(locals_):
        buf[c, :, :len(loc)] = m_all[:, off:off + len(loc)]
        off += len(loc)
        if on_device:
            buf[c, mine_dev, kmax] = f_dev[c * len(mine):(c + 1) * len(mine)]
        else:
            buf[c, mine, kmax] = f_all[c * len(mine):(c + 1) * len(mine)]
    parts = comm.all_gather_f64(buf)
    out = []
    for c in range(len(chains)):
        cols = {t: parts[k][c][:, j].copy() for k in range(world) for j, t in enumerate(per_rank[c][k])}
        fresh = np.array([parts[r % world][c][r, kmax] for r in range(R)])
        out.append((cols, fresh, keys[c]))
    h0.timers["score"] += t1 - t0
    h0.timers["gather"] += time.perf_counter() - t1
    return out


class ChainSet:
    """Several independent seeded chains of one HIRM over the same observations and the same ranks (BASELINE config 5:
    "8 chains").  The chains share the replicated COO arrays; per iteration the relation steps run chain by chain (small),
    then the entity sweeps of EVERY chain's local IRMs go out in ONE launch -- chains are what fills a GPU when an HIRM
    has only a few IRMs per rank."""

    def __init__(self, domains, schema, observations, n_chains, seed=0, max_tables=64, comm=None, backend=None, **kw):
        first = ShardedHIRM(domains, schema, observations, seed=seed, max_tables=max_tables, comm=comm, backend=backend, chain=0, **kw)
        self.chains = [first] + [ShardedHIRM(domains, schema, observations, seed=seed, max_tables=max_tables, comm=first.comm,
                                             backend=first.backend, chain=c, **kw) for c in range(1, n_chains)]
        self.backend, self.comm = first.backend, first.comm

    def transition_relations(self):
        scored = score_columns_batched(self.chains)
        return [h.transition_relations(scored=sc) for h, sc in zip(self.chains, scored)]

    def transition_irms(self):
        hs = [h.irms[t]["handle"] for h in self.chains for t in h._local()]
        self.backend.step(hs)
        for h in self.chains:
            for info in h.irms.values():
                info["steps"] += 1

    def iterate(self, n=1):
        for _ in range(n):
            self.transition_relations()
            self.transition_irms()

    def logp_scores(self):
        return [h.logp_score() for h in self.chains]
