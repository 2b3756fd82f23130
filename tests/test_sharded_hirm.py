"""hirm.sharded.ShardedHIRM -- the relation-cluster IRMs of an HIRM sharded over ranks, the relation move
(HIRM::transition_cluster_assignment_relation, cxx/hirm.hh:833-906) from one all-gathered score matrix.

CPU (gloo, world size 2): the host logic on the oracle backend -- results do not depend on the number of ranks.
GPU: the batched relation-score / fresh-table / CRP-prior kernels against numpy restatements; the device CSR build
against host-built expectations; the engine backend against the oracle backend move for move; the relation moves of
the UNMODIFIED reference on animals.unary (tests/golden/animals_unary_hirm.json.gz) replayed through the score matrix.
"""
import os
import socket

import numpy as np
import pytest

import golden_util as G
import sharded_util as SU
from hirm.sharded import Comm, ShardedHIRM


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(backend_kind, iters, comm=None, hseed=3, device=None, **gen):
    domains, schema, obs, truth = SU.synthetic_hirm(**gen)
    names = list(schema)
    if backend_kind == "oracle":
        n_ent = [domains[d] for d in domains]
        dindex = {d: k for k, d in enumerate(domains)}
        backend = SU.OracleBackend(n_ent, [[dindex[d] for d in schema[r]] for r in names], [obs[r] for r in names], 32, hseed)
        h = ShardedHIRM(domains, schema, obs, seed=hseed, max_tables=32, comm=comm, backend=backend)
    else:
        h = ShardedHIRM(domains, schema, obs, seed=hseed, max_tables=32, comm=comm, host_order=True, device=device)
    trail = []
    for _ in range(iters):
        h.transition_relations()
        h.transition_irms()
        st = h.state()
        trail.append((dict((r, h.relation_to_table(r)) for r in names), st, h.logp_score()))
    return h, trail


def _same_trail(a, b, rtol=0.0):
    assert len(a) == len(b)
    for (ra, sa, la), (rb, sb, lb) in zip(a, b):
        assert ra == rb
        assert sorted(sa) == sorted(sb)
        for t in sa:
            assert sa[t]["relations"] == sb[t]["relations"]
            for d in sa[t]["z"]:
                np.testing.assert_array_equal(sa[t]["z"][d], sb[t]["z"][d])
                assert abs(sa[t]["alpha"][d] - sb[t]["alpha"][d]) <= 1e-12 * sa[t]["alpha"][d]
        assert abs(la - lb) <= max(rtol, 1e-12) * max(1.0, abs(la)), (la, lb)


def _check_invariants(h, trail):
    rel_tab, st, score = trail[-1]
    assert sorted(st) == sorted(set(rel_tab.values()))                     # one IRM per occupied table
    for t, info in st.items():
        assert sorted(info["relations"]) == sorted(r for r, x in rel_tab.items() if x == t)
        for d, z in info["z"].items():
            assert len(z) == h.n_ent[h.dnames.index(d)] and z.min() >= 0   # every entity seated in every IRM
    assert np.isfinite(score)


def test_oracle_backend_single_rank_invariants_and_recovery():
    h, trail = _run("oracle", 6, n_ent=(40, 30), n_unary=9, n_binary=3, n_clusters=3, seed=1)
    _check_invariants(h, trail)
    assert trail[-1][2] > trail[0][2] - 1e-9 or len(trail) < 2            # the chain climbs from its prior start
    assert h.comm.n_allgather == 6                                        # ONE all-gather per relation sweep


def test_native_reseat_walk_equals_python_walk():
    """hirm_relation_reseat (C++) against the Python walk (relation_candidates + choose_index + apply_relation_choice): same
    relation partitions, same entity partitions, same scores over whole iterations -- fresh tables, emptied tables and
    singletons included (CPU, oracle backend)."""
    gen = dict(n_ent=(30, 24), n_unary=10, n_binary=4, n_clusters=3, seed=5)
    domains, schema, obs, _ = SU.synthetic_hirm(**gen)
    names, dn = list(schema), list(domains)
    trails, stats = [], []
    for native in (False, True):
        backend = SU.OracleBackend([domains[d] for d in dn], [[dn.index(d) for d in schema[r]] for r in names], [obs[r] for r in names], 32, 3)
        h = ShardedHIRM(domains, schema, obs, seed=3, max_tables=32, backend=backend)
        trail = []
        for _ in range(6):
            h.transition_relations(native=native)
            h.transition_irms()
            trail.append((dict((r, h.relation_to_table(r)) for r in names), h.state(), h.logp_score()))
        trails.append(trail); stats.append(dict(h.stats))
    _same_trail(trails[0], trails[1])
    assert stats[0] == stats[1] and stats[0]["fresh_irms"] > 0 and stats[0]["irms_deleted"] > 0 and stats[0]["relations_moved"] > 10


def test_multi_domain_logp_score_counts_used_domains_only():
    """HIRM::logp_score (cxx/hirm.hh:998-1005) on a schema with three domains where most relation clusters use only some
    of them: the CRP term of an IRM covers the domains its relations use (the reference's IRM holds no other domain,
    cxx/hirm.hh:742-782), checked against an independent numpy evaluation of the whole score."""
    import math
    rng = np.random.default_rng(0)
    domains = {"a": 14, "b": 11, "c": 9}
    schema = {"ua": ["a"], "ub": ["b"], "uc": ["c"], "ab": ["a", "b"], "cc": ["c", "c"], "abc": ["a", "b", "c"]}
    obs = {}
    for r, ds in schema.items():
        grid = np.stack(np.meshgrid(*[np.arange(domains[d]) for d in ds], indexing="ij"), -1).reshape(-1, len(ds))
        keep = grid[rng.random(len(grid)) < (1.0 if len(ds) == 1 else 0.5)].astype(np.int32)
        obs[r] = (keep, rng.integers(0, 2, len(keep)).astype(np.uint8))
    names, dn = list(schema), list(domains)
    backend = SU.OracleBackend([domains[d] for d in dn], [[dn.index(d) for d in schema[r]] for r in names], [obs[r] for r in names], 16, 1)
    h = ShardedHIRM(domains, schema, obs, seed=1, max_tables=16, backend=backend,
                    relation_tables={"ua": 0, "ab": 0, "ub": 1, "uc": 2, "cc": 2, "abc": 3})
    for it in range(3):
        if it:
            h.iterate(1)
        st = h.state()
        want = h.crp_logp_score()
        for t, info in st.items():
            used = {d for r in info["relations"] for d in schema[r]}
            assert it or len(used) < 3 or t == 3
            for d in used:
                sizes = np.bincount(np.unique(info["z"][d], return_inverse=True)[1])
                al = info["alpha"][d]
                want += len(sizes) * math.log(al) + sum(math.lgamma(c) for c in sizes) + math.lgamma(al) - math.lgamma(sizes.sum() + al)
            for r in info["relations"]:
                want += SU.relation_score_numpy(*obs[r], [info["z"][d] for d in schema[r]])
        got = h.logp_score()
        assert abs(got - want) <= 1e-9 * max(1.0, abs(want)), (it, got, want)


def test_chain_set_equals_stand_alone_chains():
    """ChainSet batches the entity sweeps of several chains; every chain walks the trajectory it walks alone."""
    from hirm.sharded import ChainSet
    gen = dict(n_ent=(24, 18), n_unary=6, n_binary=2, n_clusters=2, seed=5)
    domains, schema, obs, _ = SU.synthetic_hirm(**gen)
    names = list(schema)
    n_ent = [domains[d] for d in domains]
    dindex = {d: k for k, d in enumerate(domains)}
    mk = lambda: SU.OracleBackend(n_ent, [[dindex[d] for d in schema[r]] for r in names], [obs[r] for r in names], 32, 7)
    cs = ChainSet(domains, schema, obs, 3, seed=7, max_tables=32, backend=mk())
    cs.iterate(3)
    for c in range(3):
        solo = ShardedHIRM(domains, schema, obs, seed=7, max_tables=32, backend=mk(), chain=c)
        solo.iterate(3)
        a, b = cs.chains[c].state(), solo.state()
        assert sorted(a) == sorted(b)
        for t in a:
            assert a[t]["relations"] == b[t]["relations"]
            for d in a[t]["z"]:
                np.testing.assert_array_equal(a[t]["z"][d], b[t]["z"][d])
    s = cs.logp_scores()
    assert len(set(round(x, 6) for x in s)) > 1                           # different seeds, different chains


def _gloo_worker(rank, world, port, out, iters, gen):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    h, trail = _run("oracle", iters, comm=Comm(), **gen)
    owners = sorted({info["owner"] for info in h.irms.values()})
    dist.barrier()
    if rank == 0:
        out.put((trail, owners, h.comm.n_allgather, h.comm.n_broadcast, dict(h.stats)))
    dist.destroy_process_group()


def test_two_ranks_gloo_match_one_rank():
    """World size 2 (gloo, CPU): same relation partition, same entity partitions, same scores as one rank."""
    import torch.multiprocessing as mp
    gen = dict(n_ent=(30, 24), n_unary=8, n_binary=3, n_clusters=3, seed=2)
    iters = 4
    _, ref = _run("oracle", iters, **gen)
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q, iters, gen)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    trail, owners, n_ag, n_bc, stats = q.get()
    _same_trail(ref, trail)
    assert n_ag == iters
    # the exchange step really happened: relations changed tables, a fresh IRM was opened (one broadcast each), IRMs were
    # rebuilt on their owner and at least one migrated between the ranks
    assert stats["relations_moved"] > 0 and stats["fresh_irms"] > 0 and n_bc >= stats["fresh_irms"]
    assert stats["irms_rebuilt"] > 0 and stats["irms_migrated"] > 0, stats
    assert owners == [0, 1] or len(trail[-1][1]) == 1                      # IRMs spread over both ranks


# ---------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------
def _torch_coo(ids, val, dev=0):
    import torch
    ids = np.ascontiguousarray(ids, np.int32)
    return torch.from_numpy(ids.reshape(-1)).to("cuda:%d" % dev), torch.from_numpy(np.ascontiguousarray(val, np.uint8)).to("cuda:%d" % dev)


@pytest.mark.gpu
def test_device_coo_store_equals_host_coo_store():
    """hirm_irm_set_observations_device (caller's device buffers) against hirm_irm_set_observations: same count tables,
    same candidates and choices over a sweep (the CSR of both is built on the device; the host-fed path is what every
    golden trace runs through)."""
    from hirm import engine as E
    spec = G.IrmSpec(G.load("synth_mixed_irm")["irms"][0])
    eng = E.Engine(0)
    import engine_util as U
    ha = U.engine_from_spec(eng, spec)
    hb = eng.irm_create(spec.n_entities, spec.rel_domains, max_tables=24)
    keep = []
    for r in range(len(spec.rnames)):
        ti, tv = _torch_coo(spec.obs_ids[r], spec.obs_val[r])
        keep.append((ti, tv))
        eng.set_observations_device(hb, r, len(spec.obs_val[r]), ti.data_ptr() if len(tv) else None, tv.data_ptr() if len(tv) else None)
    eng.finalize(hb)
    for d in range(len(spec.dnames)):
        eng.set_assignments(hb, d, spec.init_z[d]); eng.set_alpha(hb, d, spec.init_alpha[d])
    for r in range(len(spec.rnames)):
        np.testing.assert_array_equal(eng.get_counts(ha, r), eng.get_counts(hb, r))
    rng = np.random.default_rng(5)
    md = np.concatenate([np.full(n, d, np.int32) for d, n in enumerate(spec.n_entities)])
    mi = np.concatenate([rng.permutation(n).astype(np.int32) for n in spec.n_entities])
    res = eng.sweep([ha, hb], [(md, mi), (md, mi)], seed=9, streams=[4, 4], counters=[0, 0], trace_cap=32)
    np.testing.assert_array_equal(res["choice"][0], res["choice"][1])
    assert G.close(res["trace"][0][2], res["trace"][1][2], rtol=1e-12)
    for r in range(len(spec.rnames)):
        np.testing.assert_array_equal(eng.get_counts(ha, r), eng.get_counts(hb, r))
    bad_i, bad_v = _torch_coo(np.array([[0]], np.int32), np.array([3], np.uint8))
    hc = eng.irm_create([4], [[0]], max_tables=8)
    with pytest.raises(E.EngineError):
        eng.set_observations_device(hc, 0, 1, bad_i.data_ptr(), bad_v.data_ptr())
    eng.close()


@pytest.mark.gpu
def test_unary_block_store_equals_coo_store():
    """Unary relations as columns of a relation-owned bit block (hirm_unary_block_create / hirm_irm_set_observations_unary)
    against the same relations as CSR entries: same count tables, same candidates, log-probabilities equal to 1e-12 (the same
    terms in another order), same choices; partially observed unary relations, an IRM using a subset of the block's columns,
    block errors."""
    from hirm import engine as E
    rng = np.random.default_rng(8)
    n, m = 90, 40
    unary = []
    for k in range(37):                                           # more than one mask word
        keep = np.flatnonzero(rng.random(n) < (1.0 if k % 3 else 0.6)).astype(np.int32)
        unary.append((keep[:, None], rng.integers(0, 2, len(keep)).astype(np.uint8)))
    grid = np.stack(np.meshgrid(np.arange(n), np.arange(n), indexing="ij"), -1).reshape(-1, 2)
    b_ee = grid[rng.random(len(grid)) < 0.2].astype(np.int32)
    grid2 = np.stack(np.meshgrid(np.arange(n), np.arange(m), indexing="ij"), -1).reshape(-1, 2)
    b_ef = grid2[rng.random(len(grid2)) < 0.5].astype(np.int32)
    others = [(b_ee, rng.integers(0, 2, len(b_ee)).astype(np.uint8)), (b_ef, rng.integers(0, 2, len(b_ef)).astype(np.uint8))]
    eng = E.Engine(0)
    dev = [_torch_coo(i, v) for i, v in unary]
    blk = eng.unary_block_create(n, [(len(v), ti.data_ptr(), tv.data_ptr()) for (i, v), (ti, tv) in zip(unary, dev)])
    pick = [0, 3, 4, 7, 11, 12, 20, 31, 32, 33, 36]               # the IRM holds a subset of the block's relations
    rel_domains = [[0]] * len(pick) + [[0, 0], [0, 1]]
    z = [rng.integers(0, 5, n).astype(np.int32), rng.integers(0, 3, m).astype(np.int32)]
    hs = []
    for use_block in (False, True):
        h = eng.irm_create([n, m], rel_domains, max_tables=24)
        for k, col in enumerate(pick):
            if use_block:
                eng.set_observations_unary(h, k, blk, col)
            else:
                eng.set_observations(h, k, *unary[col])
        for k, (i, v) in enumerate(others):
            eng.set_observations(h, len(pick) + k, i, v)
        eng.finalize(h)
        for d in range(2):
            eng.set_assignments(h, d, z[d])
        hs.append(h)
    for r in range(len(rel_domains)):
        np.testing.assert_array_equal(eng.get_counts(hs[0], r), eng.get_counts(hs[1], r))
    for sweep in range(3):
        md = np.concatenate([np.full(k, d, np.int32) for d, k in enumerate((n, m))])
        mi = np.concatenate([rng.permutation(k).astype(np.int32) for k in (n, m)])
        res = eng.sweep(hs, [(md, mi), (md, mi)], seed=9, streams=[4, 4], counters=[sweep * len(mi)] * 2, trace_cap=32)
        np.testing.assert_array_equal(res["choice"][0], res["choice"][1])
        (na, la, pa), (nb, lb, pb) = res["trace"]
        np.testing.assert_array_equal(na, nb)
        valid = np.arange(la.shape[1])[None, :] < na[:, None]      # entries beyond a move's candidates are not written
        np.testing.assert_array_equal(la[valid], lb[valid])
        # the block's relations are scored from the count matrix in column order, the CSR entries as groups in cell order:
        # the same terms added in another order
        assert np.all(np.abs(pa[valid] - pb[valid]) <= 1e-12 * np.maximum(1.0, np.abs(pa[valid])))
        for r in range(len(rel_domains)):
            np.testing.assert_array_equal(eng.get_counts(hs[0], r), eng.get_counts(hs[1], r))
    assert len(np.unique(eng.get_assignments(hs[1], 0))) > 1
    # relation scores under another IRM's partitions read the block columns' COO arrays
    sc = eng.relation_score_matrix([([0], len(unary[c][1]), dev[c][0].data_ptr(), dev[c][1].data_ptr()) for c in pick[:4]], [hs[0], hs[1]])
    np.testing.assert_array_equal(sc[:, 0], sc[:, 1])
    dup_i, dup_v = _torch_coo(np.array([[1], [1]], np.int32), np.array([0, 1], np.uint8))
    with pytest.raises(E.EngineError, match="twice"):
        eng.unary_block_create(n, [(2, dup_i.data_ptr(), dup_v.data_ptr())])
    eng.close()


@pytest.mark.gpu
def test_relation_score_matrix_fresh_scores_and_prior_draw_vs_restatements():
    from hirm import engine as E
    eng = E.Engine(0)
    rng = np.random.default_rng(11)
    n_ent, kcap, seed = [300, 70, 1500], 16, 21
    rel_domains = [[0], [1], [0, 1], [0, 0], [2, 0, 1], [2], [1, 1]]
    obs = []
    for ds in rel_domains:
        nobs = int(rng.integers(150, 3000)) if len(ds) > 1 else n_ent[ds[0]]
        if len(ds) == 1:
            ids = np.arange(nobs, dtype=np.int32)[:, None]
        else:
            ids = np.unique(np.stack([rng.integers(0, n_ent[d], nobs) for d in ds], 1), axis=0).astype(np.int32)
        obs.append((ids, (rng.random(len(ids)) < 0.4).astype(np.uint8)))
    obs.append((np.zeros((0, 1), np.int32), np.zeros(0, np.uint8)))          # a relation without observations
    rel_domains.append([1])
    bufs = [_torch_coo(i, v) for i, v in obs]
    rels = [(ds, len(v), ti.data_ptr() if len(v) else None, tv.data_ptr() if len(v) else None) for ds, (i, v), (ti, tv) in zip(rel_domains, obs, bufs)]
    # candidate IRMs: partitions only
    cands, zs = [], []
    for c in range(5):
        h = eng.irm_create(n_ent, [], max_tables=kcap)
        eng.finalize(h)
        z = [rng.integers(0, int(rng.integers(1, kcap)), n).astype(np.int32) * 3 for n in n_ent]   # labels need not be dense
        for d in range(3):
            eng.set_assignments(h, d, z[d])
        cands.append(h); zs.append(z)
    m = eng.relation_score_matrix(rels, cands)
    assert m.shape == (len(rels), 5)
    for i, (ds, (ids, val)) in enumerate(zip(rel_domains, obs)):
        for j in range(5):
            want = SU.relation_score_numpy(ids, val, [zs[j][d] for d in ds])
            assert G.close(m[i, j], want), (i, j, m[i, j], want)
    # a large relation goes through several work items and the global-table path (kcap^3 cells > shared memory)
    big = np.unique(np.stack([rng.integers(0, n_ent[2], 200000), rng.integers(0, n_ent[0], 200000)], 1), axis=0).astype(np.int32)
    bval = (rng.random(len(big)) < 0.5).astype(np.uint8)
    tb = _torch_coo(big, bval)
    mb = eng.relation_score_matrix([([2, 0], len(bval), tb[0].data_ptr(), tb[1].data_ptr())], cands)
    for j in range(5):
        assert G.close(mb[0, j], SU.relation_score_numpy(big, bval, [zs[j][2], zs[j][0]]))
    # CRP prior draws: the kernel against its sequential restatement (n spans several 1024-customer rounds)
    for n, key in ((1, 5), (37, 6), (1500, 7), (3000, 1 << 41)):
        np.testing.assert_array_equal(eng.crp_prior_draw(n, key, seed, max_tables=64), SU.crp_prior_restated(n, 1.0, seed, key))
    with pytest.raises(E.EngineError):
        eng.crp_prior_draw(3000, 7, seed, max_tables=2)
    # fresh-table scores = score under the keyed draws
    keys = (1000 + np.arange(len(rels) * 3, dtype=np.uint64)).reshape(len(rels), 3)
    f = eng.relation_scores_fresh(rels, n_ent, [64] * 3, keys, seed)
    for i, (ds, (ids, val)) in enumerate(zip(rel_domains, obs)):
        z = {d: SU.crp_prior_restated(n_ent[d], 1.0, seed, int(keys[i, d])) for d in set(ds)}
        assert G.close(f[i], SU.relation_score_numpy(ids, val, [z[d] for d in ds])), i
    eng.close()


@pytest.mark.gpu
def test_engine_backend_matches_oracle_backend_move_for_move():
    """Same seeds => the CUDA engine and the CPU oracle walk the same HIRM trajectory: relation partition, every IRM's
    entity partitions (bit-exact), alphas, logp_score."""
    gen = dict(n_ent=(40, 30), n_unary=9, n_binary=4, n_clusters=3, seed=4)
    _, a = _run("oracle", 4, **gen)
    h, b = _run("engine", 4, **gen)
    _same_trail(a, b, rtol=1e-9)
    _check_invariants(h, b)
    # count tables of every IRM = histogram of (observations, z)
    domains, schema, obs, _ = SU.synthetic_hirm(**gen)
    st = h.state()
    for t, info in h.irms.items():
        for k, r in enumerate(sorted(info["rels"])):
            name = h.rnames[r]
            cells = h.backend.get_counts(info["handle"], k)
            ids, val = obs[name]
            assert cells[:, -2].sum() == len(val) and cells[:, -1].sum() == int(val.sum())
            zsel = [st[t]["z"][d] for d in schema[name]]
            for row in cells[:5]:
                sel = np.ones(len(val), bool)
                for p, z in enumerate(zsel):
                    sel &= z[ids[:, p]] == row[p]
                assert sel.sum() == row[-2] and val[sel].sum() == row[-1]


def _fixture_hirm(fx, **kw):
    """ShardedHIRM in the state the reference had before its traced relation moves (animals.unary: one shared,
    fully observed domain)."""
    st = fx["hirm"]["init"]
    rels = {}
    for irm in fx["irms"]:
        for r in irm["relations"]:
            a = np.asarray(r["obs"], dtype=np.int64).reshape(-1, len(r["domains"]) + 1)
            rels[r["name"]] = (list(r["domains"]), a[:, :-1].astype(np.int32), a[:, -1].astype(np.uint8))
    dnames = sorted({d for v in rels.values() for d in v[0]})
    n_codes = {d: 1 + max(int(v[1][:, v[0].index(d)].max()) for v in rels.values() if d in v[0]) for d in dnames}
    schema = {r: rels[r][0] for r in sorted(rels)}
    obs = {r: (rels[r][1], rels[r][2]) for r in schema}
    parts, alphas = {}, {}
    for ts, doms in st["irm_z"].items():
        parts[int(ts)], alphas[int(ts)] = {}, {}
        for d in dnames:
            z = np.full(n_codes[d], -1, np.int32)
            for item, label in doms[d]["items"]:
                z[item] = label
            assert z.min() >= 0
            parts[int(ts)][d], alphas[int(ts)][d] = z, doms[d]["alpha"]
    return ShardedHIRM(n_codes, schema, obs, relation_tables=st["relation_table"], partitions=parts, alphas=alphas, **kw), n_codes


@pytest.mark.gpu
def test_reference_relation_moves_replayed_through_the_score_matrix():
    fx = G.load("animals_unary_hirm")
    h, n_codes = _fixture_hirm(fx, seed=1, max_tables=32)
    assert G.close(h.logp_score(), fx["hirm"]["init"]["logp_score"])
    cols, fresh, keys = h.score_columns()                 # ONE batched launch + all-gather for the whole sweep of 85 moves
    changed = 0
    for mv in fx["hirm"]["rmoves"]:
        r = h.rindex[mv["r"]]
        assert h.rel_table[r] == mv["cur"]
        parts, fscore = None, 0.0
        if mv.get("fresh_z"):                             # the reference's own CRP-prior seats of its fresh IRM
            parts = {}
            for d, seats in mv["fresh_z"].items():
                z = np.full(n_codes[d], -1, np.int32)
                for it, lb in seats:
                    z[it] = lb
                parts[h.dnames.index(d)] = z
            ids, val = h.backend._buf[r][0].cpu().numpy().reshape(-1, len(h.rel_domains[r])), h.backend._buf[r][1].cpu().numpy()
            fscore = SU.relation_score_numpy(ids, val, [parts[d] for d in h.rel_domains[r]])
        tables, lp, fresh_label = h.relation_candidates(r, cols, fscore)
        assert tables == mv["tables"], (mv["r"], tables)
        assert G.close(lp, mv["lp"]), (mv["r"], lp, mv["lp"])
        from hirm.model import choose_index
        choice = tables[choose_index(lp, mv["u"])]
        assert choice == mv["choice"]
        changed += int(choice != mv["cur"])
        h.apply_relation_choice(r, choice, cols, parts=parts)
        if choice == fresh_label:
            assert G.close(cols[choice][r], fscore)        # the new IRM's column agrees with the fresh-table score
    assert changed > 0
    h.materialize()
    fin = fx["hirm"]["final"]
    assert {name: h.relation_to_table(name) for name in h.rnames} == fin["relation_table"]
    assert G.close(h.logp_score(), fin["logp_score"])
    h.iterate(2)                                           # and the chain keeps running with its own draws
    assert np.isfinite(h.logp_score())


def _nccl_worker(rank, world, port, out, iters, gen):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    h, trail = _run("engine", iters, comm=Comm(), device=rank, **gen)
    owners = sorted({info["owner"] for info in h.irms.values()})
    dist.barrier()
    if rank == 0:
        out.put((trail, owners, dict(h.stats), h.comm.n_broadcast))
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpus_nccl_match_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    gen = dict(n_ent=(40, 30), n_unary=9, n_binary=4, n_clusters=3, seed=4)
    iters = 4
    _, ref = _run("engine", iters, **gen)
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q, iters, gen)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    trail, owners, stats, n_bc = q.get()
    _same_trail(ref, trail, rtol=1e-9)
    # not a planted start: relations move, fresh IRMs are opened (broadcast of their score column), IRMs are rebuilt from the
    # replicated COO arrays and migrate between the GPUs
    assert stats["relations_moved"] > 0 and stats["fresh_irms"] > 0 and n_bc >= stats["fresh_irms"], stats
    assert stats["irms_rebuilt"] > 0 and stats["irms_migrated"] > 0, stats
