"""Rows grouped ahead of the moves (csrc/pregroup.cuh, hirm_engine_last_rounds): for a run of moves of a domain that every CSR
relation holds once, the group lists of all its moves are built by a separate kernel and the sequential kernel skips its
phases A and B1.  The lists are the ones phase B1 builds (ascending accumulator cell), so candidates, log-probabilities, choices,
count tables and scores must be bit-identical with the ordinary path -- for any cut of the move lists into pieces, with domains
that do not qualify (held twice by a relation) in the same list, with unary-block relations, with hashed tables, with clusters,
and the reference's traces must replay through it."""
import os

import numpy as np
import pytest

import engine_util as U
import golden_util as G
from hirm import engine as E

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = E.Engine()
    yield e
    e.close()


class _env:
    def __init__(self, **kw):
        self.kw = {k: str(v) for k, v in kw.items() if v is not None}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update(self.kw)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _sparse(rng, shape, p):
    grid = np.stack(np.meshgrid(*[np.arange(k) for k in shape], indexing="ij"), -1).reshape(-1, len(shape))
    ids = grid[rng.random(len(grid)) < p].astype(np.int32)
    return ids, rng.integers(0, 2, len(ids)).astype(np.uint8)


def _irms(eng, n_irms, rng, with_unary=True, max_tables=16):
    """a, b once everywhere (their rows can be grouped ahead); c is held twice by T(c, c): it cannot."""
    import torch
    na, nb, nc = 70, 50, 40
    rels = [([0, 1, 2], _sparse(rng, (na, nb, nc), 0.05)), ([0, 1], _sparse(rng, (na, nb), 0.5)), ([1, 0], _sparse(rng, (nb, na), 0.3)),
            ([2, 2], _sparse(rng, (nc, nc), 0.4)), ([2, 0], _sparse(rng, (nc, na), 0.3))]
    unary, dev, blk = [], [], None
    if with_unary:
        for k in range(12):
            keep = np.flatnonzero(rng.random(na) < 0.8).astype(np.int32)
            unary.append((keep[:, None], rng.integers(0, 2, len(keep)).astype(np.uint8)))
        dev = [(torch.from_numpy(i.reshape(-1)).cuda(), torch.from_numpy(v).cuda()) for i, v in unary]
        blk = eng.unary_block_create(na, [(len(v), ti.data_ptr(), tv.data_ptr()) for (i, v), (ti, tv) in zip(unary, dev)])
    rel_domains = [[0]] * len(unary) + [ds for ds, _ in rels]
    hs = []
    for k in range(n_irms):
        h = eng.irm_create([na, nb, nc], rel_domains, max_tables=max_tables)
        for c in range(len(unary)):
            eng.set_observations_unary(h, c, blk, c)
        for j, (ds, (ids, val)) in enumerate(rels):
            eng.set_observations(h, len(unary) + j, ids, val)
        eng.finalize(h)
        hs.append(h)
    return hs, (na, nb, nc), len(rel_domains), dev


def _snapshot(eng, hs, res, n_rel, trace=True):
    out = []
    for k, h in enumerate(hs):
        item = [res["choice"][k].copy()]
        if trace:
            nn, ll, pp = res["trace"][k]
            valid = np.arange(ll.shape[1])[None, :] < nn[:, None]
            item += [nn.copy(), ll[valid].copy(), pp[valid].copy()]
        item += [[eng.get_assignments(h, d).copy() for d in range(3)], [eng.get_counts(h, r).copy() for r in range(n_rel)], eng.logp_score(h)]
        out.append(item)
    return out


def _same(a, b):
    for x, y in zip(a, b):
        for u, v in zip(x[:-3], y[:-3]):
            np.testing.assert_array_equal(u, v)                 # choices, candidates, log-probabilities bit for bit
        for u, v in zip(x[-3], y[-3]):
            np.testing.assert_array_equal(u, v)
        for u, v in zip(x[-2], y[-2]):
            np.testing.assert_array_equal(u, v)
        assert x[-1] == y[-1]


@pytest.mark.parametrize("with_unary,hashed,max_tables", [(True, False, 16), (False, True, 16), (False, False, 60)])
def test_pregrouped_rounds_give_identical_trajectories(eng, with_unary, hashed, max_tables):
    """(60 table slots: an accumulator space of ~3800 cells, four per thread of the pre-grouping kernel -- its vector path)"""
    rng = np.random.default_rng(23)
    extra = {"HIRM_HASHED_TABLES": 1} if hashed else {}
    with _env(**extra):
        hs, (na, nb, nc), n_rel, keep = _irms(eng, 4, rng, with_unary=with_unary, max_tables=max_tables)
    z0 = [[rng.integers(0, 5, na).astype(np.int32), rng.integers(0, 4, nb).astype(np.int32), rng.integers(0, 3, nc).astype(np.int32)] for _ in hs]
    # per IRM its own list: domain blocks (two sweeps), one IRM with a short run and an interleaved stretch in the middle
    lists = []
    for k in range(len(hs)):
        md = np.concatenate([np.full(na, 0), np.full(nb, 1), np.full(nc, 2), np.full(nb, 1), np.full(na, 0)]).astype(np.int32)
        mi = np.concatenate([rng.permutation(na), rng.permutation(nb), rng.permutation(nc), rng.permutation(nb), rng.permutation(na)]).astype(np.int32)
        if k == 1:
            md = np.concatenate([md[:30], np.array([1, 0, 2, 0, 1, 1, 0], np.int32), md[30:]])
            mi = np.concatenate([mi[:30], np.array([3, 4, 5, 4, 3, 7, 9], np.int32), mi[30:]])
        if k == 2:
            md, mi = md[:na + 10], mi[:na + 10]
        lists.append((md, mi))
    runs = {}
    for name, env in [("plain", dict(HIRM_PREGROUP=0)), ("one piece per run", dict(HIRM_PREGROUP=1)), ("pieces of 17", dict(HIRM_PREGROUP=1, HIRM_PREGROUP_SEG=17)),
                      ("pieces of 17, clusters of 4", dict(HIRM_PREGROUP=1, HIRM_PREGROUP_SEG=17, HIRM_CLUSTER_SIZE=4))]:
        with _env(**env):
            for h, z in zip(hs, z0):
                for d in range(3):
                    eng.set_assignments(h, d, z[d])
            res = eng.sweep(hs, lists, seed=9, streams=list(range(len(hs))), counters=[0] * len(hs), trace_cap=20)
            info = eng.last_sweep_info()
        assert info["pregrouped"] == (name != "plain")
        if "17" in name:
            assert info["rounds"] >= 2 * (na // 17)
        runs[name] = _snapshot(eng, hs, res, n_rel)
    assert any(not np.array_equal(o[4][0], z0[k][0]) for k, o in enumerate(runs["plain"]))
    for name in runs:
        _same(runs["plain"], runs[name])
    for h in hs:
        eng.irm_destroy(h)


def test_pregrouped_sweep_all_matches_plain(eng):
    """hirm_engine_sweep_all (order generated on the device, Philox uniforms): same moves, same choices, same state."""
    rng = np.random.default_rng(5)
    hs, (na, nb, nc), n_rel, keep = _irms(eng, 3, rng)
    z0 = [[rng.integers(0, 5, na).astype(np.int32), rng.integers(0, 4, nb).astype(np.int32), rng.integers(0, 3, nc).astype(np.int32)] for _ in hs]
    outs = []
    for env in (dict(HIRM_PREGROUP=0), dict(HIRM_PREGROUP=1, HIRM_PREGROUP_SEG=25)):
        with _env(**env):
            for k, (h, z) in enumerate(zip(hs, z0)):
                for d in range(3):
                    eng.set_assignments(h, d, z[d])
                eng.set_rng(h, 100 + k, 0)
                eng.set_rng_counters(h, 0, 0)
            n, dom, item, choice = eng.sweep_all(hs, n_sweeps=3, seed=77, want_moves=True)
            info = eng.last_sweep_info()
        assert info["pregrouped"] == ("HIRM_PREGROUP_SEG" in env)
        outs.append((dom.copy(), item.copy(), choice.copy(), [[eng.get_assignments(h, d).copy() for d in range(3)] for h in hs],
                     [[eng.get_counts(h, r).copy() for r in range(n_rel)] for h in hs], [eng.logp_score(h) for h in hs]))
    a, b = outs
    for u, v in zip(a[:3], b[:3]):
        np.testing.assert_array_equal(u, v)
    for x, y in zip(a[3], b[3]):
        for u, v in zip(x, y):
            np.testing.assert_array_equal(u, v)
    for x, y in zip(a[4], b[4]):
        for u, v in zip(x, y):
            np.testing.assert_array_equal(u, v)
    assert a[5] == b[5]
    for h in hs:
        eng.irm_destroy(h)


@pytest.mark.parametrize("case", ["animals_unary_irm", "nations_binary_py", "synth_mixed_irm", "synth_ternary_irm"])
def test_reference_traces_replay_with_pregrouping_forced(eng, case):
    with _env(HIRM_PREGROUP=1, HIRM_PREGROUP_SEG=8):
        for irm in G.load(case)["irms"]:
            spec = G.IrmSpec(irm)
            h = U.engine_from_spec(eng, spec)
            res = G.replay(E.IrmBackend(eng, h), spec)
            assert res["moves"] > 0 and res["changed"] > 0
            eng.irm_destroy(h)
