"""Thread-block cluster per IRM (north_star subsystem 3, HIRM_SWEEP_CLUSTER): up to 16 CTAs share one move's row (phase A)
and its (candidate x group) scoring units (phase D) through distributed shared memory.  The sums are made in the same
order for every cluster size, so the cluster launch must give bit-identical candidates, log-probabilities, choices and
count tables -- and replay the reference's traces like the plain launch."""
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


@pytest.mark.parametrize("case", ["animals_unary_irm", "nations_binary_py", "synth_mixed_irm", "synth_ternary_irm"])
def test_reference_traces_replay_on_cluster_launches(eng, case):
    for irm in G.load(case)["irms"]:
        spec = G.IrmSpec(irm)
        h = U.engine_from_spec(eng, spec)
        res = G.replay(E.IrmBackend(eng, h, flags=E.SWEEP_CLUSTER), spec)
        assert res["moves"] > 0 and res["changed"] > 0
        assert eng.last_sweep_info()["cluster_size"] > 1
        eng.irm_destroy(h)


def _mixed_irms(eng, n_irms, rng, max_tables=24):
    """IRMs with unary-block relations, a sparse e x e relation (repeated domain), e x f and a ternary f e e relation."""
    import torch
    n, m = 120, 30
    unary = []
    for k in range(40):
        keep = np.flatnonzero(rng.random(n) < 0.9).astype(np.int32)
        unary.append((keep[:, None], rng.integers(0, 2, len(keep)).astype(np.uint8)))
    dev = [(torch.from_numpy(i.reshape(-1)).cuda(), torch.from_numpy(v).cuda()) for i, v in unary]
    blk = eng.unary_block_create(n, [(len(v), ti.data_ptr(), tv.data_ptr()) for (i, v), (ti, tv) in zip(unary, dev)])

    def sparse(shape, p):
        grid = np.stack(np.meshgrid(*[np.arange(k) for k in shape], indexing="ij"), -1).reshape(-1, len(shape))
        ids = grid[rng.random(len(grid)) < p].astype(np.int32)
        return ids, rng.integers(0, 2, len(ids)).astype(np.uint8)
    others = [([0, 0], sparse((n, n), 0.3)), ([0, 1], sparse((n, m), 0.6)), ([1, 0, 0], sparse((m, n, n), 0.01))]
    rel_domains = [[0]] * len(unary) + [ds for ds, _ in others]
    hs = []
    for k in range(n_irms):
        h = eng.irm_create([n, m], rel_domains, max_tables=max_tables)
        for c in range(len(unary)):
            eng.set_observations_unary(h, c, blk, c)
        for j, (ds, (ids, val)) in enumerate(others):
            eng.set_observations(h, len(unary) + j, ids, val)
        eng.finalize(h)
        hs.append(h)
    return hs, (n, m), len(rel_domains), dev


@pytest.mark.parametrize("n_irms", [1, 3])
def test_cluster_sizes_give_identical_trajectories(eng, n_irms):
    rng = np.random.default_rng(11)
    hs, (n, m), n_rel, keep = _mixed_irms(eng, n_irms, rng)
    z0 = [[rng.integers(0, 6, n).astype(np.int32), rng.integers(0, 4, m).astype(np.int32)] for _ in hs]
    md = np.concatenate([np.full(n, 0, np.int32), np.full(m, 1, np.int32)] * 2)
    mi = np.concatenate([rng.permutation(n), rng.permutation(m), rng.permutation(n), rng.permutation(m)]).astype(np.int32)
    runs = {}
    for csize in (1, 2, 4, 8, 16):
        os.environ["HIRM_CLUSTER_SIZE"] = str(csize)
        try:
            for h, z in zip(hs, z0):
                for d in range(2):
                    eng.set_assignments(h, d, z[d])
            res = eng.sweep(hs, [(md, mi)] * len(hs), seed=5, streams=list(range(len(hs))), counters=[0] * len(hs), trace_cap=32)
            assert eng.last_sweep_info()["cluster_size"] == csize
        finally:
            del os.environ["HIRM_CLUSTER_SIZE"]
        out = []
        for k, h in enumerate(hs):
            nn, ll, pp = res["trace"][k]
            valid = np.arange(ll.shape[1])[None, :] < nn[:, None]
            out.append((res["choice"][k].copy(), nn.copy(), ll[valid].copy(), pp[valid].copy(),
                        [eng.get_assignments(h, d).copy() for d in range(2)], [eng.get_counts(h, r).copy() for r in range(n_rel)],
                        eng.logp_score(h)))
        runs[csize] = out
    assert any(len(np.unique(o[4][0])) != len(np.unique(z0[k][0])) or not np.array_equal(o[4][0], z0[k][0]) for k, o in enumerate(runs[1]))
    for csize in (2, 4, 8, 16):
        for a, b in zip(runs[1], runs[csize]):
            np.testing.assert_array_equal(a[0], b[0])
            np.testing.assert_array_equal(a[1], b[1])
            np.testing.assert_array_equal(a[2], b[2])
            np.testing.assert_array_equal(a[3], b[3])           # log-probabilities bit for bit
            for d in range(2):
                np.testing.assert_array_equal(a[4][d], b[4][d])
            for r in range(n_rel):
                np.testing.assert_array_equal(a[5][r], b[5][r])
            assert a[6] == b[6]
    for h in hs:
        eng.irm_destroy(h)


@pytest.mark.parametrize("n_rel,nobs", [(40, 8000), (110, 3000)])
def test_cluster_launch_with_big_cell_spaces(eng, n_rel, nobs):
    """Domains with a big accumulator cell space (kcap 255 on e x e relations: 2 * 255 + 1 cells per relation): 40 relations
    still fit the shared accumulators of a cluster CTA, 110 do not and take the IRM's global accumulator."""
    rng = np.random.default_rng(3)
    n = 3000
    rels = []
    for k in range(n_rel):
        ii = rng.integers(0, n, nobs); jj = rng.integers(0, n, nobs)
        key = np.unique(ii.astype(np.int64) * n + jj)
        ids = np.stack([key // n, key % n], 1).astype(np.int32)
        rels.append((ids, rng.integers(0, 2, len(ids)).astype(np.uint8)))
    z = rng.integers(0, 9, n).astype(np.int32)
    order = rng.permutation(n)[:400].astype(np.int32)
    outs = []
    for csize in (1, 8):
        os.environ["HIRM_CLUSTER_SIZE"] = str(csize)
        try:
            h = eng.irm_create([n], [[0, 0]] * len(rels), max_tables=255)
            for k, (ids, val) in enumerate(rels):
                eng.set_observations(h, k, ids, val)
            eng.finalize(h)
            eng.set_assignments(h, 0, z)
            res = eng.sweep([h], [(np.zeros_like(order), order)], seed=2, streams=[1], counters=[0], trace_cap=24)
        finally:
            del os.environ["HIRM_CLUSTER_SIZE"]
        nn, ll, pp = res["trace"][0]
        valid = np.arange(ll.shape[1])[None, :] < nn[:, None]
        outs.append((res["choice"][0].copy(), pp[valid].copy(), eng.get_assignments(h, 0).copy(), [eng.get_counts(h, r).copy() for r in (0, 17, n_rel - 1)]))
        eng.irm_destroy(h)
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    np.testing.assert_array_equal(outs[0][1], outs[1][1])
    np.testing.assert_array_equal(outs[0][2], outs[1][2])
    for a, b in zip(outs[0][3], outs[1][3]):
        np.testing.assert_array_equal(a, b)


def test_candidate_batches_give_identical_results(eng):
    """More (candidate x chunk) scoring units than the partial-sum buffer holds (an IRM with dozens of tables early in a chain): the
    candidates are scored in batches that alternate between the halves of the buffer.  Forced here with 64-thread CTAs (a buffer
    of 256 units) and 30 tables; results must equal the single-batch launch bit for bit, with and without a cluster."""
    rng = np.random.default_rng(21)
    hs, (n, m), n_rel, keep = _mixed_irms(eng, 2, rng, max_tables=40)
    z0 = [[rng.integers(0, 30, n).astype(np.int32), rng.integers(0, 12, m).astype(np.int32)] for _ in hs]
    md = np.concatenate([np.full(n, 0, np.int32), np.full(m, 1, np.int32)])
    mi = np.concatenate([rng.permutation(n), rng.permutation(m)]).astype(np.int32)
    runs = []
    for env in ({}, {"HIRM_GENERIC_THREADS": "64"}, {"HIRM_GENERIC_THREADS": "64", "HIRM_CLUSTER_SIZE": "4"}):
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            for h, z in zip(hs, z0):
                for d in range(2):
                    eng.set_assignments(h, d, z[d])
            res = eng.sweep(hs, [(md, mi)] * len(hs), seed=8, streams=list(range(len(hs))), counters=[0] * len(hs), trace_cap=48)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        out = []
        for k, h in enumerate(hs):
            nn, ll, pp = res["trace"][k]
            assert nn.max() >= 25                               # dozens of candidates: 64-thread CTAs need several batches
            valid = np.arange(ll.shape[1])[None, :] < nn[:, None]
            out.append((res["choice"][k].copy(), nn.copy(), ll[valid].copy(), pp[valid].copy(), eng.get_assignments(h, 0).copy(),
                        [eng.get_counts(h, r).copy() for r in range(n_rel)], eng.logp_score(h)))
        runs.append(out)
    for other in runs[1:]:
        for a, b in zip(runs[0], other):
            for u, v in zip(a[:5], b[:5]):
                np.testing.assert_array_equal(u, v)
            for u, v in zip(a[5], b[5]):
                np.testing.assert_array_equal(u, v)
            assert a[6] == b[6]
    for h in hs:
        eng.irm_destroy(h)
