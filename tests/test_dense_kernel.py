"""GPU parity tests of the dense R(e,e) kernel's less-travelled paths, each against the CPU oracle on
the same seeded inputs: the private-bins accumulation (> 8 live tables) with a mid-row flush, table
births / deaths with slot compaction, the two-tier hand-over (a chain that outgrows the small
shared-memory table is resumed by the second launch), repeated and out-of-range moves."""
import os

import numpy as np
import pytest

import engine_util as U
import golden_util as G
from hirm import engine as E
from oracle import hirm_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = E.Engine()
    yield e
    e.close()


class env:
    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update({k: str(v) for k, v in self.kw.items()})

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def build(eng, n, z0, seed, missing=0.0, kcap=32):
    x = U.planted_ree(n, seed=seed, missing=missing)
    ids, val = U.dense_to_coo(x)
    o = O.Oracle([n], [[0, 0]])
    o.set_observations(0, ids, val)
    o.set_assignments(0, z0)
    h = eng.irm_create([n], [[0, 0]], max_tables=kcap)
    eng.set_observations_dense_host(h, 0, x)
    eng.finalize(h)
    eng.set_assignments(h, 0, z0)
    return o, h


def run_and_compare(eng, o, h, order, seed, stream, trace_cap=72):
    dom = np.zeros_like(order)
    res = eng.sweep([h], [(dom, order)], seed=seed, streams=[stream], counters=[0], trace_cap=trace_cap)
    assert eng.last_sweep_info()["kernel_kind"] == 1
    un = [O.uniform(seed, stream, m) for m in range(len(order))]
    U.assert_traces_match(res["trace"][0], o, dom, order, un)
    np.testing.assert_array_equal(eng.get_assignments(h, 0), o.get_assignments(0))
    np.testing.assert_array_equal(eng.get_counts(h, 0), o.get_counts(0))
    assert G.close(eng.logp_score(h), o.logp_score())
    return res


@pytest.mark.parametrize("missing", [0.0, 0.2])
def test_many_tables_bins_path_births_and_deaths(eng, missing):
    """20 random tables over 500 entities collapse towards the planted blocks: the bins path (K > 8),
    then the dp4a paths (K <= 8, 4, 2), many table deaths (slot compaction) and a few births."""
    n = 500
    z0 = np.random.default_rng(5).integers(0, 20, n).astype(np.int32) * 3 + 1     # labels with holes
    o, h = build(eng, n, z0, seed=11, missing=missing)
    rng = np.random.default_rng(8)
    order = np.concatenate([rng.permutation(n) for _ in range(3)]).astype(np.int32)
    res = run_and_compare(eng, o, h, order, seed=3, stream=9)
    assert len(np.unique(eng.get_assignments(h, 0))) < 20
    eng.irm_destroy(h)


def test_bins_flush_inside_a_row(eng):
    """40 tables, 64-byte ring chunks: the 8-bit private bins are flushed in the middle of a row."""
    n = 1200
    z0 = np.random.default_rng(6).integers(0, 40, n).astype(np.int32)
    with env(HIRM_DENSE_CHUNK=64):
        o, h = build(eng, n, z0, seed=12, kcap=48)
        order = np.random.default_rng(9).permutation(n).astype(np.int32)[:300]
        run_and_compare(eng, o, h, order, seed=4, stream=2)
        assert eng.last_dense_plan()["chunk_bytes"] == 64
    eng.irm_destroy(h)


@pytest.mark.parametrize("tier_slots", [3, 6])
def test_tier_handover_matches_oracle(eng, tier_slots):
    """A first tier with `tier_slots` shared-memory table slots: chains stop when they need one more table and
    the second launch resumes them at the same move with the same Philox counters -- the trajectory is the
    oracle's, bit for bit."""
    n = 300
    z0 = U.crp_prior_labels(n, 1.5, np.random.default_rng(14))
    with env(HIRM_DENSE_TIER_SLOTS=tier_slots):
        o, h = build(eng, n, z0, seed=13)
        rng = np.random.default_rng(10)
        order = np.concatenate([rng.permutation(n) for _ in range(2)]).astype(np.int32)
        run_and_compare(eng, o, h, order, seed=5, stream=1)
        assert eng.last_sweep_info()["launches"] == 2
    eng.irm_destroy(h)


def test_tier_handover_many_chains_device_resident(eng):
    """Eight chains over one dataset, on-device Philox, tier hand-over in the middle: each chain equals its
    own oracle chain."""
    n, chains = 160, 8
    x = U.planted_ree(n, seed=3)
    ids, val = U.dense_to_coo(x)
    rng = np.random.default_rng(4)
    hs, zs = [], []
    for c in range(chains):
        h = eng.irm_create([n], [[0, 0]], max_tables=24)
        if c == 0:
            eng.set_observations_dense_host(h, 0, x)
            eng.finalize(h)
        else:
            eng.share_observations(h, hs[0])
        z = rng.integers(0, 2 + c, n).astype(np.int32)
        eng.set_assignments(h, 0, z)
        hs.append(h); zs.append(z)
    order = np.concatenate([rng.permutation(n) for _ in range(2)]).astype(np.int32)
    mv = (np.zeros_like(order), order)
    with env(HIRM_DENSE_TIER_SLOTS=4):
        eng.sweep(hs, [mv] * chains, seed=21, streams=np.arange(chains, dtype=np.uint64) + 5, counters=np.full(chains, 7, np.uint64))
    for c in range(chains):
        o = O.Oracle([n], [[0, 0]])
        o.set_observations(0, ids, val)
        o.set_assignments(0, zs[c])
        o.sweep(mv[0], mv[1], seed=21, stream=5 + c, offset=7)
        np.testing.assert_array_equal(eng.get_assignments(hs[c], 0), o.get_assignments(0))
        np.testing.assert_array_equal(eng.get_counts(hs[c], 0), o.get_counts(0))
    for h in hs:
        eng.irm_destroy(h)


def test_repeated_and_bad_moves(eng):
    """The same entity moved several times in a row (the fix-up of the previous move IS the diagonal), and
    an out-of-range move in the middle of a list (reported, the other moves are still made)."""
    n = 64
    z0 = np.random.default_rng(1).integers(0, 5, n).astype(np.int32)
    o, h = build(eng, n, z0, seed=2, missing=0.1)
    order = np.asarray([5, 5, 5, 6, 5, 6, 6, 0, 63, 63, 1, 0, 0], dtype=np.int32)
    run_and_compare(eng, o, h, order, seed=8, stream=4)
    bad = np.asarray([3, 64, 4], dtype=np.int32)
    with pytest.raises(E.EngineError):
        eng.sweep([h], [(np.zeros_like(bad), bad)], seed=9, streams=[1], counters=[0])
    o.move(0, 3, O.uniform(9, 1, 0)); o.move(0, 4, O.uniform(9, 1, 2))
    np.testing.assert_array_equal(eng.get_assignments(h, 0), o.get_assignments(0))
    np.testing.assert_array_equal(eng.get_counts(h, 0), o.get_counts(0))
    eng.irm_destroy(h)


def test_max_tables_exhausted_is_reported(eng):
    n = 40
    z0 = (np.arange(n) % 4).astype(np.int32)
    o, h = build(eng, n, z0, seed=3, kcap=4)
    with pytest.raises(E.EngineError) as ei:
        eng.sweep([h], [(np.zeros(3, np.int32), np.arange(3, dtype=np.int32))], seed=1, streams=[0], counters=[0])
    assert "max_tables" in str(ei.value)
    eng.irm_destroy(h)


def test_sweep_all_device_generated_order_and_batched_alpha_and_scores(eng):
    """hirm_engine_sweep_all (order generated on the device), hirm_engine_transition_alpha (Philox) and
    hirm_engine_logp_scores over several chains: the oracle, fed the generated order and the same Philox
    positions, ends in the same state; every sweep visits every entity exactly once."""
    n, chains, sweeps = 120, 3, 2
    x = U.planted_ree(n, seed=31, missing=0.05)
    ids, val = U.dense_to_coo(x)
    rng = np.random.default_rng(32)
    hs, os_ = [], []
    for c in range(chains):
        h = eng.irm_create([n], [[0, 0]], max_tables=32)
        if c == 0:
            eng.set_observations_dense_host(h, 0, x)
            eng.finalize(h)
        else:
            eng.share_observations(h, hs[0])
        z = rng.integers(0, 3 + c, n).astype(np.int32)
        eng.set_assignments(h, 0, z)
        eng.set_rng(h, 100 + c, 17 * c)
        o = O.Oracle([n], [[0, 0]])
        o.set_observations(0, ids, val)
        o.set_assignments(0, z)
        hs.append(h); os_.append(o)
    for rnd in range(2):
        nm, dom, item, choice = eng.sweep_all(hs, n_sweeps=sweeps, seed=55, want_moves=True)
        assert nm == chains * sweeps * n
        for c in range(chains):
            sl = slice(c * sweeps * n, (c + 1) * sweeps * n)
            for s in range(sweeps):
                assert sorted(item[sl][s * n:(s + 1) * n]) == list(range(n))        # a permutation per sweep
            os_[c].sweep(dom[sl], item[sl], seed=55, stream=100 + c, offset=17 * c + rnd * sweeps * n)
            np.testing.assert_array_equal(eng.get_assignments(hs[c], 0), os_[c].get_assignments(0))
            np.testing.assert_array_equal(eng.get_counts(hs[c], 0), os_[c].get_counts(0))
        # alpha moves: Philox(seed ^ "alpha", stream, counter * MAX_DOMAINS + d)
        alphas = eng.transition_alpha(hs, ngrid=20, seed=55)
        for c in range(chains):
            u = O.uniform(55 ^ 0x616c706861, 100 + c, rnd * E.MAX_DOMAINS + 0)
            a, _, _ = os_[c].transition_alpha(0, u, 20)
            assert abs(alphas[c, 0] - a) <= 1e-12 * a
            assert eng.get_alpha(hs[c], 0) == alphas[c, 0]
        sc = eng.logp_scores(hs)
        for c in range(chains):
            assert G.close(sc[c], os_[c].logp_score())
            lab, cnt = eng.get_tables(hs[c], 0)
            zz = os_[c].get_assignments(0)
            np.testing.assert_array_equal(lab, np.unique(zz))
            np.testing.assert_array_equal(cnt, [int((zz == t).sum()) for t in lab])
    for h in hs:
        eng.irm_destroy(h)


def test_sweep_all_generic_kernel_multi_domain(eng):
    """sweep_all on a mixed-arity CSR IRM (generic kernel): per sweep every domain in index order, every entity once."""
    spec = G.IrmSpec(G.load("synth_mixed_irm")["irms"][0])
    h = U.engine_from_spec(eng, spec)
    o = U.oracle_from_spec(spec)
    eng.set_rng(h, 9, 0)
    nm, dom, item, choice = eng.sweep_all([h], n_sweeps=2, seed=4, want_moves=True)
    per = sum(spec.n_entities)
    assert nm == 2 * per and list(dom[:per]) == sorted(dom[:per])
    o.sweep(dom, item, seed=4, stream=9, offset=0)
    for d in range(len(spec.dnames)):
        np.testing.assert_array_equal(eng.get_assignments(h, d), o.get_assignments(d))
    for r in range(len(spec.rnames)):
        np.testing.assert_array_equal(eng.get_counts(h, r), o.get_counts(r))
    eng.irm_destroy(h)


def test_full_size_properties_and_scores():
    """BASELINE config 3 at full size (n = 20000, 4e8 cells), where the oracle cannot follow: (1) after a sweep the
    count table is still the histogram of (observations, z), bit for bit; (2) cell counts add up to n^2; (3) the
    candidate log-probabilities of single moves equal an independent numpy / scipy evaluation of
    logp_gibbs_exact's closed form from the table and the entity's row and column (1e-9 relative: this is where
    the kernel's own fp64 log runs on counts of 1e8); (4) logp_score equals the sum over the table."""
    import torch
    from scipy.special import gammaln
    import bench
    n, kcap = 20000, 64
    pitch = (n + 15) // 16 * 16
    xh = torch.full((n, pitch), 0xFF, dtype=torch.uint8).pin_memory()
    bench.planted_ree_into(xh.numpy(), n, seed=3)
    x = xh.numpy()[:, :n]
    e = E.Engine()
    h = e.irm_create([n], [[0, 0]], max_tables=kcap)
    e.set_observations_dense_host_ptr(h, 0, xh.data_ptr(), pitch)
    e.finalize(h)
    rng = np.random.default_rng(5)
    z0 = (np.arange(n) >= n // 2).astype(np.int32)
    z0[rng.choice(n, 40, replace=False)] = 2          # a small third table
    z0[rng.choice(n, 300, replace=False)] = rng.integers(0, 2, 300)   # some entities in the wrong block
    e.set_assignments(h, 0, z0)

    def table_from(z):
        K = z.max() + 1
        oh = np.zeros((n, K), dtype=np.float64); oh[np.arange(n), z] = 1
        cnt_ = oh.sum(0)
        N = (cnt_[:, None] * cnt_[None, :]).astype(np.int64)
        S = np.zeros((K, K))
        for r0 in range(0, n, 1000):                   # float64 sums of 0/1 values are exact
            S += oh[r0:r0 + 1000].T @ (x[r0:r0 + 1000].astype(np.float64) @ oh)
        return N, np.rint(S).astype(np.int64)

    def S_(N, s):
        return gammaln(s + 1) + gammaln(N - s + 1) - gammaln(N + 2)

    import mpmath
    mpmath.mp.dps = 40

    def S_mp(N, s):          # the same sum in 40 digits: differences of 1e9-sized sums need more than fp64
        tot = mpmath.mpf(0)
        for a, b in zip(np.ravel(N), np.ravel(s)):
            a, b = int(a), int(b)
            tot += mpmath.loggamma(b + 1) + mpmath.loggamma(a - b + 1) - mpmath.loggamma(a + 2)
        return tot

    # (3) single moves, score only
    z = e.get_assignments(h, 0)
    N, S = table_from(z)
    cnt = np.bincount(z)
    for item in [0, 7, int(np.flatnonzero(z == 2)[0]), n - 1]:
        labels, lp = E.IrmBackend(e, h).score(0, item)
        c = z[item]
        zi = np.delete(z, item)
        row, col = np.delete(x[item, :], item), np.delete(x[:, item], item)
        K = N.shape[0]
        rn = np.bincount(zi, minlength=K); rs = np.bincount(zi, weights=row, minlength=K).astype(np.int64)
        cs = np.bincount(zi, weights=col, minlength=K).astype(np.int64)
        d = int(x[item, item])
        Nr, Sr = N.copy(), S.copy()                    # the table without the entity's observations
        Nr[c, :] -= rn; Sr[c, :] -= rs; Nr[:, c] -= rn; Sr[:, c] -= cs; Nr[c, c] -= 1; Sr[c, c] -= d
        want = []
        for t in labels:
            if t < K:
                Nn, Sn = Nr.copy(), Sr.copy()
                Nn[t, :] += rn; Sn[t, :] += rs; Nn[:, t] += rn; Sn[:, t] += cs; Nn[t, t] += 1; Sn[t, t] += d
                w = cnt[t] - (t == c)
                want.append(np.log(w if w > 0 else 1.0) + float(S_mp(Nn, Sn) - S_mp(Nr, Sr)))
            else:                                      # the fresh table: row and column into empty cells
                want.append(0.0 + float(S_mp(rn, rs) + S_mp(rn, cs) + S_mp([1], [d])))
        want = np.asarray(want)
        k = list(labels).index(c)
        assert G.close((lp - lp[k]), (want - want[k]), rtol=1e-9), (item, lp, want)
    # (1), (2), (4) after a sweep of 3000 moves
    order = rng.permutation(n).astype(np.int32)[:3000]
    e.sweep([h], [(np.zeros_like(order), order)], seed=2, streams=[0], counters=[0])
    z = e.get_assignments(h, 0)
    got = e.get_counts(h, 0)
    lab, zc = np.unique(z, return_inverse=True)
    N, S = table_from(zc)
    want = np.asarray([[lab[a], lab[b], N[a, b], S[a, b]] for a in range(N.shape[0]) for b in range(N.shape[1]) if N[a, b] > 0], dtype=np.int64)
    np.testing.assert_array_equal(got.astype(np.int64), want)
    assert int(got[:, 2].sum()) == n * n
    crp = len(lab) * np.log(1.0) + gammaln(np.bincount(zc)).sum() + gammaln(1.0) - gammaln(n + 1.0)
    assert G.close(e.logp_score(h), S_(got[:, 2].astype(np.int64), got[:, 3].astype(np.int64)).sum() + crp, rtol=1e-12)
    e.close()
