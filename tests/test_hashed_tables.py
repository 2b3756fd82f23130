"""Hashed count tables (north_star subsystem 1; Relation::clusters is a sparse map in the reference, cxx/hirm.hh:246, with
create-on-demand :540-543 and erase-on-zero :534-537): the same data through the dense and the hashed store must give
bit-identical count tables and trajectories; relations whose index space exceeds the dense budget (2^27 cells) load and
sweep on the hashed store automatically."""
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


def _build(eng, n_entities, rel_domains, obs, z, mode, max_tables):
    h = eng.irm_create(n_entities, rel_domains, max_tables=max_tables)
    for r, (ids, val) in enumerate(obs):
        eng.set_table_mode(h, r, mode)
        eng.set_observations(h, r, ids, val)
    eng.finalize(h)
    for d in range(len(n_entities)):
        eng.set_assignments(h, d, z[d])
    return h


@pytest.mark.parametrize("case", ["synth_ternary_irm", "synth_mixed_irm", "nations_binary_py"])
def test_reference_traces_replay_on_hashed_tables(eng, case):
    """The committed reference traces (candidate labels, log-probabilities, same-u choices, count tables at every checkpoint)
    with every relation on the hashed store."""
    for irm in G.load(case)["irms"]:
        spec = G.IrmSpec(irm)
        h = _build(eng, spec.n_entities, spec.rel_domains, list(zip(spec.obs_ids, spec.obs_val)), spec.init_z, E.TABLE_HASHED, 24)
        for d in range(len(spec.dnames)):
            eng.set_alpha(h, d, spec.init_alpha[d])
        res = G.replay(E.IrmBackend(eng, h), spec)
        assert res["moves"] > 0 and res["changed"] > 0
        eng.irm_destroy(h)


def _ternary(rng, n=(60, 50, 40), nobs=9000):
    cells = rng.choice(n[0] * n[1] * n[2], size=nobs, replace=False)
    ids = np.stack([cells // (n[1] * n[2]), (cells // n[2]) % n[1], cells % n[2]], axis=1).astype(np.int32)
    return ids, rng.integers(0, 2, nobs).astype(np.uint8)


def test_dense_and_hashed_tables_give_identical_trajectories(eng):
    """Config 4's shape: the same ternary data, the same Philox positions, many tables per domain (so that cells are born
    and emptied all the time) -- choices, assignments and count tables identical move for move; the hashed table goes
    through its rebuild (more than half full of emptied cells) and growth paths on the way."""
    rng = np.random.default_rng(0)
    n = [60, 50, 40]
    ids, val = _ternary(rng, n)
    z = [rng.integers(0, 14, k).astype(np.int32) for k in n]
    import os
    hd = _build(eng, n, [[0, 1, 2]], [(ids, val)], z, E.TABLE_DENSE, 40)
    os.environ["HIRM_HASH_INITIAL_ENTRIES"] = "256"              # start far too small: the build overflows, doubles, rebuilds
    try:
        hh = _build(eng, n, [[0, 1, 2]], [(ids, val)], z, E.TABLE_HASHED, 40)
        np.testing.assert_array_equal(eng.get_counts(hd, 0), eng.get_counts(hh, 0))
    finally:
        del os.environ["HIRM_HASH_INITIAL_ENTRIES"]
    assert len(eng.get_counts(hh, 0)) > 2000                     # far more cells than the initial 256 entries: the table grew
    for sweep in range(6):
        md = np.concatenate([np.full(k, d, np.int32) for d, k in enumerate(n)])
        mi = np.concatenate([rng.permutation(k) for k in n]).astype(np.int32)
        ra = eng.sweep([hd], [(md, mi)], seed=3, streams=[7], counters=[sweep * len(mi)], trace_cap=48)
        rb = eng.sweep([hh], [(md, mi)], seed=3, streams=[7], counters=[sweep * len(mi)], trace_cap=48)
        np.testing.assert_array_equal(ra["choice"][0], rb["choice"][0])
        np.testing.assert_array_equal(ra["trace"][0][1], rb["trace"][0][1])
        np.testing.assert_array_equal(ra["trace"][0][2], rb["trace"][0][2])     # log-probabilities bit-identical
        for d in range(3):
            np.testing.assert_array_equal(eng.get_assignments(hd, d), eng.get_assignments(hh, d))
        np.testing.assert_array_equal(eng.get_counts(hd, 0), eng.get_counts(hh, 0))
        assert eng.logp_score(hd) == eng.logp_score(hh) or G.close(eng.logp_score(hd), eng.logp_score(hh), rtol=1e-12)
    # several chains in one launch, device-generated order
    eng.set_rng(hd, 5, 0); eng.set_rng(hh, 5, 0)
    eng.sweep_all([hd], n_sweeps=2, seed=9)
    eng.sweep_all([hh], n_sweeps=2, seed=9)
    np.testing.assert_array_equal(eng.get_counts(hd, 0), eng.get_counts(hh, 0))
    eng.irm_destroy(hd); eng.irm_destroy(hh)


def test_quaternary_relation_beyond_the_dense_budget(eng):
    """A 4-ary relation at 128 tables per domain: 2^28 cells of index space, more than the dense budget -- hashed
    automatically.  It loads, sweeps, and agrees with the CPU oracle move for move (labels, log-probabilities 1e-9,
    choices, count tables)."""
    rng = np.random.default_rng(1)
    n = [30, 28, 26, 24]
    nobs = 6000
    cells = rng.choice(int(np.prod(n)), size=nobs, replace=False)
    ids = np.stack(np.unravel_index(cells, n), axis=1).astype(np.int32)
    val = rng.integers(0, 2, nobs).astype(np.uint8)
    z = [rng.integers(0, 5, k).astype(np.int32) for k in n]
    h = eng.irm_create(n, [[0, 1, 2, 3]], max_tables=128)         # TABLE_AUTO
    eng.set_observations(h, 0, ids, val)
    eng.finalize(h)
    o = O.Oracle(n, [[0, 1, 2, 3]])
    o.set_observations(0, ids, val)
    for d in range(4):
        eng.set_assignments(h, d, z[d]); o.set_assignments(d, z[d])
    np.testing.assert_array_equal(eng.get_counts(h, 0), o.get_counts(0))
    md = np.concatenate([np.full(k, d, np.int32) for d, k in enumerate(n)])
    mi = np.concatenate([rng.permutation(k) for k in n]).astype(np.int32)
    res = eng.sweep([h], [(md, mi)], seed=21, streams=[2], counters=[0], trace_cap=40)
    assert eng.last_sweep_info()["kernel_kind"] == 0
    U.assert_traces_match(res["trace"][0], o, md, mi, [O.uniform(21, 2, m) for m in range(len(mi))])
    for d in range(4):
        np.testing.assert_array_equal(eng.get_assignments(h, d), o.get_assignments(d))
    np.testing.assert_array_equal(eng.get_counts(h, 0), o.get_counts(0))
    assert G.close(eng.logp_score(h), o.logp_score())
    # a dense table of that size is refused with a clear message
    h2 = eng.irm_create(n, [[0, 1, 2, 3]], max_tables=128)
    eng.set_table_mode(h2, 0, E.TABLE_DENSE)
    eng.set_observations(h2, 0, ids, val)
    with pytest.raises(E.EngineError, match="dense count table too large"):
        eng.finalize(h2)
    eng.irm_destroy(h); eng.irm_destroy(h2)
