"""GPU parity tests: the CUDA engine, through the C ABI, against (1) the committed traces of
the unmodified reference and (2) the pinned CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

import golden_util as G
import engine_util as U
from hirm import engine as E
from oracle import hirm_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = E.Engine()
    yield e
    e.close()


@pytest.mark.parametrize("case", G.CASES)
def test_engine_replays_reference_trace(eng, case):
    """Per move: identical candidate label sets, log-probabilities within 1e-9 relative of the
    reference (brute force for repeated domains), same uniform => same table, and bit-exact
    assignments / count tables at every checkpoint."""
    fx = G.load(case)
    total = dict(moves=0, changed=0, repeated=0, alpha_moves=0)
    for irm in fx["irms"]:
        spec = G.IrmSpec(irm)
        h = U.engine_from_spec(eng, spec)
        res = G.replay(E.IrmBackend(eng, h), spec)
        for k in total:
            total[k] += res[k]
        eng.irm_destroy(h)
    assert total["moves"] > 0 and total["changed"] > 0


def test_multi_irm_single_launch_matches_oracle(eng):
    """All relation-cluster IRMs of the traced HIRM swept concurrently in ONE launch with on-device
    Philox; the oracle replays each IRM with the same (seed, stream, counter)."""
    fx = G.load("animals_unary_hirm")
    specs = [G.IrmSpec(i) for i in fx["irms"]]
    hs = [U.engine_from_spec(eng, s) for s in specs]
    rng = np.random.default_rng(3)
    moves = []
    for s in specs:
        order = np.concatenate([rng.permutation(s.n_entities[0]) for _ in range(3)]).astype(np.int32)
        moves.append((np.zeros_like(order), order))
    streams = np.arange(len(hs), dtype=np.uint64) + 11
    counters = np.full(len(hs), 5, dtype=np.uint64)
    res = eng.sweep(hs, moves, seed=1234, streams=streams, counters=counters, trace_cap=32)
    assert eng.last_sweep_info()["launches"] == 1
    for k, s in enumerate(specs):
        o = U.oracle_from_spec(s)
        un = [O.uniform(1234, int(streams[k]), 5 + m) for m in range(len(moves[k][1]))]
        U.assert_traces_match(res["trace"][k], o, moves[k][0], moves[k][1], un)
        np.testing.assert_array_equal(eng.get_assignments(hs[k], 0), o.get_assignments(0))
        for r in range(len(s.rnames)):
            np.testing.assert_array_equal(eng.get_counts(hs[k], r), o.get_counts(r))
        assert G.close(eng.logp_score(hs[k]), o.logp_score())
    for h in hs:
        eng.irm_destroy(h)


def _ree_models(eng, n, seed, missing=0.0, kcap=32, dense=True):
    import torch
    x = U.planted_ree(n, seed=seed, missing=missing)
    rng = np.random.default_rng(seed + 1)
    z0 = U.crp_prior_labels(n, 1.0, rng)
    ids, val = U.dense_to_coo(x)
    o = O.Oracle([n], [[0, 0]])
    o.set_observations(0, ids, val)
    o.set_assignments(0, z0)
    h = eng.irm_create([n], [[0, 0]], max_tables=kcap)
    keep = None
    if dense:
        pitch = (n + 15) // 16 * 16
        s0 = torch.full((n, pitch), 0xFF, dtype=torch.uint8, device="cuda")
        s0[:, :n] = torch.from_numpy(x).cuda()
        s1 = torch.full((n, pitch), 0xFF, dtype=torch.uint8, device="cuda")
        s1[:, :n] = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
        torch.cuda.synchronize()
        eng.set_observations_dense(h, 0, s0.data_ptr(), pitch, s1.data_ptr(), pitch)
        keep = (s0, s1)
    else:
        eng.set_observations(h, 0, ids, val)
    eng.finalize(h)
    eng.set_assignments(h, 0, z0)
    return x, z0, o, h, keep


@pytest.mark.parametrize("n,missing", [(37, 0.0), (200, 0.0), (333, 0.15), (1024, 0.0)])
def test_dense_ree_fast_path_matches_oracle(eng, n, missing):
    """Config-3 shape on the dense TMA kernel vs the oracle: every move's candidates and
    log-probabilities (1e-9 rel.), every sampled table, final z and count tables bit-exact."""
    x, z0, o, h, keep = _ree_models(eng, n, seed=n, missing=missing)
    np.testing.assert_array_equal(eng.get_counts(h, 0), o.get_counts(0))
    rng = np.random.default_rng(7)
    order = np.concatenate([rng.permutation(n) for _ in range(2)]).astype(np.int32)
    res = eng.sweep([h], [(np.zeros_like(order), order)], seed=99, streams=[3], counters=[0], trace_cap=40)
    assert eng.last_sweep_info()["kernel_kind"] == 1
    un = [O.uniform(99, 3, m) for m in range(len(order))]
    U.assert_traces_match(res["trace"][0], o, np.zeros_like(order), order, un)
    np.testing.assert_array_equal(eng.get_assignments(h, 0), o.get_assignments(0))
    np.testing.assert_array_equal(eng.get_counts(h, 0), o.get_counts(0))
    assert G.close(eng.logp_score(h), o.logp_score())
    eng.irm_destroy(h)


def test_dense_and_generic_kernels_agree(eng):
    """The same R(e,e) data through the CSR/generic kernel and the dense kernel: identical trajectories."""
    n = 150
    _, _, _, hd, keep = _ree_models(eng, n, seed=5, missing=0.1, dense=True)
    _, _, _, hg, _ = _ree_models(eng, n, seed=5, missing=0.1, dense=False)
    order = np.random.default_rng(1).permutation(n).astype(np.int32)
    mv = (np.zeros_like(order), order)
    rd = eng.sweep([hd], [mv], seed=5, streams=[1], counters=[0], trace_cap=40)
    assert eng.last_sweep_info()["kernel_kind"] == 1
    rg = eng.sweep([hg], [mv], seed=5, streams=[1], counters=[0], trace_cap=40)
    assert eng.last_sweep_info()["kernel_kind"] == 0
    np.testing.assert_array_equal(rd["choice"][0], rg["choice"][0])
    np.testing.assert_array_equal(rd["trace"][0][1], rg["trace"][0][1])
    nn = rd["trace"][0][0]
    for m in range(n):
        assert G.close(rd["trace"][0][2][m, :nn[m]], rg["trace"][0][2][m, :nn[m]])
    np.testing.assert_array_equal(eng.get_counts(hd, 0), eng.get_counts(hg, 0))
    eng.irm_destroy(hd); eng.irm_destroy(hg)


def test_many_chains_share_one_dataset(eng):
    """Independent seeded chains over shared slabs in one launch: each equals its own oracle chain."""
    n, chains = 96, 5
    x, z0, _, h0, keep = _ree_models(eng, n, seed=21)
    hs = [h0]
    for c in range(1, chains):
        h = eng.irm_create([n], [[0, 0]], max_tables=32)
        eng.share_observations(h, h0)
        eng.set_assignments(h, 0, z0)
        hs.append(h)
    order = np.random.default_rng(2).permutation(n).astype(np.int32)
    mv = (np.zeros_like(order), order)
    res = eng.sweep(hs, [mv] * chains, seed=77, streams=np.arange(chains, dtype=np.uint64), counters=np.zeros(chains, np.uint64))
    assert eng.last_sweep_info()["launches"] == 2      # all chains in ONE launch per shared-memory tier (8 / 32 table slots)
    ids, val = U.dense_to_coo(x)
    finals = []
    for c in range(chains):
        o = O.Oracle([n], [[0, 0]])
        o.set_observations(0, ids, val)
        o.set_assignments(0, z0)
        o.sweep(mv[0], mv[1], seed=77, stream=c, offset=0)
        np.testing.assert_array_equal(eng.get_assignments(hs[c], 0), o.get_assignments(0))
        np.testing.assert_array_equal(eng.get_counts(hs[c], 0), o.get_counts(0))
        finals.append(tuple(o.get_assignments(0)))
    assert len(set(finals)) > 1       # the chains really are different
    for h in hs:
        eng.irm_destroy(h)


def test_error_paths(eng):
    h = eng.irm_create([4], [[0, 0]], max_tables=2)
    eng.set_observations(h, 0, [[0, 1], [1, 2], [2, 3]], [1, 0, 1])
    eng.finalize(h)
    with pytest.raises(E.EngineError):          # three tables > max_tables
        eng.set_assignments(h, 0, [0, 1, 2, 0])
    eng.set_assignments(h, 0, [0, 1, 0, 1])
    with pytest.raises(E.EngineError) as ei:    # fresh-table candidate needs a free slot
        eng.sweep([h], [([0], [0])], seed=1)
    assert "max_tables" in str(ei.value)
    with pytest.raises(E.EngineError):
        eng.set_observations(h, 0, [[0, 9]], [1])
    eng.irm_destroy(h)
