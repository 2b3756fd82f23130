"""BASELINE configs 4 and 5 at their stated sizes, where the CPU oracle cannot follow: size-independent
properties of the generic (CSR) kernel through the C ABI --

  (1) after a run of moves the count tables are still the histogram of (observations, z), bit for bit, computed
      independently with torch.bincount over the COO arrays (Relation::clusters is a pure function of data and
      assignments, cxx/hirm.hh:281-289);
  (2) the candidate log-probabilities of single moves equal a brute-force evaluation from the definition -- really re-seat
      the entity, rebuild the relation's table, difference BetaBernoulli::logp_score over the cells that changed
      (cxx/hirm.hh:52-56, :516-522) -- to 1e-9 relative;
  (3) IRM::logp_score equals the sum over those tables plus the CRP terms.

Rows are ~1000 (C4) and ~450 (C5 shape) entries long, i.e. longer than the kernel's prefetched row head and its 512-entry
CTA-width switch; domains hold 1e5 / 2e5 entities."""
import math

import numpy as np
import pytest

import golden_util as G
from hirm import engine as E

pytestmark = pytest.mark.gpu


def _gammaln(x):
    from scipy.special import gammaln
    return gammaln(np.asarray(x, dtype=np.float64))


def _S(N, s):
    N, s = np.asarray(N, dtype=np.float64), np.asarray(s, dtype=np.float64)
    return _gammaln(s + 1) + _gammaln(N - s + 1) - _gammaln(N + 2)


def _table(ids, val, zs, K):
    """(N, s) int64 tensors [K ** arity] of a relation's COO arrays under per-position slot tensors zs."""
    import torch
    key = torch.zeros(val.numel(), dtype=torch.int64, device=val.device)
    for p, z in enumerate(zs):
        key = key * K + z[ids[:, p].long()]
    N = torch.bincount(key, minlength=K ** len(zs))
    s = torch.bincount(key, weights=val.double(), minlength=K ** len(zs)).round().long()
    return N, s


def _delta_score(tabs_before, tabs_after):
    """sum over relations of Relation::logp_score(after) - (before), over the cells that differ only -- in 40 digits: the
    cells hold counts of 1e4..1e5, so each lgamma is ~1e6 and a float64 difference of such sums keeps ~1e-9 absolute at best."""
    import mpmath
    mpmath.mp.dps = 40
    lg = mpmath.loggamma

    def S(N, s):
        return lg(s + 1) + lg(N - s + 1) - lg(N + 2)
    tot = mpmath.mpf(0)
    for (N0, s0), (N1, s1) in zip(tabs_before, tabs_after):
        diff = ((N0 != N1) | (s0 != s1)).nonzero().flatten()
        if diff.numel() == 0:
            continue
        for a, b, cc, dd in zip(N1[diff].tolist(), s1[diff].tolist(), N0[diff].tolist(), s0[diff].tolist()):
            tot += S(a, b) - S(cc, dd)
    return float(tot)


def _crp_terms(z, alpha=1.0):
    sizes = np.bincount(z)
    sizes = sizes[sizes > 0]
    return len(sizes) * math.log(alpha) + float(np.sum(_gammaln(sizes))) + math.lgamma(alpha) - math.lgamma(sizes.sum() + alpha)


def _check_moves(eng, h, rels, doms_of_rel, z_dev, K, picks):
    """picks: [(domain, entity)].  Brute force of every candidate of each pick against the engine's score-only trace."""
    import torch
    bk = E.IrmBackend(eng, h)
    for d, item in picks:
        labels, lp = bk.score(d, item)
        z_host = z_dev[d].cpu().numpy()
        c = int(z_host[item])
        sizes = np.bincount(z_host, minlength=K)
        touching = [r for r, ds in enumerate(doms_of_rel) if d in ds]
        before = [_table(rels[r][0], rels[r][1], [z_dev[x] for x in doms_of_rel[r]], K) for r in touching]
        want = []
        for t in labels:
            t = int(t)
            w = sizes[t] - (1 if t == c else 0) if t < K and sizes[t] > 0 else 0
            logw = math.log(w) if w > 0 else 0.0              # fresh table / singleton's own table: alpha = 1
            if t == c:
                want.append(logw)
                continue
            z_dev[d][item] = t
            after = [_table(rels[r][0], rels[r][1], [z_dev[x] for x in doms_of_rel[r]], K) for r in touching]
            z_dev[d][item] = c
            want.append(logw + _delta_score(before, after))
        want = np.asarray(want)
        k = list(labels).index(c)
        assert G.close(lp - lp[k], want - want[k], rtol=1e-9), (d, item, labels, lp - lp[k], want - want[k])


def _check_tables(eng, h, rels, doms_of_rel, K):
    import torch
    z_dev = []
    nd = max(max(ds) for ds in doms_of_rel) + 1
    zs = [eng.get_assignments(h, d) for d in range(nd)]
    total = sum(_crp_terms(z) for z in zs)
    z_dev = [torch.from_numpy(z.astype(np.int64)).cuda() for z in zs]
    for r, ds in enumerate(doms_of_rel):
        N, s = _table(rels[r][0], rels[r][1], [z_dev[x] for x in ds], K)
        nz = N.nonzero().flatten()
        want = []
        for cell in nz.cpu().numpy():
            key, digits = int(cell), []
            for _ in ds:
                digits.append(key % K)
                key //= K
            want.append(digits[::-1] + [int(N[cell]), int(s[cell])])
        want = np.asarray(sorted(want), dtype=np.int32).reshape(-1, len(ds) + 2)
        np.testing.assert_array_equal(eng.get_counts(h, r), want)          # bit-exact count tables
        total += float(np.sum(_S(want[:, -2], want[:, -1])))
    got = eng.logp_score(h)
    assert abs(got - total) <= 1e-9 * abs(total), (got, total)
    return z_dev


def test_config4_full_size_properties_and_scores():
    """3 x 100 000 entities, 1e8 observed cells, rows of ~1000 entries."""
    import torch
    import bench_workloads as W
    dev = torch.device("cuda", 0)
    n, K, KCAP = 100_000, 64, 16          # K: label space of the independent tables (labels are never compacted); KCAP: device slots
    ids, val, zt = W.c4_ternary(dev, n=n, nobs=100_000_000, seed=1)
    assert ids.shape == (100_000_000, 3) and int(val.max()) == 1
    eng = E.Engine()
    h = eng.irm_create([n, n, n], [[0, 1, 2]], max_tables=KCAP)
    eng.set_observations_device(h, 0, val.numel(), ids.data_ptr(), val.data_ptr())
    eng.finalize(h)
    rng = np.random.default_rng(2)
    for d in range(3):
        z = zt[d].astype(np.int32).copy()
        z[rng.choice(n, 2000, replace=False)] = rng.integers(0, 8, 2000)    # some entities in the wrong table
        z[rng.choice(n, 30, replace=False)] = 8                              # a small ninth table
        z[int(rng.integers(n))] = 9                                          # and a singleton
        eng.set_assignments(h, d, z)
    rels, doms = [(ids, val)], [[0, 1, 2]]
    z_dev = _check_tables(eng, h, rels, doms, K)
    single = [int(np.flatnonzero(eng.get_assignments(h, d) == 9)[0]) for d in range(3)]
    _check_moves(eng, h, rels, doms, z_dev, K, [(0, 0), (1, 77), (2, n - 1), (0, single[0]), (2, single[2])])
    m = 2000
    md = np.repeat(np.arange(3, dtype=np.int32), m)
    mi = np.concatenate([rng.permutation(n)[:m] for _ in range(3)]).astype(np.int32)
    res = eng.sweep([h], [(md, mi)], seed=11, streams=[5], counters=[0])
    assert eng.last_sweep_info()["kernel_kind"] == 0
    assert len(res["choice"][0]) == 3 * m
    z_dev = _check_tables(eng, h, rels, doms, K)
    _check_moves(eng, h, rels, doms, z_dev, K, [(0, int(mi[0])), (1, int(mi[m])), (2, 12345)])
    eng.close()


def test_config5_shape_full_size_properties_and_scores():
    """One IRM of config 5's shape: 200 unary (fully observed) + 8 binary e x e relations of 1e7 cells over 200 000 entities
    (8.4e7 observations, rows of ~1000 entries; the binary relations exercise the repeated-domain rule (i, i) -> (t, t))."""
    import torch
    import bench_workloads as W
    dev = torch.device("cuda", 0)
    n, K, KCAP = 200_000, 64, 24
    domains, schema, obs, truth = W.c5_hirm(dev, n=n, n_unary=200, n_binary=8, bobs=10_000_000, n_clusters=2, tables=12, seed=3)
    names = list(schema)
    doms = [[0] * len(schema[r]) for r in names]
    rels = [(obs[r][0].reshape(-1, len(schema[r])), obs[r][1]) for r in names]
    eng = E.Engine()
    h = eng.irm_create([n], doms, max_tables=KCAP)
    for k, r in enumerate(names):
        eng.set_observations_device(h, k, rels[k][1].numel(), rels[k][0].data_ptr(), rels[k][1].data_ptr())
    eng.finalize(h)
    rng = np.random.default_rng(4)
    z = rng.integers(0, 12, n).astype(np.int32)
    z[rng.choice(n, 50, replace=False)] = 12
    z[7] = 13                                                                # a singleton
    eng.set_assignments(h, 0, z)
    z_dev = _check_tables(eng, h, rels, doms, K)
    _check_moves(eng, h, rels, doms, z_dev, K, [(0, 0), (0, 7), (0, n - 1)])
    mi = rng.permutation(n)[:3000].astype(np.int32)
    eng.sweep([h], [(np.zeros_like(mi), mi)], seed=13, streams=[9], counters=[0])
    assert eng.last_sweep_info()["kernel_kind"] == 0
    z_dev = _check_tables(eng, h, rels, doms, K)
    _check_moves(eng, h, rels, doms, z_dev, K, [(0, int(mi[0])), (0, 4242)])
    # the whole domain once through the device-generated sweep order, then the tables again
    nm = eng.sweep_all([h], n_sweeps=1, seed=17)
    assert nm == n
    _check_tables(eng, h, rels, doms, K)
    eng.close()
