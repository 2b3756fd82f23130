"""Helpers shared by the GPU tests: build engine / oracle models from array specs."""
import numpy as np

from hirm import engine as E
from oracle import hirm_oracle as O


def engine_from_spec(eng, spec, max_tables=24):
    irm = eng.irm_create(spec.n_entities, spec.rel_domains, max_tables=max_tables)
    for r in range(len(spec.rnames)):
        eng.set_observations(irm, r, spec.obs_ids[r], spec.obs_val[r])
    eng.finalize(irm)
    for d in range(len(spec.dnames)):
        eng.set_assignments(irm, d, spec.init_z[d])
        eng.set_alpha(irm, d, spec.init_alpha[d])
    return irm


def oracle_from_spec(spec):
    o = O.Oracle(spec.n_entities, spec.rel_domains)
    for r in range(len(spec.rnames)):
        o.set_observations(r, spec.obs_ids[r], spec.obs_val[r])
    for d in range(len(spec.dnames)):
        o.set_assignments(d, spec.init_z[d])
        o.set_alpha(d, spec.init_alpha[d])
    return o


def planted_ree(n, seed=0, p_hi=0.9, p_lo=0.1, missing=0.0):
    """BASELINE config 3 generator: two planted blocks, x_ij ~ Bern(.9 if z_i != z_j else .1)."""
    rng = np.random.default_rng(seed)
    zs = (np.arange(n) >= n // 2)
    p = np.where(zs[:, None] != zs[None, :], p_hi, p_lo)
    x = (rng.random((n, n)) < p).astype(np.uint8)
    if missing > 0:
        x[rng.random((n, n)) < missing] = E.DENSE_MISSING
    return x


def crp_prior_labels(n, alpha, rng):
    """Sequential CRP prior seating (what Domain::incorporate does, cxx/hirm.hh:194-203)."""
    z = np.zeros(n, dtype=np.int32)
    counts = []
    for i in range(n):
        w = np.array(counts + [alpha], dtype=np.float64)
        k = int(rng.choice(len(w), p=w / w.sum()))
        if k == len(counts):
            counts.append(0)
        counts[k] += 1
        z[i] = k
    return z


def dense_to_coo(x):
    """All observed cells of a dense matrix as COO (ids, val)."""
    ii, jj = np.nonzero(x != E.DENSE_MISSING)
    ids = np.stack([ii, jj], axis=1).astype(np.int32)
    return ids, x[ii, jj].astype(np.uint8)


def assert_traces_match(trace, o, dom_of_move, items, uniforms, rtol=1e-9):
    """Step the oracle through the same moves and compare candidates, log-probabilities, choices."""
    n, labels, lp = trace
    for m in range(len(items)):
        ol, olp = o.score(dom_of_move[m], items[m])
        assert n[m] == len(ol), (m, n[m], ol)
        np.testing.assert_array_equal(labels[m, :n[m]], ol)
        a, b = lp[m, :n[m]], olp
        assert np.all(np.abs(a - b) <= rtol * np.maximum(1.0, np.abs(b))), (m, a, b)
        o.move(dom_of_move[m], items[m], uniforms[m])
