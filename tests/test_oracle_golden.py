"""CPU tests: the plain-C oracle (oracle/hirm_oracle.c) against traces of the unmodified
reference (tests/golden/, produced by oracle/_ref/ref_trace) -- this is what pins the oracle."""
import numpy as np
import pytest

import golden_util as G
from oracle import hirm_oracle as O


class OracleBackend(O.Oracle):
    def transition_alpha(self, dom, u, ngrid):
        a, grid, lp = O.Oracle.transition_alpha(self, dom, u, ngrid)
        return a, lp


def make_oracle(spec):
    o = OracleBackend(spec.n_entities, spec.rel_domains)
    for r in range(len(spec.rnames)):
        o.set_observations(r, spec.obs_ids[r], spec.obs_val[r])
    for d in range(len(spec.dnames)):
        o.set_assignments(d, spec.init_z[d])
        o.set_alpha(d, spec.init_alpha[d])
    return o


@pytest.mark.parametrize("case", G.CASES)
def test_oracle_replays_reference_trace(case):
    fx = G.load(case)
    total = dict(moves=0, changed=0, repeated=0, alpha_moves=0)
    for irm in fx["irms"]:
        spec = G.IrmSpec(irm)
        res = G.replay(make_oracle(spec), spec)
        for k in total:
            total[k] += res[k]
    assert total["moves"] > 0 and total["changed"] > 0
    if fx["mode"] == "irm":
        assert total["alpha_moves"] > 0  # CRP::transition_alpha is pinned against the reference's grid and scores
    if case == "nations_binary_py":
        # BASELINE config 2 against the Python backend: 10 sweeps x 181 entities, 30-point alpha grid per domain and sweep
        assert fx["backend"].startswith("python") and total["moves"] == 1810 and total["alpha_moves"] == 30
    if case in ("nations_binary_irm", "nations_binary_py", "synth_ree40_irm", "synth_mixed_irm"):
        assert total["repeated"] > 0     # the Q1 (repeated-domain) path is exercised


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    assert O.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    u = [O.uniform(1, 2, c) for c in range(1000)]
    assert 0 <= min(u) and max(u) < 1 and 0.45 < np.mean(u) < 0.55


def test_choose_matches_python_choices_rule():
    # random.Random.choices: bisect_right(cum_weights, u * total), capped at n-1
    import bisect
    import itertools
    rng = np.random.default_rng(0)
    o = None
    for _ in range(200):
        n = int(rng.integers(1, 9))
        lp = rng.normal(size=n) * 5
        u = float(rng.random())
        Z = np.log(np.sum(np.exp(lp - lp.max()))) + lp.max()
        p = [float(np.exp(x - Z)) for x in lp]
        cum = list(itertools.accumulate(p))
        want = bisect.bisect_right(cum, u * cum[-1], 0, n - 1)
        got = O.lib().hirm_oracle_choose(np.ascontiguousarray(lp).ctypes.data_as(
            __import__("ctypes").POINTER(__import__("ctypes").c_double)), n, u)
        assert got == want
