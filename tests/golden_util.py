"""Load tests/golden/*.json.gz (traces of the unmodified reference -- its C++ backend through
tests/golden/make_golden.py, its Python backend through tests/golden/make_golden_py.py) and replay them against any backend exposing
score/move/get_assignments/get_counts/logp_score -- the CPU oracle or the CUDA engine."""
import gzip
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["animals_unary_irm", "nations_binary_irm", "two_relations_irm", "animals_unary_hirm",
         "synth_ree40_irm", "synth_ternary_irm", "synth_mixed_irm",
         # the reference's PYTHON backend (src/hirm.py) on BASELINE config 2, 10 sweeps from random.Random(12):
         # tests/golden/make_golden_py.py
         "nations_binary_py"]
RTOL = 1e-9   # north_star: candidate log-probabilities within 1e-9 relative (fp64)


def load(name):
    with gzip.open(os.path.join(GOLDEN_DIR, name + ".json.gz"), "rt") as f:
        return json.load(f)


class IrmSpec:
    """Array form of one traced IRM."""

    def __init__(self, irm):
        self.raw = irm
        self.dnames = sorted(irm["domains"])
        self.dindex = {d: k for k, d in enumerate(self.dnames)}
        self.n_entities = [irm["domains"][d] for d in self.dnames]
        self.rnames = [r["name"] for r in irm["relations"]]
        self.rel_domains = [[self.dindex[d] for d in r["domains"]] for r in irm["relations"]]
        self.obs_ids, self.obs_val = [], []
        for r in irm["relations"]:
            a = np.asarray(r["obs"], dtype=np.int64).reshape(-1, len(r["domains"]) + 1)
            self.obs_ids.append(np.ascontiguousarray(a[:, :-1], dtype=np.int32))
            self.obs_val.append(np.ascontiguousarray(a[:, -1], dtype=np.uint8))
        self.init_z = [np.asarray(irm["init"]["z"][d], dtype=np.int32) for d in self.dnames]
        self.init_alpha = [irm["init"]["alpha"][d] for d in self.dnames]
        self.moves = irm["moves"]


def close(a, b, rtol=RTOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return bool(np.all(np.abs(a - b) <= rtol * np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))))


def check_state(backend, spec, state):
    for k, d in enumerate(spec.dnames):
        np.testing.assert_array_equal(backend.get_assignments(k), np.asarray(state["z"][d], dtype=np.int32))
    for r, name in enumerate(spec.rnames):
        want = np.asarray(state["counts"][name], dtype=np.int32).reshape(-1, len(spec.rel_domains[r]) + 2)
        got = backend.get_counts(r)
        np.testing.assert_array_equal(got, want)          # bit-exact count tables
    ls = backend.logp_score()
    assert close(ls, state["logp_score"]), (ls, state["logp_score"])


def replay(backend, spec, check_candidates=True):
    """Drive `backend` through the reference's trace; returns summary counters."""
    check_state(backend, spec, spec.raw["init"])
    n_moves = n_changed = n_repeated = n_alpha = 0
    for mv in spec.moves:
        if "ckpt" in mv:
            check_state(backend, spec, mv["state"])
            continue
        if "amove" in mv:
            # CRP::transition_alpha: the reference's grid (log_linspace) and per-point CRP::logp_score, same
            # uniform => same grid point
            d = spec.dindex[mv["amove"]]
            alpha, lp = backend.transition_alpha(d, mv["u"], mv["ngrid"])
            assert close(lp, mv["lp"]), (mv, lp)
            assert abs(alpha - mv["alpha"]) <= 1e-12 * mv["alpha"], (mv, alpha)
            n_alpha += 1
            continue
        d = spec.dindex[mv["d"]]
        backend.set_alpha(d, mv["alpha"])
        if check_candidates:
            labels, lp = backend.score(d, mv["i"])
            assert list(labels) == mv["labels"], (mv, labels)
            k = mv["labels"].index(mv["cur"])
            bf = np.asarray(mv["bf_lp"])
            # exact conditional: brute force over the reference's own Relation::logp_score
            assert close(lp - lp[k], bf - bf[k]), (mv, lp)
            if not mv["repeated"]:
                # domain occurs once in every relation touched: the unmodified
                # reference's logp_gibbs_exact is the per-move oracle
                assert close(lp, mv["ref_lp"]), (mv, lp)
            else:
                n_repeated += 1
        choice = backend.move(d, mv["i"], mv["u"])
        assert choice == mv["choice"], (mv, choice)
        n_moves += 1
        n_changed += int(choice != mv["cur"])
    return dict(moves=n_moves, changed=n_changed, repeated=n_repeated, alpha_moves=n_alpha)
