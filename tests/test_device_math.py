"""The sweep kernels' own fp64 score code (score_delta_bf: branch-free log / reciprocal / Stirling) against 40-digit
mpmath over the whole operand range the dense path meets: table-sized counts, the table / Stirling seam at 256,
and cells of 1e8 observations with groups of 1 .. 4e4."""
import numpy as np
import pytest

from hirm import engine as E

pytestmark = pytest.mark.gpu


def exact(N, s, gn, gs):
    import mpmath
    mpmath.mp.dps = 40
    lg = mpmath.loggamma

    def S(N, s):
        return lg(s + 1) + lg(N - s + 1) - lg(N + 2)
    return float(S(int(N) + int(gn), int(s) + int(gs)) - S(int(N), int(s)))


def test_score_delta_matches_mpmath():
    rng = np.random.default_rng(0)
    N, s, gn, gs = [], [], [], []
    for scale, gscale in [(40, 10), (250, 20), (300, 300), (5000, 50), (20000, 300), (3 * 10 ** 5, 3000), (10 ** 6, 2000), (10 ** 8, 1), (10 ** 8, 40000), (4 * 10 ** 8, 20000)]:
        for _ in range(250):
            n_ = int(rng.integers(0, scale + 1)); s_ = int(rng.integers(0, n_ + 1))
            g_ = int(rng.integers(0, gscale + 1)); t_ = int(rng.integers(0, g_ + 1))
            N.append(n_); s.append(s_); gn.append(g_); gs.append(t_)
    N += [0, 0, 255, 256, 254, 10 ** 8, 10 ** 8]; s += [0, 0, 0, 256, 127, 0, 10 ** 8]
    gn += [0, 1, 1, 1, 3, 20000, 20000]; gs += [0, 1, 1, 0, 2, 20000, 0]
    e = E.Engine()
    got = e.debug_score_delta(N, s, gn, gs)
    e.close()
    want = np.asarray([exact(*q) for q in zip(N, s, gn, gs)])
    err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
    k = int(np.argmax(err))
    assert err.max() < 1e-12, (err.max(), N[k], s[k], gn[k], gs[k], got[k], want[k])
