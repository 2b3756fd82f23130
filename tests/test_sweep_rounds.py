"""Host logic of the rows-grouped-ahead sweep (csrc/sweep_rounds.cc, hirm_plan_rounds; no GPU): every move list is cut into
pieces -- at most `seg` moves of one qualifying run (pre = 1), or everything between two qualifying runs, merged (pre = 0) -- and
round r of a sweep is piece r of every list.  The pieces of a list must tile it exactly and in order, whatever the runs."""
import numpy as np
import pytest

from hirm import engine as E


def _check(offsets, runs, seg, rounds):
    n = len(offsets) - 1
    for k in range(n):
        pieces = [rd[k] for rd in rounds if rd[k][0] < rd[k][1]]
        pos = offsets[k]
        for lo, hi, pre in pieces:
            assert lo == pos and hi > lo                        # tiles the list, in order
            pos = hi
            if pre:
                assert hi - lo <= seg
                # inside ONE qualifying run
                inside = [(st, q) for j, (st, q) in enumerate(runs[k])
                          if st <= lo and hi <= (runs[k][j + 1][0] if j + 1 < len(runs[k]) else offsets[k + 1])]
                assert len(inside) == 1 and inside[0][1]
        assert pos == offsets[k + 1]
        # no two ordinary pieces in a row (they are merged), empty rounds only at the end
        for a, b in zip(pieces, pieces[1:]):
            assert a[2] or b[2]
        flags = [rd[k][0] < rd[k][1] for rd in rounds]
        assert flags == sorted(flags, reverse=True)
    assert len(rounds) == max([1] + [sum(1 for rd in rounds if rd[k][0] < rd[k][1]) for k in range(n)])


def test_domain_blocks_become_one_piece_per_block_and_segment():
    # two lists: three domain blocks each (config 4's step), the middle domain of list 1 does not qualify
    offsets = [0, 300, 520]
    runs = [[(0, 1), (100, 1), (200, 1)], [(300, 1), (380, 0), (450, 1)]]
    rounds = E.plan_rounds(offsets, runs, seg=64)
    _check(offsets, runs, 64, rounds)
    assert [rd[0] for rd in rounds[:6]] == [(0, 64, 1), (64, 100, 1), (100, 164, 1), (164, 200, 1), (200, 264, 1), (264, 300, 1)]
    assert [rd[1] for rd in rounds[:5]] == [(300, 364, 1), (364, 380, 1), (380, 450, 0), (450, 514, 1), (514, 520, 1)]
    assert rounds[5][1][0] == rounds[5][1][1]                   # list 1 has five pieces: nothing left in round 5


def test_unqualified_runs_merge_and_unknown_lists_stay_whole():
    offsets = [0, 50, 50, 170]
    runs = [[(0, 0), (10, 0), (20, 1), (40, 0), (45, 0)], [], None]
    rounds = E.plan_rounds(offsets, runs, seg=1000)
    _check(offsets, [runs[0], [], [(50, 0)]], 1000, rounds)
    assert [rd[0] for rd in rounds] == [(0, 20, 0), (20, 40, 1), (40, 50, 0)]
    assert rounds[0][2] == (50, 170, 0) and all(rd[1][0] == rd[1][1] for rd in rounds)


def test_random_lists_are_tiled_exactly():
    rng = np.random.default_rng(0)
    for trial in range(200):
        n = int(rng.integers(1, 6))
        lens = rng.integers(0, 400, n)
        offsets = np.concatenate([[0], np.cumsum(lens)]).tolist()
        runs = []
        for k in range(n):
            if lens[k] == 0:
                runs.append([]); continue
            cuts = np.unique(np.concatenate([[0], rng.integers(0, lens[k], int(rng.integers(0, 12)))]))
            runs.append([(int(offsets[k] + c), int(rng.integers(0, 2))) for c in cuts])
        seg = int(rng.integers(1, 200))
        _check(offsets, runs, seg, E.plan_rounds(offsets, runs, seg))


def test_bad_runs_are_refused():
    with pytest.raises(E.EngineError):
        E.plan_rounds([0, 10], [[(3, 1)]], 4)                   # the first run must start the list
    with pytest.raises(E.EngineError):
        E.plan_rounds([0, 10], [[(0, 1), (12, 1)]], 4)          # a run beyond the list


def test_cluster_sizes_follow_work_and_budget():
    """hirm_plan_clusters: powers of two up to 16, the budget is never exceeded, more work never gets fewer CTAs, live tables count
    like entries, and an IRM with too little work per CTA stays on one CTA unless the caller forces clusters."""
    rng = np.random.default_rng(3)
    for trial in range(200):
        n = int(rng.integers(1, 60))
        ent = rng.choice([5.0, 80.0, 400.0, 2500.0, 7000.0], n) * rng.uniform(0.5, 1.5, n)
        tab = rng.integers(1, 60, n).astype(np.int32)
        budget = int(rng.integers(n, 200))
        cs, total = E.plan_clusters(ent, tab, budget)
        assert total == cs.sum() <= budget and set(cs.tolist()) <= {1, 2, 4, 8, 16}
        work = ent * np.maximum(1.0, tab / 8.0)
        order = np.argsort(work)
        assert np.all(np.diff(cs[order]) >= 0) or np.all(np.diff((work / cs)[order]) > -1e-9 * work.max() - 1)   # heavier: never fewer CTAs ...
        for i in range(n):
            if cs[i] > 1:
                assert work[i] / (cs[i] // 2) >= 192.0             # ... and only where a CTA keeps enough work
        # nobody could still double: out of budget, at the cap, or too little work
        left = budget - total
        for i in range(n):
            assert cs[i] >= 16 or cs[i] > left or work[i] / cs[i] < 192.0
    cs, total = E.plan_clusters([100.0, 100.0], [8, 8], 148)
    assert cs.tolist() == [1, 1]
    cs, total = E.plan_clusters([100.0, 100.0], [8, 8], 148, forced=True)
    assert cs.tolist() == [16, 16]
    # config 5's shape: a young IRM with 47 tables outweighs an old one with twice the entries
    cs, total = E.plan_clusters([3400.0, 6000.0] + [300.0] * 40, [47, 9] + [15] * 40, 72)
    assert cs[0] == 16 and cs[0] >= cs[1] and total <= 72
