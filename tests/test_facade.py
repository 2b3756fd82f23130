"""The host-side mirror of the reference's Python interface (hirm.model.IRM, after src/hirm.py:378-455).
CPU part: schema and buffering logic.  GPU part: the reference's own known-answer tests restated on the
facade -- tests/test_basic.py:21-43 (add/remove relations), :45-57 (exact recovery of the two planted
blocks within 20 sweeps) -- and the dense R(e, e) layout."""
import random

import numpy as np
import pytest

from hirm import IRM


def two_clusters():
    """The 30 x 40 noise-free block relation of the reference's fixture (tests/util_test.py:4-21)."""
    d1 = [list(range(0, 10)) + list(range(20, 30)), list(range(10, 20))]
    d2 = [list(range(0, 20)), list(range(20, 40))]
    data = [((i, j), int(a != b)) for a in range(2) for b in range(2) for i in d1[a] for j in d2[b]]
    return {"R1": ("D1", "D2")}, d1, d2, data


def test_schema_and_buffering_need_no_gpu():
    m = IRM({}, prng=random.Random(3))
    m.add_relation("Arctic", ["Animal"])
    m.add_relation("Hunts", ["Animal", "Animal"])
    m.incorporate("Arctic", ("Bear",), 1)
    m.incorporate("Hunts", ("Bear", "Fish"), 1)
    assert m.domains["Animal"].items == {"Bear", "Fish"}
    assert m.relations["Hunts"].data == {("Bear", "Fish"): 1}
    assert m.relations["Hunts"].data_r["Animal"]["Bear"] == {("Bear", "Fish")}
    assert m.domains["Animal"].crp.N == 2 and set(m.domains["Animal"].crp.assignments) == {"Bear", "Fish"}
    with pytest.raises(AssertionError):
        m.incorporate("Hunts", ("Bear", "Fish"), 0)          # src/hirm.py:177
    with pytest.raises(NotImplementedError):
        m.unincorporate("Hunts", ("Bear", "Fish"))           # src/hirm.py:390
    for r in ("Arctic", "Hunts"):
        m.remove_relation(r)
    assert not m.relations and not m.domains and not m.domain_to_relations


@pytest.mark.gpu
def test_irm_add_relation_and_sweep():
    schema = {"Flippers": ["Animal"], "Strain Teeth": ["Animal"], "Swims": ["Animal"], "Arctic": ["Animal"],
              "Hunts": ["Animal", "Animal"]}
    model = IRM({}, prng=random.Random(1))
    for relation, domains in schema.items():
        model.add_relation(relation, domains)
    model.incorporate("Arctic", ("Bear",), 1)
    model.incorporate("Hunts", ("Bear", "Bear"), 0)
    model.incorporate("Hunts", ("Bear", "Fish"), 1)
    model.transition_cluster_assignments()
    assert set(model.domains["Animal"].crp.assignments) == {"Bear", "Fish"}
    assert np.isfinite(model.logp_score())
    for relation in schema:
        model.remove_relation(relation)
    assert not model.relations and not model.domains and not model.domain_to_relations


@pytest.mark.gpu
def test_irm_two_clusters_recovered():
    schema, d1, d2, data = two_clusters()
    irm = IRM(schema, prng=random.Random(1))
    for (i, j), v in data:
        irm.incorporate("R1", (i, j), v)
    s0 = irm.logp_score()
    for _ in range(20):
        irm.transition_cluster_assignments()
    t1, t2 = irm.domains["D1"].crp.tables, irm.domains["D2"].crp.tables
    assert len(t1) == 2 and set(d1[0]) in t1.values() and set(d1[1]) in t1.values()
    assert set(d2[0]) in t2.values() and set(d2[1]) in t2.values()
    assert irm.logp_score() > s0
    # the count table read back through Relation.clusters: four cells, noise free
    cl = irm.relations["R1"].clusters
    assert len(cl) == 4 and sorted((c.N, c.s) for c in cl.values()) == sorted([(400, 0), (400, 400), (200, 200), (200, 0)])
    # later incorporate()s extend the model (state is rebuilt on the device), single moves and alpha moves work
    irm.incorporate("R1", (30, 0), 0)
    irm.transition_cluster_assignment_item("D1", 30)
    irm.transition_crp_alphas()
    assert 30 in irm.domains["D1"].crp.assignments and irm.domains["D1"].crp.alpha > 0


@pytest.mark.gpu
def test_dense_ree_through_the_facade():
    """A fully observed R(e, e) goes to the dense slab layout and the TMA kernel; the planted blocks come back."""
    n = 48
    rng = np.random.default_rng(0)
    zs = np.arange(n) >= n // 2
    irm = IRM({"R": ("e", "e")}, prng=random.Random(2))
    for i in range(n):
        for j in range(n):
            irm.incorporate("R", (i, j), int(rng.random() < (0.9 if zs[i] != zs[j] else 0.1)))
    for _ in range(10):
        irm.transition_cluster_assignments()
    assert irm._eng.last_sweep_info()["kernel_kind"] == 1
    tabs = irm.domains["e"].crp.tables
    for members in tabs.values():                      # no table mixes the two planted blocks
        assert len({m >= n // 2 for m in members}) == 1, tabs
    assert len(tabs) <= 6 and sum(len(v) for v in tabs.values()) == n
