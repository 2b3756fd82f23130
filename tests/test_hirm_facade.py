"""HIRM on the engine (hirm.model.HIRM, after src/hirm.py:457-571 / cxx/hirm.hh:784-1014).
CPU: relation-CRP bookkeeping.  GPU: the relation moves of the UNMODIFIED reference on animals.unary
(tests/golden/animals_unary_hirm.json.gz, section "hirm": traced by oracle/ref_trace.cc around the
reference's own add_relation / incorporate / Relation::logp_score calls) replayed on the facade: same candidate
tables, log-probabilities within 1e-9 relative, same uniform => same table, same final relation partition."""
import random

import numpy as np
import pytest

import golden_util as G
from hirm import HIRM
from hirm.model import RelationCRP, choose_index


def test_relation_crp_rules():
    c = RelationCRP(random.Random(1))
    assert c.tables_weights() == {0: 1.0}                      # N == 0 (cxx/hirm.hh:135-138)
    for r, t in (("a", 0), ("b", 0), ("c", 3)):
        c.incorporate(r, t)
    assert c.tables_weights() == {0: 2.0, 3: 1.0, 4: 1.0}      # fresh table = max + 1, weight alpha
    assert c.tables_weights_gibbs(0) == {0: 1.0, 3: 1.0, 4: 1.0}
    assert c.tables_weights_gibbs(3) == {0: 2.0, 3: 1.0}       # singleton: own table at weight alpha, no fresh one
    c.unincorporate("c")
    assert 3 not in c.tables and c.N == 2
    assert choose_index([0.0, -1.0, -2.0], 0.9) == 1 and choose_index([0.0], 0.3) == 0


def test_hirm_schema_without_gpu():
    h = HIRM({"black": ["animal"], "white": ["animal"], "eats": ["animal", "animal"]}, prng=random.Random(4))
    assert set(h.crp.assignments) == {"black", "white", "eats"} and set(h.irms) == set(h.crp.tables)
    h.incorporate("black", ("bear",), 1)
    h.incorporate("eats", ("bear", "fish"), 1)
    assert h.get_relation("eats").data == {("bear", "fish"): 1}
    assert "bear" in h.relation_to_irm("black").domains["animal"].items
    h.remove_relation("white")
    assert "white" not in h.schema and set(h.irms) == set(h.crp.tables)


def build_from_fixture(fx):
    """HIRM facade in the state the reference had before its traced relation moves."""
    st = fx["hirm"]["init"]
    rel_obs, rel_domains = {}, {}
    for irm in fx["irms"]:
        for r in irm["relations"]:
            a = np.asarray(r["obs"], dtype=np.int64).reshape(-1, len(r["domains"]) + 1)
            rel_obs[r["name"]], rel_domains[r["name"]] = a, list(r["domains"])
    h = HIRM({}, prng=random.Random(0))
    for table_s, doms in st["irm_z"].items():
        table = int(table_s)
        irm = h.irms[table] = h._new_irm({})
        rels = sorted(r for r, t in st["relation_table"].items() if t == table)
        for r in rels:
            irm.add_relation(r, rel_domains[r])
            h.schema[r] = rel_domains[r]
            h.crp.incorporate(r, table)
        for d, info in doms.items():
            for item, label in info["items"]:
                irm.domains[d].incorporate(item, label)        # forced seats (Domain.incorporate(item, table))
            irm._alpha[d] = info["alpha"]
        for r in rels:
            for row in rel_obs[r]:
                irm.incorporate(r, tuple(int(x) for x in row[:-1]), int(row[-1]))
    return h


@pytest.mark.gpu
def test_relation_moves_replay_reference_trace():
    fx = G.load("animals_unary_hirm")
    h = build_from_fixture(fx)
    assert G.close(h.logp_score(), fx["hirm"]["init"]["logp_score"])
    changed = 0
    for mv in fx["hirm"]["rmoves"]:
        assert h.relation_to_table(mv["r"]) == mv["cur"]
        h._seat_hook = {(mv["fresh"], d): {it: lb for it, lb in seats} for d, seats in mv.get("fresh_z", {}).items()}
        tables, logps, choice = h.transition_cluster_assignment_relation(mv["r"], u=mv["u"])
        assert tables == mv["tables"], (mv["r"], tables)
        assert G.close(logps, mv["lp"]), (mv["r"], logps, mv["lp"])
        assert choice == mv["choice"]
        changed += int(choice != mv["cur"])
    assert changed > 0
    fin = fx["hirm"]["final"]
    assert dict(h.crp.assignments) == fin["relation_table"]
    assert G.close(h.logp_score(), fin["logp_score"])
    # and the model keeps running: entity sweeps of all IRMs in one launch, relation moves with the facade's own prng
    s0 = h.logp_score()
    for _ in range(3):
        h.transition_cluster_assignments()
        h.transition_irms()
    assert np.isfinite(h.logp_score()) and set(h.irms) == set(h.crp.tables) and s0 < 0
