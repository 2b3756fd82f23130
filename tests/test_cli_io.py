"""The reference's text formats and command line on the engine (hirm.util_io, hirm.cli after cxx/util_io.cc and
cxx/hirm.cc:92-195).  CPU: parsers and writers.  GPU: the CLI end to end on animals.unary rebuilt from the golden
fixture -- messages, output file format, --load restart (the reference's own smoke lines, cxx/Makefile:45-47)."""
import os
import random

import numpy as np
import pytest

import golden_util as G
from hirm import IRM, cli, util_io


def write_dataset(tmp, case):
    fx = G.load(case)
    rels = {}
    for irm in fx["irms"]:
        for r in irm["relations"]:
            rels[r["name"]] = (list(r["domains"]), np.asarray(r["obs"], dtype=np.int64).reshape(-1, len(r["domains"]) + 1))
    base = os.path.join(str(tmp), case)
    with open(base + ".schema", "w") as f:
        for r, (ds, _) in sorted(rels.items()):
            f.write("bernoulli %s %s\n" % (r, " ".join(ds)))
    with open(base + ".obs", "w") as f:
        for r, (ds, a) in sorted(rels.items()):
            for row in a:
                f.write("%d %s %s\n" % (row[-1], r, " ".join("%s%d" % (d[:2], x) for d, x in zip(ds, row[:-1]))))
    return base, rels


def test_parsers_and_irm_writer_roundtrip(tmp_path):
    base, rels = write_dataset(tmp_path, "two_relations_irm")
    schema = util_io.load_schema(base + ".schema")
    obs = util_io.load_observations(base + ".obs")
    assert set(schema) == set(rels) and all(schema[r] == rels[r][0] for r in rels)
    assert len(obs) == sum(len(a) for _, a in rels.values()) and all(v in (0, 1) for _, _, v in obs)
    irm = IRM(schema, prng=random.Random(2))
    for r, items, v in obs:
        irm.incorporate(r, items, v)
    path = base + ".2.irm"
    util_io.to_txt_irm(path, irm)                       # no device needed: the seats are the host CRP-prior draws
    lines = open(path).read().splitlines()
    assert all(len(l.split()) >= 3 and l.split()[1].isdigit() for l in lines)
    clusters = util_io.load_clusters_irm(path)
    for d, dom in irm.domains.items():
        got = {t: set(items) for t, items in clusters[d].items()}
        assert got == {t: {str(i) for i in members} for t, members in dom.crp.tables.items()}
    again = util_io.from_txt_irm(base + ".schema", base + ".obs", path, prng=random.Random(3))
    for d in irm.domains:                               # check_irms_agree, tests/test_basic.py:96-103
        assert again.domains[d].crp.assignments == irm.domains[d].crp.assignments
    for r in irm.relations:
        assert again.relations[r].data == irm.relations[r].data and again.relations[r].data_r == irm.relations[r].data_r


@pytest.mark.gpu
def test_cli_hirm_and_irm_end_to_end(tmp_path, capsys):
    base, rels = write_dataset(tmp_path, "animals_unary_hirm")
    assert cli.main(["--facade", "--seed", "1", "--iters", "5", base]) == 0
    out = capsys.readouterr().out.splitlines()
    assert out[0] == "setting seed to 1" and out[1] == "loading schema from %s.schema" % base
    assert "selected model is HIRM" in out and "inferring 5 iters; timeout 0" in out and out[-1] == "saving to %s.1.hirm" % base
    relations, irms = util_io.load_clusters_hirm(base + ".1.hirm")
    assert sorted(r for rs in relations.values() for r in rs) == sorted(rels)           # every relation in exactly one cluster
    assert set(irms) == set(relations)
    for t, doms in irms.items():                                                        # every animal seated once per IRM
        members = [e for items in doms["animal"].values() for e in items]
        assert len(members) == len(set(members)) == 50
    # restart from the saved clusters (cxx/Makefile:47)
    assert cli.main(["--facade", "--seed", "2", "--iters", "2", "--load", base + ".1.hirm", base]) == 0
    assert os.path.exists(base + ".2.hirm")
    # IRM mode with --verbose
    assert cli.main(["--mode", "irm", "--iters", "3", "--verbose", base]) == 0
    out = capsys.readouterr().out.splitlines()
    assert "selected model is IRM" in out and out[-1] == "saving to %s.10.irm" % base
    clusters = util_io.load_clusters_irm(base + ".10.irm")
    assert sum(len(v) for v in clusters["animal"].values()) == 50
    assert cli.main(["--mode", "nope", base]) == 1
    # the sharded HIRM from the command line (a world of one here; torchrun gives one GPU per rank), and a restart from its file
    assert cli.main(["--seed", "4", "--iters", "3", "--verbose", base]) == 0            # the default HIRM path
    out = capsys.readouterr().out.splitlines()
    assert "selected model is HIRM" in out and out[-1] == "saving to %s.4.hirm" % base
    relations, irms = util_io.load_clusters_hirm(base + ".4.hirm")
    assert sorted(r for rs in relations.values() for r in rs) == sorted(rels) and set(irms) == set(relations)
    for t, doms in irms.items():
        members = [e for items in doms["animal"].values() for e in items]
        assert len(members) == len(set(members)) == 50
    assert cli.main(["--sharded", "--seed", "5", "--iters", "1", "--load", base + ".4.hirm", base]) == 0
    assert os.path.exists(base + ".5.hirm")


def test_native_reader_matches_reference_encoding(tmp_path):
    """csrc/ingest.cc against load_schema + load_observations + the first-appearance encoding of cxx/util_io.cc:63-96
    (restated here with the Python readers); no GPU needed."""
    import numpy as np
    from hirm import engine as E
    from hirm import util_io
    base = tmp_path / "toy"
    (tmp_path / "toy.schema").write_text("bernoulli likes person item\nbernoulli tall person\n\nbernoulli knows person person\n")
    lines = ["1 likes ann tea", "0 likes bob tea", "1.0 tall bob", "0 knows bob ann", "1 likes ann jam", "  1 knows cat cat  ", "0.0 tall cat"]
    (tmp_path / "toy.obs").write_text("\n".join(lines) + "\n")
    domains, schema, obs, names = E.load_dataset(str(base) + ".schema", str(base) + ".obs")
    assert schema == util_io.load_schema(str(base) + ".schema") == {"likes": ["person", "item"], "tall": ["person"], "knows": ["person", "person"]}
    codes = {d: {} for d in domains}
    want = {r: [] for r in schema}
    for r, items, v in util_io.load_observations(str(base) + ".obs"):
        want[r].append([codes[d].setdefault(it, len(codes[d])) for d, it in zip(schema[r], items)] + [v])
    assert domains == {d: len(c) for d, c in codes.items()} == {"person": 3, "item": 2}
    assert names == {d: sorted(c, key=c.get) for d, c in codes.items()}
    for r in schema:
        w = np.asarray(want[r]).reshape(-1, len(schema[r]) + 1)
        np.testing.assert_array_equal(obs[r][0], w[:, :-1])
        np.testing.assert_array_equal(obs[r][1], w[:, -1])
    (tmp_path / "bad.obs").write_text("2 tall ann\n")
    try:
        E.load_dataset(str(base) + ".schema", str(tmp_path / "bad.obs"))
        assert False
    except E.EngineError as err:
        assert "0 or 1" in str(err)
    # tabs, CRLF line ends, blank lines; arity and relation errors are reported with the line number
    (tmp_path / "ws.obs").write_text("1\tlikes\tann\ttea\r\n\r\n0 tall  ann\r\n")
    d3, _, o3, n3 = E.load_dataset(str(base) + ".schema", str(tmp_path / "ws.obs"))
    assert d3 == {"person": 1, "item": 1} and n3 == {"person": ["ann"], "item": ["tea"]} and o3["tall"][1].tolist() == [0]
    assert o3["knows"][0].shape == (0, 2)
    for text, needle in (("1 likes ann\n", "too few"), ("1 tall ann bob\n", "too many"), ("1 hates ann\n", "unknown relation"),
                         ("1 tall ann\n", "listed twice")):
        (tmp_path / "e.obs").write_text("0 tall ann\n" + text)
        try:
            E.load_dataset(str(base) + ".schema", str(tmp_path / "e.obs"))
            assert False
        except E.EngineError as err:
            assert needle in str(err) and "line 2" in str(err)
    ref = "/root/reference/cxx/assets/animals.unary"
    import os
    if os.path.exists(ref + ".obs"):
        d2, s2, o2, _ = E.load_dataset(ref + ".schema", ref + ".obs", with_names=False)
        assert d2 == {"animal": 50} and len(s2) == 85 and sum(len(v[1]) for v in o2.values()) == 4250


def test_sharded_state_writer_and_loader_roundtrip(tmp_path):
    """.hirm file of a ShardedHIRM state (oracle backend, CPU) parsed back: same relation clusters, same partitions."""
    import numpy as np
    import sharded_util as SU
    from hirm import util_io
    from hirm.sharded import ShardedHIRM
    domains, schema, obs, _ = SU.synthetic_hirm(n_ent=(12, 9), n_unary=4, n_binary=2, n_clusters=2, seed=3)
    names = {d: ["%s_%d" % (d, k) for k in range(n)] for d, n in domains.items()}
    rn, dn = list(schema), list(domains)
    backend = SU.OracleBackend([domains[d] for d in dn], [[dn.index(d) for d in schema[r]] for r in rn], [obs[r] for r in rn], 16, 2)
    h = ShardedHIRM(domains, schema, obs, seed=2, max_tables=16, backend=backend)
    h.iterate(2)
    st = h.state()
    path = str(tmp_path / "m.2.hirm")
    util_io.to_txt_hirm_state(path, st, schema, names)
    rel_tables, parts = util_io.sharded_init_from_txt(path, domains, schema, names)
    assert rel_tables == {r: h.relation_to_table(r) for r in rn}
    for t in st:
        used = {d for r in st[t]["relations"] for d in schema[r]}
        for d in used:
            np.testing.assert_array_equal(parts[t][d], st[t]["z"][d])
    backend2 = SU.OracleBackend([domains[d] for d in dn], [[dn.index(d) for d in schema[r]] for r in rn], [obs[r] for r in rn], 16, 2)
    h2 = ShardedHIRM(domains, schema, obs, seed=2, max_tables=16, backend=backend2, relation_tables=rel_tables, partitions=parts)
    assert {r: h2.relation_to_table(r) for r in rn} == rel_tables
    h2.iterate(1)
