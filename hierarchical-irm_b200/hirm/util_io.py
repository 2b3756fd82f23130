"""Text formats of the reference, unchanged (cxx/util_io.cc; src/util_io.py:14-93,195-256):

  <base>.schema   one relation per line:   bernoulli <relation> <domain> [<domain> ...]       (cxx/util_io.cc:11-34)
  <base>.obs      one observation per line: <0|1> <relation> <entity> [<entity> ...]           (:36-61)
  <base>.<seed>.irm    one table per line:  <domain> <table id> <entity> ...                   (:131-150)
  <base>.<seed>.hirm   relation clusters "<table id> <relation> ...", a blank line, then one
                       "irm=<table id>" section in the .irm format per cluster, blank-line separated (:152-180)

Within a line the order of entities / relations is unspecified in the reference (hash order); here it is the order of
first appearance.  Entity names that are numerals become ints, as src/util_io.py:8-12 (`intify`) makes them.

Also here: the JSON-compatible dict serialisers of src/util_io.py:95-193 (to_dict_* / from_dict_*), with the same keys; a
model rebuilt by from_dict_* is a fresh engine-backed model holding the same observations and the same seats."""
import ast


def intify(x):
    """src/util_io.py:8-12."""
    if x.isnumeric():
        return int(x)
    return x


def load_schema(path):
    schema = {}
    with open(path) as f:
        for line in f:
            w = line.split()
            if not w:
                continue
            assert len(w) >= 3, "schema line needs a distribution, a relation and at least one domain"
            if w[0] != "bernoulli":
                raise ValueError("only bernoulli relations are supported (README.md:122-123), got %r" % w[0])
            schema[w[1]] = w[2:]
    return schema


def load_observations(path):
    obs = []
    with open(path) as f:
        for line in f:
            w = line.split()
            if not w:
                continue
            assert len(w) >= 3
            obs.append((w[1], tuple(intify(x) for x in w[2:]), int(float(w[0]))))
    return obs


def _irm_lines(irm):
    lines = []
    for d, domain in irm.domains.items():
        tables = domain.crp.tables
        order = {it: k for k, it in enumerate(domain._names)}
        for table in sorted(tables):
            items = sorted(tables[table], key=order.get)
            lines.append("%s %d %s" % (d, table, " ".join(str(i) for i in items)))
    return lines


def to_txt_irm(path, irm):
    with open(path, "w") as f:
        f.write("".join(l + "\n" for l in _irm_lines(irm)))


def to_txt_hirm(path, hirm):
    tables = sorted(hirm.crp.tables)
    order = {r: k for k, r in enumerate(hirm.schema)}
    with open(path, "w") as f:
        for t in tables:
            f.write("%d %s\n" % (t, " ".join(sorted(hirm.crp.tables[t], key=order.get))))
        f.write("\n")
        for j, t in enumerate(tables):
            f.write("irm=%d\n" % t)
            f.write("".join(l + "\n" for l in _irm_lines(hirm.irms[t])))
            if j < len(tables) - 1:
                f.write("\n")


def load_clusters_irm(path):
    """{domain: {table id: (entity, ...)}}   (cxx/util_io.cc:196-232, src/util_io.py:41-55).  A table listed twice for a
    domain is an error (cxx/util_io.cc:216 asserts it)."""
    clusters = {}
    with open(path) as f:
        for ln, line in enumerate(f, 1):
            w = line.split()
            if not w:
                continue
            assert len(w) >= 3, "%s:%d: a cluster line needs a domain, a table id and at least one entity" % (path, ln)
            tabs = clusters.setdefault(w[0], {})
            assert int(w[1]) not in tabs, "%s:%d: table %s of domain %s listed twice" % (path, ln, w[1], w[0])
            tabs[int(w[1])] = tuple(intify(x) for x in w[2:])
    return clusters


def load_clusters_hirm(path):
    """({table id: (relation, ...)}, {table id: irm clusters})   (cxx/util_io.cc:234-307, src/util_io.py:57-93), with the
    reference's consistency checks: no relation table, IRM section, domain table listed twice (:258,:270,:292), every line
    inside a section, and the relation tables and the IRM sections name the same ids (:297-304)."""
    relations, irms, cur = {}, {}, None
    with open(path) as f:
        head = True
        for ln, line in enumerate(f, 1):
            w = line.split()
            where = "%s:%d: " % (path, ln)
            if not w:
                head, cur = False, None
                continue
            if head and w[0].isnumeric():
                assert len(w) >= 2, where + "a relation table needs at least one relation"
                assert int(w[0]) not in relations, where + "relation table %s listed twice" % w[0]
                relations[int(w[0])] = tuple(w[1:])
            elif len(w) == 1 and w[0].startswith("irm="):
                assert cur is None, where + "an irm= section must follow a blank line"
                cur = int(w[0][4:])
                assert cur not in irms, where + "irm=%d listed twice" % cur
                irms[cur] = {}
            else:
                assert cur is not None and len(w) >= 3, where + "cluster line outside an irm= section, or too short"
                tabs = irms[cur].setdefault(w[0], {})
                assert int(w[1]) not in tabs, where + "table %s of domain %s listed twice in irm=%d" % (w[1], w[0], cur)
                tabs[int(w[1])] = tuple(intify(x) for x in w[2:])
    assert set(relations) == set(irms), "%s: relation tables %s and irm sections %s differ" % (path, sorted(relations), sorted(irms))
    return relations, irms


def from_txt_irm(path_schema, path_obs, path_clusters, prng=None, **irm_kwargs):
    """src/util_io.py:227-238 / cxx/util_io.cc:309-340: a new IRM over the schema, forced seats from the clusters file,
    then the observations.  Same positional arguments as the reference; `prng` and the engine options are extras."""
    from .model import IRM
    schema, obs, clusters = load_schema(path_schema), load_observations(path_obs), load_clusters_irm(path_clusters)
    irm = IRM(schema, prng=prng, **irm_kwargs)
    for d, tabs in clusters.items():
        for t, items in tabs.items():
            for it in items:
                irm.domains[d].incorporate(it, t)
    for r, items, v in obs:
        irm.incorporate(r, items, v)
    return irm


def from_txt_hirm(path_schema, path_obs, path_clusters, prng=None, **hirm_kwargs):
    """src/util_io.py:240-256 / cxx/util_io.cc:342-392: relations into their tables, forced seats per IRM, then the
    observations.  Same positional arguments as the reference."""
    from .model import HIRM
    hirm = HIRM({}, prng=prng, **hirm_kwargs)
    schema, obs = load_schema(path_schema), load_observations(path_obs)
    relations, irms = load_clusters_hirm(path_clusters)
    for t, rs in relations.items():
        irm = hirm.irms[t] = hirm._new_irm({})
        for r in rs:
            irm.add_relation(r, schema[r])
            hirm.schema[r] = schema[r]
            hirm.crp.incorporate(r, t)
        for d, tabs in irms.get(t, {}).items():
            for tab, items in tabs.items():
                for it in items:
                    if d in irm.domains:
                        irm.domains[d].incorporate(it, tab)
    for r, items, v in obs:
        hirm.incorporate(r, items, v)
    return hirm


# ---- the sharded HIRM (hirm.sharded.ShardedHIRM works on integer codes; names come from hirm.engine.load_dataset) ----
def to_txt_hirm_state(path, state, schema, names):
    """<base>.<seed>.hirm (cxx/util_io.cc:152-180) from ShardedHIRM.state(): relation clusters, a blank line, one
    "irm=<table>" section per cluster listing the domains its relations use -- every IRM of the sharded model carries every
    domain, the reference's only those of its relations (cxx/hirm.hh:742-757)."""
    tables = sorted(state)
    with open(path, "w") as f:
        for t in tables:
            f.write("%d %s\n" % (t, " ".join(state[t]["relations"])))
        f.write("\n")
        for j, t in enumerate(tables):
            f.write("irm=%d\n" % t)
            used = []
            for r in state[t]["relations"]:
                for d in schema[r]:
                    if d not in used:
                        used.append(d)
            for d in used:
                z = state[t]["z"][d]
                by_table = {}
                for code, tab in enumerate(z):
                    by_table.setdefault(int(tab), []).append(names[d][code])
                for tab in sorted(by_table):
                    f.write("%s %d %s\n" % (d, tab, " ".join(str(x) for x in by_table[tab])))
            if j < len(tables) - 1:
                f.write("\n")


def sharded_init_from_txt(path_clusters, domains, schema, names):
    """--load for the sharded HIRM (cxx/util_io.cc:342-392): (relation_tables, partitions) for ShardedHIRM(...).  Domains a
    cluster's file section does not list (none of its relations uses them) start in one table."""
    import numpy as np
    relations, irms = load_clusters_hirm(path_clusters)
    rel_tables = {r: t for t, rs in relations.items() for r in rs}
    code = {d: {str(n): k for k, n in enumerate(ns)} for d, ns in names.items()}
    parts = {}
    for t in relations:
        parts[t] = {}
        for d in domains:
            z = np.zeros(domains[d], dtype=np.int32)
            for tab, items in irms.get(t, {}).get(d, {}).items():
                for it in items:
                    z[code[d][str(it)]] = tab
            parts[t][d] = z
    return rel_tables, parts


# ---- JSON-compatible dicts (src/util_io.py:95-193): same keys, same caveats (dict keys become strings: repr / literal_eval) --
def to_dict_BetaBernoulli(x):
    return {"alpha": x.alpha, "beta": x.beta, "N": x.N, "s": x.s}


def from_dict_BetaBernoulli(d, prng=None):
    from .model import BetaBernoulli
    return BetaBernoulli(alpha=d["alpha"], beta=d["beta"], prng=prng, N=d["N"], s=d["s"])


def to_dict_CRP(x):
    return {"alpha": x.alpha, "N": x.N, "tables": {repr(t): list(v) for t, v in x.tables.items()},
            "assignments": {repr(i): t for i, t in x.assignments.items()}}


def to_dict_Domain(x):
    return {"name": x.name, "items": list(x.items), "crp": to_dict_CRP(x.crp)}


def to_dict_Relation(x):
    return {"name": x.name, "domains": [d.name for d in x.domains],
            "clusters": {repr(c): to_dict_BetaBernoulli(v) for c, v in x.clusters.items()},
            "data": {repr(i): v for i, v in x.data.items()},
            "data_r": {repr(k): {repr(i): [list(t) for t in v] for i, v in m.items()} for k, m in x.data_r.items()}}


def to_dict_IRM(x):
    return {"schema": {r: list(ds) for r, ds in x.schema.items()},
            "domains": {k: to_dict_Domain(v) for k, v in x.domains.items()},
            "relations": {k: to_dict_Relation(v) for k, v in x.relations.items()},
            "domain_to_relations": {k: list(v) for k, v in x.domain_to_relations.items()}}


def from_dict_IRM(d, prng=None, **irm_kwargs):
    """A fresh engine-backed IRM with the serialised schema, seats (domains[..].crp.assignments, alpha) and observations
    (relations[..].data); the count tables follow from those (they are a histogram)."""
    from .model import IRM
    irm = IRM({r: tuple(ds) for r, ds in d["schema"].items()}, prng=prng, **irm_kwargs)
    for name, dd in d["domains"].items():
        dom = irm.domains[name]
        for item, table in dd["crp"]["assignments"].items():
            dom.incorporate(ast.literal_eval(item), int(table))
        dom.crp.alpha = dd["crp"]["alpha"]
    for name, rd in d["relations"].items():
        for items, value in rd["data"].items():
            irm.incorporate(name, tuple(ast.literal_eval(items)), int(value))
    return irm


def to_dict_HIRM(x):
    crp = {"alpha": x.crp.alpha, "N": x.crp.N, "tables": {repr(t): list(v) for t, v in x.crp.tables.items()},
           "assignments": {repr(r): t for r, t in x.crp.assignments.items()}}
    return {"schema": {r: list(ds) for r, ds in x.schema.items()}, "crp": crp,
            "irms": {k: to_dict_IRM(v) for k, v in x.irms.items()}}


def from_dict_HIRM(d, prng=None, **hirm_kwargs):
    from .model import HIRM
    hirm = HIRM({}, prng=prng, **hirm_kwargs)
    hirm.crp.alpha = d["crp"]["alpha"]
    for t, di in d["irms"].items():
        t = int(t)
        irm = hirm.irms[t] = hirm._new_irm({})
        for r, ds in di["schema"].items():
            irm.add_relation(r, tuple(ds))
            hirm.schema[r] = tuple(ds)
            hirm.crp.incorporate(r, t)
        for name, dd in di["domains"].items():
            for item, table in dd["crp"]["assignments"].items():
                irm.domains[name].incorporate(ast.literal_eval(item), int(table))
            irm.domains[name].crp.alpha = dd["crp"]["alpha"]
        for name, rd in di["relations"].items():
            for items, value in rd["data"].items():
                irm.incorporate(name, tuple(ast.literal_eval(items)), int(value))
    return hirm
