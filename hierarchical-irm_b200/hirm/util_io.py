"""Text formats of the reference, unchanged (cxx/util_io.cc; src/util_io.py:14-93,195-256):

  <base>.schema   one relation per line:   bernoulli <relation> <domain> [<domain> ...]       (cxx/util_io.cc:11-34)
  <base>.obs      one observation per line: <0|1> <relation> <entity> [<entity> ...]           (:36-61)
  <base>.<seed>.irm    one table per line:  <domain> <table id> <entity> ...                   (:131-150)
  <base>.<seed>.hirm   relation clusters "<table id> <relation> ...", a blank line, then one
                       "irm=<table id>" section in the .irm format per cluster, blank-line separated (:152-180)

Within a line the order of entities / relations is unspecified in the reference (hash order); here it is the order of
first appearance."""


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
            obs.append((w[1], tuple(w[2:]), int(float(w[0]))))
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
    """{domain: {table id: [entity, ...]}}   (cxx/util_io.cc:196-232)"""
    clusters = {}
    with open(path) as f:
        for line in f:
            w = line.split()
            if w:
                clusters.setdefault(w[0], {})[int(w[1])] = w[2:]
    return clusters


def load_clusters_hirm(path):
    """({table id: [relation, ...]}, {table id: irm clusters})   (cxx/util_io.cc:234-307)"""
    relations, irms, cur = {}, {}, None
    with open(path) as f:
        head = True
        for line in f:
            w = line.split()
            if not w:
                head = False
                continue
            if head:
                relations[int(w[0])] = w[1:]
            elif w[0].startswith("irm="):
                cur = int(w[0][4:])
                irms[cur] = {}
            else:
                irms[cur].setdefault(w[0], {})[int(w[1])] = w[2:]
    return relations, irms


def from_txt_irm(irm, path_schema, path_obs, path_clusters):
    """Forced seats, then the observations (cxx/util_io.cc:309-340)."""
    schema, obs, clusters = load_schema(path_schema), load_observations(path_obs), load_clusters_irm(path_clusters)
    for r, ds in schema.items():
        irm.add_relation(r, ds)
    for d, tabs in clusters.items():
        for t, items in tabs.items():
            for it in items:
                irm.domains[d].incorporate(it, t)
    for r, items, v in obs:
        irm.incorporate(r, items, v)
    return irm


def from_txt_hirm(hirm, path_schema, path_obs, path_clusters):
    """cxx/util_io.cc:342-392: relations into their tables, forced seats per IRM, then the observations."""
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
                used += [d for d in schema[r] if d not in used]
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
