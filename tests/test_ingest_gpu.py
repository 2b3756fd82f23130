"""GPU reader of the reference's text formats (csrc/ingest_gpu.cu) against the host reader (csrc/ingest.cc, which follows
cxx/util_io.cc:11-96): identical domains, identical entity codes (first appearance), identical names, the same multiset
of observations per relation; the same errors with the same line numbers."""
import os

import numpy as np
import pytest

from hirm import engine as E

pytestmark = pytest.mark.gpu


def _write(tmp, schema, lines, name="d"):
    base = os.path.join(str(tmp), name)
    with open(base + ".schema", "w") as f:
        f.write("".join(s + "\n" for s in schema))
    with open(base + ".obs", "w") as f:
        f.write("".join(l + "\n" for l in lines))
    return base


def _same(host, gpu):
    d1, s1, o1, n1 = host
    d2, s2, o2, n2, _ = gpu
    assert d1 == d2 and s1 == s2 and n1 == n2 and list(d1) == list(d2) and list(s1) == list(s2)
    for r in s1:
        a = np.concatenate([o1[r][0], o1[r][1][:, None].astype(np.int32)], axis=1)
        b = np.concatenate([o2[r][0].cpu().numpy(), o2[r][1].cpu().numpy()[:, None].astype(np.int32)], axis=1)
        assert a.shape == b.shape, r
        np.testing.assert_array_equal(a[np.lexsort(a.T[::-1])], b[np.lexsort(b.T[::-1])])


def test_gpu_reader_equals_host_reader(tmp_path):
    rng = np.random.default_rng(0)
    schema = ["bernoulli likes person item", "bernoulli knows person person", "bernoulli tall person", "bernoulli buys person item shop",
              "bernoulli empty shop"]
    people = ["p%d" % k for k in range(700)] + ["Ann-Marie", "x", "7", "007"]
    items = ["item_%x" % k for k in range(300)]
    shops = ["s%03d" % k for k in range(40)]
    lines, seen = [], set()
    for _ in range(60000):
        k = rng.integers(0, 4)
        if k == 0:
            key = ("likes", rng.choice(people), rng.choice(items))
        elif k == 1:
            key = ("knows", rng.choice(people), rng.choice(people))
        elif k == 2:
            key = ("tall", rng.choice(people))
        else:
            key = ("buys", rng.choice(people), rng.choice(items), rng.choice(shops))
        if key in seen:
            continue
        seen.add(key)
        v = ["0", "1", "1.0", "0.0", "1."][rng.integers(0, 5)]
        sep = [" ", "\t", "  "][rng.integers(0, 3)]
        lines.append(v + sep + key[0] + sep + sep.join(key[1:]) + ("\r" if rng.random() < 0.1 else ""))
        if rng.random() < 0.01:
            lines.append("")
    base = _write(tmp_path, schema, lines)
    host = E.load_dataset(base + ".schema", base + ".obs")
    gpu = E.load_dataset_gpu(base + ".schema", base + ".obs")
    _same(host, gpu)
    assert gpu[0] == {"person": 704, "item": 300, "shop": 40} and gpu[2]["empty"][1].numel() == 0
    # no trailing newline, a single line, an empty file
    with open(base + ".obs", "w") as f:
        f.write("1 tall ann\n0 likes bob tea")
    _same(E.load_dataset(base + ".schema", base + ".obs"), E.load_dataset_gpu(base + ".schema", base + ".obs"))
    open(base + ".obs", "w").close()
    _same(E.load_dataset(base + ".schema", base + ".obs"), E.load_dataset_gpu(base + ".schema", base + ".obs"))


def test_gpu_reader_reports_the_same_errors(tmp_path):
    schema = ["bernoulli likes person item", "bernoulli tall person"]
    good = ["1 tall ann", "0 likes ann tea", "1 likes bob tea"]
    for bad, needle in (("1 likes ann", "too few"), ("1 tall ann bob", "too many"), ("1 hates ann", "unknown relation"),
                        ("2 tall bob", "0 or 1"), ("0 likes ann tea", "twice")):
        base = _write(tmp_path, schema, good + [bad], name="e")
        for loader in (E.load_dataset, E.load_dataset_gpu):
            with pytest.raises(E.EngineError) as err:
                loader(base + ".schema", base + ".obs")
            assert needle in str(err.value) and "line 4" in str(err.value), (loader.__name__, str(err.value))


def test_gpu_reader_feeds_the_engine(tmp_path):
    """The device COO arrays go straight into an IRM (hirm_irm_set_observations_device): same count tables as the host path."""
    rng = np.random.default_rng(1)
    lines = ["%d R a%d b%d c%d" % (rng.integers(0, 2), i, j, k) for i in range(12) for j in range(10) for k in range(9) if rng.random() < 0.4]
    base = _write(tmp_path, ["bernoulli R a b c"], lines, name="t")
    dh, sh, oh, _ = E.load_dataset(base + ".schema", base + ".obs")
    dg, sg, og, _, times = E.load_dataset_gpu(base + ".schema", base + ".obs")
    assert set(times) == {"h2d_s", "split_parse_s", "dictionary_s", "coo_s"}
    eng = E.Engine()
    z = [rng.integers(0, 3, dh[d]).astype(np.int32) for d in dh]
    hs = []
    for obs, dev in ((oh, False), (og, True)):
        h = eng.irm_create([dh[d] for d in dh], [[0, 1, 2]], max_tables=8)
        if dev:
            eng.set_observations_device(h, 0, og["R"][1].numel(), og["R"][0].data_ptr(), og["R"][1].data_ptr())
        else:
            eng.set_observations(h, 0, *oh["R"])
        eng.finalize(h)
        for d in range(3):
            eng.set_assignments(h, d, z[d])
        hs.append(h)
    np.testing.assert_array_equal(eng.get_counts(hs[0], 0), eng.get_counts(hs[1], 0))
    eng.close()
