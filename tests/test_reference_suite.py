"""Drop-in check of the Python package: the reference's own tests/test_basic.py, UNCHANGED,
against this repository's `hirm` package.

The reference tree is never copied into the repository.  Where it exists (HIRM_REFERENCE, default /root/reference -- the
build container) the tests stage a throw-away directory

    <tmp>/hirm          -> this repository's package (symlinks to hierarchical-irm_b200/hirm/*)
    <tmp>/hirm/tests    -> the reference's tests/ (symlink)

and (1) on CPU check that every name the reference's tests and examples import from `hirm.*` resolves in the package
with a compatible signature; (2) with a GPU run `pytest --pyargs hirm.tests` there -- the reference's tests
as they are, engine underneath.  A GPU box without the reference tree skips (2); tests/test_facade.py holds the same four
known-answer tests restated, so that they run there too."""
import ast
import inspect
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "hierarchical-irm_b200", "hirm")
REF = os.environ.get("HIRM_REFERENCE", "/root/reference")
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "tests", "test_basic.py")),
                               reason="the reference tree is not on this box")


def stage(tmp):
    pkg = os.path.join(str(tmp), "hirm")
    os.makedirs(pkg)
    for f in os.listdir(PKG):
        if f != "__pycache__" and f != "tests":
            os.symlink(os.path.join(PKG, f), os.path.join(pkg, f))
    os.symlink(os.path.join(REF, "tests"), os.path.join(pkg, "tests"))
    return str(tmp)


def imported_names(path):
    """{module: {name, ...}} for every `from hirm... import ...` / `import hirm...` of a reference source file."""
    out = {}
    for node in ast.walk(ast.parse(open(path).read())):
        if isinstance(node, ast.ImportFrom) and node.module and node.module.split(".")[0] == "hirm":
            out.setdefault(node.module, set()).update(a.name for a in node.names)
    return out


@needs_ref
def test_every_name_the_reference_imports_exists():
    import importlib
    files = [os.path.join(REF, "tests", "test_basic.py")] + \
            [os.path.join(REF, "examples", f) for f in sorted(os.listdir(os.path.join(REF, "examples"))) if f.endswith(".py")]
    checked = 0
    for path in files:
        for module, names in imported_names(path).items():
            if module in ("hirm.util_plot", "hirm.tests.util_test"):      # plotting: out of scope; the reference's own test helper
                continue
            mod = importlib.import_module(module)
            for n in names:
                assert hasattr(mod, n), "%s imports %s.%s" % (os.path.basename(path), module, n)
                checked += 1
    assert checked >= 20


@needs_ref
def test_signatures_match_the_reference():
    """Same positional parameters as src/util_io.py and src/hirm.py for everything the tests and examples call."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import importlib.util
    import tempfile
    import shutil
    import hirm as mine
    from hirm import util_io as my_io, util_math as my_math
    with tempfile.TemporaryDirectory() as tmp:
        shutil.copytree(os.path.join(REF, "src"), os.path.join(tmp, "refhirm"))
        sys.path.insert(0, tmp)
        try:
            ref_io = importlib.import_module("refhirm.util_io")
            ref_math = importlib.import_module("refhirm.util_math")
            ref_h = importlib.import_module("refhirm.hirm")
        finally:
            sys.path.remove(tmp)

        def positional(f):
            return [p.name for p in inspect.signature(f).parameters.values()
                    if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD) and p.name != "self"]

        for name in ("load_schema", "load_observations", "load_clusters_irm", "load_clusters_hirm", "to_txt_irm", "to_txt_hirm",
                     "from_txt_irm", "from_txt_hirm", "to_dict_IRM", "from_dict_IRM", "to_dict_HIRM", "from_dict_HIRM",
                     "to_dict_BetaBernoulli", "to_dict_CRP", "to_dict_Domain", "to_dict_Relation"):
            want, got = positional(getattr(ref_io, name)), positional(getattr(my_io, name))
            assert got[:len(want)] == want, (name, want, got)
        for name in ("logsumexp", "log_normalize", "log_choices", "log_linspace", "linspace", "logmeanexp"):
            assert positional(getattr(my_math, name)) == positional(getattr(ref_math, name)), name
        for cls, methods in (("IRM", ["__init__", "incorporate", "unincorporate", "transition_cluster_assignments",
                                      "transition_cluster_assignment_item", "transition_crp_alphas", "logp", "logp_score",
                                      "add_relation", "remove_relation"]),
                             ("HIRM", ["__init__", "incorporate", "unincorporate", "relation_to_table", "relation_to_irm", "relation",
                                       "transition_cluster_assignments", "transition_cluster_assignment_relation",
                                       "set_cluster_assignment_gibbs", "transition_crp_alpha", "add_relation", "remove_relation",
                                       "logp", "logp_score"])):
            for m in methods:
                want, got = positional(getattr(getattr(ref_h, cls), m)), positional(getattr(getattr(mine, cls), m))
                assert got[:len(want)] == want, (cls, m, want, got)
        # numeric agreement of the helpers with the reference's on random inputs
        import random
        rng = random.Random(0)
        for _ in range(50):
            xs = [rng.gauss(0, 30) for _ in range(rng.randint(1, 9))]
            assert abs(my_math.logsumexp(xs) - ref_math.logsumexp(xs)) <= 1e-12 * max(1.0, abs(ref_math.logsumexp(xs)))
            assert my_math.log_linspace(0.1, 7.0, 30) == ref_math.log_linspace(0.1, 7.0, 30)
        assert my_math.logsumexp([]) == ref_math.logsumexp([]) and my_math.logmeanexp([-my_math.inf, 0.0]) == ref_math.logmeanexp([-my_math.inf, 0.0])


@needs_ref
@pytest.mark.gpu
def test_reference_test_suite_runs_unchanged(tmp_path):
    """`pytest --pyargs hirm.tests` with hirm = this package and hirm/tests = the reference's tests directory."""
    top = stage(tmp_path)
    env = dict(os.environ, PYTHONPATH=top + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--pyargs", "hirm.tests"], cwd=top, env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "4 passed" in r.stdout, r.stdout[-2000:]
