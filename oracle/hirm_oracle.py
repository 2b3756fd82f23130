"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the CPU oracle (oracle/hirm_oracle.c).

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg only;
the product package (hierarchical-irm_b200/) never imports it.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libhirm_oracle.so")
MAX_ARITY = 4
_lib = None


def build(force=False):
    """Compile the C restatement (and, when /root/reference is present, oracle/_ref)."""
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "hirm_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        P = ctypes.c_void_p
        i32p = ctypes.POINTER(ctypes.c_int32)
        L.hirm_oracle_create.restype = P
        L.hirm_oracle_create.argtypes = [ctypes.c_int, i32p, ctypes.c_int, i32p, i32p]
        L.hirm_oracle_destroy.argtypes = [P]
        L.hirm_oracle_set_observations.argtypes = [P, ctypes.c_int, ctypes.c_int64, i32p,
                                                   ctypes.POINTER(ctypes.c_uint8)]
        L.hirm_oracle_set_assignments.argtypes = [P, ctypes.c_int, i32p]
        L.hirm_oracle_get_assignments.argtypes = [P, ctypes.c_int, i32p]
        L.hirm_oracle_set_alpha.argtypes = [P, ctypes.c_int, ctypes.c_double]
        L.hirm_oracle_score.argtypes = [P, ctypes.c_int, ctypes.c_int, i32p,
                                        ctypes.POINTER(ctypes.c_double), ctypes.c_int]
        L.hirm_oracle_choose.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_double]
        L.hirm_oracle_move.argtypes = [P, ctypes.c_int, ctypes.c_int, ctypes.c_double]
        L.hirm_oracle_sweep.argtypes = [P, ctypes.c_int64, i32p, i32p, ctypes.POINTER(ctypes.c_double),
                                        ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64]
        L.hirm_oracle_get_counts.restype = ctypes.c_int64
        L.hirm_oracle_get_counts.argtypes = [P, ctypes.c_int, i32p, ctypes.c_int64]
        L.hirm_oracle_logp_score.restype = ctypes.c_double
        L.hirm_oracle_logp_score.argtypes = [P]
        L.hirm_oracle_transition_alpha.restype = ctypes.c_double
        L.hirm_oracle_transition_alpha.argtypes = [P, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                                   ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        L.hirm_oracle_philox4x32_10.argtypes = [ctypes.POINTER(ctypes.c_uint32)] * 3
        L.hirm_oracle_uniform.restype = ctypes.c_double
        L.hirm_oracle_uniform.argtypes = [ctypes.c_uint64] * 3
        _lib = L
    return _lib


def _i32(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


class Oracle:
    """One IRM: `n_entities[d]`, relations as lists of domain indices."""

    def __init__(self, n_entities, rel_domains):
        self.L = lib()
        self.n_entities = [int(n) for n in n_entities]
        self.rel_domains = [list(map(int, ds)) for ds in rel_domains]
        ne = np.asarray(self.n_entities, dtype=np.int32)
        ar = np.asarray([len(ds) for ds in self.rel_domains], dtype=np.int32)
        rd = np.zeros((max(1, len(self.rel_domains)), MAX_ARITY), dtype=np.int32)
        for r, ds in enumerate(self.rel_domains):
            rd[r, :len(ds)] = ds
        self.h = self.L.hirm_oracle_create(len(ne), _i32(ne), len(self.rel_domains), _i32(ar), _i32(rd))
        if not self.h:
            raise ValueError("hirm_oracle_create failed")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.hirm_oracle_destroy(self.h)
            self.h = None

    def set_observations(self, rel, ids, val):
        ids = np.ascontiguousarray(ids, dtype=np.int32).reshape(-1, len(self.rel_domains[rel]))
        val = np.ascontiguousarray(val, dtype=np.uint8)
        rc = self.L.hirm_oracle_set_observations(self.h, rel, len(val), _i32(ids),
                                                 val.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
        if rc:
            raise ValueError("set_observations rc=%d" % rc)

    def set_assignments(self, dom, labels):
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        assert len(labels) == self.n_entities[dom]
        self.L.hirm_oracle_set_assignments(self.h, dom, _i32(labels))

    def get_assignments(self, dom):
        out = np.empty(self.n_entities[dom], dtype=np.int32)
        self.L.hirm_oracle_get_assignments(self.h, dom, _i32(out))
        return out

    def set_alpha(self, dom, alpha):
        self.L.hirm_oracle_set_alpha(self.h, dom, float(alpha))

    def score(self, dom, item, cap=4096):
        labels = np.empty(cap, dtype=np.int32)
        lp = np.empty(cap, dtype=np.float64)
        n = self.L.hirm_oracle_score(self.h, dom, item, _i32(labels),
                                     lp.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), cap)
        if n < 0:
            raise ValueError("score rc=%d" % n)
        return labels[:n].copy(), lp[:n].copy()

    def choose(self, lp, u):
        lp = np.ascontiguousarray(lp, dtype=np.float64)
        return self.L.hirm_oracle_choose(lp.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(lp), float(u))

    def move(self, dom, item, u):
        rc = self.L.hirm_oracle_move(self.h, dom, item, float(u))
        if rc < 0:
            raise ValueError("move rc=%d" % rc)
        return rc

    def sweep(self, move_dom, move_item, uniforms=None, seed=0, stream=0, offset=0):
        md = np.ascontiguousarray(move_dom, dtype=np.int32)
        mi = np.ascontiguousarray(move_item, dtype=np.int32)
        up = None
        if uniforms is not None:
            uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
            up = uniforms.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        rc = self.L.hirm_oracle_sweep(self.h, len(md), _i32(md), _i32(mi), up, seed, stream, offset)
        if rc:
            raise ValueError("sweep rc=%d" % rc)

    def get_counts(self, rel):
        w = len(self.rel_domains[rel]) + 2
        n = self.L.hirm_oracle_get_counts(self.h, rel, None, 0)
        out = np.empty((max(n, 1), w), dtype=np.int32)
        n = self.L.hirm_oracle_get_counts(self.h, rel, _i32(out), n)
        return out[:n]

    def logp_score(self):
        return self.L.hirm_oracle_logp_score(self.h)

    def transition_alpha(self, dom, u, ngrid=20):
        """Returns (alpha, grid, grid log-probabilities)."""
        grid = np.zeros(ngrid, np.float64); lp = np.zeros(ngrid, np.float64)
        dp = ctypes.POINTER(ctypes.c_double)
        a = self.L.hirm_oracle_transition_alpha(self.h, dom, ngrid, float(u), grid.ctypes.data_as(dp), lp.ctypes.data_as(dp))
        return a, grid, lp


def philox4x32_10(ctr, key):
    c = (ctypes.c_uint32 * 4)(*ctr)
    k = (ctypes.c_uint32 * 2)(*key)
    o = (ctypes.c_uint32 * 4)()
    lib().hirm_oracle_philox4x32_10(c, k, o)
    return list(o)


def uniform(seed, stream, counter):
    return lib().hirm_oracle_uniform(seed, stream, counter)
