"""ctypes binding of the C ABI in include/hirm_engine.h (libhirm_engine.so).

This is the only way Python reaches the CUDA kernels; there is no CPU implementation
behind it -- importing works anywhere (so that the symbol table can be checked on a
CPU box) but creating an Engine without a CUDA device raises.
"""
import ctypes
import os
import sys

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
import build_engine as _build  # noqa: E402  (hierarchical-irm_b200/build_engine.py)

MAX_ARITY = 4
MAX_DOMAINS = 8
DENSE_MISSING = 0xFF
SWEEP_SCORE_ONLY = 1
SWEEP_FORCE_GENERIC = 2
SWEEP_CLUSTER = 4
TABLE_AUTO, TABLE_DENSE, TABLE_HASHED = 0, 1, 2

EXPORTS = [
    "hirm_last_error", "hirm_engine_abi_version", "hirm_engine_create", "hirm_engine_destroy",
    "hirm_engine_set_stream", "hirm_engine_synchronize", "hirm_irm_create", "hirm_irm_destroy",
    "hirm_irm_set_observations", "hirm_irm_set_observations_dense", "hirm_irm_set_observations_dense_host",
    "hirm_irm_share_observations", "hirm_irm_finalize", "hirm_irm_set_assignments", "hirm_irm_get_assignments",
    "hirm_irm_set_alpha", "hirm_irm_get_counts", "hirm_irm_logp_score", "hirm_engine_sweep",
    "hirm_engine_sweep_device", "hirm_engine_last_sweep_info", "hirm_engine_transpose_u8",
    "hirm_engine_debug_phase_cycles", "hirm_engine_last_dense_plan", "hirm_engine_sweep_all", "hirm_irm_set_rng",
    "hirm_engine_logp_scores", "hirm_engine_transition_alpha", "hirm_irm_get_alpha", "hirm_irm_get_tables",
    "hirm_engine_relation_scores", "hirm_irm_seat_entities", "hirm_engine_debug_score_delta",
    "hirm_irm_set_observations_device", "hirm_engine_relation_score_matrix", "hirm_engine_relation_scores_fresh",
    "hirm_engine_crp_prior_draw", "hirm_irm_set_rng_counters", "hirm_engine_clear_relation_cache",
    "hirm_engine_seat_prior", "hirm_irm_set_table_mode", "hirm_unary_block_create", "hirm_unary_block_destroy",
    "hirm_irm_set_observations_unary", "hirm_engine_last_cluster_size", "hirm_engine_last_rounds", "hirm_plan_rounds", "hirm_plan_clusters", "hirm_relation_reseat",
    "hirm_dataset_last_error", "hirm_dataset_load", "hirm_dataset_free", "hirm_dataset_counts", "hirm_dataset_domain",
    "hirm_dataset_relation", "hirm_dataset_observations", "hirm_dataset_entity",
    "hirm_gpu_dataset_last_error", "hirm_gpu_dataset_load", "hirm_gpu_dataset_free", "hirm_gpu_dataset_counts", "hirm_gpu_dataset_domain",
    "hirm_gpu_dataset_relation", "hirm_gpu_dataset_entity", "hirm_gpu_dataset_copy_relation", "hirm_gpu_dataset_timings",
]

_lib = None
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_u64p = ctypes.POINTER(ctypes.c_uint64)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_f64p = ctypes.POINTER(ctypes.c_double)


class EngineError(RuntimeError):
    pass


def lib_path():
    return _build.LIB


def load():
    """dlopen the engine (building it with nvcc if the .so is missing or stale)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build()
    L = ctypes.CDLL(path)
    V = ctypes.c_void_p
    I = ctypes.c_int
    L.hirm_last_error.restype = ctypes.c_char_p
    L.hirm_engine_create.argtypes = [ctypes.POINTER(V), I]
    L.hirm_engine_destroy.argtypes = [V]
    L.hirm_engine_set_stream.argtypes = [V, V]
    L.hirm_engine_synchronize.argtypes = [V]
    L.hirm_irm_create.argtypes = [V, I, c_i32p, c_i32p, I, c_i32p, c_i32p, c_i32p]
    L.hirm_irm_destroy.argtypes = [V, ctypes.c_int32]
    L.hirm_irm_set_observations.argtypes = [V, ctypes.c_int32, I, ctypes.c_int64, c_i32p, c_u8p]
    L.hirm_irm_set_observations_dense.argtypes = [V, ctypes.c_int32, I, V, ctypes.c_int64, V, ctypes.c_int64]
    L.hirm_irm_set_observations_dense_host.argtypes = [V, ctypes.c_int32, I, V, ctypes.c_int64]
    L.hirm_irm_share_observations.argtypes = [V, ctypes.c_int32, ctypes.c_int32]
    L.hirm_irm_finalize.argtypes = [V, ctypes.c_int32]
    L.hirm_irm_set_assignments.argtypes = [V, ctypes.c_int32, I, c_i32p]
    L.hirm_irm_get_assignments.argtypes = [V, ctypes.c_int32, I, c_i32p]
    L.hirm_irm_set_alpha.argtypes = [V, ctypes.c_int32, I, ctypes.c_double]
    L.hirm_irm_get_counts.argtypes = [V, ctypes.c_int32, I, c_i32p, ctypes.c_int64, c_i64p]
    L.hirm_irm_logp_score.argtypes = [V, ctypes.c_int32, c_f64p]
    L.hirm_engine_sweep.argtypes = [V, I, c_i32p, c_i64p, c_i32p, c_i32p, c_f64p, ctypes.c_uint64, c_u64p, c_u64p,
                                    ctypes.c_uint32, c_i32p, I, c_i32p, c_i32p, c_f64p]
    L.hirm_engine_sweep_device.argtypes = [V, I, c_i32p, V, V, V, ctypes.c_uint64, V, V, ctypes.c_uint32, V]
    L.hirm_engine_last_sweep_info.argtypes = [V, c_i32p, c_i32p, ctypes.POINTER(ctypes.c_float)]
    L.hirm_engine_transpose_u8.argtypes = [V, V, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, V, ctypes.c_int64]
    L.hirm_engine_debug_phase_cycles.argtypes = [V, V]
    L.hirm_engine_last_dense_plan.argtypes = [V, c_i32p, c_i32p, c_i32p, c_i32p]
    L.hirm_engine_sweep_all.argtypes = [V, I, c_i32p, I, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int64, c_i32p, c_i32p, c_i32p, c_i64p]
    L.hirm_irm_set_rng.argtypes = [V, ctypes.c_int32, ctypes.c_uint64, ctypes.c_uint64]
    L.hirm_engine_logp_scores.argtypes = [V, I, c_i32p, c_f64p]
    L.hirm_engine_transition_alpha.argtypes = [V, I, c_i32p, I, ctypes.c_uint64, c_f64p, c_f64p, c_f64p]
    L.hirm_irm_get_alpha.argtypes = [V, ctypes.c_int32, I, c_f64p]
    L.hirm_irm_get_tables.argtypes = [V, ctypes.c_int32, I, c_i32p, c_i32p, I, c_i32p]
    L.hirm_engine_relation_scores.argtypes = [V, ctypes.c_int32, I, I, c_i32p, c_i32p, c_f64p]
    L.hirm_irm_seat_entities.argtypes = [V, ctypes.c_int32, I, I, c_i32p, c_i32p]
    L.hirm_engine_debug_score_delta.argtypes = [V, I, V, V, V, V, c_f64p]
    L.hirm_irm_set_observations_device.argtypes = [V, ctypes.c_int32, I, ctypes.c_int64, V, V]
    VP = ctypes.POINTER(V)
    L.hirm_engine_relation_score_matrix.argtypes = [V, I, c_i32p, c_i32p, c_i64p, VP, VP, I, c_i32p, V]
    L.hirm_engine_relation_scores_fresh.argtypes = [V, I, c_i32p, c_i32p, c_i64p, VP, VP, I, c_i32p, c_i32p, ctypes.c_double,
                                                    ctypes.c_uint64, c_u64p, c_f64p]
    L.hirm_engine_clear_relation_cache.argtypes = [V]
    L.hirm_engine_seat_prior.argtypes = [V, I, c_i32p, ctypes.c_uint64]
    L.hirm_irm_set_table_mode.argtypes = [V, ctypes.c_int32, I, I]
    L.hirm_unary_block_create.argtypes = [V, ctypes.c_int32, I, c_i64p, ctypes.POINTER(V), ctypes.POINTER(V), c_i32p]
    L.hirm_unary_block_destroy.argtypes = [V, ctypes.c_int32]
    L.hirm_irm_set_observations_unary.argtypes = [V, ctypes.c_int32, I, ctypes.c_int32, I]
    L.hirm_engine_last_cluster_size.argtypes = [V, c_i32p]
    L.hirm_engine_last_rounds.argtypes = [V, c_i32p, c_i32p]
    L.hirm_plan_clusters.argtypes = [ctypes.c_int, c_f64p, c_i32p, ctypes.c_int, ctypes.c_int, c_i32p]
    L.hirm_plan_rounds.argtypes = [ctypes.c_int, c_i64p, c_i32p, ctypes.c_int, c_i64p, c_i32p, ctypes.c_int64, c_i32p, c_i64p, c_i64p, c_i32p, ctypes.c_int]
    L.hirm_relation_reseat.argtypes = [I, I, c_i32p, c_i32p, c_i32p, c_f64p, ctypes.c_int64, I, c_i32p, c_f64p, I, ctypes.c_double,
                                       c_i32p, c_i32p, c_i32p]
    L.hirm_irm_set_rng_counters.argtypes = [V, ctypes.c_int32, ctypes.c_uint64, ctypes.c_uint64]
    L.hirm_engine_crp_prior_draw.argtypes = [V, I, ctypes.c_double, I, ctypes.c_uint64, ctypes.c_uint64, c_i32p]
    CP = ctypes.POINTER(ctypes.c_char_p)
    L.hirm_dataset_last_error.restype = ctypes.c_char_p
    L.hirm_dataset_load.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(V)]
    L.hirm_dataset_free.argtypes = [V]
    L.hirm_dataset_counts.argtypes = [V, c_i32p, c_i32p]
    L.hirm_dataset_domain.argtypes = [V, I, CP, c_i32p]
    L.hirm_dataset_relation.argtypes = [V, I, CP, c_i32p, c_i32p, c_i64p]
    L.hirm_dataset_observations.argtypes = [V, I, ctypes.POINTER(c_i32p), ctypes.POINTER(c_u8p)]
    L.hirm_dataset_entity.argtypes = [V, I, ctypes.c_int32, CP]
    L.hirm_gpu_dataset_last_error.restype = ctypes.c_char_p
    L.hirm_gpu_dataset_load.argtypes = [ctypes.c_char_p, ctypes.c_char_p, I, ctypes.POINTER(V)]
    L.hirm_gpu_dataset_free.argtypes = [V]
    L.hirm_gpu_dataset_counts.argtypes = [V, c_i32p, c_i32p]
    L.hirm_gpu_dataset_domain.argtypes = [V, I, CP, c_i32p]
    L.hirm_gpu_dataset_relation.argtypes = [V, I, CP, c_i32p, c_i32p, c_i64p, ctypes.POINTER(V), ctypes.POINTER(V)]
    L.hirm_gpu_dataset_entity.argtypes = [V, I, ctypes.c_int32, CP]
    L.hirm_gpu_dataset_copy_relation.argtypes = [V, I, V, V]
    L.hirm_gpu_dataset_timings.argtypes = [V, c_f64p]
    _lib = L
    return L


def load_dataset(schema_path, obs_path, with_names=True):
    """Native reader of <base>.schema / <base>.obs (csrc/ingest.cc; cxx/util_io.cc:11-96): returns
    (domains {name: n_entities}, schema {relation: [domain names]}, observations {relation: (ids [nobs, arity] int32,
    val [nobs] uint8)}, names {domain: [entity name by code]} or None) -- the inputs of hirm.sharded.ShardedHIRM."""
    L = load()
    ds = ctypes.c_void_p()
    if L.hirm_dataset_load(os.fsencode(schema_path), os.fsencode(obs_path), ctypes.byref(ds)):
        raise EngineError(L.hirm_dataset_last_error().decode())
    try:
        nd, nr = ctypes.c_int32(), ctypes.c_int32()
        L.hirm_dataset_counts(ds, ctypes.byref(nd), ctypes.byref(nr))
        name, n = ctypes.c_char_p(), ctypes.c_int32()
        dnames, domains = [], {}
        for d in range(nd.value):
            L.hirm_dataset_domain(ds, d, ctypes.byref(name), ctypes.byref(n))
            dnames.append(name.value.decode()); domains[dnames[-1]] = n.value
        schema, observations = {}, {}
        ar, nobs, doms = ctypes.c_int32(), ctypes.c_int64(), (ctypes.c_int32 * MAX_ARITY)()
        pi, pv = c_i32p(), c_u8p()
        for r in range(nr.value):
            L.hirm_dataset_relation(ds, r, ctypes.byref(name), ctypes.byref(ar), doms, ctypes.byref(nobs))
            rn = name.value.decode()
            schema[rn] = [dnames[doms[k]] for k in range(ar.value)]
            L.hirm_dataset_observations(ds, r, ctypes.byref(pi), ctypes.byref(pv))
            if nobs.value:
                ids = np.ctypeslib.as_array(pi, shape=(nobs.value * ar.value,)).reshape(nobs.value, ar.value).copy()
                val = np.ctypeslib.as_array(pv, shape=(nobs.value,)).copy()
            else:
                ids, val = np.zeros((0, ar.value), np.int32), np.zeros(0, np.uint8)
            observations[rn] = (ids, val)
        names = None
        if with_names:
            names = {}
            for d, dn in enumerate(dnames):
                names[dn] = []
                for c in range(domains[dn]):
                    L.hirm_dataset_entity(ds, d, c, ctypes.byref(name))
                    names[dn].append(name.value.decode())
        return domains, schema, observations, names
    finally:
        L.hirm_dataset_free(ds)


def load_dataset_gpu(schema_path, obs_path, device=-1, with_names=True):
    """GPU reader of <base>.schema / <base>.obs (csrc/ingest_gpu.cu): the same encoding as load_dataset, the COO arrays as
    torch CUDA tensors -- (domains, schema, observations {relation: (ids int32 [nobs, arity], val uint8 [nobs])}, names or
    None, timings dict).  The observations of a relation come in an unspecified order."""
    import torch
    L = load()
    ds = ctypes.c_void_p()
    if L.hirm_gpu_dataset_load(os.fsencode(schema_path), os.fsencode(obs_path), int(device), ctypes.byref(ds)):
        raise EngineError(L.hirm_gpu_dataset_last_error().decode())
    try:
        dev = torch.device("cuda", torch.cuda.current_device() if device < 0 else int(device))
        nd, nr = ctypes.c_int32(), ctypes.c_int32()
        L.hirm_gpu_dataset_counts(ds, ctypes.byref(nd), ctypes.byref(nr))
        name, n = ctypes.c_char_p(), ctypes.c_int32()
        dnames, domains = [], {}
        for d in range(nd.value):
            L.hirm_gpu_dataset_domain(ds, d, ctypes.byref(name), ctypes.byref(n))
            dnames.append(name.value.decode()); domains[dnames[-1]] = n.value
        schema, observations = {}, {}
        ar, nobs, doms = ctypes.c_int32(), ctypes.c_int64(), (ctypes.c_int32 * MAX_ARITY)()
        for r in range(nr.value):
            L.hirm_gpu_dataset_relation(ds, r, ctypes.byref(name), ctypes.byref(ar), doms, ctypes.byref(nobs), None, None)
            rn = name.value.decode()
            schema[rn] = [dnames[doms[k]] for k in range(ar.value)]
            ids = torch.empty((nobs.value, ar.value), dtype=torch.int32, device=dev)
            val = torch.empty(nobs.value, dtype=torch.uint8, device=dev)
            if L.hirm_gpu_dataset_copy_relation(ds, r, ids.data_ptr(), val.data_ptr()):
                raise EngineError(L.hirm_gpu_dataset_last_error().decode())
            observations[rn] = (ids, val)
        names = None
        if with_names:
            names = {}
            for d, dn in enumerate(dnames):
                names[dn] = []
                for c in range(domains[dn]):
                    L.hirm_gpu_dataset_entity(ds, d, c, ctypes.byref(name))
                    names[dn].append(name.value.decode())
        t = (ctypes.c_double * 4)()
        L.hirm_gpu_dataset_timings(ds, t)
        return domains, schema, observations, names, dict(h2d_s=t[0], split_parse_s=t[1], dictionary_s=t[2], coo_s=t[3])
    finally:
        L.hirm_gpu_dataset_free(ds)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def relation_reseat(labels, sizes, rel_table, scores, fresh_col, order, uniforms, alpha, pos):
    """hirm_relation_reseat (csrc/relation_moves.cc; host only): sizes and rel_table (int32 arrays) are updated in place.
    Returns (status, pos, moved, fresh_label): status 0 = done, 1 = the relation at order[pos] chose the fresh table."""
    L = load()
    assert labels.dtype == np.int32 and sizes.dtype == np.int32 and rel_table.dtype == np.int32 and order.dtype == np.int32
    assert scores.dtype == np.float64 and scores.flags.c_contiguous and uniforms.dtype == np.float64
    p, moved, fresh = ctypes.c_int32(int(pos)), ctypes.c_int32(0), ctypes.c_int32(-1)
    rc = L.hirm_relation_reseat(len(rel_table), len(labels), _p(labels, c_i32p), _p(sizes, c_i32p), _p(rel_table, c_i32p),
                                _p(scores, c_f64p), scores.shape[1], int(fresh_col), _p(order, c_i32p), _p(uniforms, c_f64p), len(order),
                                float(alpha), ctypes.byref(p), ctypes.byref(moved), ctypes.byref(fresh))
    if rc < 0:
        raise EngineError("hirm_relation_reseat: bad argument (%d)" % rc)
    return rc, p.value, moved.value, fresh.value


def plan_clusters(entries, tables, budget, forced=False):
    """hirm_plan_clusters (csrc/sweep_rounds.cc; host only): CTAs per IRM of a generic launch."""
    L = load()
    ent = np.ascontiguousarray(entries, dtype=np.float64)
    tab = None if tables is None else np.ascontiguousarray(tables, dtype=np.int32)
    out = np.ones(max(len(ent), 1), np.int32)
    total = L.hirm_plan_clusters(len(ent), _p(ent, c_f64p), None if tab is None else _p(tab, c_i32p), int(budget), int(bool(forced)), _p(out, c_i32p))
    if total < 0:
        raise EngineError("hirm_plan_clusters: bad argument")
    return out[:len(ent)].copy(), total


def plan_rounds(offsets, runs, seg, runcap=64):
    """hirm_plan_rounds (csrc/sweep_rounds.cc; host only): `offsets` [n + 1] of the move lists, `runs` = per list a sequence of
    (start, qualifies) pairs of its runs of equal domain, or None for "unknown".  Returns a list of rounds, each a list of
    (lo, hi, pre) per move list (lo == hi: nothing to do in that round)."""
    L = load()
    n = len(offsets) - 1
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    n_runs = np.zeros(max(n, 1), np.int32)
    start = np.zeros((max(n, 1), runcap), np.int64); ok = np.zeros((max(n, 1), runcap), np.int32)
    for k in range(n):
        if runs[k] is None:
            n_runs[k] = -1
            continue
        n_runs[k] = len(runs[k])
        for j, (st, q) in enumerate(runs[k][:runcap]):
            start[k, j] = st; ok[k, j] = int(bool(q))
    nr = ctypes.c_int32(0)
    if L.hirm_plan_rounds(n, _p(off, c_i64p), _p(n_runs, c_i32p), runcap, _p(start, c_i64p), _p(ok, c_i32p), int(seg), ctypes.byref(nr),
                          None, None, None, 0):
        raise EngineError("hirm_plan_rounds: bad argument")
    lo = np.zeros((nr.value, max(n, 1)), np.int64); hi = np.zeros_like(lo); pre = np.zeros((nr.value, max(n, 1)), np.int32)
    if L.hirm_plan_rounds(n, _p(off, c_i64p), _p(n_runs, c_i32p), runcap, _p(start, c_i64p), _p(ok, c_i32p), int(seg), ctypes.byref(nr),
                          _p(lo, c_i64p), _p(hi, c_i64p), _p(pre, c_i32p), nr.value):
        raise EngineError("hirm_plan_rounds: bad argument")
    return [[(int(lo[r, k]), int(hi[r, k]), int(pre[r, k])) for k in range(n)] for r in range(nr.value)]


class Engine:
    """One engine per GPU (hirm_engine_create)."""

    def __init__(self, device=-1):
        self.L = load()
        h = ctypes.c_void_p()
        self._h = None
        self._ck(self.L.hirm_engine_create(ctypes.byref(h), device))
        self._h = h
        self._schemas = {}

    def _ck(self, rc):
        if rc != 0:
            raise EngineError(self.L.hirm_last_error().decode())

    def close(self):
        if self._h is not None:
            self.L.hirm_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- schema / observations -------------------------------------------------------
    def irm_create(self, n_entities, rel_domains, max_tables=32):
        ne = np.ascontiguousarray(n_entities, dtype=np.int32)
        if np.isscalar(max_tables):
            max_tables = [max_tables] * len(ne)
        mt = np.ascontiguousarray(max_tables, dtype=np.int32)
        ar = np.ascontiguousarray([len(ds) for ds in rel_domains], dtype=np.int32)
        rd = np.zeros((max(1, len(rel_domains)), MAX_ARITY), dtype=np.int32)
        for r, ds in enumerate(rel_domains):
            rd[r, :len(ds)] = ds
        out = ctypes.c_int32()
        self._ck(self.L.hirm_irm_create(self._h, len(ne), _p(ne, c_i32p), _p(mt, c_i32p), len(rel_domains),
                                        _p(ar, c_i32p), _p(rd, c_i32p), ctypes.byref(out)))
        self._schemas[out.value] = ([int(x) for x in ne], [list(map(int, ds)) for ds in rel_domains])
        return out.value

    def irm_destroy(self, irm):
        self._ck(self.L.hirm_irm_destroy(self._h, irm))
        self._schemas.pop(irm, None)

    def set_observations(self, irm, rel, ids, val):
        a = len(self._schemas[irm][1][rel])
        ids = np.ascontiguousarray(ids, dtype=np.int32).reshape(-1, a)
        val = np.ascontiguousarray(val, dtype=np.uint8)
        assert len(ids) == len(val)
        self._ck(self.L.hirm_irm_set_observations(self._h, irm, rel, len(val), _p(ids, c_i32p), _p(val, c_u8p)))

    def set_observations_dense(self, irm, rel, slab0_ptr, pitch0, slab1_ptr, pitch1):
        self._ck(self.L.hirm_irm_set_observations_dense(self._h, irm, rel, slab0_ptr, pitch0, slab1_ptr, pitch1))

    def set_observations_dense_host(self, irm, rel, x):
        x = np.ascontiguousarray(x, dtype=np.uint8)
        self._ck(self.L.hirm_irm_set_observations_dense_host(self._h, irm, rel, x.ctypes.data, x.strides[0]))
        self.synchronize()

    def set_observations_dense_host_ptr(self, irm, rel, ptr, pitch):
        self._ck(self.L.hirm_irm_set_observations_dense_host(self._h, irm, rel, ptr, pitch))

    def set_table_mode(self, irm, rel, mode):
        """0 auto, 1 dense, 2 hashed count table for a CSR-stored relation (before finalize)."""
        self._ck(self.L.hirm_irm_set_table_mode(self._h, irm, rel, mode))

    def unary_block_create(self, n_entities, cols):
        """cols: list of (nobs, d_ids_ptr, d_val_ptr), one per unary relation -> block handle (bit rows built on the device)."""
        n = len(cols)
        nobs = np.ascontiguousarray([c[0] for c in cols], dtype=np.int64)
        ids = (ctypes.c_void_p * n)(*[c[1] or None for c in cols])
        val = (ctypes.c_void_p * n)(*[c[2] or None for c in cols])
        out = ctypes.c_int32()
        self._ck(self.L.hirm_unary_block_create(self._h, int(n_entities), n, _p(nobs, c_i64p), ids, val, ctypes.byref(out)))
        return out.value

    def unary_block_destroy(self, block):
        self._ck(self.L.hirm_unary_block_destroy(self._h, block))

    def set_observations_unary(self, irm, rel, block, col):
        self._ck(self.L.hirm_irm_set_observations_unary(self._h, irm, rel, block, col))

    def share_observations(self, dst, src):
        self._ck(self.L.hirm_irm_share_observations(self._h, dst, src))

    def finalize(self, irm):
        self._ck(self.L.hirm_irm_finalize(self._h, irm))

    # -- state -------------------------------------------------------------------------
    def set_assignments(self, irm, dom, labels):
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        assert len(labels) == self._schemas[irm][0][dom]
        self._ck(self.L.hirm_irm_set_assignments(self._h, irm, dom, _p(labels, c_i32p)))

    def seat_prior(self, irms, seed=0):
        """Seat every entity of every domain of the listed IRMs by sequential CRP(alpha) prior draws made on the device."""
        irms = np.ascontiguousarray(irms, dtype=np.int32)
        self._ck(self.L.hirm_engine_seat_prior(self._h, len(irms), _p(irms, c_i32p), int(seed)))

    def get_assignments(self, irm, dom):
        out = np.empty(max(1, self._schemas[irm][0][dom]), dtype=np.int32)
        self._ck(self.L.hirm_irm_get_assignments(self._h, irm, dom, _p(out, c_i32p)))
        return out[:self._schemas[irm][0][dom]]

    def set_alpha(self, irm, dom, alpha):
        self._ck(self.L.hirm_irm_set_alpha(self._h, irm, dom, float(alpha)))

    def get_counts(self, irm, rel):
        a = len(self._schemas[irm][1][rel])
        n = ctypes.c_int64()
        self._ck(self.L.hirm_irm_get_counts(self._h, irm, rel, None, 0, ctypes.byref(n)))
        out = np.empty((max(1, n.value), a + 2), dtype=np.int32)
        self._ck(self.L.hirm_irm_get_counts(self._h, irm, rel, _p(out, c_i32p), n.value, ctypes.byref(n)))
        return out[:n.value]

    def logp_score(self, irm):
        out = ctypes.c_double()
        self._ck(self.L.hirm_irm_logp_score(self._h, irm, ctypes.byref(out)))
        return out.value

    # -- the hot path --------------------------------------------------------------------
    def sweep(self, irms, moves, uniforms=None, seed=0, streams=None, counters=None, flags=0, trace_cap=0):
        """moves: list (one per irm) of (move_dom, move_item) arrays.
        Returns dict(choice=[...per irm], trace=[(n, labels, logps) per irm] or None)."""
        irms = np.ascontiguousarray(irms, dtype=np.int32)
        off = np.zeros(len(irms) + 1, dtype=np.int64)
        for k, (md, mi) in enumerate(moves):
            assert len(md) == len(mi)
            off[k + 1] = off[k] + len(md)
        nm = int(off[-1])
        md = np.ascontiguousarray(np.concatenate([np.asarray(m[0], dtype=np.int32) for m in moves]) if nm else np.zeros(0, np.int32))
        mi = np.ascontiguousarray(np.concatenate([np.asarray(m[1], dtype=np.int32) for m in moves]) if nm else np.zeros(0, np.int32))
        un = None
        if uniforms is not None:
            un = np.ascontiguousarray(np.concatenate([np.asarray(u, dtype=np.float64) for u in uniforms]))
            assert len(un) == nm
        st = None if streams is None else np.ascontiguousarray(streams, dtype=np.uint64)
        ct = None if counters is None else np.ascontiguousarray(counters, dtype=np.uint64)
        choice = np.empty(max(nm, 1), dtype=np.int32)
        tn = tl = tp = None
        if trace_cap > 0:
            tn = np.zeros(max(nm, 1), dtype=np.int32)
            tl = np.zeros((max(nm, 1), trace_cap), dtype=np.int32)
            tp = np.zeros((max(nm, 1), trace_cap), dtype=np.float64)
        self._ck(self.L.hirm_engine_sweep(self._h, len(irms), _p(irms, c_i32p), _p(off, c_i64p), _p(md, c_i32p),
                                          _p(mi, c_i32p), _p(un, c_f64p), seed, _p(st, c_u64p), _p(ct, c_u64p),
                                          flags, _p(choice, c_i32p), trace_cap, _p(tn, c_i32p), _p(tl, c_i32p),
                                          _p(tp, c_f64p)))
        res = dict(choice=[choice[off[k]:off[k + 1]] for k in range(len(irms))], trace=None)
        if trace_cap > 0:
            res["trace"] = [(tn[off[k]:off[k + 1]], tl[off[k]:off[k + 1]], tp[off[k]:off[k + 1]]) for k in range(len(irms))]
        return res

    def sweep_device(self, irms, d_move_off, d_move_dom, d_move_item, seed, d_stream, d_counter, flags=0, d_out_choice=None):
        irms = np.ascontiguousarray(irms, dtype=np.int32)
        self._ck(self.L.hirm_engine_sweep_device(self._h, len(irms), _p(irms, c_i32p), d_move_off, d_move_dom,
                                                 d_move_item, seed, d_stream, d_counter, flags, d_out_choice))

    def last_sweep_info(self):
        n, k, ms = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_float()
        self._ck(self.L.hirm_engine_last_sweep_info(self._h, ctypes.byref(n), ctypes.byref(k), ctypes.byref(ms)))
        c = ctypes.c_int32()
        self._ck(self.L.hirm_engine_last_cluster_size(self._h, ctypes.byref(c)))
        r, g = ctypes.c_int32(), ctypes.c_int32()
        self._ck(self.L.hirm_engine_last_rounds(self._h, ctypes.byref(r), ctypes.byref(g)))
        return dict(launches=n.value, kernel_kind=k.value, kernel_ms=ms.value, cluster_size=c.value, rounds=r.value, pregrouped=bool(g.value))

    def sweep_all(self, irms, n_sweeps=1, seed=0, flags=0, want_moves=False):
        """transition_cluster_assignments_all for every listed IRM, order generated on the device.
        Returns the number of moves, or (n_moves, dom, item, choice) when want_moves."""
        irms = np.ascontiguousarray(irms, dtype=np.int32)
        n = ctypes.c_int64()
        if not want_moves:
            self._ck(self.L.hirm_engine_sweep_all(self._h, len(irms), _p(irms, c_i32p), n_sweeps, seed, flags, 0, None, None, None, ctypes.byref(n)))
            return n.value
        total = sum(sum(self._schemas[int(h)][0]) for h in irms) * n_sweeps
        dom = np.empty(max(total, 1), np.int32); item = np.empty(max(total, 1), np.int32); choice = np.empty(max(total, 1), np.int32)
        self._ck(self.L.hirm_engine_sweep_all(self._h, len(irms), _p(irms, c_i32p), n_sweeps, seed, flags, total, _p(dom, c_i32p),
                                              _p(item, c_i32p), _p(choice, c_i32p), ctypes.byref(n)))
        return n.value, dom[:n.value], item[:n.value], choice[:n.value]

    def set_rng(self, irm, stream, move_counter=0):
        self._ck(self.L.hirm_irm_set_rng(self._h, irm, stream, move_counter))

    def set_rng_counters(self, irm, sweep_counter, alpha_counter):
        self._ck(self.L.hirm_irm_set_rng_counters(self._h, irm, int(sweep_counter), int(alpha_counter)))

    def logp_scores(self, irms):
        irms = np.ascontiguousarray(irms, dtype=np.int32)
        out = np.empty(max(len(irms), 1), np.float64)
        self._ck(self.L.hirm_engine_logp_scores(self._h, len(irms), _p(irms, c_i32p), _p(out, c_f64p)))
        return out[:len(irms)]

    def transition_alpha(self, irms, ngrid=20, seed=0, uniforms=None, trace=False):
        """CRP::transition_alpha for every domain of every listed IRM; returns alpha [n_irms, MAX_DOMAINS]
        (and the grid log-probabilities [n_irms, MAX_DOMAINS, ngrid] when trace)."""
        irms = np.ascontiguousarray(irms, dtype=np.int32)
        un = None
        if uniforms is not None:
            un = np.zeros((len(irms), MAX_DOMAINS), np.float64)
            u = np.asarray(uniforms, np.float64).reshape(len(irms), -1)
            un[:, :u.shape[1]] = u
        out = np.empty((max(len(irms), 1), MAX_DOMAINS), np.float64)
        tr = np.zeros((max(len(irms), 1), MAX_DOMAINS, ngrid), np.float64) if trace else None
        self._ck(self.L.hirm_engine_transition_alpha(self._h, len(irms), _p(irms, c_i32p), ngrid, seed, _p(un, c_f64p), _p(out, c_f64p), _p(tr, c_f64p)))
        return (out[:len(irms)], tr[:len(irms)]) if trace else out[:len(irms)]

    def get_alpha(self, irm, dom):
        out = ctypes.c_double()
        self._ck(self.L.hirm_irm_get_alpha(self._h, irm, dom, ctypes.byref(out)))
        return out.value

    def get_tables(self, irm, dom, cap=256):
        lab = np.empty(cap, np.int32); cnt = np.empty(cap, np.int32); n = ctypes.c_int32()
        self._ck(self.L.hirm_irm_get_tables(self._h, irm, dom, _p(lab, c_i32p), _p(cnt, c_i32p), cap, ctypes.byref(n)))
        return lab[:n.value].copy(), cnt[:n.value].copy()

    def relation_scores(self, src_irm, rel, cand_irms, cand_doms):
        """Relation.logp_score of (src_irm, rel) under each candidate IRM's partitions; cand_doms[k] = the candidate's
        domain index for every position of the relation."""
        ci = np.ascontiguousarray(cand_irms, dtype=np.int32)
        cd = np.zeros((max(len(ci), 1), MAX_ARITY), dtype=np.int32)
        for k, ds in enumerate(cand_doms):
            cd[k, :len(ds)] = ds
        out = np.empty(max(len(ci), 1), np.float64)
        self._ck(self.L.hirm_engine_relation_scores(self._h, src_irm, rel, len(ci), _p(ci, c_i32p), _p(cd, c_i32p), _p(out, c_f64p)))
        return out[:len(ci)]

    def set_observations_device(self, irm, rel, nobs, d_ids_ptr, d_val_ptr):
        """COO arrays already on the device (caller-owned buffers, e.g. torch tensors' data_ptr(); never copied)."""
        self._ck(self.L.hirm_irm_set_observations_device(self._h, irm, rel, int(nobs), d_ids_ptr, d_val_ptr))

    @staticmethod
    def _rel_arrays(rels):
        """rels: list of (domain indices per position, nobs, d_ids_ptr, d_val_ptr)."""
        n = len(rels)
        ar = np.ascontiguousarray([len(r[0]) for r in rels], dtype=np.int32)
        rd = np.zeros((max(n, 1), MAX_ARITY), dtype=np.int32)
        for k, r in enumerate(rels):
            rd[k, :len(r[0])] = r[0]
        nobs = np.ascontiguousarray([r[1] for r in rels], dtype=np.int64)
        ids = (ctypes.c_void_p * max(n, 1))(*[r[2] or None for r in rels])
        val = (ctypes.c_void_p * max(n, 1))(*[r[3] or None for r in rels])
        return ar, rd, nobs, ids, val

    def relation_score_matrix(self, rels, cand_irms, d_out_ptr=None):
        """Relation.logp_score of every relation under every candidate IRM's partitions -> [len(rels), len(cand_irms)]
        (numpy), or written to the device buffer `d_out_ptr` (float64, row-major) when given."""
        ci = np.ascontiguousarray(cand_irms, dtype=np.int32)
        out = None if d_out_ptr else np.zeros((len(rels), len(ci)), np.float64)
        if len(rels) == 0 or len(ci) == 0:
            return out
        ar, rd, nobs, ids, val = self._rel_arrays(rels)
        self._ck(self.L.hirm_engine_relation_score_matrix(self._h, len(rels), _p(ar, c_i32p), _p(rd, c_i32p), _p(nobs, c_i64p), ids, val,
                                                          len(ci), _p(ci, c_i32p), d_out_ptr if d_out_ptr else out.ctypes.data))
        return out

    def relation_scores_fresh(self, rels, n_entities, max_tables, keys, seed, alpha=1.0):
        """Relation.logp_score of every relation under its own fresh CRP(alpha)-prior partitions; keys [len(rels), n_domains]."""
        out = np.zeros(len(rels), np.float64)
        if len(rels) == 0:
            return out
        ar, rd, nobs, ids, val = self._rel_arrays(rels)
        ne = np.ascontiguousarray(n_entities, dtype=np.int32); mt = np.ascontiguousarray(max_tables, dtype=np.int32)
        ky = np.ascontiguousarray(keys, dtype=np.uint64).reshape(len(rels), len(ne))
        self._ck(self.L.hirm_engine_relation_scores_fresh(self._h, len(rels), _p(ar, c_i32p), _p(rd, c_i32p), _p(nobs, c_i64p), ids, val,
                                                          len(ne), _p(ne, c_i32p), _p(mt, c_i32p), float(alpha), int(seed), _p(ky, c_u64p),
                                                          _p(out, c_f64p)))
        return out

    def crp_prior_draw(self, n, key, seed, alpha=1.0, max_tables=64):
        out = np.empty(max(int(n), 1), np.int32)
        self._ck(self.L.hirm_engine_crp_prior_draw(self._h, int(n), float(alpha), int(max_tables), int(seed), int(key), _p(out, c_i32p)))
        return out[:int(n)]

    def clear_relation_cache(self):
        self._ck(self.L.hirm_engine_clear_relation_cache(self._h))

    def seat_entities(self, irm, dom, items, labels):
        it = np.ascontiguousarray(items, dtype=np.int32); lb = np.ascontiguousarray(labels, dtype=np.int32)
        assert len(it) == len(lb)
        self._ck(self.L.hirm_irm_seat_entities(self._h, irm, dom, len(it), _p(it, c_i32p), _p(lb, c_i32p)))

    def debug_score_delta(self, N, s, gn, gs):
        a = [np.ascontiguousarray(v, dtype=np.uint32) for v in (N, s, gn, gs)]
        out = np.empty(max(len(a[0]), 1), np.float64)
        self._ck(self.L.hirm_engine_debug_score_delta(self._h, len(a[0]), a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data,
                                                      a[3].ctypes.data, _p(out, c_f64p)))
        return out[:len(a[0])]

    def last_dense_plan(self):
        a, b, c, d = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        self._ck(self.L.hirm_engine_last_dense_plan(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d)))
        return dict(ctas_per_sm=a.value, stages=b.value, chunk_bytes=c.value, smem_bytes=d.value)

    def transpose_u8(self, d_in, rows, cols, pitch_in, d_out, pitch_out):
        self._ck(self.L.hirm_engine_transpose_u8(self._h, d_in, rows, cols, pitch_in, d_out, pitch_out))

    def debug_phase_cycles(self, d_buf_ptr):
        self._ck(self.L.hirm_engine_debug_phase_cycles(self._h, d_buf_ptr))

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.L.hirm_engine_set_stream(self._h, cuda_stream_ptr))

    def synchronize(self):
        self._ck(self.L.hirm_engine_synchronize(self._h))


class IrmBackend:
    """score/move/get_* view of one engine IRM (what tests/golden_util.replay drives)."""

    def __init__(self, engine, irm, flags=0):
        self.e, self.irm, self.flags = engine, irm, flags      # flags: e.g. SWEEP_CLUSTER (same results, cluster launch)

    def set_alpha(self, dom, alpha):
        self.e.set_alpha(self.irm, dom, alpha)

    def get_assignments(self, dom):
        return self.e.get_assignments(self.irm, dom)

    def get_counts(self, rel):
        return self.e.get_counts(self.irm, rel)

    def logp_score(self):
        return self.e.logp_score(self.irm)

    def transition_alpha(self, dom, u, ngrid):
        """CRP::transition_alpha of ONE domain with an injected uniform (the engine call moves every domain
        of the IRM: the others are put back).  Returns (alpha, grid log-probabilities)."""
        nd = len(self.e._schemas[self.irm][0])
        before = [self.e.get_alpha(self.irm, k) for k in range(nd)]
        un = np.zeros(MAX_DOMAINS); un[dom] = u
        alpha, lp = self.e.transition_alpha([self.irm], ngrid=ngrid, uniforms=[un], trace=True)
        for k in range(nd):
            if k != dom:
                self.e.set_alpha(self.irm, k, before[k])
        return float(alpha[0, dom]), lp[0, dom]

    def score(self, dom, item, cap=256):
        r = self.e.sweep([self.irm], [([dom], [item])], flags=SWEEP_SCORE_ONLY | self.flags, trace_cap=cap)
        n, labels, lp = r["trace"][0]
        return labels[0, :n[0]].copy(), lp[0, :n[0]].copy()

    def move(self, dom, item, u):
        r = self.e.sweep([self.irm], [([dom], [item])], uniforms=[[u]], flags=self.flags)
        return int(r["choice"][0][0])
