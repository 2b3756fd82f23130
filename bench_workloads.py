"""Synthetic workloads of BASELINE.json configs 4 and 5 (SURVEY.md 8d generators), made on the GPU with torch (buffers only),
and the measurement of each through the engine's C ABI.  Used by bench.py (sub-records `config.extra_workloads`) and by the
full-size property tests (tests/test_full_size.py).

C4: sparse ternary `bernoulli R a b c`, 3 domains x 100 000 entities, 1e8 distinct observed cells uniform over the cube,
    planted 8 tables per domain, theta[8,8,8] ~ Beta(.5,.5).  Algorithmic bytes per move = deg * 9 (uint8 value + two int32
    other-entity ids per CSR visit): 2.7e9 B per sweep of 3e5 moves.
C5: 2000 unary (fully observed) + 50 binary `e x e` relations (1e7 random cells each) over 200 000 entities, 8 planted relation
    clusters each with its own 16-table entity partition; HIRM with its relation-cluster IRMs sharded over the ranks
    (hirm.sharded.ChainSet), 8 chains.  Algorithmic bytes per pass over all IRMs of one chain = 4e8 * 1 + 50 * 2 * 1e7 * 5
    = 5.4e9 B.
"""
import time

import numpy as np


def c4_ternary(dev, n=100_000, nobs=100_000_000, tables=8, seed=0):
    """-> (ids int32 [nobs, 3] on `dev`, val uint8 [nobs], planted z per domain (numpy))."""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + seed)
    rng = np.random.default_rng(seed)
    zt = [rng.integers(0, tables, n) for _ in range(3)]
    theta = torch.from_numpy(rng.beta(0.5, 0.5, size=(tables, tables, tables))).to(dev)
    key = torch.randint(0, n ** 3, (int(nobs * 1.01) + 1024,), device=dev, generator=g, dtype=torch.int64)
    key = torch.unique(key)                                  # distinct cells (Relation::incorporate asserts it, cxx/hirm.hh:272)
    key = key[torch.randperm(key.numel(), device=dev, generator=g)[:nobs]]
    a, b, c = key // (n * n), (key // n) % n, key % n
    del key
    z = [torch.from_numpy(x).to(dev) for x in zt]
    p = theta[z[0][a], z[1][b], z[2][c]]
    val = (torch.rand(p.shape, device=dev, generator=g, dtype=torch.float64) < p).to(torch.uint8)
    ids = torch.stack([a, b, c], dim=1).to(torch.int32).contiguous()
    return ids, val, zt


def c5_hirm(dev, n=200_000, n_unary=2000, n_binary=50, bobs=10_000_000, n_clusters=8, tables=16, seed=0):
    """-> (domains, schema, observations {relation: (ids tensor, val tensor)}, truth {relation: planted cluster}).
    The unary relations are fully observed and share ONE ids buffer (arange(n))."""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(4321 + seed)
    rng = np.random.default_rng(seed)
    zc = [torch.from_numpy(rng.integers(0, tables, n)).to(dev) for _ in range(n_clusters)]
    schema, obs, truth = {}, {}, {}
    ar = torch.arange(n, dtype=torch.int32, device=dev)
    for j in range(n_unary):
        c = j % n_clusters
        theta = torch.from_numpy(rng.beta(0.3, 0.3, size=tables)).to(dev)
        name = "u%04d" % j
        schema[name] = ["e"]
        obs[name] = (ar, (torch.rand(n, device=dev, generator=g, dtype=torch.float64) < theta[zc[c]]).to(torch.uint8))
        truth[name] = c
    for j in range(n_binary):
        c = j % n_clusters
        key = torch.unique(torch.randint(0, n * n, (int(bobs * 1.01) + 1024,), device=dev, generator=g, dtype=torch.int64))
        key = key[torch.randperm(key.numel(), device=dev, generator=g)[:bobs]]
        a, b = key // n, key % n
        theta = torch.from_numpy(rng.beta(0.3, 0.3, size=(tables, tables))).to(dev)
        val = (torch.rand(a.shape, device=dev, generator=g, dtype=torch.float64) < theta[zc[c][a], zc[c][b]]).to(torch.uint8)
        name = "b%04d" % j
        schema[name] = ["e", "e"]
        obs[name] = (torch.stack([a, b], dim=1).to(torch.int32).contiguous(), val)
        truth[name] = c
    return {"e": n}, schema, obs, truth


# ------------------------------------------------------------------------------------------------------------------
def bench_c4(eng, dev, peak_gbs, chains=0, moves_per_domain=0, steps=2, warmup=1, n=100_000, nobs=100_000_000, max_tables=64,
             seed=0, world=1, rank=0, barrier=None):
    """`chains` independent chains over one C4 data set (replicas across GPUs, no collective), started from device-drawn
    CRP(1) prior seats; one step = `moves_per_domain` entity moves in each of the three domains of every chain (explicit
    move lists resident in HBM, Philox uniforms).  Returns the sub-record."""
    import torch
    props = torch.cuda.get_device_properties(dev)
    chains = chains or 8 * props.multi_processor_count
    m = moves_per_domain or 4096
    t0 = time.perf_counter()
    ids, val, _ = c4_ternary(dev, n=n, nobs=nobs, seed=seed)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    h0 = eng.irm_create([n, n, n], [[0, 1, 2]], max_tables=max_tables)
    eng.set_observations_device(h0, 0, val.numel(), ids.data_ptr(), val.data_ptr())
    eng.finalize(h0)
    eng.synchronize()
    ingest_s = time.perf_counter() - t0
    hs = [h0]
    for c in range(1, chains):
        h = eng.irm_create([n, n, n], [[0, 1, 2]], max_tables=max_tables)
        eng.share_observations(h, h0)
        hs.append(h)
    t0 = time.perf_counter()
    for c, h in enumerate(hs):
        eng.set_rng(h, (rank * chains + c) | (1 << 40), 0)
    eng.seat_prior(hs, seed=seed)
    eng.logp_scores(hs)                                       # builds every chain's count table (histogram of the COO arrays)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    per = 3 * m
    g = torch.Generator(device=dev)
    g.manual_seed(99 + seed + rank)
    d_off = torch.arange(chains + 1, dtype=torch.int64, device=dev) * per
    d_dom = torch.arange(3, dtype=torch.int32, device=dev).repeat_interleave(m).repeat(chains).contiguous()
    d_stream = torch.tensor([(rank * chains + c) | (1 << 40) for c in range(chains)], dtype=torch.int64, device=dev)
    irm_ids = np.asarray(hs, dtype=np.int32)
    items, ctrs = [], []
    for s in range(warmup + steps):
        items.append(torch.randint(0, n, (chains * per,), device=dev, generator=g, dtype=torch.int32))
        ctrs.append(torch.full((chains,), s * per, dtype=torch.int64, device=dev))

    def step(s):
        eng.sweep_device(irm_ids, d_off.data_ptr(), d_dom.data_ptr(), items[s].data_ptr(), seed, d_stream.data_ptr(), ctrs[s].data_ptr())

    for s in range(warmup):
        step(s)
    if barrier:
        barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s in range(warmup, warmup + steps):
        step(s)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    info = eng.last_sweep_info()
    eng.logp_scores(hs[:1])                                   # surfaces a kernel error of the timed launches, if any
    tabs = [int(len(np.unique(eng.get_assignments(hs[0], d)))) for d in range(3)]
    deg = nobs / n
    bytes_per_move = deg * 9.0
    moves = chains * per
    achieved = moves * bytes_per_move / (ms * 1e-3) / 1e9
    rec = {"workload": "C4: sparse 'bernoulli R a b c', 3 x %d entities, %.1e cells (deg %.0f), planted 8 tables/domain, theta~Beta(.5,.5)" % (n, nobs, deg),
           "metric": "gibbs_entity_moves_per_s", "value_per_gpu": moves / (ms * 1e-3), "unit": "moves/s", "ms_per_step": ms,
           "steps": steps, "warmup": warmup, "chains_per_gpu": chains, "moves_per_chain_per_step": per,
           "step": "%d moves in each of the 3 domains of every chain (explicit move lists in HBM, uniform random entities)" % m,
           "init": "CRP(1) prior seats drawn on the device (hirm_engine_seat_prior); no burn-in", "tables_alive_chain0": tabs,
           "max_tables": max_tables, "kernel_kind": info["kernel_kind"], "gpu_launches": int(info["launches"] * steps),
           "rounds_per_step": int(info.get("rounds", 1)), "rows_grouped_ahead": bool(info.get("pregrouped", False)),
           "gen_s": gen_s, "ingest_s": ingest_s, "state_build_s": build_s,
           "roofline": {"bound": "hbm", "kernel": "pregroup_rows_kernel + generic_sweep_kernel" if info.get("pregrouped") else "generic_sweep_kernel",
                        "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                        "frac": achieved / peak_gbs, "traffic": None,
                        "algorithmic_bytes_per_move": bytes_per_move, "algorithmic_bytes_per_launch": moves * bytes_per_move,
                        "launch_ms": ms}}
    for h in hs:
        eng.irm_destroy(h)
    del ids, val
    torch.cuda.empty_cache()
    return rec


def bench_c5(dev, peak_gbs, comm, chains=8, steps=2, warmup=1, n=200_000, n_unary=2000, n_binary=50, bobs=10_000_000,
             max_tables=64, seed=0, barrier=None, sync_max=None):
    """BASELINE config 5: the HIRM sharded over the ranks of `comm`, `chains` chains started from a CRP prior seating of
    the relations (NOT the planted clusters); one step = one inference_hirm iteration of every chain (cxx/hirm.cc:61-90):
    relation moves from ONE all-gathered score matrix, then the entity sweeps + alpha moves of every IRM."""
    import torch
    from hirm.sharded import ChainSet
    t0 = time.perf_counter()
    domains, schema, obs, truth = c5_hirm(dev, n=n, n_unary=n_unary, n_binary=n_binary, bobs=bobs, seed=seed)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    h = ChainSet(domains, schema, obs, chains, seed=seed, max_tables=max_tables, comm=comm, device=dev.index)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    sync_max = sync_max or (lambda x: x)
    rows = []
    for it in range(warmup + steps):
        if barrier:
            barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        moved = sum(h.transition_relations())
        torch.cuda.synchronize()
        t_rel = sync_max(time.perf_counter() - t0)
        n_irms = sum(len(c.irms) for c in h.chains)
        if barrier:
            barrier()
        t0 = time.perf_counter()
        h.transition_irms()
        torch.cuda.synchronize()
        t_ent = sync_max(time.perf_counter() - t0)
        rows.append(dict(relation_step_s=t_rel, entity_step_s=t_ent, irms=n_irms, relations_moved=moved, entity_moves=n_irms * n))
    timed = rows[warmup:]
    ent_s = sum(r["entity_step_s"] for r in timed)
    rel_s = sum(r["relation_step_s"] for r in timed)
    moves = sum(r["entity_moves"] for r in timed)
    bytes_pass = n_unary * n * 1.0 + n_binary * 2.0 * bobs * 5.0          # SURVEY 8d, per chain
    achieved = chains * bytes_pass * steps / ent_s / 1e9 / comm.world      # per GPU
    timers = {k: sum(c.timers[k] for c in h.chains) for k in h.chains[0].timers}
    stats = {k: sum(c.stats[k] for c in h.chains) for k in h.chains[0].stats}
    scores = [float(s) for s in h.logp_scores()]
    rec = {"workload": "C5: HIRM, %d unary + %d binary (%.0e cells each) Bernoulli relations over %d entities, %d planted relation clusters"
                       % (n_unary, n_binary, bobs, n, 8),
           "metric": "gibbs_entity_moves_per_s", "value": moves / ent_s, "unit": "moves/s", "n_gpus": comm.world, "chains": chains,
           "scaling": "strong", "steps": steps, "warmup": warmup,
           "step": "one inference_hirm iteration of every chain: %d relation moves per chain from one all-gathered score matrix, then the "
                   "entity sweeps and alpha moves of every relation-cluster IRM" % (n_unary + n_binary),
           "init": "relations seated by the sequential CRP(1) prior (not the planted clusters); entity partitions CRP(1) prior draws",
           "ms_per_step": 1e3 * (ent_s + rel_s) / steps, "entity_step_ms": 1e3 * ent_s / steps, "relation_step_ms": 1e3 * rel_s / steps,
           "iterations_per_s": steps / (ent_s + rel_s), "irms_total": [r["irms"] for r in rows],
           "relations_moved": [r["relations_moved"] for r in rows], "split_s": timers, "events": stats,
           "collectives": {"all_gather": comm.n_allgather, "broadcast": comm.n_broadcast},
           "logp_score_per_chain": scores, "gen_s": gen_s, "setup_s": setup_s, "max_tables": max_tables,
           "roofline": {"bound": "hbm", "kernel": "generic_sweep_kernel", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                        "frac": achieved / peak_gbs, "traffic": None, "algorithmic_bytes_per_pass_per_chain": bytes_pass,
                        "note": "per GPU: bytes of all chains' entity passes / entity-step time / n_gpus"}}
    return rec


def bench_ksweep(eng, dev, peak_gbs, ks=(2, 4, 8, 16, 32, 64), n=20000, chains=0, moves=1000, max_tables=64, seed=0):
    """Dense R(e, e) kernel against the number of live tables (BASELINE config 3: "up to 64 clusters"): for every K a data
    set with K planted blocks (theta[K, K] ~ Beta(.5, .5)), chains started AT the planted partition (so K tables are alive),
    `moves` entity moves per chain from resident move lists.  Algorithmic bytes per move = 2 n (row + column of uint8 cells)."""
    import torch
    props = torch.cuda.get_device_properties(dev)
    chains = chains or 8 * props.multi_processor_count
    pitch = (n + 15) // 16 * 16
    out = []
    g = torch.Generator(device=dev)
    hs = []
    for K in ks:
        g.manual_seed(777 + seed + K)
        rng = np.random.default_rng(seed + K)
        zt = np.sort(rng.integers(0, K, n)).astype(np.int32)
        zt[:K] = np.arange(K)                                  # every block is populated
        theta = torch.from_numpy(rng.beta(0.5, 0.5, size=(K, K))).to(dev)
        zd = torch.from_numpy(zt.astype(np.int64)).to(dev)
        s0 = torch.full((n, pitch), 0xFF, dtype=torch.uint8, device=dev)
        for r0 in range(0, n, 2000):
            p = theta[zd[r0:r0 + 2000]][:, zd]
            s0[r0:r0 + 2000, :n] = (torch.rand(p.shape, device=dev, generator=g, dtype=torch.float32) < p).to(torch.uint8)
        s1 = torch.full((n, pitch), 0xFF, dtype=torch.uint8, device=dev)
        s1[:, :n] = s0[:, :n].t()
        torch.cuda.synchronize()
        h0 = eng.irm_create([n], [[0, 0]], max_tables=max_tables)
        eng.set_observations_dense(h0, 0, s0.data_ptr(), pitch, s1.data_ptr(), pitch)
        eng.finalize(h0)
        hs = [h0]
        for c in range(1, chains):
            h = eng.irm_create([n], [[0, 0]], max_tables=max_tables)
            eng.share_observations(h, h0)
            hs.append(h)
        for c, h in enumerate(hs):
            eng.set_assignments(h, 0, zt)
            eng.set_rng(h, c | (2 << 40), 0)
        eng.logp_scores(hs)
        d_off = torch.arange(chains + 1, dtype=torch.int64, device=dev) * moves
        d_dom = torch.zeros(chains * moves, dtype=torch.int32, device=dev)
        d_stream = torch.arange(chains, dtype=torch.int64, device=dev) + (2 << 40)
        irm_ids = np.asarray(hs, dtype=np.int32)
        ms = []
        for rep in range(3):
            items = torch.randint(0, n, (chains * moves,), device=dev, generator=g, dtype=torch.int32)
            ctr = torch.full((chains,), rep * moves, dtype=torch.int64, device=dev)
            eng.sweep_device(irm_ids, d_off.data_ptr(), d_dom.data_ptr(), items.data_ptr(), seed, d_stream.data_ptr(), ctr.data_ptr())
            torch.cuda.synchronize()
            ms.append(eng.last_sweep_info()["kernel_ms"])
        best = min(ms[1:])
        alive = int(len(np.unique(eng.get_assignments(hs[0], 0))))
        rate = chains * moves / (best * 1e-3)
        out.append({"planted_tables": int(K), "tables_alive_chain0_after": alive, "moves_per_s": rate, "kernel_ms": best,
                    "algorithmic_gbs": rate * 2.0 * n / 1e9, "frac_of_hbm_peak": rate * 2.0 * n / 1e9 / peak_gbs})
        for h in hs:
            eng.irm_destroy(h)
        del s0, s1
        torch.cuda.empty_cache()
    return {"workload": "C3 shape, n=%d, K planted blocks, chains started at the planted partition; %d chains x %d moves, kernel time (CUDA events), best of 2 after a warm-up launch"
                        % (n, chains, moves), "kernel": "dense_ree_sweep_kernel", "peak_gbs": peak_gbs, "rows": out}


def bench_c3_at(eng, dev, n, chains=0, burn_in=3, sweeps=5, max_tables=64, seed=0):
    """The headline workload's generator and state preparation at ANOTHER size `n` -- the size of the reference arm's bounded
    sample (`bench.py --impl reference` runs the unmodified reference at n = 1000: it cannot load n = 20 000), so that the
    two arms can also be compared like for like.  Same steps: planted two-block R(e, e), CRP(1)-prior seats, `burn_in` untimed
    sweeps, then `sweeps` timed full sweeps of every chain (device-generated order, Philox), kernel + launch time by CUDA events."""
    import torch
    props = torch.cuda.get_device_properties(dev)
    chains = chains or 32 * props.multi_processor_count
    rng = np.random.default_rng(seed)
    zs = np.arange(n) >= n // 2
    x = (rng.random((n, n)) < np.where(zs[:, None] != zs[None, :], 0.9, 0.1)).astype(np.uint8)
    h0 = eng.irm_create([n], [[0, 0]], max_tables=max_tables)
    eng.set_observations_dense_host(h0, 0, x)
    eng.finalize(h0)
    hs = [h0]
    for c in range(1, chains):
        h = eng.irm_create([n], [[0, 0]], max_tables=max_tables)
        eng.share_observations(h, h0)
        hs.append(h)
    for c, h in enumerate(hs):
        eng.set_rng(h, c | (3 << 40), 0)
    eng.seat_prior(hs, seed=seed)
    for _ in range(burn_in):
        eng.sweep_all(hs, n_sweeps=1, seed=seed)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(sweeps):
        eng.sweep_all(hs, n_sweeps=1, seed=seed)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    tabs = [int(len(np.unique(eng.get_assignments(h, 0)))) for h in hs[:8]]
    for h in hs:
        eng.irm_destroy(h)
    return {"workload": "C3 generator at n=%d (%d cells): the size of the reference arm's sample" % (n, n * n), "n_entities": n,
            "chains_per_gpu": chains, "sweeps_timed": sweeps, "burn_in_sweeps": burn_in, "value_per_gpu": chains * n * sweeps / (ms * 1e-3),
            "unit": "moves/s", "ms_per_sweep": ms / sweeps, "tables_alive_first_chains": tabs,
            "note": "through hirm_engine_sweep_all (host call per sweep, order generated on the device); compare with `--impl reference` at the same n"}
