#!/usr/bin/env python
"""Benchmark of the HIRM entity-move hot path on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE.json config 3 -- one dense `bernoulli R e e` relation,
n = 20000 entities (4e8 observed cells), planted two-block structure (SURVEY.md 8d generator,
numpy default_rng(0)), `chains` independent seeded Gibbs chains per GPU sharing the
read-only observation slabs (800 MB in HBM, > L2), each chain making `moves_per_step`
sequential entity moves per step (default: one full sweep).  One step = one launch of the persistent
sweep kernel per shared-memory tier.

    python bench.py [--gpus N] [--steps K] [--warmup W]           # this engine
    python bench.py --impl reference ...                          # the reference's CPU code

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "hierarchical-irm_b200"))

METRIC = "gibbs_entity_moves_per_s"
UNIT = "moves/s"
REF_BENCH = os.path.join(ROOT, "oracle", "_ref", "ref_bench")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=20000)
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (0 = 32 per SM: five resident, the rest queued)")
    ap.add_argument("--moves-per-step", type=int, default=0, help="moves per chain per step (0 = n: one full sweep)")
    ap.add_argument("--max-tables", type=int, default=64)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--ref-n", type=int, default=1000, help="entities of the reference arm's bounded sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--burn-in", type=int, default=3, help="untimed sweeps from the CRP-prior start before the warm-up steps (state preparation)")
    ap.add_argument("--extra", default="c4,c5,ksweep,refsize", help="further records measured into config.extra_workloads: comma list of c4, c5 (BASELINE configs 4, 5), ksweep (dense kernel against live tables), refsize (config 3 at the reference arm's sample size); 'none' to skip")
    ap.add_argument("--c4-chains", type=int, default=0)
    ap.add_argument("--c4-moves", type=int, default=0, help="moves per domain per chain per step of the C4 record")
    ap.add_argument("--c5-chains", type=int, default=8)
    ap.add_argument("--c5-scale", type=float, default=1.0, help="scale factor on C5's entities / relations (1.0 = the stated size)")
    return ap.parse_args()


def workload_name(n):
    return "C3: dense 'bernoulli R e e', n=%d (%.1e cells), planted 2 blocks p=.9/.1, default_rng(0)" % (n, float(n) * n)


# ------------------------------------------------------------------------------------------
# data
# ------------------------------------------------------------------------------------------
def planted_ree_into(x, n, seed=0, chunk=1024):
    """SURVEY.md 8d / tests.engine_util.planted_ree, written row-chunked into x[:n, :n] (uint8)."""
    rng = np.random.default_rng(seed)
    zs = np.arange(n) >= n // 2
    for r0 in range(0, n, chunk):
        r1 = min(n, r0 + chunk)
        p = np.where(zs[r0:r1, None] != zs[None, :], 0.9, 0.1)
        x[r0:r1, :n] = rng.random((r1 - r0, n)) < p


def crp_prior_labels_fast(n, alpha, rng):
    """A draw from the sequential CRP(alpha) prior (Domain::incorporate, cxx/hirm.hh:194-203), vectorised:
    customer i opens a table w.p. alpha/(i+alpha), else copies the table of a uniformly random earlier customer."""
    i = np.arange(n)
    new = rng.random(n) < alpha / (i + alpha)
    new[0] = True
    parent = np.where(new, i, (rng.random(n) * np.maximum(i, 1)).astype(np.int64))
    root = parent.copy()
    while True:
        nxt = root[root]
        if np.array_equal(nxt, root):
            break
        root = nxt
    _, z = np.unique(root, return_inverse=True)
    return z.astype(np.int32)


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_id):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_id), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.3:
                continue
            f = [c.strip() for c in line.split(",")]
            try:
                sm.append(float(f[0])); smax = float(f[1])
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# the reference's CPU implementation (oracle/_ref/ref_bench links the unmodified reference)
# ------------------------------------------------------------------------------------------
def run_reference(n_ref, steps, warmup, per_round_s=4.0, procs=None):
    """steps+warmup rounds of the reference's own IRM::transition_cluster_assignment_item on the
    bounded sample, one single-threaded chain per host core (the reference has no threads)."""
    cores = os.cpu_count() or 1
    mem_gb = 420e-9 * n_ref * n_ref * 1.3
    try:
        avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 1e9
    except Exception:
        avail = 32.0
    procs = procs or max(1, min(cores, 32, int(avail * 0.5 / max(mem_gb, 1e-3))))
    if os.path.exists(REF_BENCH):
        kind = "reference"
        per = 100000
        cmds = [[REF_BENCH, str(n_ref), str(1 + p), str(steps + warmup), str(per), str(per_round_s)] for p in range(procs)]
        ps = [subprocess.Popen(c, stdout=subprocess.PIPE, text=True) for c in cmds]
        outs = [p.communicate()[0] for p in ps]
        moves = secs = 0.0
        rate = 0.0
        tables = []
        for out in outs:
            rows = [l.split() for l in out.strip().splitlines() if l.startswith("round")]
            rows = rows[warmup:]
            m = sum(int(r[3]) for r in rows); s = sum(float(r[5]) for r in rows)
            tables += [int(r[7]) for r in rows]
            moves += m; secs = max(secs, s)
            rate += m / s if s > 0 else 0.0
        ms_per_step = 1e3 * secs / max(steps, 1)
        sample = ("unmodified reference (cxx/hirm.hh IRM::transition_cluster_assignment_item) on the config-3 generator "
                  "at n=%d (%d cells): the reference needs ~420 B/observation and >0.2 s/move at n=20000, so the full "
                  "workload cannot be loaded; %d independent single-threaded chains, %.0f s per step; K~%d tables"
                  % (n_ref, n_ref * n_ref, procs, per_round_s, int(statistics.median(tables)) if tables else 0))
    else:
        # the compiled reference did not travel: time the C oracle port instead
        kind = "port"
        from oracle import hirm_oracle as O
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import engine_util as U
        x = U.planted_ree(n_ref, seed=0)
        ids, val = U.dense_to_coo(x)
        o = O.Oracle([n_ref], [[0, 0]])
        o.set_observations(0, ids, val)
        o.set_assignments(0, crp_prior_labels_fast(n_ref, 1.0, np.random.default_rng(1)))
        order = np.random.default_rng(2).permutation(n_ref).astype(np.int32)
        m = 0; t0 = time.time()
        while time.time() - t0 < per_round_s * (steps + warmup) / 2 and m < len(order):
            o.move(0, int(order[m]), 0.5); m += 1
        secs = time.time() - t0
        procs, rate, ms_per_step = 1, m / secs, 1e3 * secs / max(steps, 1)
        sample = "plain-C oracle port (oracle/hirm_oracle.c) at n=%d, 1 thread, %d moves" % (n_ref, m)
    return dict(value=rate, unit=UNIT, cores=procs, kind=kind, sample=sample), ms_per_step


def main_reference(args, rank, world):
    if rank != 0:
        return 0
    cb, ms = run_reference(args.ref_n, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.n), "sample_n": args.ref_n},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------
# this engine
# ------------------------------------------------------------------------------------------
def main_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from hirm import engine as E

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    props = torch.cuda.get_device_properties(dev)
    n = args.n
    M = args.moves_per_step or n                  # one step = one full sweep of every chain
    # 5 chains are resident per SM, the rest queue behind them: the launch ends with its slowest chain, so the deeper the queue the
    # smaller the tail (measured: 8 per SM 121.8 M moves/s, 16: 124.7 M, 25: 135.1 M, 32: 139.1 M, 40: 138.9 M, 50: 140.2 M)
    chains = args.chains or 32 * props.multi_processor_count
    pitch = (n + 15) // 16 * 16

    # ---- ingest: observations from pinned host memory (timed once, reported in config) -------
    x_host = torch.full((n, pitch), 0xFF, dtype=torch.uint8).pin_memory()
    planted_ree_into(x_host.numpy(), n, seed=0)
    eng = E.Engine(local_rank)
    h0 = eng.irm_create([n], [[0, 0]], max_tables=args.max_tables)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.set_observations_dense_host_ptr(h0, 0, x_host.data_ptr(), pitch)
    eng.finalize(h0)
    eng.synchronize()
    ingest_s = time.perf_counter() - t0
    hs = [h0]
    for c in range(1, chains):
        h = eng.irm_create([n], [[0, 0]], max_tables=args.max_tables)
        eng.share_observations(h, h0)
        hs.append(h)
    rng = np.random.default_rng(1000 + args.seed + rank)
    t0 = time.perf_counter()
    from hirm import shard
    stream_ids = shard.chain_streams(world * chains, rank, world)          # global chain index = Philox stream, whatever N
    for c, h in enumerate(hs):
        eng.set_rng(h, stream_ids[c], 0)
    eng.seat_prior(hs, seed=args.seed)            # Domain::incorporate: one sequential CRP(1) prior draw per chain, made on the device
    k0 = [len(np.unique(eng.get_assignments(h, 0))) for h in hs[:8]]
    eng.logp_scores(hs)                           # forces every chain's count-table build
    build_s = time.perf_counter() - t0

    # ---- move lists resident in HBM: every chain walks its own fresh random permutation per sweep ----------
    B = max(0, args.burn_in)
    nsteps = B + args.warmup + args.steps
    gen = torch.Generator(device=dev); gen.manual_seed(7 + args.seed + rank)
    d_items = []
    for s in range(nsteps):
        perm = torch.argsort(torch.rand(chains, n, device=dev, generator=gen), dim=1).to(torch.int32)
        d_items.append(perm[:, :M].contiguous().view(-1))
    irm_ids = np.asarray(hs, dtype=np.int32)
    d_off = torch.arange(chains + 1, dtype=torch.int64, device=dev) * M
    d_dom = torch.zeros(chains * M, dtype=torch.int32, device=dev)
    d_stream = torch.tensor(stream_ids, dtype=torch.int64, device=dev)
    d_ctr = [torch.full((chains,), s * M, dtype=torch.int64, device=dev) for s in range(nsteps)]
    torch.cuda.synchronize()

    def device_step(s):
        eng.sweep_device(irm_ids, d_off.data_ptr(), d_dom.data_ptr(), d_items[s].data_ptr(), args.seed,
                         d_stream.data_ptr(), d_ctr[s].data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- HBM-resident timing: W warm-up sweeps (the first one is the burn-in from the CRP prior), then exactly
    #      K sweeps between CUDA events ----
    # burn-in (state preparation, untimed): the chains leave the CRP-prior start (~10 tables) for the posterior (2-3 tables)
    burn_ms, burn_all_ms = None, []
    for s in range(B + args.warmup):
        device_step(s)
        if s < max(B, 1):
            torch.cuda.synchronize()
            burn_all_ms.append(eng.last_sweep_info()["kernel_ms"])
            if s == 0:
                burn_ms = burn_all_ms[0]
    barrier()
    launches_per_step = eng.last_sweep_info()["launches"] if args.warmup else 1
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.time()
    ev0.record()
    for s in range(B + args.warmup, nsteps):
        device_step(s)
    ev1.record()
    torch.cuda.synchronize()
    tw1 = time.time()
    elapsed_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(tw0, tw1)
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    info = eng.last_sweep_info()
    plan = eng.last_dense_plan()
    moves_per_gpu_step = chains * M
    value = world * moves_per_gpu_step * args.steps / (elapsed_ms * 1e-3)

    # roofline of the dominant kernel: algorithmic bytes = 2 rows of n bytes per move
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    launch_ms = elapsed_ms / args.steps
    alg_bytes = 2.0 * n * moves_per_gpu_step
    achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel": "dense_ree_sweep_kernel",
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": launch_ms,
                "launch_timing": "CUDA events around the K timed steps on the engine's stream; one step = the sweep kernel's "
                                 "tier launches (the second tier exits at once when no chain overflowed)",
                "frac_of_nominal_8000": achieved / 8000.0}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            # NOT measured in this run: DRAM bytes per move of an earlier `ncu --set full` capture of the same kernel
            # (profiles/traffic.json) times this launch's move count
            tj = json.load(open(tpath))
            roofline["traffic"] = tj.get("dram_bytes_per_move", 0) * moves_per_gpu_step or None
            roofline["traffic_source"] = "extrapolated from " + str(tj.get("source"))
            if roofline["traffic"]:
                roofline["dram_gbs_from_traffic"] = roofline["traffic"] / (launch_ms * 1e-3) / 1e9
                roofline["dram_frac_from_traffic"] = roofline["dram_gbs_from_traffic"] / peak
        except Exception:
            pass

    # ---- end to end through the public C ABI, as the reference's loop makes it (cxx/hirm.cc:39-59) -------------
    # per step: transition_cluster_assignments_all of every chain (hirm_engine_sweep_all: HOST irm ids in, sweep
    # order generated on the device), CRP::transition_alpha of every chain (alphas back to the host), and
    # IRM::logp_score of every chain (scores back to the host) -- the per-iteration work and read-backs of
    # inference_irm with --verbose.  Observations were ingested once (IRM.incorporate precedes the loop).
    def host_step(s):
        eng.sweep_all(hs, n_sweeps=1, seed=args.seed + 1)
        al = eng.transition_alpha(hs, ngrid=20, seed=args.seed + 1)
        return al, eng.logp_scores(hs)

    host_step(0)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        al, sc = host_step(s)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": world * chains * n * args.steps / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": int(3 * chains * 4 + (chains + 1) * 8 + 3 * 3 * chains * 8),
           "d2h_bytes_per_step": int(chains * 8 * 8 + chains * 8 + 2 * chains * 4),
           "ms_per_step": 1e3 * e2e_s / args.steps,
           "note": "public API with host buffers: hirm_engine_sweep_all + hirm_engine_transition_alpha + hirm_engine_logp_scores "
                   "per step (irm ids / Philox positions in; alphas, scores and kernel error words out); the 400 MB observation "
                   "matrix was ingested once before the loop (H2D + device transpose: %.3f s)" % ingest_s}

    # the same step through the explicit-move-list entry point (host move lists in, every chosen table out)
    host_moves = [(np.zeros(M, np.int32), np.ascontiguousarray(rng.permutation(n).astype(np.int32)[:M])) for _ in hs]
    streams = np.asarray(stream_ids, dtype=np.uint64)
    eng.sweep(hs, host_moves, seed=args.seed + 2, streams=streams, counters=np.zeros(chains, np.uint64))
    t0 = time.perf_counter()
    eng.sweep(hs, host_moves, seed=args.seed + 2, streams=streams, counters=np.full(chains, M, np.uint64))
    lists_s = time.perf_counter() - t0

    ktabs = [len(np.unique(eng.get_assignments(h, 0))) for h in hs[:8]]
    for h in hs:
        eng.irm_destroy(h)
    del d_items, d_dom, x_host
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs, each with its own roofline record (bench_workloads.py) ------------------------------
    extra = {}
    want = [w for w in args.extra.split(",") if w and w != "none"]

    def sync_max(x):
        torch.cuda.synchronize()
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    import bench_workloads as W
    if "c4" in want:
        rec = W.bench_c4(eng, dev, peak, chains=args.c4_chains, moves_per_domain=args.c4_moves, steps=max(1, min(args.steps, 2)), warmup=1,
                         seed=args.seed, world=world, rank=rank, barrier=barrier)
        ms = sync_max(rec["ms_per_step"])                     # replicas: the job's step ends with the slowest rank
        rec["n_gpus"], rec["scaling"] = world, "weak"
        rec["value"] = world * rec["chains_per_gpu"] * rec["moves_per_chain_per_step"] / (ms * 1e-3)
        extra["c4"] = rec
    if "refsize" in want and rank == 0:
        extra["c3_at_reference_sample_size"] = W.bench_c3_at(eng, dev, args.ref_n, seed=args.seed)
    if "ksweep" in want and rank == 0:
        extra["dense_k_sweep"] = W.bench_ksweep(eng, dev, peak, seed=args.seed)
    if world > 1:
        dist.barrier()
    if "c5" in want:
        from hirm.sharded import Comm
        sc = args.c5_scale
        rec = W.bench_c5(dev, peak, Comm(), chains=args.c5_chains, steps=max(1, min(args.steps, 2)), warmup=1,
                         n=int(200_000 * sc), n_unary=int(2000 * sc), n_binary=max(1, int(50 * sc)), bobs=int(10_000_000 * sc),
                         seed=args.seed, barrier=barrier, sync_max=sync_max)
        extra["c5"] = rec
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline, _ = run_reference(args.ref_n, 1, 1, per_round_s=8.0)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": launch_ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(n), "n_entities": n, "chains_per_gpu": chains,
                           "step": "one full Gibbs sweep (%d entity moves) of each of the %d independent chains per GPU" % (M, chains),
                           "moves_per_chain_per_step": M, "max_tables": args.max_tables,
                           "init": "sequential CRP(alpha=1) prior draw per chain, then %d untimed burn-in sweeps (state preparation), then the warm-up steps" % B,
                           "burn_in_sweeps": B, "burn_in_sweep_ms": burn_all_ms,
                           "tables_at_start_first_chains": k0, "tables_alive_first_chains": ktabs,
                           "burn_in_first_sweep_moves_per_s": (moves_per_gpu_step / (burn_ms * 1e-3)) if burn_ms else None,
                           "l2": "inputs (2 x %d MB slabs) larger than L2; every chain streams its own random permutation of rows" % (n * pitch // 1000000),
                           "parallelism": "%d independent chains per GPU (%d per SM), replicas across GPUs (no collective)" % (chains, plan["ctas_per_sm"]),
                           "sweeps_per_s": value / n, "ingest_s": ingest_s, "state_build_s": build_s,
                           "kernel_kind": info["kernel_kind"], "dense_plan": plan,
                           "explicit_move_lists_moves_per_s": moves_per_gpu_step / lists_s,
                           "extra_workloads": extra},
                "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
                "gpu_launches": int(launches_per_step * args.steps), "clocks": clocks}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return main_reference(args, rank, world)
    return main_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
