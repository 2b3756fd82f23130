#!/usr/bin/env python
"""Benchmark of the HIRM entity-move hot path on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE.json config 3 -- one dense `bernoulli R e e` relation,
n = 20000 entities (4e8 observed cells), planted two-block structure (SURVEY.md 8d generator,
numpy default_rng(0)), `chains` independent seeded Gibbs chains per GPU sharing the
read-only observation slabs (800 MB in HBM, > L2), each chain making `moves_per_step`
sequential entity moves per step.  One step = one launch of the persistent sweep kernel.

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
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (0 = one per SM)")
    ap.add_argument("--moves-per-step", type=int, default=1000, help="moves per chain per step")
    ap.add_argument("--max-tables", type=int, default=64)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--ref-n", type=int, default=1000, help="entities of the reference arm's bounded sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    n, M = args.n, args.moves_per_step
    chains = args.chains or props.multi_processor_count
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
    for c, h in enumerate(hs):
        eng.set_assignments(h, 0, crp_prior_labels_fast(n, 1.0, rng))
    k0 = len(eng.get_counts(hs[0], 0))           # forces the count-table build of chain 0
    build_s = time.perf_counter() - t0

    # ---- move lists: every chain walks its own random permutation, M moves per step ------------
    total_steps = 2 * (args.warmup + args.steps) + 2
    perms = [rng.permutation(n).astype(np.int32) for _ in hs]

    def step_moves(s):
        idx = (np.arange(M) + s * M) % n
        return [p[idx] for p in perms]

    irm_ids = np.asarray(hs, dtype=np.int32)
    d_off = torch.arange(chains + 1, dtype=torch.int64, device=dev) * M
    d_dom = torch.zeros(chains * M, dtype=torch.int32, device=dev)
    d_stream = (torch.arange(chains, dtype=torch.int64, device=dev) + rank * chains)
    d_items = [torch.from_numpy(np.concatenate(step_moves(s))).to(dev) for s in range(args.warmup + args.steps)]
    d_ctr = [torch.full((chains,), s * M, dtype=torch.int64, device=dev) for s in range(args.warmup + args.steps)]
    torch.cuda.synchronize()

    def device_step(s):
        eng.sweep_device(irm_ids, d_off.data_ptr(), d_dom.data_ptr(), d_items[s].data_ptr(), args.seed,
                         d_stream.data_ptr(), d_ctr[s].data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- HBM-resident timing: W warm-up launches, then exactly K launches between CUDA events ----
    for s in range(args.warmup):
        device_step(s)
    barrier()
    sampler = ClockSampler(getattr(props, "uuid", None) and "GPU-" + str(props.uuid) or local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.time()
    ev0.record()
    for s in range(args.warmup, args.warmup + args.steps):
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
    kind = eng.last_sweep_info()["kernel_kind"]
    moves_per_gpu_step = chains * M
    value = world * moves_per_gpu_step * args.steps / (elapsed_ms * 1e-3)

    # roofline of the dominant (only) kernel: algorithmic bytes = 2 rows of n bytes per move
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
                "frac_of_nominal_8000": achieved / 8000.0}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            roofline["traffic"] = tj.get("dram_bytes_per_move", 0) * moves_per_gpu_step or None
            roofline["traffic_source"] = tj.get("source")
        except Exception:
            pass

    # ---- end to end through the C ABI with HOST buffers --------------------------------------
    # per step: host move lists -> device (inside hirm_engine_sweep), sweep, every chain's
    # assignments + every move's chosen table back to the host.
    base = args.warmup + args.steps
    host_moves = [[(np.zeros(M, np.int32), mv) for mv in step_moves(base + s)] for s in range(args.steps + 1)]
    streams = np.arange(chains, dtype=np.uint64) + rank * chains

    def host_step(s):
        r = eng.sweep(hs, host_moves[s], seed=args.seed, streams=streams,
                      counters=np.full(chains, (base + s) * M, dtype=np.uint64))
        zs = [eng.get_assignments(h, 0) for h in hs]
        return r, zs

    host_step(0)
    barrier()
    t0 = time.perf_counter()
    for s in range(1, args.steps + 1):
        host_step(s)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": world * moves_per_gpu_step * args.steps / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": int(moves_per_gpu_step * 8 + chains * 24 + (chains + 1) * 8),
           "d2h_bytes_per_step": int(moves_per_gpu_step * 4 + chains * (n * 4 + args.max_tables * 4)),
           "ms_per_step": 1e3 * e2e_s / args.steps,
           "note": "observations ingested once before (as IRM.incorporate precedes transition_cluster_assignments); "
                   "ingest of the 400 MB host matrix (H2D + device transpose) took %.3f s" % ingest_s}

    # number of tables now alive (explains the fp64 work per move)
    ktabs = [len(np.unique(eng.get_assignments(h, 0))) for h in hs[:8]]

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline, _ = run_reference(args.ref_n, 1, 1, per_round_s=8.0)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": launch_ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(n), "n_entities": n, "chains_per_gpu": chains,
                           "moves_per_chain_per_step": M, "max_tables": args.max_tables,
                           "tables_alive_first_chains": ktabs, "tables_at_start": int(k0),
                           "l2": "inputs (2 x %d MB slabs) larger than L2; every chain streams different rows" % (n * pitch // 1000000),
                           "parallelism": "%d independent chains per GPU, replicas across GPUs (no collective)" % chains,
                           "sweeps_per_s": value / n, "ingest_s": ingest_s, "state_build_s": build_s,
                           "kernel_kind": kind},
                "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
                "gpu_launches": args.steps, "clocks": clocks}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
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
