"""Sharded HIRM on a scaled BASELINE config 5: `NU` unary + `NB` binary Bernoulli relations over one domain of `NE`
entities, `NC` planted relation clusters each with its own planted entity partition; the relation-cluster IRMs are sharded
over the ranks (hirm.sharded.ShardedHIRM), relation moves from one all-gathered score matrix per iteration.

    python profiles/hirm_scale.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/hirm_scale.py
    env: NE NU NB BOBS (cells per binary relation) NC ITERS KCAP SEED CHAINS PLANTED (1: start from the planted relation clusters) BURN (entity-only steps before the timed iterations)

Prints one JSON line per rank-0 run: entity moves/s of the whole job (max over ranks of the device-synchronised wall
time of the entity step), the time split of the relation step, the number of collectives."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hierarchical-irm_b200"))
import numpy as np
import torch
import torch.distributed as dist

from hirm.sharded import ChainSet, Comm


def env(name, default):
    return int(os.environ.get(name, default))


def make(ne, nu, nb, bobs, nc, seed, tables=8):
    rng = np.random.default_rng(seed)
    zc = [rng.integers(0, tables, ne) for _ in range(nc)]
    schema, obs, truth = {}, {}, {}
    ar = np.arange(ne, dtype=np.int32)[:, None]
    for j in range(nu):
        c = j % nc
        theta = rng.beta(0.3, 0.3, size=tables)
        schema["u%04d" % j] = ["e"]
        obs["u%04d" % j] = (ar, (rng.random(ne) < theta[zc[c]]).astype(np.uint8))
        truth["u%04d" % j] = c
    for j in range(nb):
        c = j % nc
        cells = np.unique(rng.integers(0, ne * ne, size=int(bobs * 1.02), dtype=np.int64))[:bobs]
        ids = np.stack([cells // ne, cells % ne], axis=1).astype(np.int32)
        theta = rng.beta(0.3, 0.3, size=(tables, tables))
        schema["b%04d" % j] = ["e", "e"]
        obs["b%04d" % j] = (ids, (rng.random(len(ids)) < theta[zc[c][ids[:, 0]], zc[c][ids[:, 1]]]).astype(np.uint8))
        truth["b%04d" % j] = c
    return {"e": ne}, schema, obs, truth


def main():
    rank, world, local = env("RANK", 0), env("WORLD_SIZE", 1), env("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ne, nu, nb, bobs, nc = env("NE", 20000), env("NU", 200), env("NB", 8), env("BOBS", 1000000), env("NC", 8)
    iters, kcap, seed = env("ITERS", 3), env("KCAP", 32), env("SEED", 0)
    domains, schema, obs, truth = make(ne, nu, nb, bobs, nc, seed)
    t0 = time.perf_counter()
    rt = {r: truth[r] for r in schema} if env("PLANTED", 1) else None
    chains = env("CHAINS", 1)
    h = ChainSet(domains, schema, obs, chains, seed=seed, max_tables=kcap, comm=Comm(), relation_tables=rt, device=local)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    n_obs = sum(len(v[1]) for v in obs.values())

    def sync_max(x):
        torch.cuda.synchronize()
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(env("BURN", 2)):                         # entity-only steps first: the partitions of the planted clusters become informative
        h.transition_irms()
    rows = []
    for it in range(iters):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        moved = sum(h.transition_relations())
        t_rel = sync_max(time.perf_counter() - t0)
        n_irms = sum(len(c.irms) for c in h.chains)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        h.transition_irms()
        t_ent = sync_max(time.perf_counter() - t0)
        rows.append(dict(iter=it, relation_step_s=t_rel, entity_step_s=t_ent, irms=n_irms, relations_moved=moved,
                         entity_moves=n_irms * ne, loads=h.chains[0].loads()))
    score = h.logp_scores()
    if rank == 0:
        steady = rows[1:] or rows
        ent = sum(r["entity_moves"] for r in steady) / sum(r["entity_step_s"] for r in steady)
        print(json.dumps(dict(workload="scaled C5: %d unary + %d binary (%d cells each) relations over %d entities, %d planted clusters"
                                       % (nu, nb, bobs, ne, nc), n_gpus=world, chains=chains, observations=n_obs, setup_s=setup_s,
                              entity_moves_per_s=ent, relation_step_s=float(np.mean([r["relation_step_s"] for r in steady])),
                              timers={k: sum(c.timers[k] for c in h.chains) for k in h.chains[0].timers},
                              collectives=dict(all_gather=h.comm.n_allgather, broadcast=h.comm.n_broadcast),
                              logp_score=score, rows=rows)))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
