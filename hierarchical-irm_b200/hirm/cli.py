"""`python -m hirm.cli` -- the reference's `hirm.out` command line on the GPU engine (cxx/hirm.cc:92-195):
same flags, same messages, same input and output files.

    python -m hirm.cli [--mode {irm,hirm}] [--seed 10] [--iters 10] [--verbose] [--timeout 0] [--load FILE] <path>

HIRM mode runs on hirm.sharded.ShardedHIRM (animals.unary, 10 iterations: 0.06 s against the reference's 0.11 s; under
torchrun one GPU per rank); --facade selects the Python mirror of the reference's classes instead.
"""
import argparse
import random
import sys
import time

from .model import HIRM, IRM
from . import util_io


def inference_irm(irm, iters, timeout, verbose):
    """cxx/hirm.cc:39-59: per iteration a sweep over every entity of every domain, then every domain's alpha."""
    t0 = time.time()
    for i in range(iters):
        if timeout > 0 and time.time() - t0 > timeout:           # CHECK_TIMEOUT, cxx/hirm.cc:15-24
            print("timeout after %1.2fs " % (time.time() - t0,))
            break
        irm.transition_cluster_assignments()
        irm.transition_crp_alphas(ngrid=20)
        if verbose:                                              # REPORT_SCORE, :26-37 (per sweep here, per move there)
            print("%f %f" % (time.time() - t0, irm.logp_score()))


def inference_hirm(hirm, iters, timeout, verbose):
    """cxx/hirm.cc:61-90: relation moves, then the entity sweeps and alpha moves of every IRM."""
    t0 = time.time()
    for i in range(iters):
        if timeout > 0 and time.time() - t0 > timeout:
            print("timeout after %1.2fs " % (time.time() - t0,))
            break
        hirm.transition_cluster_assignments(list(hirm.schema))
        hirm.transition_irms(ngrid=20)
        if verbose:
            print("%f %f" % (time.time() - t0, hirm.logp_score()))


def main_sharded(args, path_schema, path_obs, path_save):
    """inference_hirm (cxx/hirm.cc:61-90) on the sharded model; launched with torchrun it uses one GPU per rank."""
    import os
    from . import engine as E
    from .sharded import Comm, ShardedHIRM
    if args.mode != "hirm":
        print("--sharded is an HIRM mode")
        return 1
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    say = print if rank == 0 else (lambda *a, **k: None)
    import torch
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    say("setting seed to %d" % args.seed)
    say("loading schema from %s" % path_schema)
    say("loading observations from %s" % path_obs)
    domains, schema, obs, names = E.load_dataset(path_schema, path_obs)
    say("selected model is HIRM")
    kw = {}
    if args.load:
        say("loading clusters from %s" % args.load)
        kw["relation_tables"], kw["partitions"] = util_io.sharded_init_from_txt(args.load, domains, schema, names)
    else:
        say("incorporating observations")
    model = ShardedHIRM(domains, {r: schema[r] for r in sorted(schema)}, obs, seed=args.seed, max_tables=args.max_tables,
                        comm=Comm(), device=local, **kw)
    say("inferring %d iters; timeout %d" % (args.iters, args.timeout))
    t0 = time.time()
    for i in range(args.iters):
        # CHECK_TIMEOUT (cxx/hirm.cc:15-24); every rank must take the same branch, so rank 0's clock decides
        if args.timeout > 0 and model.comm.broadcast_obj(time.time() - t0 > args.timeout, 0):
            say("timeout after %1.2fs " % (time.time() - t0,))
            break
        model.iterate(1)
        if args.verbose:
            score = model.logp_score()
            say("%f %f" % (time.time() - t0, score))
    state = model.state()
    if rank == 0:
        print("saving to %s.hirm" % path_save)
        util_io.to_txt_hirm_state(path_save + ".hirm", state, schema, names)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main(argv=None):
    ap = argparse.ArgumentParser(prog="hirm", description="Run a hierarchical infinite relational model.")
    ap.add_argument("--mode", default="hirm", help="options are {irm, hirm}")
    ap.add_argument("--seed", type=int, default=10, help="random seed")
    ap.add_argument("--iters", type=int, default=10, help="number of inference iterations")
    ap.add_argument("--verbose", action="store_true", help="report results to terminal")
    ap.add_argument("--timeout", type=int, default=0, help="number of seconds of inference")
    ap.add_argument("--load", default="", help="path to .[h]irm file with initial clusters")
    ap.add_argument("--sharded", action="store_true",
                    help="HIRM mode on hirm.sharded.ShardedHIRM (the default for --mode hirm): the relation-cluster IRMs spread over "
                         "the ranks of a torchrun launch (one GPU each), relation moves from one all-gathered score matrix; native "
                         ".obs reader")
    ap.add_argument("--facade", action="store_true",
                    help="HIRM mode on the Python mirror of the reference's classes (hirm.model.HIRM): one relation move at a time, "
                         "entities seated only in the IRMs that observe them")
    ap.add_argument("--max-tables", type=int, default=64, help="(--sharded) table slots per domain kept on the device")
    ap.add_argument("path", nargs="?", help="base name of the .schema file")
    args = ap.parse_args(argv)
    if args.path is None:
        ap.print_help()
        return 1
    if args.mode not in ("hirm", "irm"):
        ap.print_help()
        print("unknown mode %s" % args.mode)
        return 1
    path_obs, path_schema = args.path + ".obs", args.path + ".schema"
    path_save = "%s.%d" % (args.path, args.seed)
    if args.sharded or (args.mode == "hirm" and not args.facade):
        return main_sharded(args, path_schema, path_obs, path_save)
    print("setting seed to %d" % args.seed)
    prng = random.Random(args.seed)
    print("loading schema from %s" % path_schema)
    schema = util_io.load_schema(path_schema)
    print("loading observations from %s" % path_obs)
    observations = util_io.load_observations(path_obs)
    if args.mode == "irm":
        print("selected model is IRM")
        if not args.load:
            model = IRM(schema, prng=prng)
            print("incorporating observations")
            for r, items, v in observations:
                model.incorporate(r, items, v)
        else:
            print("loading clusters from %s" % args.load)
            model = util_io.from_txt_irm(path_schema, path_obs, args.load, prng=prng)
        print("inferring %d iters; timeout %d" % (args.iters, args.timeout))
        inference_irm(model, args.iters, args.timeout, args.verbose)
        path_save += ".irm"
        print("saving to %s" % path_save)
        util_io.to_txt_irm(path_save, model)
        return 0
    print("selected model is HIRM")
    if not args.load:
        model = HIRM({r: schema[r] for r in sorted(schema)}, prng=prng)      # std::map order, cxx/hirm.hh:795
        print("incorporating observations")
        for r, items, v in observations:
            model.incorporate(r, items, v)
    else:
        print("loading clusters from %s" % args.load)
        model = util_io.from_txt_hirm(path_schema, path_obs, args.load, prng=prng)
    print("inferring %d iters; timeout %d" % (args.iters, args.timeout))
    inference_hirm(model, args.iters, args.timeout, args.verbose)
    path_save += ".hirm"
    print("saving to %s" % path_save)
    util_io.to_txt_hirm(path_save, model)
    return 0


if __name__ == "__main__":
    sys.exit(main())
