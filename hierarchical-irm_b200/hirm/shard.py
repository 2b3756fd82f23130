"""Multi-GPU partition of the entity-move path: one process per GPU, no data-path collective on the entity
sweeps (the one exchange step of an HIRM iteration, the relation move, is in hirm.sharded).

Moves inside one IRM are sequentially dependent; different IRMs (the relation clusters of an HIRM) and
different seeded chains are independent.  So ranks own disjoint sets of chains / IRMs:
  * chains: contiguous, balanced blocks; chain c always uses Philox stream c, whatever the world size,
    so a run is reproducible across 1, 2, 4, 8 GPUs;
  * IRMs of an HIRM: longest-processing-time bin packing by observation count (the sweep cost of an IRM
    is proportional to the observations its entities touch).
Timing follows the bench contract: barrier, device-timed region, MAX over ranks.
"""


def chain_block(n_chains, rank, world):
    """Chains [lo, hi) owned by `rank`: sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n_chains, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def chain_streams(n_chains, rank, world):
    """Philox stream id of every chain owned by `rank` (= the global chain index)."""
    lo, hi = chain_block(n_chains, rank, world)
    return list(range(lo, hi))


def pack_irms(weights, world):
    """LPT bin packing: returns owner[k] for every IRM k and the per-rank load."""
    load = [0] * world
    owner = [0] * len(weights)
    for k in sorted(range(len(weights)), key=lambda k: (-weights[k], k)):
        r = min(range(world), key=lambda r: (load[r], r))
        owner[k] = r
        load[r] += weights[k]
    return owner, load


def max_over_ranks(value, group=None):
    """MAX of a host float over the ranks of `group` (torch.distributed); the value itself without a group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
