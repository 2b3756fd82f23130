"""World-size-2 gloo test (CPU) of the multi-GPU host logic: the rank -> chain / Philox-stream partition is a
disjoint cover that does not depend on the world size, IRMs are bin-packed, timing is the max over ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hirm import shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_chains, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    streams = shard.chain_streams(n_chains, rank, world)
    mine = torch.full((n_chains,), -1, dtype=torch.int64)
    mine[:len(streams)] = torch.tensor(streams, dtype=torch.int64)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    allv = sorted(int(v) for g in gathered for v in g if v >= 0)
    t = shard.max_over_ranks(10.0 + rank)            # rank 1 is the slow one
    owner, load = shard.pack_irms([5, 1, 9, 3, 3, 7], world)
    dist.barrier()
    if rank == 0:
        out.put((allv, t, owner, load))
    dist.destroy_process_group()


def test_two_ranks_partition_and_timing():
    world, n_chains = 2, 11
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_chains, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    allv, t, owner, load = q.get()
    assert allv == list(range(n_chains))              # disjoint cover, stream id = global chain index
    assert t == 11.0                                  # max over ranks
    assert sum(load) == 28 and max(load) <= 15 and len(owner) == 6 and set(owner) == {0, 1}


def test_blocks_are_balanced_and_world_size_independent():
    for n in (0, 1, 7, 592, 2368):
        for world in (1, 2, 4, 8):
            blocks = [shard.chain_block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
            assert sum((shard.chain_streams(n, r, world) for r in range(world)), []) == list(range(n))
