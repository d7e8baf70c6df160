"""world_size-2 gloo test of the rank plumbing bench.py uses for N > 1 (replica sharding, max-over-ranks timing)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from thinshelllab_b200 import dist as td
    mine = td.shard_sheets(5, rank, world_size)
    td.barrier()
    units, (t_dev, t_e2e) = td.aggregate(1000 * len(mine), [10.0 + 5 * rank, 30.0 - 7 * rank])
    q.put((rank, mine, units, t_dev, t_e2e))
    dist.destroy_process_group()


def test_two_rank_aggregation():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0][1] == [0, 2, 4] and out[1][1] == [1, 3]
    for _, _, units, t_dev, t_e2e in out:
        assert units == 5000.0              # SUM of units over ranks
        assert t_dev == 15.0 and t_e2e == 30.0   # MAX over ranks, per timer


def test_single_process_passthrough():
    from thinshelllab_b200 import dist as td
    assert td.world() == 1
    assert td.aggregate(7, [1.5, 2.5]) == (7.0, [1.5, 2.5])
