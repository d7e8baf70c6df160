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


def test_strip_partition_geometry_tiles_the_global_sheet():
    """synthetic.strip_spec (host side of the strip partition): the owned rows of all ranks tile the global
    sheet, ghost rows replicate the neighbours' first / last owned rows (what tsl_dist.cu's halo exchange moves), strips start on even
    grid rows (the alternating triangulation of Cloth.init_mesh stays aligned) and every strip's table slab covers it"""
    import numpy as np
    from thinshelllab_b200.synthetic import strip_spec
    R, M, world = 16, 24, 4
    glob = strip_spec(world * R, M, 0, 1)
    gx = glob["cloth_pos"].copy(); gx[:, 0] += glob["x_shift"]
    row = M + 1
    specs = [strip_spec(R, M, r, world) for r in range(world)]
    owned = []
    for r, sp in enumerate(specs):
        x = sp["cloth_pos"].copy(); x[:, 0] += sp["x_shift"]
        assert sp["ghost_lo"] == (2 if r > 0 else 0) and sp["ghost_hi"] == (2 if r < world - 1 else 0)
        assert (r * R - sp["ghost_lo"]) % 2 == 0
        assert sp["own1"] - sp["own0"] == R * row and sp["rows"] == R + sp["ghost_lo"] + sp["ghost_hi"]
        owned.append(x[sp["own0"]:sp["own1"]])
        if r > 0:      # my lower ghost rows = the lower neighbour's last owned rows (send offset rows_local - 2 ghost_hi there)
            nb = specs[r - 1]; y = nb["cloth_pos"].copy(); y[:, 0] += nb["x_shift"]
            lo = (nb["rows"] - 2 * nb["ghost_hi"]) * row
            assert np.abs(x[:sp["ghost_lo"] * row] - y[lo:lo + nb["ghost_hi"] * row]).max() < 1e-15
        if r < world - 1:
            nb = specs[r + 1]; y = nb["cloth_pos"].copy(); y[:, 0] += nb["x_shift"]
            assert np.abs(x[(sp["rows"] - sp["ghost_hi"]) * row:] - y[nb["ghost_lo"] * row:2 * nb["ghost_lo"] * row]).max() < 1e-15
        tnx, tny, _ = sp["table_N"]
        ext = np.abs(sp["cloth_pos"][:, :2]).max(0)
        assert ext[0] <= 0.5 * (tnx - 1) * 0.003 and ext[1] <= 0.5 * (tny - 1) * 0.003
    assert np.abs(np.concatenate(owned) - gx).max() < 1e-15
    assert specs[0]["n_tris_global"] == 2 * (world * R - 1) * M
