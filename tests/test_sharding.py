"""CPU: pair-list sharding and the final gather, world_size 2 over gloo (the N > 1 host logic)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vlg_b200.sharding import gather_results, shard_range
    lo, hi = shard_range(n_total, rank, world)
    # every rank "optimises" its shard: result rows are a function of the GLOBAL curve id only
    ids = torch.arange(lo, hi)
    omega = torch.stack([ids.float() * 10 + j for j in range(10)], dim=1).view(-1, 5, 2)
    energy = ids.float() ** 2
    full_omega = gather_results(omega, n_total)
    full_energy = gather_results(energy, n_total)
    if rank == 0:
        out["omega"] = full_omega
        out["energy"] = full_energy
    else:
        assert full_omega is None and full_energy is None
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [45, 8778, 2, 1])
def test_shard_and_gather_world2(n_total):
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), n_total, out), nprocs=2, join=True)
    ids = torch.arange(n_total).float()
    assert torch.equal(out["energy"], ids ** 2)
    assert out["omega"].shape == (n_total, 5, 2)
    assert torch.equal(out["omega"][:, 0, 0], ids * 10)


def test_shard_ranges_cover_and_balance():
    from vlg_b200.sharding import shard_range, shard_sizes
    for n in (0, 1, 7, 45, 8778, 100000):
        for world in (1, 2, 4, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = shard_sizes(n, world)
            assert max(sizes) - min(sizes) <= 1
