"""CPU, world_size 2 over gloo: the exchange a sharded get_close needs (positives summed, first
maximum of f0 in global row order) gives exactly what one rank would have computed alone."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from meshclust_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _single_rank_scan(f0, flags):
    # Trainer::get_close serial semantics: strict > from (NULL, -1): first maximum wins
    best, row = -1.0, -1
    for i, v in enumerate(f0):
        if v > best:
            best, row = v, i
    return (len(f0), int(flags.sum()), row, best)


def _worker(rank, world, port, f0_all, flags_all, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = f0_all.shape[1]
    lo, hi = sharding.shard_bounds(n, world, rank)
    local = [_single_rank_scan(f0_all[s, lo:hi], flags_all[s, lo:hi]) for s in range(f0_all.shape[0])]
    got = sharding.combine_scan_results(local, lo, torch.device("cpu"))
    # device-record path (one all-gather): same answer.  It assumes equal blocks per rank, so it is
    # fed blocks of the first `world * (n // world)` rows only.
    per = n // world
    rec = torch.zeros((f0_all.shape[0], 4), dtype=torch.int64)
    for s in range(f0_all.shape[0]):
        t = _single_rank_scan(f0_all[s, rank * per:(rank + 1) * per], flags_all[s, rank * per:(rank + 1) * per])
        rec[s, 0], rec[s, 1], rec[s, 2] = t[0], t[1], t[2]
        rec[s, 3] = int(np.float64(t[3]).view(np.int64))
    got2 = sharding.combine_scan_records(rec, per)
    if rank == 0:
        q.put((got, got2))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_scan_exchange(world):
    rng = np.random.default_rng(3)
    n, scans = 1001, 6
    f0 = rng.normal(0.2, 0.5, (scans, n))
    f0[1, [10, 700]] = 5.0          # tie across shards: the smaller global row must win
    f0[2, :] = -3.0                 # nothing above -1: no seed (row -1)
    f0[3, 900] = 7.0                # maximum in the last shard
    f0[4, 5] = np.nan               # NaN never wins
    flags = (f0 > 0.9).astype(np.uint8)
    want = [_single_rank_scan(f0[s], flags[s]) for s in range(scans)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, f0, flags, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, got2 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got == want
    per = n // world
    want2 = [_single_rank_scan(f0[s, :per * world], flags[s, :per * world]) for s in range(scans)]
    assert got2 == want2
    assert want[1][2] == 10 and want[2][2] == -1 and want[3][2] == 900


def test_shard_bounds_cover():
    for n in (1, 7, 100000, 1000003):
        for w in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
