"""CPU, world_size 2 over gloo: the N>1 host logic of supermc_b200/launch.py -- event-id sharding, the
sum-allreduce of averaged-profile accumulators + event counters, and the rank-ordered merge of the
operation-9 tables (which must reproduce the 1-GPU file)."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_shard_ranges_partition_the_run():
    from supermc_b200.launch import shard_range
    for nev in (1, 7, 1000, 10**6 + 3):
        for world in (1, 2, 3, 8):
            r = [shard_range(nev, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == nev
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from supermc_b200.launch import allreduce_sums, shard_range, merge_rank_tables
    # accumulators: every rank holds the sum of f(event id) over its shard
    nev, G = 1001, 64
    lo, hi = shard_range(nev, rank, world)
    ids = np.arange(lo, hi, dtype=np.float64)
    buf = torch.from_numpy(np.outer(ids, np.ones(G)).sum(0).copy())
    total = allreduce_sums(buf, hi - lo, dist)
    assert total == nev
    assert np.allclose(buf.numpy() / total, np.arange(nev).mean())
    # operation-9 tables: rank r writes its rows; rank 0 merges in rank order
    d = os.path.join(tmp, "data") if rank == 0 else os.path.join(tmp, "data_rank%d" % rank)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "sn_ecc_eccp_10.dat"), "w") as f:
        for k in range(lo, hi):
            f.write("%16.8g\n" % k)
    dist.barrier()
    if rank == 0:
        merge_rank_tables(os.path.join(tmp, "data"), world)
        rows = np.loadtxt(os.path.join(tmp, "data", "sn_ecc_eccp_10.dat"))
        assert np.array_equal(rows, np.arange(nev))
        assert not os.path.exists(os.path.join(tmp, "data_rank1"))
    dist.destroy_process_group()


def test_world2_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
