"""N > 1 path on CPU (gloo, world_size 2): the shard plan of the benchmark / scan driver (keys sorted by cost class and
dealt round-robin), the gather of the ranks' tables to rank 0 and the timing reduction."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from qboot_b200 import workload as wl

    spin, delta, S, P = bench.make_shard(rank, world, n * world)
    assert len(spin) == n
    # every rank can compute every other rank's share: the unsharded list, sorted by cost class, dealt round-robin
    keys = wl.sweep_keys(bench.PREC, bench.LAM, bench.NUMAX, bench.MAXSPIN, n * world)
    order = sorted(range(len(keys)), key=lambda i: (keys[i][0] != 0, i))
    fs, fd, fS, fP = wl.to_records([keys[i] for i in order[rank::world]], bench.PREC)
    assert np.array_equal(spin, fs) and np.array_equal(delta, fd) and np.array_equal(S, fS) and np.array_equal(P, fP)
    # cost classes are balanced: the ranks' counts of spin-0 tables differ by at most one
    z = torch.tensor([int((spin == 0).sum())], dtype=torch.int64)
    zs = [torch.zeros_like(z) for _ in range(world)]
    dist.all_gather(zs, z)
    assert max(int(x) for x in zs) - min(int(x) for x in zs) <= 1
    # the benchmark's data-path collective: every rank's (stand-in) tables gathered into rank 0's buffer, in rank order
    mine = torch.from_numpy(delta.view(np.uint8).reshape(-1).copy())
    got = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, got, dst=0)
    if rank == 0:
        for r in range(world):
            want = wl.to_records([keys[i] for i in order[r::world]], bench.PREC)[1]
            assert np.array_equal(got[r].numpy().view(np.uint32).reshape(want.shape), want)
    # checksum of checksums: the union of the shards is the unsharded list, no overlap
    h = torch.tensor([int(np.bitwise_xor.reduce(delta.view(np.uint32).ravel().astype(np.uint64) * np.uint64(2654435761) % np.uint64(2**61)))], dtype=torch.int64)
    hs = [torch.zeros_like(h) for _ in range(world)]
    dist.all_gather(hs, h)
    # the benchmark's reduction: max over ranks of the time, sum of the tables
    t = torch.tensor([10.0 + rank, float(n)], dtype=torch.float64)
    mx, sm = t.clone(), t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put((float(mx[0]), float(sm[1]), [int(x) for x in hs]))
    dist.barrier()
    dist.destroy_process_group()


def test_shards_are_disjoint_slices_and_reduction_is_max():
    world, n = 2, 200
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    mx, total, hashes = q.get(timeout=10)
    assert mx == 11.0 and total == 2 * n
    assert hashes[0] != hashes[1]


def test_workload_counts_match_the_reference_model():
    """the mixed-Ising key set at the north-star size: 6137 tables (SURVEY.md section 6 measured the same count)"""
    from qboot_b200 import workload as wl

    keys = wl.mixed_ising_keys(1024, 19, 20, 50)
    assert 6000 < len(keys) < 6300  # the reference build needs 6137 incl. its discrete sectors and finite ranges
    assert 1200 < len(wl.single_ising_keys(1024, 19, 20, 50)) < 1300  # reference: 1292
    w = wl.work_per_table(400, 19, 32, 2)
    assert abs(w["total"] - 2.34e7) / 2.34e7 < 0.01
    assert abs(wl.work_per_table(400, 19, 32, 0)["total"] - 2.20e7) / 2.20e7 < 0.01
    assert abs(wl.work_per_table(100, 11, 19, 2)["total"] - 1.90e6) / 1.90e6 < 0.01
    assert abs(wl.work_per_table(800, 27, 48, 2)["total"] - 1.31e8) / 1.31e8 < 0.01
