"""N>1 host path on CPU: world_size-2 gloo processes run the same plumbing bench.py uses under NCCL --
LUT broadcast from rank 0, item sharding, result reduction.  No GPU, no kernels."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_pkg
    pkg = load_pkg()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blob = pkg.parallel.broadcast_lut(torch, dist, rank, world, torch.device("cpu"))
    local = pkg.lut_blob()
    ok_blob = bool(np.array_equal(blob.numpy(), local)) and pkg.parallel.check_lut(blob.numpy())
    b, e = pkg.parallel.shard_range(n_items, rank, world)
    sums, maxes = pkg.parallel.reduce_stats(torch, dist, world, torch.device("cpu"), [e - b, (e - b) * 4960, b], [10.0 + rank])
    q.put((rank, ok_blob, b, e, sums, maxes))
    dist.destroy_process_group()


def test_world2_lut_broadcast_shard_reduce():
    world, n_items = 2, 1000003
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, ok0, b0, e0, s0, m0), (r1, ok1, b1, e1, s1, m1) = res
    assert ok0 and ok1                                   # rank 1 received exactly rank 0's blob
    assert b0 == 0 and e0 == b1 and e1 == n_items and abs((e0 - b0) - (e1 - b1)) <= 1
    assert s0 == s1 == [n_items, n_items * 4960, b1]
    assert m0 == m1 == [11.0]                            # max over ranks


def test_shard_range_partitions():
    from __graft_entry__ import load_pkg
    pkg = load_pkg()
    for n in (0, 1, 7, 8, 9, 1 << 20):
        for w in (1, 2, 4, 8):
            r = [pkg.parallel.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1
