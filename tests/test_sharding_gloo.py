"""world_size-2 check (gloo, CPU) of the multi-GPU host logic: batch sharding covers every image once,
and the bench's max-/sum-over-ranks reductions agree across ranks. The data path has no collective."""
import os
import socket
import sys

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


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from unseenobjectswithmeanshift_b200 import sharding
    r, lr, w = sharding.init_from_env("gloo")
    assert (r, w) == (rank, world)
    b, e = sharding.shard_range(total, r, w)
    owned = torch.zeros(total, dtype=torch.int64)
    owned[b:e] = 1
    dist.all_reduce(owned)
    sharding.barrier()
    mx = sharding.max_over_ranks(10.0 + rank)
    sm = sharding.sum_over_ranks(e - b)
    q.put((rank, b, e, owned.tolist(), mx, sm))
    dist.destroy_process_group()


def test_two_rank_batch_sharding():
    world, total = 2, 17
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, b0, e0, owned0, mx0, sm0), (_, b1, e1, owned1, mx1, sm1) = res
    assert (b0, e0, b1, e1) == (0, 9, 9, 17)
    assert owned0 == owned1 == [1] * total          # every image owned by exactly one rank
    assert mx0 == mx1 == 11.0 and sm0 == sm1 == float(total)


def test_shard_range_properties():
    from unseenobjectswithmeanshift_b200.sharding import shard_range
    for total in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
