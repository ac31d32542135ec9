"""CPU, world_size 2, gloo: the N>1 host logic (batch sharding, max-over-ranks timing, flat gradient bucket)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from esr_b200 import dist as ed
    r, w, _ = ed.init_from_env("gloo")
    lo, hi = ed.shard_range(11, w, r)
    slow = ed.max_over_ranks(10.0 + r)
    g = [torch.full((3, 2), float(r + 1)), torch.full((5,), float(10 * (r + 1)))]
    ed.flat_allreduce_(g, average=True)
    q.put((r, lo, hi, slow, g[0][0, 0].item(), g[1][0].item()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, s0, a0, b0), (r1, lo1, hi1, s1, a1, b1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 6, 6, 11)                # disjoint cover of the 11 units
    assert s0 == s1 == 11.0                                    # slowest rank wins
    assert a0 == a1 == 1.5 and b0 == b1 == 15.0                # averaged bucket


def test_shard_range_covers_everything():
    from esr_b200.dist import shard_range
    for n in (0, 1, 7, 8, 32, 33):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
