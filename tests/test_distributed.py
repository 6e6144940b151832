"""world_size-2 gloo tests of the multi-GPU host logic (no GPU needed): session sharding and the gather of
per-session results.  The data path itself has no collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from eks_b200.parallel import gather_session_results, shard_indices


def test_shard_indices_partition():
    for n in (1, 2, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            parts = [shard_indices(n, r, world) for r in range(world)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_sessions, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        mine = shard_indices(n_sessions, rank, world)
        # stand-in for the per-session result (s_finals of K=3 keypoints): a function of the session index
        local = torch.tensor([[i + 0.25 * k for k in range(3)] for i in mine], dtype=torch.float64)
        full = gather_session_results(local, n_sessions)
        t = torch.tensor([float(len(mine))])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)   # the bench's max-over-ranks timing reduction
        q.put((rank, full.tolist(), t.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_sessions', [4, 5])
def test_two_rank_gather_gloo(n_sessions):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_sessions, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [[i + 0.25 * k for k in range(3)] for i in range(n_sessions)]
    for rank, full, mx in got:
        assert full == expect
        assert mx == float((n_sessions + 1) // 2)
